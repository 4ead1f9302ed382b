#!/bin/bash
# Development helper: time the fused kernel for each compiled variant of launch_fused
# (needs a library built with -DF2_FUSED_TUNING: make -C f2cnn_b200/csrc clean all EXTRA=-DF2_FUSED_TUNING)
# (F2_FUSED_VARIANT: 0 = default <16 CTAs/SM, unroll 16>, 168 = <16, 8>, 208 = <20, 8>).
# profiles/r01b_tune_variants.log was made with the 128-thread kernel's variants of that time.
for v in 0 168 208; do
  F2_FUSED_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('variant $v fused %.2f ms frac %.3f step %.2f ms' % (r['kernel_ms'], r['frac'], d['ms_per_step']))"
done
