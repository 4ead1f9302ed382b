#!/bin/bash
# Development helper: time the fused kernel for each compiled (min-blocks, unroll) variant.
for v in 0 58 68 44 54 64 416 316; do
  F2_FUSED_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('variant $v fused %.2f ms frac %.3f step %.2f ms' % (r['kernel_ms'], r['frac'], d['ms_per_step']))"
done
