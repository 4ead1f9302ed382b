#!/bin/bash
# Development sweep: direct-form threshold (and kernel variant) of the fused kernel, resident bench only.
for cfg in "1e9 0" "0.25 0" "0.035 0" "0 0" "0.035 208" "0.035 1616"; do
  set -- $cfg
  echo "== F2_DIRECT_MIN_CY=$1 F2_FUSED_VARIANT=$2"
  F2_DIRECT_MIN_CY=$1 F2_FUSED_VARIANT=$2 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
