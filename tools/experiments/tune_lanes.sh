#!/bin/bash
for v in 4_8 4_4 8_8; do
  F2CNN_B200_LIB=$PWD/f2cnn_b200/libf2cnn_b200_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('lanes $v kernel %.2f ms frac %.3f step %.2f ms' % (r['kernel_ms'], r['frac'], d['ms_per_step']))"
done
