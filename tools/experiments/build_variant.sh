#!/bin/bash
# Development helper: build libf2cnn_b200_<NW>_<U>.so with a given lane-kernel shape.
set -e
NW=$1; U=$2
cd "$(dirname "$0")/../f2cnn_b200/csrc"
OUT=../libf2cnn_b200_${NW}_${U}.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DF2_LANE_WARPS=$NW -DF2_LANE_UNROLL=$U"
mkdir -p /tmp/f2v_${NW}_${U}
for f in f2_capi f2_fused f2_lanes f2_prep f2_post; do nvcc $FLAGS -c $f.cu -o /tmp/f2v_${NW}_${U}/$f.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT /tmp/f2v_${NW}_${U}/*.o -cudart shared
echo built $OUT
