#!/bin/bash
# usage: gpu_ncu_k.sh <kernel-regex> <outname> [bench args]   (env passes through)
mkdir -p gpurun_out
K=$1; O=$2; shift 2
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$K" -s 4 -c 1 -o gpurun_out/$O \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncu_$O.log 2>&1
tail -2 gpurun_out/ncu_$O.log | cut -c1-200
