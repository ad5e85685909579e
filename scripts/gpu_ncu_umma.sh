#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_score_umma" -s 6 -c 2 -o gpurun_out/prof_umma \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_umma.log 2>&1
tail -2 gpurun_out/ncu_umma.log | cut -c1-200
