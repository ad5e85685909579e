#!/bin/bash
# ncu launch list (per-launch device time) + one full capture of the top kernels.  Numbers under ncu are never bench values.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_score_umma|k_topk_store|k_scan|k_score_simt|k_count|k_fill" -s 36 -c 6 -o gpurun_out/prof_full \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/
