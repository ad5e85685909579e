#!/bin/bash
# Round-1 evidence: launch list + full ncu captures of the dominant kernels.  Numbers under ncu are never bench values.
mkdir -p gpurun_out
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 40 -c 72 --csv --log-file gpurun_out/r01_launches_cfg2.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph --pipeline 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"k_score_umma|k_topk_fast|k_tilemeta" -s 9 -c 3 -o gpurun_out/r01_cfg2_umma_topk \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph --pipeline 1 > /dev/null 2>&1
if [ -n "$ALL" ]; then
timeout 400 $NCU --set full --import-source on -k regex:"k_score_simt" -s 2 -c 1 -o gpurun_out/r01_cfg5s_simt \
    python bench.py --workload cfg5s --steps 4 --warmup 3 --no-cpu-baseline --no-graph --pipeline 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"k_tree_mask" -s 4 -c 2 -o gpurun_out/r01_cfg4_tree_mask \
    python tools/bench_mask.py > /dev/null 2>&1
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches_cfg2.csv
