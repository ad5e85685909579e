#!/bin/bash
# Round 2, first GPU call: the variants written without GPU access at the end of round 1 (ROADMAP.md "plan of record").
#   gpurun --timeout 900 -- 'bash scripts/gpu_experimental.sh'
# 1. grouped top-k, fused scoring + top-k and priority launches against the default, bit for bit (child processes, 75 s each),
#    then the variants the default test run leaves out (1 and 2 groups, fused with 5 and 3 groups)
# 2. if the fused kernel is exact: its cfg2 step time against the default schedule
# 3. the bench with and without per-launch priorities (the autotune prints both)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_zz_experimental.py -q -rxX 2>&1 | tee gpurun_out/experimental_tests.log
for v in "GDR_TOPK_GROUPS 1" "GDR_TOPK_GROUPS 2" "FUSED 5" "FUSED 3"; do
    timeout 100 python tests/_experimental_child.py $v 2>&1 | tail -1 | tee -a gpurun_out/experimental_tests.log
done
for g in 4 5 3; do
    timeout 300 python tools/bench_fused.py --groups $g 2>gpurun_out/bench_fused_g$g.err | tee gpurun_out/bench_fused_g$g.json
done
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_autotune.err | tee gpurun_out/bench_autotune.json
