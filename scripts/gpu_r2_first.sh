#!/bin/bash
# Round 2, GPU call 1: parity suite (pipeline / fused variants first), smoke, the bench with its autotune, launch list + ncu of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sharded_p2p.py -q --timeout 150 > gpurun_out/r02_pytest_pipeline.log 2>&1; tail -15 gpurun_out/r02_pytest_pipeline.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_pipeline.py --deselect tests/test_gpu_sharded_p2p.py > gpurun_out/r02_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 1500 gpurun_out/r02_bench_default.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value %.3fM q/s  us/step %.2f  e2e %.3fM  kernel %s %.2f us frac %.3f  scoring alone %.2f us  step_frac %.3f' % (
        d['value'] / 1e6, d['ms_per_step'] * 1e3, d['e2e']['value'] / 1e6, r['kernel'][:24], r['kernel_ms'] * 1e3, r['frac'],
        r['scoring_alone']['kernel_ms'] * 1e3, r['whole_step_frac']))
    print(json.dumps(d['config']['launch_autotune']))
    print(d['config']['schedule'][:80], d['e2e']['copies_alone'], d['cpu_baseline'])
except Exception as e:
    print('bench line unreadable:', e)
PY
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_cfg2.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule fused > /dev/null 2> gpurun_out/r02_ncu_launches.err
timeout 300 $NCU --set full --import-source on -k regex:"k_score_topk_fused64" -s 6 -c 1 -o gpurun_out/r02_cfg2_fused64 \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule fused > /dev/null 2> gpurun_out/r02_ncu_full.err
ls -la gpurun_out/r02_*
