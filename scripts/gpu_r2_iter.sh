#!/bin/bash
# Round 2 iteration call: variant / p2p tests, parity suite, smoke, A/B against a second checkout in _r1/ (if present), the bench with its autotune, ncu captures.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
    print(sys.argv[1], 'us/step %.2f  kernel %s %.2f us frac %.3f  step_frac %.3f  e2e %.2fM  sm_mhz %s %s' % (d['ms_per_step'] * 1e3, r['kernel'][:22], r['kernel_ms'] * 1e3, r['frac'], r['whole_step_frac'], d['e2e']['value'] / 1e6, d['clocks']['sm_mhz'], d['clocks']['reasons']))
    if d['config'].get('launch_autotune'): print('   autotune', json.dumps(d['config']['launch_autotune']))
    if 'scoring_alone' in r: print('   scoring alone %.2f us' % (r['scoring_alone']['kernel_ms'] * 1e3))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
timeout 700 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sharded_p2p.py -q --timeout 150 > gpurun_out/r02_pytest_pipeline.log 2>&1; tail -15 gpurun_out/r02_pytest_pipeline.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_pipeline.py --deselect tests/test_gpu_sharded_p2p.py > gpurun_out/r02_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
[ -d _r1 ] && (cd _r1 && timeout 300 python bench.py --no-cpu-baseline --launch-priorities off --schedule batches > ../gpurun_out/ab_r1.json 2> ../gpurun_out/ab_r1.err); show gpurun_out/ab_r1.json
timeout 300 python bench.py --no-cpu-baseline --no-autotune --schedule batches > gpurun_out/ab_cur.json 2> gpurun_out/ab_cur.err; show gpurun_out/ab_cur.json
timeout 500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 800 gpurun_out/r02_bench_default.err; show gpurun_out/r02_bench_default.json
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_cfg2.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule fused > /dev/null 2> gpurun_out/r02_ncu_launches.err
timeout 300 $NCU --set full --import-source on -k regex:"k_score_topk_fused" -s 6 -c 1 -o gpurun_out/r02_cfg2_fused \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule fused > /dev/null 2> gpurun_out/r02_ncu_full.err
timeout 300 $NCU --set full --import-source on -k regex:"k_score_umma|k_topk_fast" -s 8 -c 2 -o gpurun_out/r02_cfg2_umma_topk \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule batches --pipeline 1 > /dev/null 2> gpurun_out/ab_ncu.err
ls -la gpurun_out/*.ncu-rep | tail -3
