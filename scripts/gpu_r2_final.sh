#!/bin/bash
# Round 2 closing call (1 GPU): the whole -m gpu suite, smoke(), both bench arms, cfg1, and the ncu evidence of the timed schedule.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get('roofline') or {}
    print(sys.argv[1], 'value %.3fM  us/step %.2f' % (d['value'] / 1e6, d['ms_per_step'] * 1e3), 'kernel', str(r.get('kernel'))[:40], 'kernel_us', r.get('kernel_ms') and round(r['kernel_ms'] * 1e3, 2),
          'frac', r.get('frac') and round(r['frac'], 3), 'step_frac', r.get('whole_step_frac') and round(r['whole_step_frac'], 3), 'e2e', d.get('e2e', {}).get('value'),
          'alone', (r.get('scoring_alone') or {}).get('kernel_ms'), 'clocks', d.get('clocks'))
    c = d.get('config', {})
    print('   schedule', str(c.get('schedule'))[:90]); print('   autotune', json.dumps(c.get('launch_autotune'))[:1400]); print('   cpu', json.dumps(d.get('cpu_baseline'))[:300])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_final.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 600 gpurun_out/r02_bench_default.err; show gpurun_out/r02_bench_default.json
if [ -n "$ALL" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; show gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --workload cfg1 --steps 960 --warmup 5 --no-cpu-baseline --no-autotune --schedule batches > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err; show gpurun_out/r02_bench_cfg1.json
fi
NCU="ncu --clock-control none --cache-control none"
timeout 200 $NCU --metrics gpu__time_duration.sum -s 70 -c 70 --csv --log-file gpurun_out/r02_launches_cfg2_partitioned.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule partitioned --small-sms 56 --ctas-per-sm 2 > /dev/null 2> gpurun_out/r02_ncu_launches.err
timeout 300 $NCU --set full --import-source on -k regex:"k_score_umma_x2|k_topk_fast" -s 8 -c 2 -o gpurun_out/r02_cfg2_partitioned_x2 \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule partitioned --small-sms 56 --ctas-per-sm 2 > /dev/null 2> gpurun_out/r02_ncu_full.err
if [ -n "$ALL" ]; then
timeout 200 $NCU --set full --import-source on -k regex:"k_score_tile_f32" -s 4 -c 1 -o gpurun_out/r02_cfg1_tile_f32 \
    python bench.py --workload cfg1 --steps 16 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule batches --pipeline 1 > /dev/null 2> gpurun_out/r02_ncu_full_cfg1.err
fi
ls -la gpurun_out/*.ncu-rep | tail -4; tail -3 gpurun_out/r02_ncu_full.err
