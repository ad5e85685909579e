#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded_p2p.py tests/test_gpu_pipeline.py -q --timeout 200 > gpurun_out/dbg_pytest.log 2>&1; tail -4 gpurun_out/dbg_pytest.log
GDR_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline \
    --workload cfg5s --steps 240 --warmup 3 > gpurun_out/dbg_n${N}.json 2> gpurun_out/dbg_n${N}.err
grep "^\[rank\|Error\|error\|illegal\|gdr" gpurun_out/dbg_n${N}.err | head -30 | cut -c1-300
python - gpurun_out/dbg_n${N}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c = d['config']; e = d.get('e2e') or {}
    print(sys.argv[1], 'value %.3fM q/s  us/step %.1f  e2e %s  exchange %s schedule %s step_frac %.3f' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, ('%.3fM' % (e['value'] / 1e6)) if e else None, c.get('exchange'), c.get('schedule'), d['roofline']['whole_step_frac']))
    print('   checks', c.get('results_verified'), 'notes', c.get('notes'))
except Exception as ex:
    print('unreadable', ex)
PY
