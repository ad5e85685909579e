#!/bin/bash
# cfg5s (12.5 M docs per GPU, 131,072 clusters per GPU, 1,250 owned queries per GPU, beam 100) on N GPUs, cluster-sharded
N=${1:-2}
mkdir -p gpurun_out
timeout ${BENCH_TIMEOUT:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline \
    --workload cfg5s --steps 240 --warmup 3 > gpurun_out/r02_bench_n${N}_cfg5s.json 2> gpurun_out/r02_bench_n${N}_cfg5s.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_n${N}_cfg5s.err | tail -4
python - gpurun_out/r02_bench_n${N}_cfg5s.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c = d['config']; e = d.get('e2e') or {}
print(sys.argv[1], 'value %.3fM q/s  us/step %.1f  e2e %s  exchange %s schedule %s step_frac %.3f' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, ('%.3fM' % (e['value'] / 1e6)) if e else None, c.get('exchange'), c.get('schedule'), d['roofline']['whole_step_frac']))
print('   checks', c.get('results_verified'), 'notes', c.get('notes'))
PY
