#!/bin/bash
# multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded_p2p.py tests/test_gpu_sharded.py -q --timeout 300 > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -12 gpurun_out/r02_pytest_multi_n$N.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c = d['config']
    e = d.get('e2e') or {}
    print(sys.argv[1], 'value %.2fM q/s  us/step %.2f  e2e %s  exchange %s schedule %s graph %s step_frac %.3f' % (d['value'] / 1e6, d['ms_per_step'] * 1e3,
          ('%.2fM' % (e['value'] / 1e6)) if e else None, c.get('exchange'), str(c.get('schedule'))[:12], c.get('cuda_graph'), d['roofline']['whole_step_frac']))
    print('   checks', c.get('results_verified'), 'notes', c.get('notes'), 'e2e copies', e.get('copies'), 'e2e graph', e.get('cuda_graph'))
except Exception as ex:
    print(sys.argv[1], 'unreadable', ex)
PY
}
run() { # name, extra args
  name=$1; shift
  timeout ${BENCH_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline "$@" \
      > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  tail -c 600 gpurun_out/r02_bench_n${N}_$name.err | grep -v "^$" | tail -4; show gpurun_out/r02_bench_n${N}_$name.json
}
run p2p --exchange auto
if [ -z "$SHORT" ]; then
run nccl --exchange nccl
run replica --mode replica --no-autotune --schedule batches
fi
