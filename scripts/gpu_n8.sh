#!/bin/bash
# 8-GPU call: the NVLink tests, the sharded cfg2 line (what the driver's SCALE run measures) and the cfg5s line (12.5 M docs per GPU = 100 M docs).
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 400 python -m pytest "tests/test_gpu_sharded_p2p.py::test_p2p_sharded_over_nvlink" tests/test_gpu_sharded.py -q --timeout 300 > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -5 gpurun_out/r02_pytest_multi_n$N.log
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
run() { name=$1; shift
  timeout ${BENCH_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline "$@" \
      > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_n${N}_$name.err | tail -4; show gpurun_out/r02_bench_n${N}_$name.json
}
run p2p
run cfg5s --workload cfg5s --steps 240 --warmup 3
