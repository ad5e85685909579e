#!/bin/bash
# 2-GPU sanity run of the driver's SCALE path (cluster-sharded cfg2 per GPU, peer-to-peer exchange, every check on)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 960 --warmup 3 \
    > gpurun_out/r02_bench_n2_p2p_final.json 2> gpurun_out/r02_bench_n2_p2p_final.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_n2_p2p_final.err | tail -4
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n2_p2p_final.json') if l.startswith('{')][-1]); c = d['config']
print('value %.2fM q/s  us/step %.2f  e2e %.2fM' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, d['e2e']['value'] / 1e6), c.get('exchange'), c.get('schedule'), c.get('results_verified'))
PY
