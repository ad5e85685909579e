#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 280 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 480 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N replica: value %.3fM q/s ms/step %.4f e2e %.3fM %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['config']['parallelism']))" || tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload cfg5s --steps 20 --warmup 3 > gpurun_out/bench_cfg5s_n$N.json 2> gpurun_out/bench_cfg5s_n$N.err
tail -1 gpurun_out/bench_cfg5s_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=$N cfg5s sharded: value %.3fM q/s ms/step %.4f e2e %.3fM phases %s frac %.3f cpu %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None))" || tail -5 gpurun_out/bench_cfg5s_n$N.err
