#!/bin/bash
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.3fM q/s  ms/step %.4f  e2e %.3fM  phases %s  frac %.3f step_frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], r['whole_step_frac']))"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 90 > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
timeout 150 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; show gpurun_out/bench_small.log
GDR_TOPK_WIDE=1 timeout 150 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_wide.log 2>&1; show gpurun_out/bench_wide.log
timeout 150 python bench.py --steps 480 --warmup 5 --no-cpu-baseline --workload cfg5s > gpurun_out/bench_cfg5s.log 2>&1; show gpurun_out/bench_cfg5s.log
