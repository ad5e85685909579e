#!/bin/bash
# full GPU test suite + smoke + all bench workloads (short) -> gpurun_out/
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.3fM q/s  ms/step %.4f  e2e %.3fM %s phases %s  frac %.3f kernel_ms %.4f step_frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e'].get('segments_ms_per_step'), {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], r['kernel_ms'], r['whole_step_frac']))"; }
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; show gpurun_out/bench_default.log
for w in cfg1 cfg3 cfg5s; do
  timeout 300 python bench.py --workload $w --steps 240 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; show gpurun_out/bench_$w.log
done
