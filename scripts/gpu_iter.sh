#!/bin/bash
# parity (short timeouts) -> bench -> optional ncu of one kernel
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_parity.log 2>&1
rc=$?; tail -4 gpurun_out/pytest_parity.log
if [ $rc -ne 0 ]; then echo "parity failed/hung rc=$rc"; exit 1; fi
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.3fM q/s  ms/step %.4f  e2e %.3fM  phases %s  frac %.3f kernel_ms %.4f step_frac %.3f %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], r['kernel_ms'], r['whole_step_frac'], d['config'].get('schedule')))"; }
timeout 150 python bench.py --steps 960 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench.log 2>&1; show gpurun_out/bench.log
