#!/bin/bash
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.3fM q/s  ms/step %.4f  e2e %.3fM  phases %s  frac %.3f step_frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], r['whole_step_frac']))"; }
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_parity.log 2>&1; rc=$?; tail -3 gpurun_out/pytest_parity.log
if [ $rc -ne 0 ]; then echo "parity failed rc=$rc"; exit 1; fi
timeout 150 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.log 2>&1; show gpurun_out/bench_a.log
for n in 1 2; do echo "NTK=$n"; NTK=$n PRIO="-3,-2,0" bash scripts/gpu_timeline2.sh 4 24 | grep "us/step"; done
NTK=2 PRIO="-3,-2,0" TRACE=1 bash scripts/gpu_timeline2.sh 12 12 | tail -8
