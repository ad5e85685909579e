#!/bin/bash
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.3fM q/s  ms/step %.4f  e2e %.3fM  phases %s  frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac']))"; }
for C in 148 140 132 124 116 108; do
  GDR_UMMA_CTAS=$C timeout 100 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c$C.log 2>&1; show gpurun_out/bench_c$C.log
done
GDR_TOPK_NOREG=1 timeout 100 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_noreg.log 2>&1; show gpurun_out/bench_noreg.log
GDR_TOPK_NOREG=1 GDR_UMMA_CTAS=124 timeout 100 python bench.py --steps 960 --warmup 5 --no-cpu-baseline > gpurun_out/bench_noreg124.log 2>&1; show gpurun_out/bench_noreg124.log
for C in 148 124 108; do echo "phase-split CTAS=$C"; GDR_UMMA_CTAS=$C NTK=1 PRIO="-3,-2,0" bash scripts/gpu_timeline2.sh 4 24 | grep "us/step"; done
