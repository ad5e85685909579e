#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; print('$2 ms/step %.4f value %.3fM phases %s frac %.3f' % (d['ms_per_step'], d['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac']))
except Exception as e: print('$2 FAILED', e); print(open('$1.err').read()[-800:])
"; }
python bench.py --workload cfg1 --steps 480 --no-cpu-baseline > gpurun_out/t1.json 2> gpurun_out/t1.json.err; show gpurun_out/t1.json cfg1
python bench.py --steps 480 --no-cpu-baseline --path simt > gpurun_out/t2.json 2> gpurun_out/t2.json.err; show gpurun_out/t2.json cfg2-simt
python bench.py --workload cfg5s --steps 20 --no-cpu-baseline --pipeline 1 > gpurun_out/t3.json 2> gpurun_out/t3.json.err; show gpurun_out/t3.json cfg5s
