#!/bin/bash
mkdir -p gpurun_out
for v in $@; do
  GDR_UMMA_DEBUG=$v timeout 100 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('dbg=$v  ms/step %.4f  phases %s' % (d['ms_per_step'], {k: round(v*1000,1) for k,v in r['phase_ms'].items()}))"
done
