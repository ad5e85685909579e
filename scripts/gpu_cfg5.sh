#!/bin/bash
mkdir -p gpurun_out
for p in umma simt; do
timeout 400 python bench.py --workload cfg5s --steps 20 --warmup 3 --no-cpu-baseline --path $p --pipeline 1 > gpurun_out/bench_cfg5s_$p.json 2> gpurun_out/bench_cfg5s_$p.err
tail -1 gpurun_out/bench_cfg5s_$p.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$p value %.3fM q/s  ms/step %.4f  e2e %.3fM  phases %s  frac %.3f step_frac %.3f alg %.2f GB path %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k: round(v*1000,1) for k,v in r['phase_ms'].items()}, r['frac'], r['whole_step_frac'], r['algorithmic_bytes_per_launch']/1e9, d['path']))" || tail -3 gpurun_out/bench_cfg5s_$p.err
done
