#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_masks.py tests/test_gpu_sharded.py -q --timeout 600 > gpurun_out/r02_pytest_quick.log 2>&1; tail -8 gpurun_out/r02_pytest_quick.log
timeout 300 python tools/bench_mask.py 2> gpurun_out/r02_bench_mask.err | tee gpurun_out/r02_bench_mask.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('position_mask', d['position_mask'])"
for w in cfg3 $EXTRA; do
  timeout 400 python bench.py --workload $w --steps 240 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 300 gpurun_out/r02_bench_$w.err
  python - gpurun_out/r02_bench_$w.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'us/step %.2f  kernel %s %.2f us frac %.3f  step_frac %.3f  e2e %.2fM' % (d['ms_per_step'] * 1e3, r['kernel'][:22], r['kernel_ms'] * 1e3, r['frac'], r['whole_step_frac'], d['e2e']['value'] / 1e6))
PY
done
