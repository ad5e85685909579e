#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded_p2p.py -q --timeout 300 -x > gpurun_out/r02_pytest_fp32.log 2>&1; tail -6 gpurun_out/r02_pytest_fp32.log
timeout 400 python bench.py --workload cfg1 --steps 480 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err; tail -c 400 gpurun_out/r02_bench_cfg1.err
python - gpurun_out/r02_bench_cfg1.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'us/step %.2f  kernel %s %.2f us frac %.3f  step_frac %.3f  e2e %.2fM path %s' % (d['ms_per_step'] * 1e3, r['kernel'][:22], r['kernel_ms'] * 1e3, r['frac'], r['whole_step_frac'], d['e2e']['value'] / 1e6, d['path']))
PY
