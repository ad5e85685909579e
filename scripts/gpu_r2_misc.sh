#!/bin/bash
# Round 2: the rows beside the headline — integration test, mask / similarity kernels, cfg1 / cfg3 / cfg5s bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integration.py tests/test_gpu_masks.py tests/test_gpu_parity.py -q --timeout 600 > gpurun_out/r02_pytest_misc.log 2>&1; tail -8 gpurun_out/r02_pytest_misc.log
timeout 300 python tools/bench_mask.py > gpurun_out/r02_bench_mask.json 2> gpurun_out/r02_bench_mask.err; tail -c 1200 gpurun_out/r02_bench_mask.json; echo
timeout 300 python tools/bench_similarity.py > gpurun_out/r02_bench_similarity.json 2> gpurun_out/r02_bench_similarity.err; tail -c 900 gpurun_out/r02_bench_similarity.json; tail -3 gpurun_out/r02_bench_similarity.err
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
    print(sys.argv[1], 'us/step %.2f  kernel %s %.2f us frac %.3f  step_frac %.3f  e2e %.2fM' % (d['ms_per_step'] * 1e3, r['kernel'][:22], r['kernel_ms'] * 1e3, r['frac'], r['whole_step_frac'], d['e2e']['value'] / 1e6))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
for w in cfg1 cfg3 cfg5s; do
  timeout 400 python bench.py --workload $w --steps 240 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 300 gpurun_out/r02_bench_$w.err; show gpurun_out/r02_bench_$w.json
done
