#!/bin/bash
# full GPU suite + bench JSON lines for profiles/ (default run incl. CPU baseline, the other workloads) + ncu evidence
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/r01_bench_default.json; echo
for w in cfg1 cfg3 cfg5s; do
  timeout 300 python bench.py --workload $w --steps 480 --warmup 5 > gpurun_out/r01_bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 200 gpurun_out/r01_bench_$w.json; echo
done
timeout 300 python bench.py --pipeline 1 --no-cpu-baseline --steps 960 > gpurun_out/r01_bench_serial.json 2>/dev/null
bash scripts/gpu_profiles.sh
