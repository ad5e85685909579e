#!/bin/bash
# bench JSON lines for profiles/ (default run incl. CPU baseline, the other workloads) + ncu evidence + memcheck
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/r01_bench_default.json
for w in cfg1 cfg3 cfg5s; do
  timeout 300 python bench.py --workload $w --steps 480 --warmup 5 > gpurun_out/r01_bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/r01_bench_$w.json; echo
done
bash scripts/gpu_profiles.sh
sed -i 's/for tool in memcheck racecheck/for tool in ${TOOLS:-memcheck racecheck}/' scripts/gpu_sanitize.sh
TOOLS=memcheck bash scripts/gpu_sanitize.sh
