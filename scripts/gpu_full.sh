#!/bin/bash
# Full GPU validation: every -m gpu test, smoke, default bench (with CPU baseline), cfg3, mask bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>&1; tail -c 400 gpurun_out/bench_reference.json
timeout 300 python bench.py --workload cfg3 --steps 48 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 900 gpurun_out/bench_cfg3.json
timeout 300 python tools/bench_mask.py > gpurun_out/bench_mask.json 2> gpurun_out/bench_mask.err; cat gpurun_out/bench_mask.json | cut -c1-1200
