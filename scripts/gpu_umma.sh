#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 50 -k "seeded_synthetic" > gpurun_out/pytest_umma.log 2>&1
rc=$?; tail -8 gpurun_out/pytest_umma.log
if [ $rc -ne 0 ]; then echo "UMMA tests failed/hung rc=$rc: skipping bench"; exit 1; fi
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 90 > gpurun_out/pytest_parity.log 2>&1; tail -5 gpurun_out/pytest_parity.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 120 python bench.py --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
