#!/bin/bash
# SM-partition experiment (green contexts): parity of the `partitioned` schedule, then its step time against `batches` on the same box.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 200 python -m pytest tests/test_gpu_pipeline.py::test_partitioned_schedule_equals_default -q --timeout 150 > gpurun_out/r02_pytest_partition.log 2>&1; grep -n "Error\|error\|passed\|failed" gpurun_out/r02_pytest_partition.log | cut -c1-400 | head -20
probe() { echo "== $*"; timeout 150 python bench.py --probe --gpus 1 --steps 1920 --warmup 3 --workload cfg2 --replicas 4 "$@" 2>&1 | grep -v "^$" | tail -3 | cut -c1-700; }
probe --schedule batches --pipeline 5 --launch-priorities on
probe --schedule partitioned --pipeline 5 --small-sms 64
probe --schedule partitioned --pipeline 5 --small-sms 56
probe --schedule partitioned --pipeline 5 --small-sms 72
probe --schedule partitioned --pipeline 5 --small-sms 48
probe --schedule partitioned --pipeline 5 --small-sms 64 --big-streams 1
probe --schedule partitioned --pipeline 8 --small-sms 64
probe --schedule partitioned --pipeline 5 --small-sms 64 --no-graph
