#!/bin/bash
# SM-partition experiment, second pass: smaller top-k side, launch priorities, more scoring streams / batches in flight.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
probe() { echo "== $*"; timeout 150 python bench.py --probe --gpus 1 --steps 1920 --warmup 3 --workload cfg2 --replicas 4 "$@" 2>&1 | grep -v "^$" | tail -2 | cut -c1-500; }
probe --schedule partitioned --pipeline 5 --small-sms 48
probe --schedule partitioned --pipeline 5 --small-sms 48 --launch-priorities on
probe --schedule partitioned --pipeline 5 --small-sms 40
probe --schedule partitioned --pipeline 5 --small-sms 40 --launch-priorities on
probe --schedule partitioned --pipeline 5 --small-sms 32 --launch-priorities on
probe --schedule partitioned --pipeline 5 --small-sms 48 --big-streams 3
probe --schedule partitioned --pipeline 6 --small-sms 48 --big-streams 3 --launch-priorities on
probe --schedule partitioned --pipeline 8 --small-sms 48 --launch-priorities on
probe --schedule batches --pipeline 5 --launch-priorities on
