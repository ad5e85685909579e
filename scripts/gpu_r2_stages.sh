#!/bin/bash
# Ring-depth experiment: the library must have been built with GDR_BUILD_UM_STAGES=8 before the call (it travels as built).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
probe() { echo "== $*"; timeout 150 python bench.py --probe --gpus 1 --steps 1920 --warmup 3 --workload cfg2 --replicas 4 "$@" 2>&1 | grep -v "^$" | tail -2 | cut -c1-400; }
probe --schedule partitioned --pipeline 5 --small-sms 48
probe --schedule partitioned --pipeline 5 --small-sms 56
probe --schedule partitioned --pipeline 5 --small-sms 64
probe --schedule partitioned --pipeline 5 --small-sms 72
probe --schedule batches --pipeline 5 --launch-priorities on
timeout 100 python -m pytest tests/test_gpu_parity.py -q -x --timeout 90 -k "cfg2 or golden" 2>&1 | tail -2
