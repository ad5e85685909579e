#!/bin/bash
# Two scoring CTAs per SM with 4-stage rings: the library must have been built with GDR_BUILD_UM_STAGES=4 GDR_BUILD_UM_MIN_CTAS=2.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
PROBE_MASKS=0 PROBE_CTAS=296,248,200,168,148,100 timeout 150 python tools/probe_umma_limits.py 2>&1 | grep -v "^$" | tee gpurun_out/r02_umma_twoctas.txt | tail -8
probe() { echo "== $*"; timeout 150 python bench.py --probe --gpus 1 --steps 1920 --warmup 3 --workload cfg2 --replicas 4 "$@" 2>&1 | grep -v "^$" | tail -2 | cut -c1-400; }
probe --schedule partitioned --pipeline 5 --small-sms 48 --ctas-per-sm 2
probe --schedule partitioned --pipeline 5 --small-sms 56 --ctas-per-sm 2
probe --schedule partitioned --pipeline 5 --small-sms 64 --ctas-per-sm 2
probe --schedule partitioned --pipeline 5 --small-sms 72 --ctas-per-sm 2
probe --schedule batches --pipeline 5 --launch-priorities on --ctas-per-sm 2
probe --schedule batches --pipeline 5 --launch-priorities on
timeout 100 python -m pytest tests/test_gpu_parity.py -q -x --timeout 90 -k "cfg2 or golden" 2>&1 | tail -2
