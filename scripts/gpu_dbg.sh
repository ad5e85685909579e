#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
GDR_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline \
    --workload cfg5s --steps 240 --warmup 3 > gpurun_out/dbg_n${N}.json 2> gpurun_out/dbg_n${N}.err
grep "^\[rank\|Error\|error\|illegal\|gdr" gpurun_out/dbg_n${N}.err | head -30 | cut -c1-300
