#!/bin/bash
# ncu launch list (device time of every launch) of the partitioned schedule: only the library's own kernels (namespace gdr)
mkdir -p gpurun_out
NCU="ncu --clock-control none --cache-control none"
timeout 150 $NCU --metrics gpu__time_duration.sum -k regex:"^k_" -s 48 -c 72 --csv --log-file gpurun_out/r02_launches_cfg2_partitioned.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule partitioned --small-sms 56 --ctas-per-sm 2 > /dev/null 2> gpurun_out/r02_ncu_launches.err
grep -c "gpu__time_duration" gpurun_out/r02_launches_cfg2_partitioned.csv; grep "ERROR" gpurun_out/r02_launches_cfg2_partitioned.csv | head -3
