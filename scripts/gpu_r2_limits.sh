#!/bin/bash
# per-SM limiter of the scoring CTA: the library must be the developer build (GDR_BUILD_DEBUG_KNOBS=1) when the call is made
mkdir -p gpurun_out
timeout 200 python tools/probe_umma_limits.py 2>&1 | grep -v "^$" | tee gpurun_out/r02_umma_limits.txt | tail -40
