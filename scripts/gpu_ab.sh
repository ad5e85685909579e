#!/bin/bash
# A/B on ONE box: round-1 tree (_r1/, git archive of 28a50db) against the current tree, same schedule; then ncu of the stand-alone kernels.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
    print(sys.argv[1], 'us/step %.2f  kernel_ms %.2f us (%s)  e2e %.2fM  sm_mhz %s %s' % (d['ms_per_step'] * 1e3, r['kernel_ms'] * 1e3, r['kernel'][:20], d['e2e']['value'] / 1e6, d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
(cd _r1 && timeout 300 python bench.py --no-cpu-baseline --launch-priorities off --schedule batches > ../gpurun_out/ab_r1.json 2> ../gpurun_out/ab_r1.err); show gpurun_out/ab_r1.json
timeout 300 python bench.py --no-cpu-baseline --no-autotune --schedule batches > gpurun_out/ab_cur.json 2> gpurun_out/ab_cur.err; show gpurun_out/ab_cur.json
(cd _r1 && timeout 300 python bench.py --no-cpu-baseline --launch-priorities off --schedule batches > ../gpurun_out/ab_r1b.json 2> ../gpurun_out/ab_r1b.err); show gpurun_out/ab_r1b.json
timeout 300 python bench.py --no-cpu-baseline --no-autotune --schedule batches > gpurun_out/ab_curb.json 2> gpurun_out/ab_curb.err; show gpurun_out/ab_curb.json
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --set full --import-source on -k regex:"k_score_umma|k_topk_fast" -s 8 -c 2 -o gpurun_out/r02_cfg2_umma_topk \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph --no-autotune --schedule batches --pipeline 1 > /dev/null 2> gpurun_out/ab_ncu.err
ls -la gpurun_out/r02_cfg2_umma_topk.ncu-rep
