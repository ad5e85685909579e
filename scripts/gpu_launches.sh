#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 40 -c 48 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = [i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[hdr+2:]:
    if len(r) > vi: d[r[ki][:70]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(f"{k:70s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.2f} us  min={min(v)/1000:8.2f}")
PY
