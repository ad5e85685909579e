#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
for p in 2 3 4 6; do python bench.py --steps 384 --warmup 5 --no-cpu-baseline --pipeline $p | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(\"pipeline $p: value %.3fM  ms/step %.4f  e2e %.3fM\" % (d[\"value\"]/1e6, d[\"ms_per_step\"], d[\"e2e\"][\"value\"]/1e6))"; done
