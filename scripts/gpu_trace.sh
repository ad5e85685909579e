#!/bin/bash
# the timeline / stage masks exist only in developer builds of the library
export GDR_BUILD_DEBUG_KNOBS=1; python -m gdr_b200._build > /dev/null
for v in $@; do
echo "== dbg=$v"
GDR_UMMA_TRACE=1 GDR_UMMA_DEBUG=$v timeout 100 python - <<'PY' 2>&1 | grep "umma trace" | tail -1 | cut -c1-1500
import sys, torch
sys.path.insert(0, '.')
import bench
from gdr_b200 import ClusterStore
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda', 0)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
st = ClusterStore(emb, offsets, docid)
(q, beams), = bench.synth_batches(cfg, 1, cfg['C'], cfg['B'], 4321, dev)
for i in range(3): st.score_topk(q, beams, 100)
torch.cuda.synchronize()
st.last_stats()
PY
done
