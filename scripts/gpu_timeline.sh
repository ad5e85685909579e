#!/bin/bash
# the timeline / stage masks exist only in developer builds of the library
export GDR_BUILD_DEBUG_KNOBS=1; python -m gdr_b200._build > /dev/null
GDR_UMMA_TRACE=1 timeout 200 python - <<'PY' 2>&1 | grep timeline | python -c "
import sys
rows=[list(map(int,l.split()[1:])) for l in sys.stdin if l.startswith('[timeline]')]
rows=[r for r in rows if r[1]>0]
t0=min(r[0] for r in rows)
rows.sort()
for r in rows: print('inv %7.1f-%7.1f  umma %7.1f-%7.1f  topk %7.1f-%7.1f us | umma CTA loop start %7.1f..%7.1f end %7.1f..%7.1f' % tuple((x-t0)/1000 for x in r))"
import sys, torch
sys.path.insert(0, '.')
import bench
from gdr_b200 import ClusterStore
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda', 0)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
NP = 12
stores = [ClusterStore(emb, offsets, docid) for _ in range(NP)]
batches = bench.synth_batches(cfg, 4, cfg['C'], cfg['B'], 4321, dev)
outs = [(torch.empty((1, cfg['B'], 100), device=dev), torch.empty((1, cfg['B'], 100), dtype=torch.int32, device=dev)) for _ in range(NP)]
streams = [torch.cuda.Stream() for _ in range(3)]
def run():
    cur = torch.cuda.current_stream()
    for s in streams: s.wait_stream(cur)
    for i in range(NP):
        with torch.cuda.stream(streams[i % 3]):
            q, b = batches[i % 4]
            stores[i].score_topk(q, b, 100, out=outs[i])
    for s in streams: cur.wait_stream(s)
for _ in range(2): run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        run()
torch.cuda.current_stream().wait_stream(side)
for _ in range(3): g.replay()
torch.cuda.synchronize()
for st in stores: st.last_stats()
PY
