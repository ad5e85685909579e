#!/bin/bash
# the timeline / stage masks exist only in developer builds of the library
export GDR_BUILD_DEBUG_KNOBS=1; python -m gdr_b200._build > /dev/null
GDR_UMMA_TRACE=${TRACE:-} timeout 200 python - "$@" <<'PY' 2>&1 | grep -v "umma trace" | python -u -c "
import sys
rows=[]
for l in sys.stdin:
    if l.startswith('[timeline]'): rows.append(list(map(int,l.split()[1:])))
    else: print(l.rstrip())
rows=[r for r in rows if r[1]>0]
if rows:
    t0=min(r[0] for r in rows); rows.sort()
    for r in rows: print('inv %7.1f-%7.1f  umma %7.1f-%7.1f  topk %7.1f-%7.1f us | umma CTA loop start %7.1f..%7.1f end %7.1f..%7.1f' % tuple((x-t0)/1000 for x in r))"
import sys, torch, time
sys.path.insert(0, '.')
import bench
from gdr_b200 import ClusterStore
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda', 0)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
embs = [emb] + [emb.clone() for _ in range(3)]
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 24
stores = [ClusterStore(embs[i % 4], offsets, docid) for i in range(DEPTH)]
batches = bench.synth_batches(cfg, 8, cfg['C'], cfg['B'], 4321, dev)
outs = [(torch.empty((1, cfg['B'], 100), device=dev), torch.empty((1, cfg['B'], 100), dtype=torch.int32, device=dev)) for _ in range(DEPTH)]
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -5)
import os
P = [int(x) for x in os.environ.get('PRIO', '-1,-3,0').split(',')]
s_inv = torch.cuda.Stream(priority=P[0]); NSC = int(os.environ.get('NSC', '1')); s_scs = [torch.cuda.Stream(priority=P[1]) for _ in range(NSC)]; NTK = int(os.environ.get('NTK', '1'))
s_tks = [torch.cuda.Stream(priority=P[2]) for _ in range(NTK)]
print('priorities inv/score/topk', P, 'range', torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else None)
SK_I, SK_S, SK_T = 256, 512, 1024
def run(n):
    cur = torch.cuda.current_stream()
    for s in [s_inv] + s_scs + s_tks: s.wait_stream(cur)
    done = [None] * DEPTH
    for i in range(n):
        h = i % DEPTH; q, b = batches[i % 8]
        with torch.cuda.stream(s_inv):
            if done[h] is not None: s_inv.wait_event(done[h])
            stores[h].score_topk(q, b, 100, out=outs[h], flags=SK_S | SK_T)
            e1 = torch.cuda.Event(); e1.record(s_inv)
        s_sc = s_scs[i % NSC]
        with torch.cuda.stream(s_sc):
            s_sc.wait_event(e1)
            stores[h].score_topk(q, b, 100, out=outs[h], flags=SK_I | SK_T)
            e2 = torch.cuda.Event(); e2.record(s_sc)
        s_tk = s_tks[i % NTK]
        with torch.cuda.stream(s_tk):
            s_tk.wait_event(e2)
            stores[h].score_topk(q, b, 100, out=outs[h], flags=SK_I | SK_S)
            e3 = torch.cuda.Event(); e3.record(s_tk)
            done[h] = e3
    for s in [s_inv] + s_scs + s_tks: cur.wait_stream(s)
for h in range(DEPTH):
    q, b = batches[0]; stores[h].score_topk(q, b, 100, out=outs[h])
torch.cuda.synchronize()
run(NS); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        run(NS)
torch.cuda.current_stream().wait_stream(side)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print("phase-split pipeline depth %d: %.2f us/step" % (DEPTH, e0.elapsed_time(e1) * 1000 / (20 * NS)))
# correctness of the split call vs the fused call
ref_s, ref_d = stores[0].score_topk(batches[(NS - 1) % 8][0], batches[(NS - 1) % 8][1], 100)
h = (NS - 1) % DEPTH
print("split == fused:", bool(torch.equal(outs[h][1][0], ref_d) and torch.equal(outs[h][0][0], ref_s)))
import os
if os.environ.get("GDR_UMMA_TRACE"):
    for st in stores: st.last_stats()
PY
