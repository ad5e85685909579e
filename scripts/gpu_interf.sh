#!/bin/bash
# pairwise interference: the scoring phase launched back to back on one stream while another stream loops another phase
timeout 200 python - <<'PY' 2>&1 | grep -v "umma trace\|timeline" | tail -12
import sys, torch
sys.path.insert(0, '.')
import bench
from gdr_b200 import ClusterStore
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda', 0)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
embs = [emb, emb.clone()]
A = [ClusterStore(embs[i], offsets, docid) for i in range(2)]      # scoring stream: two replicas
Bst = [ClusterStore(embs[i], offsets, docid) for i in range(2)]    # other stream
batches = bench.synth_batches(cfg, 2, cfg['C'], cfg['B'], 4321, dev)
outA = (torch.empty((1, cfg['B'], 100), device=dev), torch.empty((1, cfg['B'], 100), dtype=torch.int32, device=dev))
outB = (torch.empty((1, cfg['B'], 100), device=dev), torch.empty((1, cfg['B'], 100), dtype=torch.int32, device=dev))
for i in range(2):
    A[i].score_topk(*batches[i], 100, out=outA); Bst[i].score_topk(*batches[i], 100, out=outB)
torch.cuda.synchronize()
SK_I, SK_S, SK_T = 256, 512, 1024
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run(other_flags, n=60, ratio=1):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(sa):
        torch.cuda._sleep(6_000_000); e[0].record(sa)
        for i in range(n): A[i % 2].score_topk(*batches[i % 2], 100, out=outA, flags=SK_I | SK_T)
        e[1].record(sa)
    if other_flags is not None:
        with torch.cuda.stream(sb):
            torch.cuda._sleep(6_000_000); e[2].record(sb)
            for i in range(n * ratio): Bst[i % 2].score_topk(*batches[i % 2], 100, out=outB, flags=other_flags)
            e[3].record(sb)
    torch.cuda.synchronize()
    a = e[0].elapsed_time(e[1]) * 1000 / n
    b = e[2].elapsed_time(e[3]) * 1000 / (n * ratio) if other_flags is not None else 0.0
    return a, b
print("scoring alone            : %.1f us" % run(None)[0])
a, b = run(SK_I | SK_S);        print("scoring || top-k loop    : scoring %.1f us, top-k %.1f us per launch" % (a, b))
a, b = run(SK_S | SK_T);        print("scoring || inversion loop: scoring %.1f us, inversion %.1f us per launch" % (a, b))
a, b = run(SK_I | SK_S, ratio=2); print("scoring || top-k loop x2 : scoring %.1f us, top-k %.1f us per launch" % (a, b))
with torch.cuda.stream(sb):
    torch.cuda._sleep(6_000_000); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(sb)
    for i in range(60): Bst[i % 2].score_topk(*batches[i % 2], 100, out=outB, flags=SK_I | SK_S)
    e1.record(sb)
torch.cuda.synchronize(); print("top-k alone              : %.1f us" % (e0.elapsed_time(e1) * 1000 / 60))
PY
