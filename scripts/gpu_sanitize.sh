#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small but representative set of calls: both scoring paths, top-k
# fast/general paths, masks, beam step, merge.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import gdr_oracle as orc
from gdr_b200 import ClusterStore, DeviceTrie, TreeBuilder, position_mask_, compute_similarity
from gdr_b200.sharded import ShardedRetriever, pack_candidates
N, C, D, Q, K, k = 3000, 24, 128, 40, 6, 50
emb, offsets, docid = orc.synth_corpus(N, C, D, seed=1); emb = emb.bfloat16().float()
q, beams, bs = orc.synth_queries(Q, C, K, D, seed=2); prob = torch.softmax(bs, -1)
for dt in (torch.bfloat16, torch.float32):
    st = ClusterStore.from_csr(emb, offsets, docid, dtype=dt)
    for flags in ((0, 2, 4) if dt == torch.bfloat16 else (0,)):
        for kk in (50, 300):
            st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), kk, prob=prob.cuda(), alphas=[0.0, 1.0], act="tanh", flags=flags)
    st.score_topk(q.repeat_interleave(K, 0).cuda(), torch.from_numpy(beams).cuda(), k, per_beam=True)
    st.centroids()
tb = TreeBuilder(); rng = np.random.RandomState(0)
paths = [[i * 30 + int(c) + 2 for i, c in enumerate(rng.randint(0, 30, 3))] + [1] for _ in range(100)]
for i, p in enumerate(paths): tb.add(p, i)
trie = DeviceTrie.from_root(tb.build())
ids = torch.zeros(24, 3, dtype=torch.int64); ids[:, 1:] = torch.tensor([p[:2] for p in paths[:24]]); ids[5, 2] = 999
for V in (32128, 301):
    sc = torch.randn(24, V)
    trie.mask_(sc.clone().cuda(), ids.cuda()); trie.mask_(sc.clone().cuda(), ids.cuda(), strict=True)
    trie.beam_step(sc.cuda(), ids.cuda(), -torch.rand(24).cuda(), 6)
position_mask_(torch.randn(3, 4, 302).cuda(), 30)
compute_similarity(q.cuda(), emb[:333].cuda())
s = torch.randn(4, 10, 50).sort(-1, descending=True).values; d = torch.randint(0, 1000, (4, 10, 50))
ShardedRetriever._cuda_merge(torch.stack([pack_candidates(s[r], d[r].int()) for r in range(4)]).cuda(), 50)
torch.cuda.synchronize(); print("SANITIZER_SCRIPT_DONE")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_SCRIPT_DONE|Invalid|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
