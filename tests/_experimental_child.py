"""Child process of tests/test_gpu_zz_experimental.py: runs an experimental launch variant of the library (selected by an
environment variable that the library reads when a store is created) against the default variant on the same inputs and
prints one JSON line.  A separate process so that a variant that faults or hangs cannot take the test session down."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gdr_oracle as orc                      # noqa: E402  (input synthesis only)
from gdr_b200 import ClusterStore             # noqa: E402


def fused(groups):
    """gdr_score_fused (scoring of batch i + top-k of batch i-1 in one launch, two handles) against gdr_score_topk."""
    os.environ["GDR_FUSED_GROUPS"] = groups
    N, C, D, Q, K, k = 20000, 128, 768, 600, 20, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=21)
    base = ClusterStore.from_csr(emb.bfloat16().float(), offsets, docid, dtype=torch.bfloat16)
    h = [ClusterStore(base.emb, torch.as_tensor(base.offsets_host), base.docid) for _ in range(2)]
    batches, refs, outs = [], [], []
    for i in range(5):
        q, beams, beam_scores = orc.synth_queries(Q - 7 * i, C, K, D, seed=30 + i)       # a different batch size every time
        batches.append((q.cuda(), torch.from_numpy(beams).cuda(), torch.softmax(beam_scores, -1).cuda()))
    for q, b, p in batches:
        s, d = base.score_topk(q, b, k, prob=p, alphas=[1.0], act="tanh")
        refs.append((s[0].clone(), d[0].clone()))
    for rep in range(2):                      # twice: the queue counters of both handles must come back to zero
        outs = []
        for i, (q, b, p) in enumerate(batches):
            h[i % 2].invert(q, b, k, prob=p, act="tanh")
            r = h[i % 2].score_fused(h[(i - 1) % 2] if i else None, alpha=1.0)
            if r is not None:
                outs.append((r[0].clone(), r[1].clone()))
        outs.append(ClusterStore.flush_fused(h[(len(batches) - 1) % 2], 1.0))
        torch.cuda.synchronize()
        same = [bool(torch.equal(a[0], b_[0]) and torch.equal(a[1], b_[1])) for a, b_ in zip(outs, refs)]
        if not all(same) or len(outs) != len(refs):
            print(json.dumps({"variant": f"fused G={groups}", "identical": same, "ok": False}))
            return
    print(json.dumps({"variant": f"fused G={groups}", "identical": same, "ok": True}))


def main():
    var, value = sys.argv[1], sys.argv[2]
    torch.cuda.set_device(0)
    if var == "FUSED":
        return fused(value)
    cases = []
    # (N, C, D, Q, K, k, with bias): the cfg2 shape scaled down, a k around n (n <= k: take-all fallback on some queries), mass ties
    for N, C, D, Q, K, k, bias, ties in ((20000, 128, 768, 300, 20, 100, True, False), (1500, 64, 128, 97, 6, 128, False, False),
                                         (4000, 16, 64, 64, 4, 50, True, True)):
        emb, offsets, docid = orc.synth_corpus(N, C, D, seed=11)
        emb = emb.bfloat16().float()
        if ties:
            emb[:] = emb[0]                   # every score equal: the boundary bin overflows, general fallback, docid order
        q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=12)
        prob = torch.softmax(beam_scores, -1).cuda() if bias else None
        os.environ.pop(var, None)
        base = ClusterStore.from_csr(emb, offsets, docid, dtype=torch.bfloat16)
        os.environ[var] = value
        test = ClusterStore.from_csr(emb, offsets, docid, dtype=torch.bfloat16)
        os.environ.pop(var, None)
        qd, bd = q.cuda(), torch.from_numpy(beams).cuda()
        for rep in range(3):                  # repeated calls: the variant's queue counters must come back to zero
            s0, d0 = base.score_topk(qd, bd, k, prob=prob, alphas=[0.0, 1.0] if bias else None)
            s1, d1 = test.score_topk(qd, bd, k, prob=prob, alphas=[0.0, 1.0] if bias else None)
            torch.cuda.synchronize()
            cases.append(bool(torch.equal(s0, s1) and torch.equal(d0, d1)))
    print(json.dumps({"variant": f"{var}={value}", "identical": cases, "ok": all(cases)}))


if __name__ == "__main__":
    main()
