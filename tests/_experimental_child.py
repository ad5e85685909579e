"""Child process of tests/test_gpu_experimental.py: runs an experimental launch variant of the library (selected by an
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


def main():
    var, value = sys.argv[1], sys.argv[2]
    torch.cuda.set_device(0)
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
