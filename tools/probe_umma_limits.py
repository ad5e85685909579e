#!/usr/bin/env python
"""What bounds the tcgen05 scoring CTA per SM?  Needs the developer build (GDR_BUILD_DEBUG_KNOBS=1 python -m gdr_b200._build).

Times the cfg2 scoring kernel alone (GDR_SKIP_INVERT | GDR_SKIP_TOPK on handles whose inversion is in place, 40 back-to-back
launches over 4 corpus copies inside a CUDA graph) for several persistent-CTA counts, with single pipeline stages switched off
through GDR_UMMA_DEBUG (1 = no L2 hint, 2 = no TMA of the embeddings, 4 = no MMA, 8 = no B fill of the query terms).  Results
are invalid under the masks - this is a timing experiment only.  One JSON line per configuration."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                               # noqa: E402
from gdr_b200 import ClusterStore                          # noqa: E402

cfg = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
embs = [emb] + [emb.clone() for _ in range(3)]
batches = bench.synth_batches(cfg, 2, cfg["C"], cfg["B"], 4321, dev)
k, B = cfg["k"], cfg["B"]
dummy = (torch.empty((1, B, k), dtype=torch.float32, device=dev), torch.empty((1, B, k), dtype=torch.int32, device=dev))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

MASKS = [int(x) for x in os.environ.get("PROBE_MASKS", "0,8,2,4,10,12,14").split(",")]
CTAS = [int(x) for x in os.environ.get("PROBE_CTAS", "148,124,100,84").split(",")]
for mask in MASKS:
    os.environ["GDR_UMMA_DEBUG"] = str(mask)             # read by gdr_store_create in developer builds
    stores = [ClusterStore(e, offsets, docid) for e in embs]
    for ctas in CTAS:
        for s in stores:
            s.set_option("umma_ctas", ctas)
        for i, s in enumerate(stores):
            s.invert(batches[i % 2][0], batches[i % 2][1], k)
        torch.cuda.synchronize()

        def run(n):
            for i in range(n):
                s = stores[i % 4]
                s.score_topk(batches[i % 2][0], batches[i % 2][1], k, out=dummy, flags=256 | 1024)

        g = bench.capture(run, 40)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 40)
        print(json.dumps({"umma_debug_mask": mask, "ctas": ctas, "us_per_launch": round(sorted(ts)[2], 2)}), flush=True)
    del stores
