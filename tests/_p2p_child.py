"""Child process of tests/test_gpu_sharded_p2p.py: the peer-to-peer candidate exchange of a cluster-sharded corpus with every
"rank" living in THIS process on ONE GPU (gdr_store_p2p_attach_local cross-wires the handles' exchange buffers), so the
ownership / offset / flag logic is exercised without NVLink or IPC.  Each simulated rank must return, for the queries it owns,
exactly what a plain call on the whole corpus returns.  A separate process: a wrong flag protocol would spin forever."""
import json
import os
import sys

# every simulated rank drives several streams and some of their kernels SPIN for a peer's kernel: with the default of 8 hardware work
# queues two such streams can share a queue, and the spinning kernel then blocks the very kernel it waits for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gdr_oracle as orc                      # noqa: E402  (input synthesis only)
from gdr_b200 import ClusterStore            # noqa: E402
from gdr_b200 import _cabi                   # noqa: E402
from gdr_b200.sharded import ShardedPipeline, partition_contiguous      # noqa: E402


def shards(emb, offsets, docid, world, dtype):
    bounds = partition_contiguous(np.diff(offsets), world)
    docid_t = torch.from_numpy(np.asarray(docid)).cuda()
    out = []
    for r in range(world):
        lo, hi = int(offsets[bounds[r]]), int(offsets[bounds[r + 1]])
        out.append(ClusterStore.shard(emb[lo:hi].to(dtype).cuda(), offsets, docid_t, int(bounds[r]), int(bounds[r + 1])))
    return out


def serial(world, path):
    """Plain gdr_score_topk calls on p2p handles, issued phase by phase: every rank inverts + scores, then every rank selects."""
    N, C, D, K, k, b_own = 24000, 96, 768, 12, 60, 80
    dtype = torch.float32 if path == "fp32" else torch.bfloat16
    flags = {"umma": _cabi.FORCE_UMMA, "simt": _cabi.FORCE_SIMT, "fp32": 0}[path]
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=61)
    emb = emb.bfloat16().float()
    full = ClusterStore.from_csr(emb, offsets, docid, dtype=dtype)
    st = shards(emb, offsets, docid, world, dtype)
    for r, s in enumerate(st):
        s.p2p_init(world, r, b_own, K)
    for s in st:
        s.p2p_attach_local(st)
    B = world * b_own
    ok = []
    for rep in range(3):                       # repeated batches: epochs advance, counters come back to zero
        q, beams, beam_scores = orc.synth_queries(B, C, K, D, seed=70 + rep)
        qd, bd, pd = q.cuda(), torch.from_numpy(beams).cuda(), torch.softmax(beam_scores, -1).cuda()
        ref_s, ref_d = full.score_topk(qd, bd, k, prob=pd, alphas=[0.0, 1.0], act="tanh", flags=flags)
        outs = [(torch.empty((2, b_own, k), dtype=torch.float32, device="cuda"), torch.empty((2, b_own, k), dtype=torch.int32, device="cuda")) for _ in st]
        for s, o in zip(st, outs):
            s.score_topk(qd, bd, k, prob=pd, alphas=[0.0, 1.0], act="tanh", flags=flags | _cabi.SKIP_TOPK, out=o)
        for s, o in zip(st, outs):
            s.score_topk(qd, bd, k, prob=pd, alphas=[0.0, 1.0], act="tanh", flags=flags | _cabi.SKIP_INVERT | _cabi.SKIP_SCORE, out=o)
        torch.cuda.synchronize()
        for r, o in enumerate(outs):
            sl = slice(r * b_own, (r + 1) * b_own)
            ok.append(bool(torch.equal(o[0], ref_s[:, sl]) and torch.equal(o[1], ref_d[:, sl])))
    print(json.dumps({"variant": f"p2p serial world={world} {path}", "identical": ok, "ok": all(ok)}))


def fused(world, schedule="fused"):
    """ShardedPipeline per simulated rank (fused schedule: the ranks' launches interleaved on one stream; batches: five streams per rank)."""
    N, C, D, K, k, b_own = 30000, 160, 768, 20, 100, 128
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=81)
    emb = emb.bfloat16().float()
    full = ClusterStore.from_csr(emb, offsets, docid, dtype=torch.bfloat16)
    st = shards(emb, offsets, docid, world, torch.bfloat16)
    pipes = [ShardedPipeline(st[r], r, world, b_own, K, k, schedule=schedule, depth=3, local_peers=True) for r in range(world)]
    ShardedPipeline.connect_local(pipes)
    B = world * b_own
    batches, refs = [], []
    for i in range(5):
        q, beams, beam_scores = orc.synth_queries(B, C, K, D, seed=90 + i)
        qd, bd, pd = q.cuda(), torch.from_numpy(beams).cuda(), torch.softmax(beam_scores, -1).cuda()
        batches.append((qd, bd, pd))
        s, d = full.score_topk(qd, bd, k, prob=pd, alphas=[0.5], act="tanh")
        refs.append((s[0].clone(), d[0].clone()))
    ok = []
    for rep in range(2):
        tickets = [[] for _ in pipes]
        for qd, bd, pd in batches:             # lock step: every rank submits batch i before anyone submits batch i+1
            for r, p in enumerate(pipes):
                tickets[r].append(p.submit(qd, bd, prob=pd, alpha=0.5, act="tanh"))
        for p in pipes:                        # one host thread drives every rank: all last scoring launches first ...
            p.pr.flush_scoring()
        for p in pipes:                        # ... then the final top-k of each (it waits for every rank's scores)
            p.flush()
        torch.cuda.synchronize()
        for r in range(world):
            sl = slice(r * b_own, (r + 1) * b_own)
            ok += [bool(torch.equal(t.scores, ref[0][sl]) and torch.equal(t.docids, ref[1][sl])) for t, ref in zip(tickets[r], refs)]
    print(json.dumps({"variant": f"p2p {schedule} world={world}", "schedule": pipes[0].schedule, "identical": ok,
                      "ok": all(ok) and pipes[0].schedule.startswith(schedule)}))


def gather(world):
    """PeerAllGather with every rank in this process: one stream per simulated rank (each rank's wait kernel needs the others' copies)."""
    from gdr_b200.sharded import PeerAllGather
    B, D, K, S = 96, 64, 12, 3
    objs = [PeerAllGather(r, world, [(B, D, torch.float32), (B, K, torch.int32)], S, torch.device("cuda", 0), local=True) for r in range(world)]
    PeerAllGather.connect_local(objs)
    streams = [torch.cuda.Stream() for _ in range(world)]
    ok = []
    for it in range(2 * S + 1):                # slots are reused, epochs advance
        g = torch.Generator().manual_seed(100 + it)
        q = torch.randn(world * B, D, generator=g)
        b = torch.randint(0, 1000, (world * B, K), generator=g, dtype=torch.int32)
        owns = []
        for r in range(world):
            own = torch.empty(objs[r].own_bytes, dtype=torch.uint8, device="cuda")
            own[:B * D * 4].view(torch.float32).view(B, D).copy_(q[r * B:(r + 1) * B])
            own[B * D * 4:].view(torch.int32).view(B, K).copy_(b[r * B:(r + 1) * B])
            owns.append(own)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                objs[r].all_gather(it % S, owns[r])
        torch.cuda.synchronize()
        for r in range(world):
            ok.append(bool(torch.equal(objs[r].gathered(it % S, 0).cpu(), q) and torch.equal(objs[r].gathered(it % S, 1).cpu(), b)))
    print(json.dumps({"variant": f"peer all-gather world={world}", "identical": ok, "ok": all(ok)}))


if __name__ == "__main__":
    torch.cuda.set_device(0)
    mode, world = sys.argv[1], int(sys.argv[2])
    if mode == "gather":
        gather(world)
    elif mode in ("fused", "batches"):
        fused(world, mode)
    else:
        serial(world, mode)
