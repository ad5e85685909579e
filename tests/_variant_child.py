"""Child process of tests/test_gpu_pipeline.py: runs one launch variant of the library (a handle option, or the pipelined
schedules of gdr_b200/pipeline.py) against the plain gdr_score_topk call on the same inputs and prints one JSON line.
A separate process so that a variant that faults or hangs costs its own timeout, not the test session."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gdr_oracle as orc                      # noqa: E402  (input synthesis only)
from gdr_b200 import ClusterStore, PipelinedRetriever            # noqa: E402


def _batches(C, K, D, sizes, seed):
    out = []
    for i, Q in enumerate(sizes):
        q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=seed + i)
        out.append((q.cuda(), torch.from_numpy(beams).cuda(), torch.softmax(beam_scores, -1).cuda()))
    return out


def fused(groups):
    """gdr_score_fused driven by hand (two handles) against gdr_score_topk: prob + tanh + alpha, a different batch size every time."""
    N, C, D, K, k = 20000, 128, 768, 20, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=21)
    base = ClusterStore.from_csr(emb.bfloat16().float(), offsets, docid, dtype=torch.bfloat16)
    h = [base.clone_handle().set_option("fused_groups", int(groups)) for _ in range(2)]
    batches = _batches(C, K, D, [600 - 7 * i for i in range(5)], 30)
    refs = []
    for q, b, p in batches:
        s, d = base.score_topk(q, b, k, prob=p, alphas=[1.0], act="tanh")
        refs.append((s[0].clone(), d[0].clone()))
    same = []
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
            break
    print(json.dumps({"variant": f"fused G={groups}", "identical": same, "ok": bool(same) and all(same) and len(outs) == len(refs)}))


def pipeline(schedule, groups):
    """PipelinedRetriever (the product schedule) against gdr_score_topk: eager, then captured in a CUDA graph and replayed,
    then through host buffers (submit_host).  cfg2-like density (20 pairs per cluster) so that `auto` picks the fused schedule."""
    N, C, D, K, k = 24000, 192, 768, 20, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=41)
    stores = [ClusterStore.from_csr(emb.bfloat16().float(), offsets, docid, dtype=torch.bfloat16) for _ in range(2)]
    batches = _batches(C, K, D, [256] * 7, 50)
    refs = []
    for i, (q, b, p) in enumerate(batches):
        s, d = stores[i % 2].score_topk(q, b, k, prob=p, alphas=[0.5], act="tanh")
        refs.append((s[0].clone(), d[0].clone()))
    if schedule == "partitioned":             # (`groups` = persistent scoring CTAs per SM here: 1 = k_score_umma, 2 = k_score_umma_x2)
        pr = PipelinedRetriever(stores, schedule=schedule, scoring_ctas_per_sm=int(groups)).reserve(256, K, k)
    else:
        pr = PipelinedRetriever(stores, schedule=schedule, fused_groups=int(groups)).reserve(256, K, k)
    checks = {}

    def run():
        ts = [pr.submit(q, b, k, prob=p, alpha=0.5, act="tanh", which=i % 2) for i, (q, b, p) in enumerate(batches)]
        pr.flush()
        return ts

    ts = run()
    torch.cuda.synchronize()
    checks["eager"] = all(torch.equal(t.scores, r[0]) and torch.equal(t.docids, r[1]) for t, r in zip(ts, refs))
    checks["schedule"] = pr.last_schedule
    run()                                     # a second pass: queue counters and scratch reuse
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            ts = run()
    torch.cuda.current_stream().wait_stream(side)
    for t in ts:
        t.scores.zero_(); t.docids.zero_()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    checks["graph"] = all(torch.equal(t.scores, r[0]) and torch.equal(t.docids, r[1]) for t, r in zip(ts, refs))
    # host buffers: one pinned input and one pinned output buffer per batch (no prob on this path: alpha unused)
    refs0 = [stores[i % 2].score_topk(q, b, k) for i, (q, b, p) in enumerate(batches)]
    refs0 = [(s.clone(), d.clone()) for s, d in refs0]
    ins, outs = [], []
    for q, b, p in batches:
        h = torch.empty(q.numel() * 4 + b.numel() * 4, dtype=torch.uint8).pin_memory()
        h[:q.numel() * 4].view(torch.float32).copy_(q.cpu().reshape(-1))
        h[q.numel() * 4:].view(torch.int32).copy_(b.cpu().reshape(-1))
        ins.append(h)
        outs.append(torch.zeros(2 * 256 * k * 4, dtype=torch.uint8).pin_memory())
    for i in range(len(batches)):
        pr.submit_host(ins[i], 256, K, k, outs[i], which=i % 2)
    pr.flush()
    torch.cuda.synchronize()
    checks["host"] = all(torch.equal(o[:256 * k * 4].view(torch.float32).view(256, k), r[0].cpu()) and
                         torch.equal(o[256 * k * 4:].view(torch.int32).view(256, k), r[1].cpu()) for o, r in zip(outs, refs0))
    ok = checks["eager"] and checks["graph"] and checks["host"] and checks["schedule"] == {"batches": "batches", "partitioned": "partitioned"}.get(schedule, "fused")
    if pr.partition is not None:
        checks["sms"] = [pr.partition.sms_big, pr.partition.sms_small]
    print(json.dumps({"variant": f"pipeline {schedule} G={groups}", "checks": checks, "ok": bool(ok)}))


def option(name, value):
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
        base = ClusterStore.from_csr(emb, offsets, docid, dtype=torch.bfloat16)
        test = base.clone_handle().set_option(name, int(value))
        qd, bd = q.cuda(), torch.from_numpy(beams).cuda()
        for rep in range(3):                  # repeated calls: the variant's queue counters must come back to zero
            s0, d0 = base.score_topk(qd, bd, k, prob=prob, alphas=[0.0, 1.0] if bias else None)
            s1, d1 = test.score_topk(qd, bd, k, prob=prob, alphas=[0.0, 1.0] if bias else None)
            torch.cuda.synchronize()
            cases.append(bool(torch.equal(s0, s1) and torch.equal(d0, d1)))
    print(json.dumps({"variant": f"{name}={value}", "identical": cases, "ok": all(cases)}))


def main():
    mode, value = sys.argv[1], sys.argv[2]
    torch.cuda.set_device(0)
    if mode == "FUSED":
        return fused(value)
    if mode.startswith("PIPELINE_"):
        try:
            return pipeline(mode[len("PIPELINE_"):].lower(), value)
        except Exception as e:                 # an SM partition needs green contexts in the driver (CUDA 12.4+)
            if getattr(e, "status", 0) == -3 and "green contexts" in str(e):
                print(json.dumps({"variant": mode, "ok": True, "skipped": str(e)}))
                return
            raise
    return option(mode, value)


if __name__ == "__main__":
    main()
