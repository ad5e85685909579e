#!/usr/bin/env python
"""Round-2 experiment harness (ROADMAP.md "plan of record"): cfg2 step time of the fused schedule — ONE launch per batch that
scores batch i and selects the top-k of batch i-1 in the same persistent CTAs (gdr_score_fused), the inversion of batch i+1
running one batch ahead on a second stream — against the default schedule of bench.py's inner loop (whole batches round-robin on
`--pipeline` streams).  Three handles = three scratch sets: while launch i scores into h[i % 3] and reads h[(i-1) % 3] for the
top-k, the inversion of batch i+1 fills h[(i+1) % 3].  Checks the fused results against gdr_score_topk first and refuses to time
a variant that differs.  Prints one JSON line.  Written without GPU access at the end of round 1; first run is round 2's.

    python tools/bench_fused.py [--steps 960] [--pipeline 5] [--groups 4]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synth_shard / synth_batches / WORKLOADS)
from gdr_b200 import ClusterStore  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=960)
    ap.add_argument("--pipeline", type=int, default=5)
    ap.add_argument("--groups", type=int, default=4, choices=[3, 4, 5])
    ap.add_argument("--replicas", type=int, default=4)
    ap.add_argument("--ctas", type=int, default=140, help="CTAs of the fused grid (GDR_UMMA_CTAS): fewer than the 148 SMs, because the "
                    "inversion's 1,024-thread k_scan CTA does not fit beside an 832-thread fused CTA and needs SMs of its own; the "
                    "default schedule keeps one scoring CTA per SM")
    args = ap.parse_args()
    os.environ["GDR_FUSED_GROUPS"] = str(args.groups)
    cfg = dict(bench.WORKLOADS["cfg2"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    k, B = cfg["k"], cfg["B"]
    emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
    embs = [emb] + [emb.clone() for _ in range(args.replicas - 1)]          # replicas cycled: every step streams a store copy not in L2
    batches = bench.synth_batches(cfg, 8, cfg["C"], B, 4321, dev)
    n_b, R = len(batches), args.replicas

    def handles(n):
        return [[ClusterStore(e, offsets, docid) for e in embs] for _ in range(n)]

    # ---- correctness first: fused sequence == gdr_score_topk, batch by batch
    ref_h = handles(1)[0]
    refs = [ref_h[i % R].score_topk(q, b, k) for i, (q, b) in enumerate(batches)]
    refs = [(s.clone(), d.clone()) for s, d in refs]
    os.environ["GDR_UMMA_CTAS"] = str(args.ctas)          # read once per handle, at creation: only the fused handles get it
    h = handles(3)
    os.environ.pop("GDR_UMMA_CTAS")
    outs = []
    for i, (q, b) in enumerate(batches):
        cur, prev = h[i % 3][i % R], (h[(i - 1) % 3][(i - 1) % R] if i else None)
        cur.invert(q, b, k)
        r = cur.score_fused(prev)
        if r is not None:
            outs.append((r[0].clone(), r[1].clone()))
    outs.append(ClusterStore.flush_fused(h[(n_b - 1) % 3][(n_b - 1) % R]))
    torch.cuda.synchronize()
    same = [bool(torch.equal(a[0], r[0]) and torch.equal(a[1], r[1])) for a, r in zip(outs, refs)]
    if not all(same):
        print(json.dumps({"ok": False, "identical": same}))
        return 1

    # ---- fused schedule: stream M runs the fused launches back to back, stream A the inversions one batch ahead
    out_s = [torch.empty((B, k), dtype=torch.float32, device=dev) for _ in range(3)]
    out_d = [torch.empty((B, k), dtype=torch.int32, device=dev) for _ in range(3)]
    s_a = torch.cuda.Stream()

    def run_fused(n, cur_stream):
        s_a.wait_stream(cur_stream)
        ev_f = {}
        for i in range(n):
            q, b = batches[i % n_b]
            cur = h[i % 3][i % R]
            with torch.cuda.stream(s_a):
                if i - 2 in ev_f:
                    s_a.wait_event(ev_f[i - 2])       # the batch that last used this scratch set has had its top-k
                cur.invert(q, b, k)
                e_inv = torch.cuda.Event()
                e_inv.record(s_a)
            cur_stream.wait_event(e_inv)
            prev = h[(i - 1) % 3][(i - 1) % R] if i else None
            cur.score_fused(prev, out=(out_s[(i - 1) % 3], out_d[(i - 1) % 3]) if i else None)
            ev_f[i] = torch.cuda.Event()
            ev_f[i].record(cur_stream)
        ClusterStore.flush_fused(h[(n - 1) % 3][(n - 1) % R], out=(out_s[(n - 1) % 3], out_d[(n - 1) % 3]))
        cur_stream.wait_stream(s_a)

    # ---- default schedule (bench.py's): whole batches round-robin on n_pipe streams, each with its own handles
    n_pipe = args.pipeline
    hp = handles(n_pipe)
    streams = [torch.cuda.Stream() for _ in range(n_pipe)]
    outs_p = [(torch.empty((1, B, k), dtype=torch.float32, device=dev), torch.empty((1, B, k), dtype=torch.int32, device=dev)) for _ in range(n_pipe)]

    def run_default(n, cur_stream):
        for s in streams:
            s.wait_stream(cur_stream)
        for i in range(n):
            q, b = batches[i % n_b]
            with torch.cuda.stream(streams[i % n_pipe]):
                hp[i % n_pipe][i % R].score_topk(q, b, k, out=outs_p[i % n_pipe])
        for s in streams:
            cur_stream.wait_stream(s)

    period = 120                                   # lcm(3 handles, 8 batches, 4 replicas, 5 pipes)
    result = {"ok": True, "groups": args.groups, "fused_ctas": args.ctas, "steps": args.steps, "period": period}
    for name, fn in (("default", run_default), ("fused", run_fused)):
        fn(period, torch.cuda.current_stream())    # warm-up: every handle allocates its scratch
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn(period, side)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(3):
            e0.record()
            for _ in range(max(1, args.steps // period)):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3 / (max(1, args.steps // period) * period))
        result[name + "_us_per_step"] = sorted(times)[1]
    result["queries_per_s"] = {n: B / (result[n + "_us_per_step"] * 1e-6) for n in ("default", "fused")}
    print(json.dumps(result))
    return 0


if __name__ == "__main__":
    sys.exit(main())
