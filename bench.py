#!/usr/bin/env python
"""bench.py — queries/sec of cluster-restricted scoring + top-100 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg5s]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (pair inversion -> gather-and-score -> per-query top-k) over one
batch of synthetic queries.  N = 1 runs BASELINE.json configs[1] ("cfg2": 109,739 x 768 bf16 corpus,
1,024 k-means-shaped clusters, batch 1,024 queries, beam 20, top-100).  N > 1 (weak scaling, one process per GPU):
cfg2's corpus is 169 MB, so every rank holds a replica and its own 1,024-query batch — independent units, no
data-path collective (`--mode replica`, the default); `--workload cfg5s` (a 12.5 M-doc slice per GPU, the shape of
a corpus that does NOT fit one GPU) runs the cluster-sharded path of SURVEY.md §8e: queries replicated, local top-k
on each rank's clusters, one NCCL all-gather of packed (score, docid) candidates, merge top-k (`--mode sharded`).
Batches are independent, so `--pipeline` of them (default 5) are kept in flight on as many streams inside one CUDA
graph; `e2e` adds one pinned-host H2D copy (q + beams) and one D2H copy (scores + docids) per step.
Launch autotune (`--launch-priorities auto`, the default): before the stores are created rank 0 times 1,920 device-resident
steps of the workload in child processes — default launches, per-launch priorities (GDR_LAUNCH_PRIORITIES=1: inversion >
scoring > top-k), priorities with 3 more batches in flight, and (cfg2 only) the experimental fused schedule `gdr_score_fused`,
whose child first checks every batch's result bit for bit against gdr_score_topk — and keeps a candidate only if it beats the
default by > 3 %; all timings (or failures) are printed in `config.launch_autotune`, so the line says what was chosen and why.
Timing: W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the
launching stream, max over ranks.  L2 hygiene: each step reads a different replica of the store
(`config.l2`: the replicas together are several times the 126 MB L2) and a different query batch.
`--impl reference` (and the `cpu_baseline` object of the default arm) time the reference's own
dense.py path — the oracle port, oracle/gdr_oracle.py — on the host cores.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N docs, C clusters, D, batch B, beam K, top-k      (BASELINE.json configs[0], [1], [2]; cfg5s = per-GPU slice of [4])
    "cfg1": dict(N=109739, C=1024, D=768, B=1024, K=10, k=100, fp32=True),     # configs[0]: fp32 store, beam 10 (the reference's CPU case)
    "cfg2": dict(N=109739, C=1024, D=768, B=1024, K=20, k=100),
    "cfg3": dict(N=73970, C=1024, D=768, B=1024, K=100, k=1000),
    "cfg5s": dict(N=12500000, C=131072, D=768, B=1250, K=100, k=100),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


def synth_shard(cfg, seed, device):
    """SURVEY.md §8d recipe, generated on the device: emb = randn(N, D) * D^-0.5 (bf16), assign = randint(0, C),
    CSR by stable sort.  Returns (emb [N, D] bf16 cluster-contiguous, offsets [C+1] cpu, docid [N])."""
    g = torch.Generator(device=device).manual_seed(seed)
    N, C, D = cfg["N"], cfg["C"], cfg["D"]
    assign = torch.randint(0, C, (N,), generator=g, device=device)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=C)
    offsets = torch.zeros(C + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    dt = torch.float32 if cfg.get("fp32") else torch.bfloat16
    emb = torch.empty((N, D), dtype=dt, device=device)
    step = 1 << 20
    for i in range(0, N, step):   # chunked so the fp32 temporary stays small at 12.5 M rows
        emb[i:i + step] = (torch.randn((min(step, N - i), D), generator=g, device=device) * D ** -0.5).to(dt)
    return emb, offsets.cpu(), order


def synth_batches(cfg, n_batches, C_total, B, seed, device):
    """q = randn(B, D); beams = K distinct clusters per query (randperm-like), on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(n_batches):
        q = torch.randn((B, cfg["D"]), generator=g, device=device)
        if C_total <= 8192:
            beams = torch.argsort(torch.rand((B, C_total), generator=g, device=device), dim=1)[:, :cfg["K"]]
        else:  # huge C: sample with replacement, duplicates are vanishingly rare and legal
            beams = torch.randint(0, C_total, (B, cfg["K"]), generator=g, device=device)
        out.append((q.contiguous(), beams.to(torch.int32).contiguous()))
    return out


def cpu_reference_qps(emb_cpu, offsets, docid, q_cpu, beams_cpu, k, min_seconds, max_queries):
    """The reference's dense.py path (oracle port) on the host cores: per query gather + q @ p.T + topk."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gdr_oracle as orc            # the one place bench.py may execute oracle/: the CPU baseline
    torch.set_num_threads(os.cpu_count())
    off = offsets.numpy()
    done, t0 = 0, time.perf_counter()
    chunk = 64
    orc.dense_topk(q_cpu[:8], emb_cpu, off, docid, beams_cpu[:8].numpy(), k)     # warm-up
    t0 = time.perf_counter()
    while done < max_queries and (time.perf_counter() - t0 < min_seconds or done == 0):
        sl = slice(done % q_cpu.shape[0], done % q_cpu.shape[0] + chunk)
        orc.dense_topk(q_cpu[sl], emb_cpu, off, docid, beams_cpu[sl].numpy(), k)
        done += min(chunk, q_cpu[sl].shape[0])
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def autotune_launch_config(args, local_rank, n_pipe_default):
    """Times the device-resident loop of this workload under a few launch configurations, each in a CHILD process (the library
    reads its launch knobs once per store, and a child that fails or hangs cannot take the bench down), and returns
    (use_priorities, pipeline, schedule, fused_ctas, fused_groups, report).  A configuration replaces the default only if it is more than 3 % faster; the
    fused schedule (gdr_score_fused: ONE launch scores batch i and selects the top-k of batch i-1) is eligible only if the
    child found its results identical, bit for bit, to gdr_score_topk's on every batch."""
    variants = [("default", "0", n_pipe_default, args.schedule, 0, 0), ("priorities", "1", n_pipe_default, args.schedule, 0, 0)]
    if args.workload == "cfg2" and args.path == "auto" and args.schedule == "auto":
        # fused grid on 140 or 132 CTAs (the SMs it leaves are where the inversion's large CTAs — k_scan, k_fill, k_tilemeta — run)
        # with 4 or 5 top-k groups per CTA (more queries in flight against fewer registers per thread)
        variants += [("fused", "0", n_pipe_default, "fused", 140, 4), ("fused_g5", "0", n_pipe_default, "fused", 140, 5),
                     ("fused_132_g5", "0", n_pipe_default, "fused", 132, 5)]
    variants.append(("priorities_deep", "1", n_pipe_default + 3, args.schedule, 0, 0))
    report, best, t_start = {}, None, time.time()
    for name, prio, n_pipe, schedule, fused_ctas, fused_groups in variants:
        if time.time() - t_start > 210:      # bound on the whole autotune (a hung candidate costs its 120 s timeout)
            report[name] = {"skipped": "autotune time budget spent", "batches_in_flight": n_pipe}
            continue
        env = {k_: v_ for k_, v_ in os.environ.items() if k_ not in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK",
                                                                     "TORCHELASTIC_RUN_ID", "MASTER_ADDR", "MASTER_PORT")}
        env.update(GDR_LAUNCH_PRIORITIES=prio, LOCAL_RANK=str(local_rank))
        if fused_ctas:
            env.update(GDR_FUSED_CTAS=str(fused_ctas), GDR_FUSED_GROUPS=str(fused_groups))
        cmd = [sys.executable, os.path.abspath(__file__), "--probe", "--gpus", "1", "--steps", "1920", "--warmup", "3", "--workload", args.workload,
               "--path", args.path, "--pipeline", str(n_pipe), "--schedule", schedule, "--replicas", str(args.replicas)]
        us = None
        try:
            out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120)
            lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            line = json.loads(lines[-1]) if out.returncode == 0 and lines else {}
            if line.get("us_per_step") and (schedule != "fused" or line.get("schedule") == "fused"):
                us = float(line["us_per_step"])
                report[name] = {"us_per_step": us, "batches_in_flight": n_pipe}
                if schedule == "fused":
                    report[name].update(verified_identical_to_default=True, fused_ctas=fused_ctas, fused_groups=fused_groups)
            else:
                report[name] = {"failed": (line.get("failed") or out.stderr or out.stdout)[-200:], "batches_in_flight": n_pipe}
        except Exception as e:      # timeout, launch failure, malformed line: the default stays
            report[name] = {"failed": repr(e)[:200], "batches_in_flight": n_pipe}
        if us and (best is None or us < best[0]):
            best = (us, name, prio == "1", n_pipe, schedule, fused_ctas, fused_groups)
    base = report["default"].get("us_per_step")
    if best is None or base is None or best[1] == "default" or best[0] > 0.97 * base:
        report["chosen"] = "default"
        return False, n_pipe_default, args.schedule, 0, 0, report
    report["chosen"] = best[1]
    return best[2], best[3], best[4], best[5], best[6], report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4800)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="gdr_b200", choices=["gdr_b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=0, help="store replicas cycled to defeat L2 (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="auto", choices=["auto", "replica", "sharded"],
                    help="N > 1: replica = every rank holds the corpus and its own query batches (no collective); sharded = clusters "
                         "sharded over ranks, queries replicated, NCCL all-gather of candidates + merge (auto: sharded for cfg5s)")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "umma"], help="force a scoring path")
    ap.add_argument("--pipeline", type=int, default=0, help="independent batches kept in flight, each with its own scratch (0 = auto: 6 for the "
                    "phase schedule, 5 for the batch schedule; 1 = strictly serial)")
    ap.add_argument("--schedule", default="auto", choices=["auto", "batches", "phases", "fused"],
                    help="batches: whole batches round-robin on one stream per batch in flight; phases: inversion / scoring / top-k on their own "
                         "(prioritised) streams, ordered with events, so scoring kernels of consecutive batches overlap (auto = batches, "
                         "which measured faster; fused = EXPERIMENT, gdr_score_fused: one launch scores batch i and selects the top-k of "
                         "batch i-1, inversion one batch ahead on a second stream — only adopted by the autotune when a child process "
                         "found it bit-identical to the default and > 3 % faster)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from the host instead of replaying a CUDA graph")
    ap.add_argument("--launch-priorities", default="auto", choices=["auto", "on", "off"],
                    help="per-launch scheduling priorities of the library (env GDR_LAUNCH_PRIORITIES, ROADMAP.md item 0: inversion > scoring > "
                         "top-k).  auto: rank 0 times a short run of this workload with and without them in child processes before the "
                         "stores are created and keeps the faster setting (like a cuDNN-style launch autotune); both timings are reported "
                         "in config.launch_autotune")
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)     # child mode of the autotune: time the device-resident loop, print one line
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    k, K, D = cfg["k"], cfg["K"], cfg["D"]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        small = dict(cfg)
        g = torch.Generator().manual_seed(1234)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gdr_oracle as orc
        emb, offsets, docid = orc.synth_corpus(small["N"] if small["N"] <= 200000 else 200000,
                                               small["C"] if small["N"] <= 200000 else 2048, D, seed=1234)
        emb = emb.bfloat16().float()
        C = offsets.size - 1
        q, beams, _ = orc.synth_queries(256, C, K, D, seed=4321)
        per_step = 64
        torch.set_num_threads(os.cpu_count())
        beams_t = torch.from_numpy(beams)
        for i in range(args.warmup):
            orc.dense_topk(q[:8], emb, offsets, docid, beams[:8], k)
        t0 = time.perf_counter()
        for i in range(args.steps):
            lo = (i * per_step) % 256
            orc.dense_topk(q[lo:lo + per_step], emb, offsets, docid, beams[lo:lo + per_step], k)
        dt = time.perf_counter() - t0
        qps = args.steps * per_step / dt
        print(json.dumps({
            "impl": "reference", "metric": "queries/sec, cluster-restricted scoring + top-k", "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "docs": int(emb.shape[0]), "clusters": int(C), "dim": D, "beam": K, "top_k": k,
                       "queries_per_step": per_step},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} queries of the {args.workload} workload, oracle/gdr_oracle.dense_topk "
                                       f"(reference dense.py:53-54 + Tensor.topk), torch {torch.__version__} CPU, {os.cpu_count()} threads"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ gdr_b200 arm (GPU)
    import torch.distributed as dist
    from gdr_b200 import ClusterStore
    from gdr_b200.sharded import ShardedRetriever

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sharded = world > 1 and (args.mode == "sharded" or (args.mode == "auto" and args.workload == "cfg5s"))
    B_global = cfg["B"] * world          # queries per step over all ranks
    B_rank = B_global if sharded else cfg["B"]      # queries each rank handles per step (sharded: all of them, replicated)
    C_total = cfg["C"] * world if sharded else cfg["C"]
    path_flags = {"auto": 0, "simt": 2, "umma": 4}[args.path]

    # ---- launch configuration: measured, not assumed (a short device-resident run per candidate, in child processes)
    autotune = None
    if args.launch_priorities == "on":
        os.environ["GDR_LAUNCH_PRIORITIES"] = "1"
    elif args.launch_priorities == "off":
        os.environ["GDR_LAUNCH_PRIORITIES"] = "0"
    elif (not args.probe and not sharded and "GDR_LAUNCH_PRIORITIES" not in os.environ and args.pipeline != 1
          and args.schedule != "phases" and not args.no_graph):
        decision = torch.zeros(5, dtype=torch.int32, device=dev)
        if rank == 0:
            use_prio, n_best, sched, f_ctas, f_groups, autotune = autotune_launch_config(args, local_rank, args.pipeline if args.pipeline > 0 else 5)
            decision = torch.tensor([int(use_prio), n_best, int(sched == "fused"), f_ctas, f_groups], dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(decision, src=0)
        use_prio, n_best, use_fused, f_ctas, f_groups = (int(x) for x in decision.tolist())
        os.environ["GDR_LAUNCH_PRIORITIES"] = "1" if use_prio else "0"      # read by the library when a store is created
        args.pipeline = n_best
        if use_fused:
            args.schedule = "fused"
            os.environ["GDR_FUSED_CTAS"] = str(f_ctas)
            os.environ["GDR_FUSED_GROUPS"] = str(f_groups)
    esize = 4 if cfg.get("fp32") else 2
    emb_bytes = cfg["N"] * D * esize
    replicas = args.replicas or max(1, min(6, -(-640 * 2 ** 20 // emb_bytes)))     # >= 640 MB of distinct store bytes in rotation
    stores = []
    base_offsets = None
    for r in range(replicas):
        emb, offsets, docid = synth_shard(cfg, 1234 + (1000 * rank if sharded else 0), dev) if r == 0 else (stores[0].emb.clone(), base_offsets, stores[0].docid.clone())
        base_offsets = offsets
        # docids are global: rank r's documents are numbered after those of ranks < r
        stores.append(ClusterStore(emb, offsets, docid + (rank * cfg["N"] if sharded else 0) if r == 0 else docid))
    n_batches = 8
    # sharded: same seed on every rank = replicated queries; replica: every rank draws its own batches
    batches = synth_batches(cfg, n_batches, C_total, B_rank, 4321 + (0 if sharded or world == 1 else rank), dev)
    if sharded:
        # rank r owns global clusters [r*C, (r+1)*C): contiguous blocks are already balanced for this synthetic corpus
        g2l = torch.full((C_total,), -1, dtype=torch.int32, device=dev)
        g2l[rank * cfg["C"]:(rank + 1) * cfg["C"]] = torch.arange(cfg["C"], dtype=torch.int32, device=dev)

    # Batches are independent, so `n_pipe` of them are kept in flight on `n_pipe` CUDA streams, each with its own
    # store handles (= its own scratch) and result buffers: the latency-bound inversion and top-k kernels of one batch
    # co-reside with, and hide under, the HBM-bound scoring kernel of its neighbours.
    # measured at cfg2 with every step streaming a store replica that is not in L2: phases 55.3 us per step, batches 51.4
    phases = not sharded and args.pipeline != 1 and args.schedule == "phases"
    n_pipe = 1 if sharded else (args.pipeline if args.pipeline > 0 else (6 if phases else 5))
    pipes = []
    for p in range(n_pipe):
        st_p = stores if p == 0 else [ClusterStore(s0.emb, torch.as_tensor(s0.offsets_host), s0.docid) for s0 in stores]
        pipes.append(dict(
            stream=torch.cuda.Stream(), stores=st_p,
            retr=[ShardedRetriever(x, g2l) for x in st_p] if sharded else None,
            q=torch.empty_like(batches[0][0]), b=torch.empty_like(batches[0][1]),
            out_s=torch.empty((1, B_rank, k), dtype=torch.float32, device=dev),
            out_d=torch.empty((1, B_rank, k), dtype=torch.int32, device=dev)))

    # phase schedule: one high-priority stream for the (tiny, latency-bound) inversion kernels so they never queue behind a
    # scoring grid that is waiting for SMs, three scoring streams (consecutive scoring kernels overlap tail to head, the
    # dynamic tile queue absorbs the staggered CTA starts), three low-priority top-k streams, one copy stream for e2e
    SK_I, SK_S, SK_T = 256, 512, 1024
    if phases:
        lo_p, hi_p = 0, -3
        try:
            lo_p, hi_p = torch.cuda.Stream.priority_range()
        except Exception:
            pass
        s_inv = torch.cuda.Stream(priority=hi_p)
        s_scs = [torch.cuda.Stream(priority=min(lo_p, hi_p + 1)) for _ in range(3)]
        s_tks = [torch.cuda.Stream(priority=lo_p) for _ in range(3)]
        s_h2d = torch.cuda.Stream(priority=hi_p)
        phase_streams = [s_inv, s_h2d] + s_scs + s_tks

    def step(i, P=None):
        """One pass of the hot path over one device-resident batch."""
        P = P or pipes[0]
        q, beams = batches[i % n_batches]
        if not sharded:
            P["stores"][i % replicas].score_topk(q, beams, k, out=(P["out_s"], P["out_d"]), flags=path_flags)
        else:
            return P["retr"][i % replicas].score_topk(q, beams, k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fork(cur):
        for P in pipes:
            P["stream"].wait_stream(cur)

    def join(cur):
        for P in pipes:
            cur.wait_stream(P["stream"])

    def run_phases(n, cur, e2e=False):
        """n steps, each issued as three calls (inversion / scoring / top-k) on the phase streams; capturable into one graph."""
        for s in phase_streams:
            s.wait_stream(cur)
        done = [None] * n_pipe
        for i in range(n):
            P = pipes[i % n_pipe]
            st = P["stores"][i % replicas]
            q, beams = (P["q_in"], P["b_in"]) if e2e else batches[i % n_batches]
            out = (P["o_s"], P["o_d"]) if e2e else (P["out_s"], P["out_d"])
            if e2e:
                with torch.cuda.stream(s_h2d):
                    if done[i % n_pipe] is not None:
                        s_h2d.wait_event(done[i % n_pipe])           # the previous batch of this slot no longer reads its inputs
                    P["in_dev"].copy_(in_host[i % n_batches], non_blocking=True)
                    e0_ = torch.cuda.Event()
                    e0_.record(s_h2d)
            with torch.cuda.stream(s_inv):
                if e2e:
                    s_inv.wait_event(e0_)
                elif done[i % n_pipe] is not None:
                    s_inv.wait_event(done[i % n_pipe])               # ... nor its scratch
                st.score_topk(q, beams, k, out=out, flags=path_flags | SK_S | SK_T)
                e1_ = torch.cuda.Event()
                e1_.record(s_inv)
            s_sc = s_scs[i % len(s_scs)]
            with torch.cuda.stream(s_sc):
                s_sc.wait_event(e1_)
                st.score_topk(q, beams, k, out=out, flags=path_flags | SK_I | SK_T)
                e2_ = torch.cuda.Event()
                e2_.record(s_sc)
            s_tk = s_tks[i % len(s_tks)]
            with torch.cuda.stream(s_tk):
                s_tk.wait_event(e2_)
                st.score_topk(q, beams, k, out=out, flags=path_flags | SK_I | SK_S)
                if e2e:
                    P["res"].copy_(P["out_dev"], non_blocking=True)
                e3_ = torch.cuda.Event()
                e3_.record(s_tk)
                done[i % n_pipe] = e3_
        for s in phase_streams:
            cur.wait_stream(s)

    def run_steps(n, cur):
        """n steps round-robin over the pipes' streams (fork/join on `cur`, so it is capturable into one graph)."""
        if fused:
            run_fused(n, cur)
            return
        if phases:
            run_phases(n, cur)
            return
        if n_pipe == 1:
            for i in range(n):
                step(i)
            return
        fork(cur)
        for i in range(n):
            P = pipes[i % n_pipe]
            with torch.cuda.stream(P["stream"]):
                step(i, P)
        join(cur)

    for P in pipes:                                   # every (pipe, replica) handle allocates its scratch once
        with torch.cuda.stream(P["stream"]):
            for i in range(max(args.warmup, replicas)):
                step(i, P)
    barrier()
    stats = stores[0].last_stats()

    # ---- fused schedule (EXPERIMENT, only on request or when the autotune's child verified it and found it faster):
    # launch i scores batch i and, in the same persistent CTAs, selects the top-k of batch i-1 (gdr_score_fused); the inversion
    # of batch i+1 runs one batch ahead on a second stream.  Three scratch sets: launch i scores into set i % 3, reads set
    # (i-1) % 3 for the top-k, and the inversion of batch i+1 fills set (i+1) % 3.  The fused grid leaves 8 SMs to the inversion
    # (k_scan's 1,024-thread CTA does not fit beside an 832-thread fused CTA; 108-140 scoring CTAs measured the same speed).
    fused = not sharded and args.schedule == "fused" and not args.no_graph
    fused_launches = None
    if fused:
        os.environ["GDR_UMMA_CTAS"] = os.environ.get("GDR_FUSED_CTAS", "140")
        fh = [[ClusterStore(s0.emb, torch.as_tensor(s0.offsets_host), s0.docid) for s0 in stores] for _ in range(3)]
        os.environ.pop("GDR_UMMA_CTAS")
        f_out = [(torch.empty((B_rank, k), dtype=torch.float32, device=dev), torch.empty((B_rank, k), dtype=torch.int32, device=dev)) for _ in range(3)]
        s_inv = torch.cuda.Stream()

        def run_fused(n, cur, keep=None):
            """n steps + the flush of the last batch; capturable.  keep: list that receives clones of every batch's result."""
            s_inv.wait_stream(cur)
            ev_f = {}
            for i in range(n):
                q, beams = batches[i % n_batches]
                h_cur = fh[i % 3][i % replicas]
                with torch.cuda.stream(s_inv):
                    if i - 2 in ev_f:
                        s_inv.wait_event(ev_f[i - 2])         # the batch that last used this scratch set has had its top-k
                    h_cur.invert(q, beams, k, flags=path_flags)
                    e_inv = torch.cuda.Event()
                    e_inv.record(s_inv)
                cur.wait_event(e_inv)
                with torch.cuda.stream(cur):
                    r = h_cur.score_fused(fh[(i - 1) % 3][(i - 1) % replicas] if i else None, out=f_out[(i - 1) % 3] if i else None)
                    if keep is not None and r is not None:
                        keep.append((r[0].clone(), r[1].clone()))
                    ev_f[i] = torch.cuda.Event()
                    ev_f[i].record(cur)
            with torch.cuda.stream(cur):
                r = ClusterStore.flush_fused(fh[(n - 1) % 3][(n - 1) % replicas], out=f_out[(n - 1) % 3])
                if keep is not None:
                    keep.append((r[0].clone(), r[1].clone()))
            cur.wait_stream(s_inv)

        # results first: every batch through the fused sequence must equal gdr_score_topk's result bit for bit
        try:
            got = []
            run_fused(3 * n_batches, torch.cuda.current_stream(), keep=got)
            torch.cuda.synchronize()
            for i, (gs, gd) in enumerate(got):
                rs, rd = stores[i % replicas].score_topk(batches[i % n_batches][0], batches[i % n_batches][1], k, flags=path_flags)
                if not (torch.equal(gs, rs) and torch.equal(gd, rd)):
                    raise RuntimeError(f"fused schedule: batch {i} differs from gdr_score_topk")
            fused_launches = int(fh[0][0].last_stats()["launches"]) + 1          # inversion kernels + the fused launch
        except Exception as e:
            if args.probe:
                print(json.dumps({"probe": True, "us_per_step": None, "failed": f"fused: {e}"[:200]}))
                return
            fused = False                             # fall back to the default schedule; said in config.schedule_note
            autotune = dict(autotune or {}, fused_rejected_in_parent=str(e)[:200])

    # CUDA graph of `period` consecutive steps (every replica / batch / pipe combination once), replayed: no host launch latency
    period = math.lcm(replicas, n_batches, 3 if fused else n_pipe)
    if n_pipe > 1 or fused:
        period *= max(2, -(-80 // period))        # the pipeline drains at every graph boundary: amortise it over >= 80 steps
    if args.steps < period:
        period = max(1, args.steps)               # short runs: one graph of exactly --steps steps
    use_graph = not sharded and not args.no_graph
    graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                run_steps(period, side)
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        barrier()
    steps = args.steps                            # exactly --steps steps are timed: whole replays + one shorter graph for the remainder
    graph_rem, rem = None, (args.steps % period if use_graph else 0)
    if rem:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph_rem = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_rem, stream=side):
                run_steps(rem, side)
        torch.cuda.current_stream().wait_stream(side)
        graph_rem.replay()
        barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    if use_graph:
        for _ in range(steps // period):
            graph.replay()
        if graph_rem is not None:
            graph_rem.replay()
    else:
        run_steps(steps, torch.cuda.current_stream())
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    qps = steps * B_global / (ms * 1e-3)
    if args.probe:                                # child of autotune_launch_config: two more timed regions, median, one line, done
        reps = [ms]
        for _ in range(2):
            e0.record()
            for _ in range(steps // period):
                graph.replay()
            if graph_rem is not None:
                graph_rem.replay()
            e1.record()
            barrier()
            reps.append(e0.elapsed_time(e1))
        print(json.dumps({"probe": True, "us_per_step": sorted(reps)[1] / steps * 1e3, "reps_us_per_step": [r / steps * 1e3 for r in reps],
                          "launch_priorities": os.environ.get("GDR_LAUNCH_PRIORITIES", "0"), "batches_in_flight": n_pipe,
                          "schedule": "fused" if fused else ("phases" if phases else "batches")}))
        return

    # ---- per-phase device time of the dominant kernel (CUDA events recorded inside the library, same stream)
    phase = {"invert": 0.0, "score_umma": 0.0, "score_simt": 0.0, "topk": 0.0}
    n_prof = min(steps, 4 * period)
    for s in stores:
        s.set_profiling(True)
    for i in range(n_prof):
        torch.cuda._sleep(2_000_000)      # ~1 ms of GPU idle-spin so the host runs ahead: events then see back-to-back kernels
        step(i)
        for key, v in stores[i % replicas].last_phase_ms().items():
            phase[key] += v / n_prof
    for s in stores:
        s.set_profiling(False)

    # ---- average launch duration of the scoring kernel: the scoring phase alone (GDR_SKIP_INVERT | GDR_SKIP_TOPK), launched
    # back to back on one stream over alternating store replicas (each launch streams a replica the previous one did not),
    # between two CUDA events; replayed from a CUDA graph so the host's launch rate cannot show up in the number.
    kernel_ms = None
    if not sharded:
        SK_I, SK_T = 256, 1024
        P0 = pipes[0]
        for r in range(replicas):                # leaves batch r's inversion in replica r's scratch
            P0["stores"][r].score_topk(batches[r % n_batches][0], batches[r % n_batches][1], k, out=(P0["out_s"], P0["out_d"]), flags=path_flags)
        torch.cuda.synchronize()
        n_rep = 20 * replicas

        def score_only(n):
            for i in range(n):
                r = i % replicas
                q_, b_ = batches[r % n_batches]
                P0["stores"][r].score_topk(q_, b_, k, out=(P0["out_s"], P0["out_d"]), flags=path_flags | SK_I | SK_T)

        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not args.no_graph:
            # the launches are replayed from a CUDA graph, so the host's launch rate cannot show up in the number
            kside = torch.cuda.Stream()
            kside.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(kside):
                kgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(kgraph, stream=kside):
                    score_only(n_rep)
            torch.cuda.current_stream().wait_stream(kside)
            kgraph.replay()
            torch.cuda.synchronize()
            times = []
            for _ in range(5):
                k0.record()
                kgraph.replay()
                k1.record()
                torch.cuda.synchronize()
                times.append(k0.elapsed_time(k1) / n_rep)
            kernel_ms = sorted(times)[len(times) // 2]
        else:
            torch.cuda._sleep(8_000_000)             # park the GPU while the host enqueues
            k0.record()
            score_only(n_rep)
            k1.record()
            torch.cuda.synchronize()
            kernel_ms = k0.elapsed_time(k1) / n_rep

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region.
    # Every step copies ITS inputs host->device and ITS results device->host; steps are issued round-robin on
    # `n_pipe` CUDA streams (each with its own device/host buffers and its own store handle = its own scratch),
    # so step i+1's H2D overlaps step i's kernels and step i-1's D2H — the way a serving loop would run it.
    # One pinned host buffer per batch holding q then beams, one device input buffer and one device/host result buffer per
    # pipe: ONE H2D and ONE D2H copy per step (each copy node costs a few microseconds on top of its bytes).
    q_bytes, b_bytes, r_bytes = B_rank * D * 4, B_rank * K * 4, B_rank * k * 4
    in_host = []
    for qb, bb in batches:
        h = torch.empty(q_bytes + b_bytes, dtype=torch.uint8).pin_memory()
        h[:q_bytes].view(torch.float32).view(B_rank, D).copy_(qb.cpu())
        h[q_bytes:].view(torch.int32).view(B_rank, K).copy_(bb.cpu())
        in_host.append(h)
    for P in pipes:
        P["in_dev"] = torch.empty(q_bytes + b_bytes, dtype=torch.uint8, device=dev)
        P["q_in"] = P["in_dev"][:q_bytes].view(torch.float32).view(B_rank, D)
        P["b_in"] = P["in_dev"][q_bytes:].view(torch.int32).view(B_rank, K)
        P["out_dev"] = torch.empty(2 * r_bytes, dtype=torch.uint8, device=dev)
        P["o_s"] = P["out_dev"][:r_bytes].view(torch.float32).view(1, B_rank, k)
        P["o_d"] = P["out_dev"][r_bytes:].view(torch.int32).view(1, B_rank, k)
        P["res"] = torch.empty(2 * r_bytes, dtype=torch.uint8).pin_memory()
        P["res_s"] = P["res"][:r_bytes].view(torch.float32).view(B_rank, k)
        P["res_d"] = P["res"][r_bytes:].view(torch.int32).view(B_rank, k)

    # The H2D copies run back to back on a stream of their own (the 3.2 MB of queries per step is what bounds the end-to-end
    # rate, so the copy engine must never wait for a batch's compute or D2H); a slot's input buffer is refilled once the
    # batch that last used it has finished.
    s_in = torch.cuda.Stream()

    def e2e_step(i, done):
        P = pipes[i % n_pipe]
        with torch.cuda.stream(s_in):
            if done[i % n_pipe] is not None:
                s_in.wait_event(done[i % n_pipe])
            P["in_dev"].copy_(in_host[i % n_batches], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        with torch.cuda.stream(P["stream"]):
            P["stream"].wait_event(ev_in)
            if not sharded:
                P["stores"][i % replicas].score_topk(P["q_in"], P["b_in"], k, out=(P["o_s"], P["o_d"]), flags=path_flags)
            else:
                s_, d_ = P["retr"][i % replicas].score_topk(P["q_in"], P["b_in"], k)
                P["o_s"][0].copy_(s_)
                P["o_d"][0].copy_(d_)
            P["res"].copy_(P["out_dev"], non_blocking=True)
            ev_done = torch.cuda.Event()
            ev_done.record(P["stream"])
            done[i % n_pipe] = ev_done

    def run_e2e(n, cur):
        if phases:
            run_phases(n, cur, e2e=True)
            return
        fork(cur)
        s_in.wait_stream(cur)
        done = [None] * n_pipe
        for i in range(n):
            e2e_step(i, done)
        join(cur)
        cur.wait_stream(s_in)

    run_e2e(2 * n_pipe * replicas, torch.cuda.current_stream())      # warm-up: every (pipe, replica) handle has its scratch
    barrier()
    cur = torch.cuda.current_stream()
    # the copies are graph nodes too (fixed pinned buffers, as a serving loop that refills them would use), so the host
    # only replays: the number is bounded by PCIe and the GPU, not by Python launch overhead
    e2e_graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            e2e_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e2e_graph, stream=side):
                run_e2e(period, side)
        cur.wait_stream(side)
        e2e_graph.replay()
        barrier()
    e2e_steps = max(period, (min(steps, 960) // period) * period) if use_graph else max(12, min(steps, 96))
    # what the copies alone cost on this box (same buffers, one stream per direction): the floor under the end-to-end number
    pcie = {}
    for name, dst, src, nbytes in (("h2d", pipes[0]["in_dev"], in_host[0], q_bytes + b_bytes), ("d2h", pipes[0]["res"], pipes[0]["out_dev"], 2 * r_bytes)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        pcie[name + "_us_per_step"] = e0.elapsed_time(e1) * 1000 / 50
        pcie[name + "_GBps"] = nbytes * 50 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # PCIe on a shared host is noisy: the timed region is repeated five times and the median segment reported
    e2e_segments = []
    for _ in range(5):
        e0.record()
        if use_graph:
            for _ in range(e2e_steps // period):
                e2e_graph.replay()
        else:
            run_e2e(e2e_steps, cur)
        e1.record()
        barrier()
        e2e_segments.append(e0.elapsed_time(e1))
    e2e_ms = sorted(e2e_segments)[len(e2e_segments) // 2]
    if not sharded:
        # the pipelined end-to-end loop must return what a plain serial call returns for the same batch
        last = (period if use_graph else e2e_steps) - 1
        P = pipes[last % n_pipe]
        chk_s, chk_d = stores[last % replicas].score_topk(batches[last % n_batches][0], batches[last % n_batches][1], k, flags=path_flags)
        torch.cuda.synchronize()
        if os.environ.get("GDR_TOPK_DEBUG"):
            pass                                  # timing experiment: the top-k is cut short, results are meaningless
        elif not torch.equal(P["res_d"], chk_d.cpu()) or not torch.equal(P["res_s"], chk_s.cpu()):
            raise RuntimeError("end-to-end pipeline result differs from the serial call")
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_qps = e2e_steps * B_global / (e2e_ms * 1e-3)
    h2d = B_rank * D * 4 + B_rank * K * 4          # per rank and step
    d2h = B_rank * k * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (every touched embedding read once + queries + results)
    peak, peak_src = peaks()
    beams0 = batches[0][1]
    lo_c = rank * cfg["C"] if sharded else 0
    local = beams0[(beams0 >= lo_c) & (beams0 < lo_c + cfg["C"])] - lo_c
    emb_touched = int(stores[0].sizes_host[torch.unique(local).cpu().numpy()].sum()) * D * esize
    alg_bytes = emb_touched + B_rank * D * 4 + B_rank * k * 8      # rank 0's launch
    dominant = max(("score_umma", "score_simt"), key=lambda n: phase[n])
    dom_ms = kernel_ms if kernel_ms else phase[dominant]     # phase[]: one event-bracketed launch inside a full call (includes launch gaps)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_ms = ms / steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath) and world == 1 and args.path == "auto":
        t = json.load(open(tpath)).get(args.workload)
        if t and t["kernel"] in {"score_umma": "k_score_umma", "score_simt": "k_score_simt"}[dominant]:
            traffic = t["dram_read_bytes"] + t["dram_write_bytes"]       # one ncu --set full capture of this workload (profiles/)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": {"score_umma": "k_score_umma (tcgen05 grouped GEMM)", "score_simt": "k_score_simt (GEMV)"}[dominant],
                "kernel_ms": dom_ms, "kernel_ms_method": "median of 5 replays of a CUDA graph of 80 back-to-back launches of the scoring phase, between two CUDA events" if kernel_ms else
                "one event-bracketed launch inside a full call", "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "phase_ms": phase, "whole_step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    if fused:
        roofline["note"] = ("the timed step launches k_score_topk_fused (the scoring CTA below plus top-k groups of the previous batch); kernel_ms / "
                            "achieved are the scoring phase alone (k_score_umma), whole_step_frac is the fused step")

    cpu = None
    if not args.no_cpu_baseline:
        nq = min(256, B_rank)
        n_cpu = min(cfg["N"], 400000)       # bounded sample of the corpus for huge shards
        c_cpu = int(np.searchsorted(stores[0].offsets_host, n_cpu, side="right") - 1)
        n_cpu = int(stores[0].offsets_host[c_cpu])
        emb_cpu = stores[0].emb[:n_cpu].float().cpu()
        lb = batches[0][1][:nq].cpu()
        if sharded or c_cpu < cfg["C"]:   # the CPU leg scores (a prefix of) the rank-0 shard only
            lb = torch.where((lb >= 0) & (lb < c_cpu), lb, torch.full_like(lb, -1))
        v, done, dt = cpu_reference_qps(emb_cpu, torch.as_tensor(stores[0].offsets_host[:c_cpu + 1]), stores[0].docid[:n_cpu].cpu().numpy().astype("int64"),
                                        batches[0][0][:nq].cpu(), lb, k, min_seconds=10.0, max_queries=20000)
        cpu = {"value": v, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{done} queries of the {args.workload} workload in {dt:.1f} s: oracle/gdr_oracle.dense_topk (reference dense.py:53-54 "
                         f"+ Tensor.topk per query), torch {torch.__version__} CPU with {os.cpu_count()} threads"}

    line = {
        "metric": "queries/sec, cluster-restricted scoring + top-%d" % k, "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if cfg.get("fp32") else "bf16", "data": "synthetic",
        "config": {"workload": args.workload + ("" if world == 1 else (f" x{world} cluster-sharded" if sharded else f" x{world} replicas, queries sharded")),
                   "precision": "fp32 embeddings x fp32 queries, fp32 FMA" if cfg.get("fp32") else "bf16 embeddings x fp32 queries (exact 3-term bf16 split), fp32 accumulate", "docs_per_gpu": cfg["N"],
                   "clusters_per_gpu": cfg["C"], "dim": D, "global_batch": B_global, "beam": K, "top_k": k,
                   "l2": f"{replicas} store replicas ({replicas * emb_bytes / 2**20:.0f} MB) and {n_batches} query batches cycled; inputs larger than L2",
                   "cuda_graph": bool(use_graph), "batches_in_flight": n_pipe, "schedule": ("fused (EXPERIMENT gdr_score_fused: one launch scores batch i and selects the top-k of batch i-1 in the same CTAs; inversion one batch "
                                "ahead on a second stream; fused grid of " + os.environ.get("GDR_FUSED_CTAS", "140") + " CTAs x " + os.environ.get("GDR_FUSED_GROUPS", "4") + " top-k groups; 3 scratch sets; results verified bit-identical to gdr_score_topk before timing; e2e and roofline "
                                "legs use the default schedule)") if fused else
                               ("phases (inversion / scoring x3 / top-k x3 streams, events)" if phases else "batches (one stream per batch in flight)"), "scoring_path": args.path,
                   "launch_priorities": os.environ.get("GDR_LAUNCH_PRIORITIES", "0") not in ("", "0"), "launch_autotune": autotune,
                   "parallelism": "single GPU" if world == 1 else (f"clusters sharded over {world} GPUs, queries replicated, NCCL all-gather of candidates + merge"
                                                                   if sharded else f"corpus replicated on {world} GPUs, queries sharded, no data-path collective")},
        "clocks": clocks, "gpu_launches": ((fused_launches * steps + -(-steps // period)) if fused else (int(stats["launches"]) * steps + (steps if sharded else 0))) * (1 if sharded else world),
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "segments_ms_per_step": [round(x / e2e_steps, 5) for x in e2e_segments], "estimator": "median of 5 timed segments",
                "copies_alone": {k_: round(v_, 2) for k_, v_ in pcie.items()},
                "pipeline": f"{n_pipe} batches in flight ({'phase' if phases else 'batch'} schedule), pinned host buffers, per step one H2D copy (q + beams) and one D2H copy (scores + docids)"},
        "roofline": roofline, "cpu_baseline": cpu,
        "path": {"simt_items": int(stats["simt_items"]), "umma_tiles": int(stats["umma_tiles"]), "clusters_touched": int(stats["clusters_touched"])},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
