#!/usr/bin/env python
"""bench.py — queries/sec of cluster-restricted scoring + top-k (BASELINE.json `metric`), measured through the product API
(`gdr_b200.PipelinedRetriever` / `gdr_b200.sharded.ShardedPipeline`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg1|cfg2|cfg3|cfg5s]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (pair inversion -> gather-and-score -> per-query top-k) over one batch of synthetic
queries.  N = 1 runs BASELINE.json configs[1] ("cfg2": 109,739 x 768 bf16 corpus, 1,024 k-means-shaped clusters, batch 1,024
queries, beam 20, top-100).  N > 1 (weak scaling, one process per GPU) runs the north-star multi-GPU path on the SAME per-GPU
shape: the corpus is N x 109,739 docs in N x 1,024 clusters SHARDED BY CLUSTER, the global batch is N x 1,024 queries
(replicated for scoring, each rank owns the results of its own 1,024), every rank scores the beams that land in its clusters,
the candidates travel to the query's owner and the owner selects the top-k — by default with peer-to-peer stores over
NVLink fused into the scoring epilogue (`exchange: p2p`), else by NCCL all-gather of (score, docid) lists + merge
(`exchange: nccl`); `--mode replica` keeps round 1's N independent replicas.  The sharded result is checked bit for bit against
a single-GPU call on the gathered corpus before anything is timed.
Timing: W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream, max over
ranks; the steps are replayed from a CUDA graph.  L2 hygiene: each step streams a different replica of the store
(`config.l2`: together several times the 126 MB L2) and a different query batch.
`e2e` is the same metric through the same API with HOST buffers: one pinned H2D and one D2H copy per step inside the timed
region, on a copy stream per direction (`PipelinedRetriever.submit_host`).
`--impl reference` (and `cpu_baseline`) time the reference's own dense.py path on the host cores: the unmodified
`DenseModel.compute_similarity` from baseline/_ref/ (kind "reference"), or the oracle port when that copy is absent (kind "port").
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

# The pipelines keep a dozen streams busy and, on a sharded corpus, some one-warp kernels spin until a peer GPU's kernel has run: with the
# default of 8 hardware work queues two streams can share a queue and a spinning kernel could sit in front of the work it waits for.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec, cluster-restricted scoring + top-k"      # ONE string for both arms (the driver refuses to divide otherwise)

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
    CSR by stable sort.  Returns (emb [N, D] cluster-contiguous, offsets [C+1] cpu, docid [N])."""
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
        if C_total <= 8192 and B * C_total <= (1 << 26):
            beams = torch.argsort(torch.rand((B, C_total), generator=g, device=device), dim=1)[:, :cfg["K"]]
        else:  # huge C: sample with replacement, duplicates are vanishingly rare and legal
            beams = torch.randint(0, C_total, (B, cfg["K"]), generator=g, device=device)
        out.append((q.contiguous(), beams.to(torch.int32).contiguous()))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU path (reference arm and cpu_baseline leg: the ONLY places bench.py executes oracle/ or baseline/_ref)
# ---------------------------------------------------------------------------------------------------------------------
def load_cpu_reference():
    """Returns (fn(q, emb, offsets, docid, beams, k) -> (scores, docids), kind, description).  kind "reference": the gather loop
    below calls the UNMODIFIED `DenseModel.compute_similarity` of the reference's dense.py (baseline/_ref/, a git-ignored copy
    shipped by __graft_entry__.build()) + `Tensor.topk` (main_models.py:1625); kind "port": oracle/gdr_oracle.dense_topk."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gdr_oracle as orc
    try:
        import ref_shims
        if not ref_shims.reference_available():
            raise FileNotFoundError("no reference tree")
        dense = ref_shims.load_ref_dense()
        sim = dense.DenseModel.compute_similarity              # dense.py:53-54, called unbound (it never touches self)

        def fn(q, emb, offsets, docid, beams, k):
            B = q.shape[0]
            out_s = torch.full((B, k), float("-inf"))
            out_d = torch.full((B, k), -1, dtype=torch.int64)
            docid_t = torch.as_tensor(np.asarray(docid).astype(np.int64))
            for b in range(B):                                # the per-query gather of main_models.py:1441-1462 (CSR slices instead of per-doc cat)
                rows = [torch.arange(int(offsets[c]), int(offsets[c + 1])) for c in beams[b].tolist() if c >= 0]
                if not rows:
                    continue
                rows = torch.cat(rows)
                if rows.numel() == 0:
                    continue
                s = sim(None, q[b:b + 1], emb[rows])[0]
                kk = min(k, s.numel())
                v, i = s.topk(kk, largest=True, sorted=True)
                out_s[b, :kk], out_d[b, :kk] = v, docid_t[rows[i]]
            return out_s, out_d

        return fn, "reference", f"unmodified DenseModel.compute_similarity of {os.path.relpath(ref_shims.REF_MODEL_DIR, ROOT)}/dense.py:53-54 + Tensor.topk per query"
    except Exception as e:        # the copy is absent (or its imports fail on this box): the restatement, said so in `kind`
        return (lambda q, emb, offsets, docid, beams, k: orc.dense_topk(q, emb, offsets, docid, beams, k)), "port", \
            f"oracle/gdr_oracle.dense_topk (restatement of dense.py:53-54 + Tensor.topk; reference copy unavailable: {type(e).__name__})"


def cpu_reference_qps(fn, emb_cpu, offsets, docid, q_cpu, beams_cpu, k, min_seconds, max_queries):
    torch.set_num_threads(os.cpu_count())
    off = offsets.numpy()
    chunk = 64
    fn(q_cpu[:8], emb_cpu, off, docid, beams_cpu[:8].numpy(), k)     # warm-up
    done, t0 = 0, time.perf_counter()
    while done < max_queries and (time.perf_counter() - t0 < min_seconds or done == 0):
        lo = done % q_cpu.shape[0]
        sl = slice(lo, lo + chunk)
        fn(q_cpu[sl], emb_cpu, off, docid, beams_cpu[sl].numpy(), k)
        done += min(chunk, q_cpu[sl].shape[0])
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def reference_arm(args, cfg):
    """bench.py --impl reference: the reference's CPU path on this box's host cores, same config / metric / unit."""
    k, K, D = cfg["k"], cfg["K"], cfg["D"]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gdr_oracle as orc
    fn, kind, what = load_cpu_reference()
    small = cfg["N"] <= 200000
    emb, offsets, docid = orc.synth_corpus(cfg["N"] if small else 200000, cfg["C"] if small else 2048, D, seed=1234)
    if not cfg.get("fp32"):
        emb = emb.bfloat16().float()
    C = offsets.size - 1
    q, beams, _ = orc.synth_queries(256, C, K, D, seed=4321)
    per_step = 64
    torch.set_num_threads(os.cpu_count())
    for _ in range(args.warmup):
        fn(q[:8], emb, offsets, docid, beams[:8], k)
    t0 = time.perf_counter()
    for i in range(args.steps):
        lo = (i * per_step) % 256
        fn(q[lo:lo + per_step], emb, offsets, docid, beams[lo:lo + per_step], k)
    dt = time.perf_counter() - t0
    qps = args.steps * per_step / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "docs": int(emb.shape[0]), "clusters": int(C), "dim": D, "beam": K, "top_k": k,
                   "queries_per_step": per_step},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{args.steps} steps x {per_step} queries of the {args.workload} workload: {what}; torch {torch.__version__} CPU, "
                                   f"{os.cpu_count()} threads"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def capture(fn, n):
    """CUDA graph of fn(n) (fn forks / joins its own streams from the current one)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn(n)
    torch.cuda.current_stream().wait_stream(side)
    return g


def probe_main(args, cfg, dev):
    """Child of the launch autotune: verify one schedule bit for bit against the serial call, time it, print one line."""
    from gdr_b200 import ClusterStore, PipelinedRetriever
    k, B = cfg["k"], cfg["B"]
    emb, offsets, docid = synth_shard(cfg, 1234, dev)
    stores = [ClusterStore(emb, offsets, docid)] + [ClusterStore(emb.clone(), offsets, docid) for _ in range(args.replicas - 1)]
    batches = synth_batches(cfg, 8, cfg["C"], B, 4321, dev)
    flags = {"auto": 0, "simt": 2, "umma": 4}[args.path]
    pr = PipelinedRetriever(stores, schedule=args.schedule, depth=args.pipeline, fused_ctas=args.fused_ctas, fused_groups=args.fused_groups,
                            launch_priorities=args.launch_priorities == "on", small_sms=args.small_sms, big_streams=args.big_streams,
                            scoring_ctas_per_sm=args.ctas_per_sm).reserve(B, cfg["K"], k, flags)
    R, nb = len(stores), len(batches)

    def run(n, keep=None):
        for i in range(n):
            t = pr.submit(batches[i % nb][0], batches[i % nb][1], k, flags=flags, which=i % R)
            if keep is not None:
                keep.append(t)
        pr.flush()

    try:
        got = []
        run(3 * nb, got)
        torch.cuda.synchronize()
        for i, t in enumerate(got):
            rs, rd = stores[i % R].score_topk(batches[i % nb][0], batches[i % nb][1], k, flags=flags)
            if not (torch.equal(t.scores, rs) and torch.equal(t.docids, rd)):
                raise RuntimeError(f"batch {i} differs from gdr_score_topk")
        period = math.lcm(R, nb, 3, args.pipeline or 5)
        period *= max(1, -(-80 // period))
        g, graph_error = None, None
        if not args.no_graph:
            inner = []

            def run_guarded(n):
                try:
                    run(n)
                except Exception as e_:         # the error that invalidates a capture is raised here; capture_end would mask it
                    inner.append(f"{type(e_).__name__}: {e_}"[:200])
                    raise
            try:
                g = capture(run_guarded, period)
                g.replay()
                torch.cuda.synchronize()
            except Exception as e:              # (reported; the caller treats a candidate that cannot be captured as failed)
                g, graph_error = None, (inner[0] if inner else f"{type(e).__name__}: {e}")[:200]
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = []
        for _ in range(3):
            e0.record()
            for _ in range(max(1, args.steps // period)):
                g.replay() if g is not None else run(period)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1) * 1e3 / (max(1, args.steps // period) * period))
        if graph_error is not None:
            raise RuntimeError(f"host-launched step {sorted(reps)[1]:.2f} us; CUDA graph capture failed: {graph_error}")
        if pr.last_schedule != args.schedule and args.schedule != "auto":
            raise RuntimeError(f"asked for schedule '{args.schedule}', the pipeline ran '{pr.last_schedule}'")
        extra = {"sms": [pr.partition.sms_big, pr.partition.sms_small]} if pr.partition is not None else {}
        print(json.dumps({"probe": True, "us_per_step": sorted(reps)[1], "reps_us_per_step": reps, "schedule": pr.last_schedule,
                          "verified_identical_to_serial": True, **extra}))
    except Exception as e:
        print(json.dumps({"probe": True, "us_per_step": None, "failed": f"{type(e).__name__}: {e}"[:300]}))


SCHEDULES = ["batches", "fused", "partitioned"]


def autotune(args, local_rank):
    """Launch autotune (like a cuDNN-style algorithm search, but over schedules): rank 0 times a short device-resident run of
    the workload under each candidate in a CHILD process (a candidate that faults or hangs costs its own timeout), each child
    first checking every batch's result bit for bit against the serial gdr_score_topk; the fastest verified candidate wins, the
    default (`batches`) stays if nothing beats it by > 3 %.  All timings / failures are reported in config.launch_autotune."""
    cands = [("batches", dict(schedule="batches", pipeline=5))]
    if args.workload == "cfg2" and args.path == "auto":
        # Candidates in the order of how they measured (the time budget cuts the tail, not the head).  SM partition (green contexts):
        # inversion + top-k on a small SM set, scoring on the rest; us per step: one 6-stage scoring CTA per SM: 40 SMs -> 52.3,
        # 48 -> 48.4, 56 -> 50.6, 64 -> 52.7, 72 -> 55.1; two 4-stage CTAs per SM (k_score_umma_x2): 48 -> 45.7, 56 -> 44.3, 64 -> 45.6, 72 -> 47.5
        cands += [("partitioned_56x2", dict(schedule="partitioned", pipeline=5, small_sms=56, ctas_per_sm=2)),
                  ("batches_priorities", dict(schedule="batches", pipeline=5, launch_priorities="on")),
                  ("partitioned_48x2", dict(schedule="partitioned", pipeline=5, small_sms=48, ctas_per_sm=2)),
                  ("partitioned_64x2", dict(schedule="partitioned", pipeline=5, small_sms=64, ctas_per_sm=2)),
                  ("fused_140", dict(schedule="fused", fused_groups=5, fused_ctas=140)),
                  ("partitioned_48", dict(schedule="partitioned", pipeline=5, small_sms=48, ctas_per_sm=1)),
                  ("fused64_140", dict(schedule="fused", fused_groups=9, fused_ctas=140))]
    else:
        cands.append(("batches_priorities", dict(schedule="batches", pipeline=5, launch_priorities="on")))
    report, best, t_start = {}, None, time.time()
    for name, opt in cands:
        if time.time() - t_start > 240:
            report[name] = {"skipped": "autotune time budget spent"}
            continue
        env = {k_: v_ for k_, v_ in os.environ.items() if k_ not in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK",
                                                                     "TORCHELASTIC_RUN_ID", "MASTER_ADDR", "MASTER_PORT")}
        env["LOCAL_RANK"] = str(local_rank)
        cmd = [sys.executable, os.path.abspath(__file__), "--probe", "--gpus", "1", "--steps", "1920", "--warmup", "3", "--workload", args.workload,
               "--path", args.path, "--replicas", str(args.replicas or 4)]
        for k_, v_ in opt.items():
            cmd += ["--" + k_.replace("_", "-"), str(v_)]
        try:
            out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120)
            lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            line = json.loads(lines[-1]) if out.returncode == 0 and lines else {}
            if line.get("us_per_step"):
                report[name] = {"us_per_step": round(line["us_per_step"], 3), "schedule": line.get("schedule"), "verified_identical_to_serial": True}
                if line.get("sms"):
                    report[name]["sms_scoring_and_topk"] = line["sms"]
                if best is None or line["us_per_step"] < best[0]:
                    best = (line["us_per_step"], name, opt)
            else:
                report[name] = {"failed": (line.get("failed") or out.stderr or out.stdout)[-300:]}
        except Exception as e:
            report[name] = {"failed": repr(e)[:200]}
    base = report.get("batches", {}).get("us_per_step")
    if best is None or (base is not None and best[1] != "batches" and best[0] > 0.97 * base):
        report["chosen"] = "batches"
        return dict(schedule="batches", pipeline=5), report
    report["chosen"] = best[1]
    return best[2], report


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
                    help="N > 1: sharded (default) = clusters sharded over ranks, global batch replicated for scoring, candidates exchanged "
                         "to the query's owner, top-k at the owner; replica = every rank holds the corpus and its own batches (no collective)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="sharded mode: p2p = scores stored straight into the owner's score buffer over NVLink by the scoring epilogue; "
                         "nccl = local top-k, NCCL all-gather of (score, docid) lists, merge; auto = p2p if it sets up and verifies, else nccl")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "umma"], help="force a scoring path")
    ap.add_argument("--pipeline", type=int, default=0, help="batches in flight of the `batches` schedule (0 = 5; 1 = strictly serial)")
    ap.add_argument("--small-sms", type=int, default=56, help="`partitioned` schedule: SMs of the inversion + top-k side")
    ap.add_argument("--big-streams", type=int, default=2, help="`partitioned` schedule: streams of the scoring side")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="persistent tcgen05 scoring CTAs per SM (0 = the schedule's default: 2 for `partitioned`, else 1)")
    ap.add_argument("--schedule", default="auto", choices=["auto", "batches", "fused", "partitioned"],
                    help="PipelinedRetriever schedule; auto = launch autotune (N = 1) / the pipeline's own choice")
    ap.add_argument("--fused-ctas", type=int, default=0)
    ap.add_argument("--fused-groups", type=int, default=0)
    ap.add_argument("--launch-priorities", default="off", choices=["on", "off"])
    ap.add_argument("--no-graph", action="store_true", help="launch every step from the host instead of replaying a CUDA graph")
    ap.add_argument("--no-autotune", action="store_true")
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)     # child mode of the autotune
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    k, K, D = cfg["k"], cfg["K"], cfg["D"]

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, cfg)
        return

    import torch.distributed as dist
    from gdr_b200 import ClusterStore, PipelinedRetriever

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.probe:
        args.replicas = args.replicas or 4
        return probe_main(args, cfg, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sharded = world > 1 and args.mode != "replica"
    if sharded:
        import bench_sharded
        try:
            return bench_sharded.run(args, cfg, rank, world, local_rank, dev)
        except BaseException:          # a rank that fails must not leave the others waiting in a collective until the launcher's timeout
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)

    B = cfg["B"]
    flags = {"auto": 0, "simt": 2, "umma": 4}[args.path]
    esize = 4 if cfg.get("fp32") else 2
    emb_bytes = cfg["N"] * D * esize
    replicas = args.replicas or max(1, min(6, -(-640 * 2 ** 20 // emb_bytes)))     # >= 640 MB of distinct store bytes in rotation
    args.replicas = replicas

    # ---- schedule: measured, not assumed
    tune_report = None
    opt = dict(schedule=args.schedule, pipeline=args.pipeline, fused_ctas=args.fused_ctas, fused_groups=args.fused_groups,
               launch_priorities=args.launch_priorities, small_sms=args.small_sms, ctas_per_sm=args.ctas_per_sm)
    if args.schedule == "auto" and not args.no_autotune and not args.no_graph and args.pipeline != 1:
        decision = torch.zeros(7, dtype=torch.int32, device=dev)
        if rank == 0:
            best, tune_report = autotune(args, local_rank)
            decision = torch.tensor([SCHEDULES.index(best.get("schedule")), best.get("pipeline", 0), best.get("fused_ctas", 0),
                                     best.get("fused_groups", 0), int(best.get("launch_priorities") == "on"), best.get("small_sms", 56), best.get("ctas_per_sm", 0)],
                                    dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(decision, src=0)
        f, p_, fc, fg, lp, ssm, cps = (int(x) for x in decision.tolist())
        opt = dict(schedule=SCHEDULES[f], pipeline=p_, fused_ctas=fc, fused_groups=fg, launch_priorities="on" if lp else "off", small_sms=ssm, ctas_per_sm=cps)
    n_pipe = opt["pipeline"] if opt["pipeline"] > 0 else 5

    stores = []
    emb, offsets, docid = synth_shard(cfg, 1234, dev)
    for r in range(replicas):
        stores.append(ClusterStore(emb if r == 0 else emb.clone(), offsets, docid))
    n_batches = 8
    batches = synth_batches(cfg, n_batches, cfg["C"], B, 4321 + (0 if world == 1 else rank), dev)
    pr = PipelinedRetriever(stores, schedule=opt["schedule"] if opt["schedule"] != "auto" else "auto", depth=n_pipe, fused_ctas=opt["fused_ctas"],
                            fused_groups=opt["fused_groups"], launch_priorities=opt["launch_priorities"] == "on", small_sms=opt["small_sms"],
                            big_streams=args.big_streams, scoring_ctas_per_sm=opt["ctas_per_sm"]).reserve(B, K, k, flags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, keep=None):
        """n device-resident steps through the product pipeline (+ the flush that ends a sequence)."""
        for i in range(n):
            t = pr.submit(batches[i % n_batches][0], batches[i % n_batches][1], k, flags=flags, which=i % replicas)
            if keep is not None:
                keep.append(t)
        pr.flush()

    # ---- results first: the pipelined schedule must return what the serial call returns, bit for bit
    got = []
    run_steps(max(args.warmup, 2 * n_batches), got)
    barrier()
    for i, t in enumerate(got):
        rs, rd = stores[i % replicas].score_topk(batches[i % n_batches][0], batches[i % n_batches][1], k, flags=flags)
        if not (torch.equal(t.scores, rs) and torch.equal(t.docids, rd)):
            raise RuntimeError(f"pipelined schedule '{pr.last_schedule}': batch {i} differs from the serial gdr_score_topk")
    schedule = pr.last_schedule
    launches_per_step = pr.launches()
    stats = stores[0].last_stats()

    period = math.lcm(replicas, n_batches, 3 if schedule == "fused" else n_pipe)
    period *= max(2, -(-80 // period))            # the pipeline drains at every graph boundary: amortise it over >= 80 steps
    if args.steps < period:
        period = max(1, args.steps)
    use_graph = not args.no_graph
    graph = graph_rem = None
    rem = args.steps % period if use_graph else 0
    if use_graph:
        graph = capture(run_steps, period)
        graph.replay()
        if rem:
            graph_rem = capture(run_steps, rem)
            graph_rem.replay()
        barrier()
    steps = args.steps

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
        run_steps(steps)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    qps = steps * B * world / (ms * 1e-3)
    step_ms = ms / steps

    # ---- average launch duration of the dominant kernel, measured live: the kernel that the timed step launches, back to back on
    # one stream over alternating store replicas, replayed from a CUDA graph (the host's launch rate cannot show up), between two
    # CUDA events.  fused schedule: k_score_topk_fused (scoring of one batch + top-k of the previous one); else k_score_umma /
    # k_score_simt alone (GDR_SKIP_INVERT | GDR_SKIP_TOPK on a handle whose inversion is in place).
    SK_I, SK_T = 256, 1024
    h_k = [[s.clone_handle() for s in stores] for _ in range(2)]
    if schedule == "fused":
        for h in (h for hs in h_k for h in hs):
            if opt["fused_ctas"]:
                h.set_option("umma_ctas", opt["fused_ctas"])
            if opt["fused_groups"]:
                h.set_option("fused_groups", opt["fused_groups"])
    h_alone = h_k
    if schedule == "partitioned":                 # the kernel as the step runs it: CTAs = the scoring side's SMs, on a stream of that side
        for h in (h for hs in h_k for h in hs):
            h.set_option("umma_ctas_per_sm", pr.scoring_ctas_per_sm)
            h.set_option("umma_ctas", pr.scoring_ctas_per_sm * pr.partition.sms_big)
        h_alone = [[s.clone_handle() for s in stores] for _ in range(2)]          # ... and on the whole device for comparison
    dummy = (torch.empty((1, B, k), dtype=torch.float32, device=dev), torch.empty((1, B, k), dtype=torch.int32, device=dev))
    for s_ in range(2):
        for r in range(replicas):
            q_, b_ = batches[(2 * r + s_) % n_batches]
            h_k[s_][r].invert(q_, b_, k, flags=flags)
            if h_alone is not h_k:
                h_alone[s_][r].invert(q_, b_, k, flags=flags)
    torch.cuda.synchronize()
    n_rep = 20 * replicas
    out2 = (dummy[0][0], dummy[1][0])

    def kernel_only(n):
        if schedule == "partitioned":
            cur_, big_ = torch.cuda.current_stream(), pr.partition.big[0]
            big_.wait_stream(cur_)
            with torch.cuda.stream(big_):
                scoring_alone(n, h_k)
            cur_.wait_stream(big_)
            return
        for i in range(n):
            r, s_ = i % replicas, i % 2
            if schedule == "fused":
                h_k[s_][r].score_fused(h_k[1 - s_][(i - 1) % replicas] if i else None, out=out2 if i else None)
            else:
                q_, b_ = batches[(2 * r + s_) % n_batches]
                h_k[s_][r].score_topk(q_, b_, k, out=dummy, flags=flags | SK_I | SK_T)

    def scoring_alone(n, handles=None):
        hh = handles if handles is not None else h_alone
        for i in range(n):
            r, s_ = i % replicas, i % 2
            q_, b_ = batches[(2 * r + s_) % n_batches]
            hh[s_][r].score_topk(q_, b_, k, out=dummy, flags=flags | SK_I | SK_T)

    def time_graph(fn):
        g_ = capture(fn, n_rep) if use_graph else None
        if g_ is not None:
            g_.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            if g_ is None:
                torch.cuda._sleep(8_000_000)
            e0.record()
            g_.replay() if g_ is not None else fn(n_rep)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / n_rep)
        return sorted(ts)[len(ts) // 2]

    kernel_ms = time_graph(kernel_only)
    scoring_ms = time_graph(scoring_alone) if schedule in ("fused", "partitioned") else kernel_ms

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region: one H2D (q + beams) and one
    # D2H (scores + docids) per step, each direction on its own copy stream (PipelinedRetriever.submit_host)
    q_bytes, b_bytes, r_bytes = B * D * 4, B * K * 4, B * k * 4
    in_host, out_host = [], []
    for qb, bb in batches:
        h = torch.empty(q_bytes + b_bytes, dtype=torch.uint8).pin_memory()
        h[:q_bytes].view(torch.float32).view(B, D).copy_(qb.cpu())
        h[q_bytes:].view(torch.int32).view(B, K).copy_(bb.cpu())
        in_host.append(h)
        out_host.append(torch.empty(2 * r_bytes, dtype=torch.uint8).pin_memory())

    e2e_steps_box = []
    pcie = {}
    d_in = torch.empty(q_bytes + b_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(2 * r_bytes, dtype=torch.uint8, device=dev)
    for name, dst, src, nbytes in (("h2d", d_in, in_host[0], q_bytes + b_bytes), ("d2h", out_host[0], d_out, 2 * r_bytes)):
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

    def measure_e2e(pr_x):
        """Median-of-7 time of e2e_steps host-buffer steps through `pr_x`, results checked against the serial call."""
        def run_e2e(n):
            for i in range(n):
                pr_x.submit_host(in_host[i % n_batches], B, K, k, out_host[i % n_batches], flags=flags, which=i % replicas)
            pr_x.flush()

        for o in out_host:
            o.zero_()
        run_e2e(2 * n_batches)
        barrier()
        e2e_period = math.lcm(period, 4)
        e2e_graph = capture(run_e2e, e2e_period) if use_graph else None
        if e2e_graph is not None:
            e2e_graph.replay()
            barrier()
        e2e_steps = max(e2e_period, (min(steps, 1920) // e2e_period) * e2e_period)
        e2e_steps_box.append(e2e_steps)
        segments = []
        for _ in range(7):            # PCIe on a shared host is noisy: seven timed segments, the median is reported
            e0.record()
            if e2e_graph is not None:
                for _ in range(e2e_steps // e2e_period):
                    e2e_graph.replay()
            else:
                run_e2e(e2e_steps)
            e1.record()
            barrier()
            segments.append(e0.elapsed_time(e1))
        ms_ = sorted(segments)[len(segments) // 2]
        # the host buffers must hold what a plain serial call returns for the same batch
        for i in range(n_batches):
            j = max(x for x in range(e2e_period) if x % n_batches == i)          # the last step of a replay that wrote out_host[i]
            cs, cd = stores[j % replicas].score_topk(batches[i][0], batches[i][1], k, flags=flags)
            torch.cuda.synchronize()
            if not (torch.equal(out_host[i][:r_bytes].view(torch.float32).view(B, k), cs.cpu()) and
                    torch.equal(out_host[i][r_bytes:].view(torch.int32).view(B, k), cd.cpu())):
                raise RuntimeError(f"end-to-end pipeline ({pr_x.last_schedule}) result differs from the serial call")
        if world > 1:
            t_ = torch.tensor([ms_], device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms_ = float(t_.item())
        return ms_, segments

    # The host-buffer path is bound by the 3.2 MB H2D copy per step (PCIe), so what counts there is a short dependency chain behind each
    # copy, not SM residency: when the device-resident loop runs on the SM partition, the whole-call `batches` schedule is measured as
    # well and the faster of the two is reported (both are listed in e2e.by_schedule).
    e2e_cands = [(schedule, pr)]
    if schedule == "partitioned":
        e2e_cands.append(("batches", PipelinedRetriever(stores, schedule="batches", depth=n_pipe, launch_priorities=True).reserve(B, K, k, flags)))
    e2e_by = {name: measure_e2e(p_) for name, p_ in e2e_cands}
    e2e_sched = min(e2e_by, key=lambda n_: e2e_by[n_][0])
    e2e_ms, e2e_segments = e2e_by[e2e_sched]
    e2e_steps = e2e_steps_box[0]
    e2e_qps = e2e_steps * B * world / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (every touched embedding read once + queries + results)
    peak, peak_src = peaks()
    beams0 = batches[0][1]
    emb_touched = int(stores[0].sizes_host[torch.unique(beams0[beams0 >= 0]).cpu().numpy()].sum()) * D * esize
    alg_bytes = emb_touched + B * D * 4 + B * k * 8
    simt = int(stats["umma_tiles"]) == 0
    kname = ("k_score_topk_fused64" if opt["fused_groups"] == 9 else "k_score_topk_fused") if schedule == "fused" else (
        "k_score_simt" if simt else ("k_score_tile_f32" if cfg.get("fp32") else (
            "k_score_umma_x2" if schedule == "partitioned" and pr.scoring_ctas_per_sm == 2 else "k_score_umma")))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath) and args.path == "auto":
        entries = json.load(open(tpath)).get(args.workload) or []
        for t in (entries if isinstance(entries, list) else [entries]):
            if t["kernel"] == kname:
                traffic = t["dram_read_bytes"] + t["dram_write_bytes"]   # one ncu --set full capture of this workload (profiles/)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kname + {"k_score_simt": " (GEMV)", "k_score_umma": " (tcgen05 grouped GEMM)", "k_score_umma_x2": " (tcgen05 grouped GEMM, 4-stage ring, two CTAs per SM)", "k_score_tile_f32": " (shared-memory-tiled fp32, fma.rn.f32x2)"}.get(
                    kname, " (tcgen05 grouped GEMM of batch i + per-query top-k of batch i-1 in one persistent CTA per SM)") + (
                    f", as the timed step runs it: confined to the scoring side's {pr.partition.sms_big} of {pr.partition.sms_big + pr.partition.sms_small} SMs "
                    "(the other SMs run the top-k and inversion kernels of the neighbouring batches); `scoring_alone` is the same kernel on the whole device"
                    if schedule == "partitioned" else ""),
                "kernel_ms": kernel_ms,
                "kernel_ms_method": f"median of 5 replays of a CUDA graph of {n_rep} back-to-back launches of this kernel over alternating store "
                                    "replicas, between two CUDA events",
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "scoring_alone": {"kernel": "k_score_simt" if simt else ("k_score_tile_f32" if cfg.get("fp32") else "k_score_umma"), "kernel_ms": scoring_ms,
                                  "frac": alg_bytes / (scoring_ms * 1e-3) / 1e9 / peak},
                "whole_step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": achieved / 8000.0}

    cpu = None
    if not args.no_cpu_baseline:
        fn, kind, what = load_cpu_reference()
        nq = min(256, B)
        n_cpu = min(cfg["N"], 400000)       # bounded sample of the corpus for huge shards
        c_cpu = int(np.searchsorted(stores[0].offsets_host, n_cpu, side="right") - 1)
        n_cpu = int(stores[0].offsets_host[c_cpu])
        emb_cpu = stores[0].emb[:n_cpu].float().cpu()
        lb = batches[0][1][:nq].cpu()
        if c_cpu < cfg["C"]:
            lb = torch.where((lb >= 0) & (lb < c_cpu), lb, torch.full_like(lb, -1))
        v, done, dt = cpu_reference_qps(fn, emb_cpu, torch.as_tensor(stores[0].offsets_host[:c_cpu + 1]), stores[0].docid[:n_cpu].cpu().numpy().astype("int64"),
                                        batches[0][0][:nq].cpu(), lb, k, min_seconds=10.0, max_queries=20000)
        cpu = {"value": v, "unit": "queries/s", "cores": os.cpu_count(), "kind": kind,
               "sample": f"{done} queries of the {args.workload} workload in {dt:.1f} s: {what}; torch {torch.__version__} CPU with {os.cpu_count()} threads"}

    sched_txt = {"fused": f"fused (gdr_score_fused via PipelinedRetriever: one launch scores batch i and selects the top-k of batch i-1 in the same CTAs; "
                          f"inversion one batch ahead on a second stream; 3 scratch sets; grid of {opt['fused_ctas'] or 140} CTAs x {opt['fused_groups'] or 5} top-k groups)",
                 "partitioned": f"partitioned (PipelinedRetriever: SM partition by CUDA green contexts - inversion and top-k of every batch on {n_pipe} streams of a "
                                f"{pr.partition.sms_small if pr.partition else 0}-SM set, scoring kernels ({pr.scoring_ctas_per_sm} persistent CTA(s) per SM) on {args.big_streams} streams of the "
                                f"other {pr.partition.sms_big if pr.partition else 0} SMs; {n_pipe} batches in flight)",
                 "batches": f"batches (PipelinedRetriever: whole gdr_score_topk calls round-robin on {n_pipe} streams)"}[schedule]
    line = {
        "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if cfg.get("fp32") else "bf16", "data": "synthetic",
        "config": {"workload": args.workload + ("" if world == 1 else f" x{world} replicas, queries sharded"),
                   "precision": "fp32 embeddings x fp32 queries, fp32 FMA accumulate" if cfg.get("fp32") else "bf16 embeddings x fp32 queries (exact 3-term bf16 split), fp32 accumulate",
                   "docs_per_gpu": cfg["N"], "clusters_per_gpu": cfg["C"], "dim": D, "global_batch": B * world, "beam": K, "top_k": k,
                   "l2": f"{replicas} store replicas ({replicas * emb_bytes / 2**20:.0f} MB) and {n_batches} query batches cycled; inputs larger than L2",
                   "cuda_graph": bool(use_graph), "api": "gdr_b200.PipelinedRetriever.submit / submit_host", "schedule": sched_txt,
                   "results_verified": "every batch of the pipelined schedule bit-identical to the serial gdr_score_topk before timing",
                   "scoring_path": args.path, "launch_priorities": opt["launch_priorities"] == "on", "launch_autotune": tune_report,
                   "parallelism": "single GPU" if world == 1 else f"corpus replicated on {world} GPUs, queries sharded, no data-path collective"},
        "clocks": clocks, "gpu_launches": (launches_per_step * steps + -(-steps // period)) * world,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": q_bytes + b_bytes, "d2h_bytes_per_step": 2 * r_bytes, "steps": e2e_steps,
                "segments_ms_per_step": [round(x / e2e_steps, 5) for x in e2e_segments], "estimator": "median of 7 timed segments",
                "copies_alone": {k_: round(v_, 2) for k_, v_ in pcie.items()},
                "schedule": e2e_sched + (" + launch priorities" if e2e_sched == "batches" and schedule == "partitioned" else ""),
                "by_schedule": {n_: {"value": e2e_steps * B * world / (v_[0] * 1e-3), "ms_per_step": v_[0] / e2e_steps} for n_, v_ in e2e_by.items()},
                "pipeline": "PipelinedRetriever.submit_host: pinned host buffers, per step one H2D copy (q + beams) and one D2H copy (scores + docids), "
                            f"one copy stream per direction, {n_pipe + 1} staging slots"},
        "roofline": roofline, "cpu_baseline": cpu,
        "path": {"simt_items": int(stats["simt_items"]), "umma_tiles": int(stats["umma_tiles"]), "clusters_touched": int(stats["clusters_touched"])},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
