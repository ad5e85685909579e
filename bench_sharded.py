"""bench.py --gpus N > 1: the north-star multi-GPU path on a cluster-sharded corpus, weak scaling.

Per GPU the shape is the workload's own (cfg2: 109,739 docs in 1,024 clusters, 1,024 owned queries per step), so the driver's
`v_N / (N * v_1)` compares like with like: the corpus is N shards = N x C clusters with GLOBAL cluster ids, the global batch is
N x B queries (identical on every rank — the scoring side of the path needs every query wherever one of its beams lives), each
rank owns the results of its B queries.  Per step and rank: invert the global batch, score the beams that land in the rank's
clusters, deliver the candidates to the owners, select the owners' top-k.

exchange `p2p` (default): the scoring epilogue stores every score straight into the owner's score buffer over NVLink and the
owner's top-k waits for per-rank arrival flags (gdr_b200.sharded.ShardedPipeline; five batches in flight).  exchange `nccl`: local top-k of
all N x B queries, NCCL all-gather of packed (score, docid) lists, merge (gdr_b200.sharded.ShardedRetriever) — the fallback when
peer mapping is unavailable, and the cross-check.  Before anything is timed, the sharded result of every rank is compared bit for
bit with a plain single-GPU call on the gathered corpus (when the gathered corpus fits: cfg1-3), and p2p with nccl.
"""
import json
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

import bench


def run(args, cfg, rank, world, local_rank, dev):
    from gdr_b200 import ClusterStore
    from gdr_b200.sharded import PeerAllGather, ShardedPipeline, ShardedRetriever

    k, K, D, B, C, N = cfg["k"], cfg["K"], cfg["D"], cfg["B"], cfg["C"], cfg["N"]
    flags = {"auto": 0, "simt": 2, "umma": 4}[args.path]
    esize = 4 if cfg.get("fp32") else 2
    emb_bytes = N * D * esize
    replicas = args.replicas or max(1, min(6, -(-640 * 2 ** 20 // emb_bytes)))
    B_g, C_g = B * world, C * world

    trace = bool(os.environ.get("GDR_BENCH_TRACE"))

    def mark(what):
        if trace:
            torch.cuda.synchronize()
            print(f"[rank {rank}] {what}", file=__import__("sys").stderr, flush=True)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def finish():
        """Leave without tearing down NCCL communicators / IPC mappings that CUDA graphs still reference (that teardown hung for the whole
        launcher timeout): every rank has reported, a last barrier, then a hard exit."""
        try:
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            import sys
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    # ---- the corpus: rank r generates shard r; the CSR metadata (cluster sizes, docids) is gathered and replicated
    emb_l, offsets_l, order_l = bench.synth_shard(cfg, 1234 + 1000 * rank, dev)
    sizes_l = torch.diff(offsets_l).to(torch.int32).to(dev)
    sizes_all = torch.empty(C_g, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(sizes_all, sizes_l)
    offsets_g = torch.zeros(C_g + 1, dtype=torch.int64)
    offsets_g[1:] = torch.cumsum(sizes_all.cpu().long(), 0)
    docid_l = (order_l + rank * N).to(torch.int32)             # docids are global: rank r's documents are numbered after those of ranks < r
    docid_g = torch.empty(N * world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(docid_g, docid_l)
    c_lo, c_hi = rank * C, (rank + 1) * C
    embs = [emb_l] + [emb_l.clone() for _ in range(replicas - 1)]
    shards = [ClusterStore.shard(e, offsets_g, docid_g, c_lo, c_hi) for e in embs]
    n_batches = 8
    batches = bench.synth_batches(cfg, n_batches, C_g, B_g, 4321, dev)      # same seed on every rank: the global batch is replicated
    own = slice(rank * B, (rank + 1) * B)

    # ---- exchange: peer-to-peer if it sets up and verifies, else NCCL lists
    notes = {}
    sp = None
    if args.exchange in ("auto", "p2p"):
        try:
            sp = ShardedPipeline(shards, rank, world, B, K, k, flags=flags, schedule="fused" if args.schedule == "fused" else "auto",
                                 depth=args.pipeline or 5, fused_ctas=args.fused_ctas, fused_groups=args.fused_groups)
        except Exception as e:             # e.g. no peer access between the GPUs of this box
            notes["p2p_setup_failed"] = f"{type(e).__name__}: {e}"[:300]
            sp = None
    ok_all = torch.tensor([1 if sp is not None else 0], device=dev)
    dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    if not int(ok_all.item()):
        if args.exchange == "p2p":
            raise RuntimeError("--exchange p2p: peer-to-peer setup failed on some rank: " + notes.get("p2p_setup_failed", "(another rank)"))
        sp = None
    # the NCCL path: a LOCAL view of the shard (local cluster ids, beams localized) + all-gather + merge
    g2l = torch.full((C_g,), -1, dtype=torch.int32, device=dev)
    g2l[c_lo:c_hi] = torch.arange(C, dtype=torch.int32, device=dev)
    local_store = ClusterStore(emb_l, offsets_l, docid_l)
    retr = ShardedRetriever(local_store, g2l)

    # (the cross-check against the p2p path must score with the SAME kernel family: the shard handle decides by the global pair density,
    # the local view would decide by its own — bf16 x 3-term tcgen05 and the fp32 GEMV differ in accumulation order, i.e. in the last bits)
    shard_umma = (not cfg.get("fp32")) and D % 64 == 0 and (args.path == "umma" or (args.path == "auto" and B_g * K >= 3 * C_g))
    nccl_flags = flags or (4 if shard_umma else 2)

    def nccl_step(i, f=0):
        q, beams = batches[i % n_batches]
        return retr.score_topk(q, beams, k, flags=f)

    mark('stores and pipelines built')
    # ---- results first
    checks = {}
    q0, b0 = batches[0]
    ns, nd = nccl_step(0, nccl_flags)
    mark('nccl cross-check call done')
    if sp is not None:
        t0 = sp.submit(q0, b0, flags=flags, which=0)
        t1 = sp.submit(batches[1][0], batches[1][1], flags=flags, which=1 % replicas)
        sp.flush()
        torch.cuda.synchronize()
        mark('p2p first batches done')
        same = bool(torch.equal(t0.scores, ns[own]) and torch.equal(t0.docids, nd[own]))
        checks["p2p_equals_nccl"] = same
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            if args.exchange == "p2p":
                raise RuntimeError("p2p-sharded result differs from the NCCL all-gather + merge result")
            notes["p2p_rejected"] = "result differed from the NCCL path on some rank"
            sp = None
    if world * emb_bytes <= 24 * 2 ** 30:
        # the whole corpus on every rank (setup only): a plain single-GPU call must give the same bits for this rank's queries
        full_emb = torch.empty((N * world, D), dtype=emb_l.dtype, device=dev)
        dist.all_gather_into_tensor(full_emb, emb_l)
        full = ClusterStore(full_emb, offsets_g, docid_g)
        fs, fd = full.score_topk(q0, b0, k, flags=flags)
        torch.cuda.synchronize()
        got_s, got_d = (t0.scores, t0.docids) if sp is not None else (ns[own], nd[own])
        same = bool(torch.equal(got_d, fd[own]) and torch.equal(got_s, fs[own]))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            raise RuntimeError("sharded result differs from the single-GPU call on the gathered corpus")
        checks["sharded_equals_single_gpu_on_gathered_corpus"] = True
        del full, full_emb, fs, fd
        torch.cuda.empty_cache()
    exchange = "p2p" if sp is not None else "nccl"

    # ---- the timed loop
    def run_steps(n):
        if sp is not None:
            for i in range(n):
                sp.submit(batches[i % n_batches][0], batches[i % n_batches][1], flags=flags, which=i % replicas)
            sp.flush()
        else:
            for i in range(n):
                nccl_step(i)

    run_steps(max(args.warmup, 2 * n_batches))
    barrier()
    mark('warm-up steps done')
    period = math.lcm(replicas, n_batches, 3, args.pipeline or 5)
    period *= max(2, -(-80 // period))
    if args.steps < period:
        period = max(1, args.steps)
    graph = graph_rem = None
    use_graph = not args.no_graph
    if use_graph:
        try:
            graph = capture(run_steps, period)
            mark('graph captured')
            graph.replay()
            mark('graph replayed once')
            rem = args.steps % period
            if rem:
                graph_rem = capture(run_steps, rem)
                graph_rem.replay()
        except Exception as e:
            notes["graph_capture_failed"] = f"{type(e).__name__}: {e}"[:200]
            graph = graph_rem = None
            use_graph = False
    ok_g = torch.tensor([int(use_graph)], device=dev)
    dist.all_reduce(ok_g, op=dist.ReduceOp.MIN)
    if not int(ok_g.item()):
        graph = graph_rem = None
        use_graph = False
    barrier()
    steps = args.steps
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
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
    mark('timed region done')
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    qps = steps * B_g / (ms * 1e-3)
    step_ms = ms / steps

    # ---- end to end with HOST buffers: every rank uploads ITS B queries (+ beams), an NCCL all-gather over NVLink assembles the
    # global batch on every rank (the collective of the data path), the pipeline runs, every rank downloads its B results
    q_bytes, b_bytes, r_bytes = B * D * 4, B * K * 4, B * k * 4
    in_host, out_host = [], []
    for qb, bb in batches:
        h = torch.empty(q_bytes + b_bytes, dtype=torch.uint8).pin_memory()
        h[:q_bytes].view(torch.float32).view(B, D).copy_(qb[own].cpu())
        h[q_bytes:].view(torch.int32).view(B, K).copy_(bb[own].cpu())
        in_host.append(h)
        out_host.append(torch.empty(2 * r_bytes, dtype=torch.uint8).pin_memory())
    n_slots = 6
    # the global batch is assembled on every rank by the copy engines over NVLink (PeerAllGather); NCCL all-gather only if that cannot be set up
    pag = None
    if sp is not None:
        try:
            pag = PeerAllGather(rank, world, [(B, D, torch.float32), (B, K, torch.int32)], n_slots, dev)
        except Exception as e:
            notes["peer_all_gather_failed"] = f"{type(e).__name__}: {e}"[:200]
            pag = None
    okp = torch.tensor([int(pag is not None)], device=dev)
    dist.all_reduce(okp, op=dist.ReduceOp.MIN)
    if not int(okp.item()):
        pag = None
    slots = [dict(own=torch.empty(q_bytes + b_bytes, dtype=torch.uint8, device=dev),
                  gq=pag.gathered(i_, 0) if pag is not None else torch.empty((B_g, D), dtype=torch.float32, device=dev),
                  gb=pag.gathered(i_, 1) if pag is not None else torch.empty((B_g, K), dtype=torch.int32, device=dev),
                  out=torch.empty(2 * r_bytes, dtype=torch.uint8, device=dev), free=None)
             for i_ in range(n_slots)]
    s_h2d, s_comm, s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        for s_ in (s_h2d, s_comm, s_d2h):
            s_.wait_stream(cur)
        for sl in slots:
            sl["free"] = None
        jobs = []

        def drain(final=False):
            while jobs and (jobs[0][0] is None or jobs[0][0].event is not None or final):
                tk, sl, oh = jobs.pop(0)
                with torch.cuda.stream(s_d2h):
                    if tk is not None:
                        s_d2h.wait_event(tk.event)
                    else:
                        s_d2h.wait_stream(cur)
                    oh.copy_(sl["out"], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_d2h)
                sl["free"] = ev

        for i in range(n):
            sl = slots[i % n_slots]
            with torch.cuda.stream(s_h2d):
                if sl["free"] is not None:
                    s_h2d.wait_event(sl["free"])
                sl["own"].copy_(in_host[i % n_batches], non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record(s_h2d)
            with torch.cuda.stream(s_comm):
                s_comm.wait_event(ev_up)
                if pag is not None:
                    pag.all_gather(i % n_slots, sl["own"])
                else:
                    dist.all_gather_into_tensor(sl["gq"], sl["own"][:q_bytes].view(torch.float32).view(B, D))
                    dist.all_gather_into_tensor(sl["gb"], sl["own"][q_bytes:].view(torch.int32).view(B, K))
                ev_g = torch.cuda.Event()
                ev_g.record(s_comm)
            cur.wait_event(ev_g)
            o_s = sl["out"][:r_bytes].view(torch.float32).view(B, k)
            o_d = sl["out"][r_bytes:].view(torch.int32).view(B, k)
            if sp is not None:
                tk = sp.submit(sl["gq"], sl["gb"], flags=flags, out=(o_s, o_d), which=i % replicas)
                jobs.append((tk, sl, out_host[i % n_batches]))
            else:
                s_, d_ = retr.score_topk(sl["gq"], sl["gb"], k)
                o_s.copy_(s_[own])
                o_d.copy_(d_[own])
                jobs.append((None, sl, out_host[i % n_batches]))
            drain()
        if sp is not None:
            sp.flush()
        drain(final=True)
        for s_ in (s_h2d, s_comm, s_d2h):
            cur.wait_stream(s_)

    e2e = None
    try:
        run_e2e(2 * n_batches)
        barrier()
        mark('e2e warm-up done')
        e2e_period = math.lcm(period, n_slots)
        e2e_graph = None
        if use_graph:
            try:
                e2e_graph = capture(run_e2e, e2e_period)
                mark('e2e graph captured')
                e2e_graph.replay()
                mark('e2e graph replayed once')
            except Exception as e:
                notes["e2e_graph_capture_failed"] = f"{type(e).__name__}: {e}"[:200]
                e2e_graph = None
            okg = torch.tensor([int(e2e_graph is not None)], device=dev)
            dist.all_reduce(okg, op=dist.ReduceOp.MIN)
            if not int(okg.item()):
                e2e_graph = None
        barrier()
        e2e_steps = max(e2e_period, (min(steps, 960) // e2e_period) * e2e_period)
        segs = []
        for _ in range(5):
            barrier()
            e0.record()
            if e2e_graph is not None:
                for _ in range(e2e_steps // e2e_period):
                    e2e_graph.replay()
            else:
                run_e2e(e2e_steps)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            segs.append(float(t.item()))
        e2e_ms = sorted(segs)[len(segs) // 2]
        # the downloaded results are the owner's slice of the single-GPU answer (checked above for batch 0 on the device path)
        chk = (sp.submit(q0, b0, flags=flags, which=0) if sp is not None else None)
        if sp is not None:
            sp.flush()
            torch.cuda.synchronize()
            ref_s, ref_d = chk.scores.cpu(), chk.docids.cpu()
        else:
            s_, d_ = nccl_step(0)
            torch.cuda.synchronize()
            ref_s, ref_d = s_[own].cpu(), d_[own].cpu()
        if not (torch.equal(out_host[0][:r_bytes].view(torch.float32).view(B, k), ref_s) and
                torch.equal(out_host[0][r_bytes:].view(torch.int32).view(B, k), ref_d)):
            raise RuntimeError("end-to-end sharded pipeline result differs from the device-resident result")
        # per-GPU copy bandwidth with all ranks copying at once (the host's PCIe / memory system is shared)
        pcie = {}
        for name, dst, src, nbytes in (("h2d", slots[0]["own"], in_host[0], q_bytes + b_bytes), ("d2h", out_host[0], slots[0]["out"], 2 * r_bytes)):
            barrier()
            e0.record()
            for _ in range(50):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            pcie[name + "_GBps_per_gpu_all_ranks_copying"] = round(nbytes * 50 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
        e2e = {"value": e2e_steps * B_g / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q_bytes + b_bytes, "d2h_bytes_per_step": 2 * r_bytes,
               "steps": e2e_steps, "segments_ms_per_step": [round(x / e2e_steps, 5) for x in segs], "estimator": "median of 5 timed segments, max over ranks",
               "cuda_graph": e2e_graph is not None, "copies": pcie,
               "all_gather": "copy engines over NVLink + arrival flags (gdr_b200.sharded.PeerAllGather)" if pag is not None else "NCCL all_gather_into_tensor",
               "pipeline": f"per rank and step: one H2D copy of the rank's own {B} queries + beams, all-gather of q ({B_g * D * 4} B) and beams over NVLink, "
                           f"the sharded pipeline, one D2H copy of the rank's {B} results; copy / collective / compute on separate streams, {n_slots} slots"}
    except Exception as e:
        notes["e2e_failed"] = f"{type(e).__name__}: {e}"[:300]
        ok_e = 0
    else:
        ok_e = 1
    flag = torch.tensor([ok_e], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not int(flag.item()):
        e2e = None

    if rank != 0:
        finish()

    peak, peak_src = bench.peaks()
    beams0 = batches[0][1]
    local = beams0[(beams0 >= c_lo) & (beams0 < c_hi)]
    emb_touched = int(shards[0].sizes_host[torch.unique(local).cpu().numpy()].sum()) * D * esize
    alg_bytes = emb_touched + B_g * D * 4 + B * k * 8      # this rank: its touched embeddings + EVERY query of the global batch + its own results
    stats = shards[0].last_stats() if sp is None else sp.handles[0].last_stats()
    launches = (sp.pr.launches() if sp is not None else int(local_store.last_stats()["launches"]) + 2)
    roofline = {"bound": "hbm", "achieved": alg_bytes / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                "traffic": None, "kernel": ("k_score_topk_fused_p2p" if sp is not None and sp.schedule == "fused" else ("k_score_umma_p2p" if sp is not None else "scoring kernel")) + " (per-rank step; the whole step is the unit here)",
                "kernel_ms": step_ms, "kernel_ms_method": "the timed region / steps (max over ranks)", "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "whole_step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                "note": "per rank; replicated queries count as algorithmic bytes (every rank reads all N x B queries once)"}
    cpu = None
    line = {
        "metric": bench.METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if cfg.get("fp32") else "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload} x{world} cluster-sharded", "docs_per_gpu": N, "clusters_per_gpu": C, "docs_total": N * world,
                   "clusters_total": C_g, "dim": D, "global_batch": B_g, "queries_owned_per_gpu": B, "beam": K, "top_k": k,
                   "precision": "fp32 embeddings x fp32 queries, fp32 FMA" if cfg.get("fp32") else "bf16 embeddings x fp32 queries (exact 3-term bf16 split), fp32 accumulate",
                   "l2": f"{replicas} shard replicas ({replicas * emb_bytes / 2**20:.0f} MB per GPU) and {n_batches} query batches cycled; inputs larger than L2",
                   "cuda_graph": bool(use_graph), "api": "gdr_b200.sharded.ShardedPipeline.submit" if sp is not None else "gdr_b200.sharded.ShardedRetriever.score_topk",
                   "exchange": exchange, "schedule": sp.schedule if sp is not None else "serial", "results_verified": checks, "notes": notes,
                   "parallelism": (f"clusters sharded over {world} GPUs (contiguous ranges of the global cluster numbering), global batch of {B_g} queries replicated for scoring, "
                                   + ("candidates exchanged by peer-to-peer stores over NVLink fused into the scoring epilogue (each score lands in its query owner's score buffer), "
                                      "per-rank arrival flags, top-k at the owner — no collective in the device-resident loop; NCCL all-gather of (score, docid) lists + merge is the verified fallback"
                                      if sp is not None else "local top-k, NCCL all-gather of (score, docid) candidates + merge top-k"))},
        "clocks": clocks, "gpu_launches": int(launches) * steps * world, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        "path": {"simt_items": int(stats["simt_items"]), "umma_tiles": int(stats["umma_tiles"]), "clusters_touched": int(stats["clusters_touched"])},
    }
    print(json.dumps(line), flush=True)
    finish()


def capture(fn, n):
    """CUDA graph of fn(n); thread-local capture mode so that NCCL's watchdog thread cannot invalidate the capture."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            fn(n)
    torch.cuda.current_stream().wait_stream(side)
    return g
