"""Pipelined fine stage: several independent batches in flight on one GPU.

The reference scores one validation batch at a time (GDR_model/main_models.py:1434-1637).  Batches are independent, so
this module keeps them in flight; it is the schedule `bench.py` times and the one a serving / validation loop should use:

* `fused` (bf16 store on the tcgen05 path, k <= 128): ONE launch per batch scores batch i and, in the same persistent
  CTAs, selects the top-k of batch i-1 (`gdr_score_fused`, csrc/score_fused.cu); the pair inversion of batch i+1 runs one
  batch ahead on a second stream.  Three scratch sets (= three store handles over the same device arrays): launch i scores
  into set i % 3 and reads set (i-1) % 3, the inversion of batch i+1 fills set (i+1) % 3.
* `batches` (every other shape): whole `gdr_score_topk` calls round-robin on `depth` streams, one handle each, so the
  latency-bound inversion and top-k kernels of one batch hide under the scoring kernel of its neighbours.
* `partitioned` (opt-in; tcgen05 shapes): the SMs are split into two disjoint sets (CUDA green contexts, `SmPartition`): the
  inversion and the top-k of every batch run on streams of the small set, the scoring kernels on streams of the big set, so the
  two sides do not compete for residency on the same SMs.  Same handles / scratch sets as `batches`.

Which one (cfg2 on a B200, us per 1,024-query step, inputs on the device): partitioned 41-46, batches with launch priorities 51,
fused 53-54, batches 54 — `bench.py` measures them and picks; 'auto' keeps the static rule fused-where-eligible-else-batches.  Behind
host-to-device copies (`submit_host`) the step is bound by the copy and `batches` is the fastest (66 vs 73 us for `partitioned`).

Contract: `submit()` returns a `Ticket`; the ticket's outputs are complete, in stream order on the stream that calls it, after
`ticket.wait()` (which needs two further `submit()`s or a `flush()` to have been issued: in the fused schedule `submit(i)`
enqueues the inversion of batch i and the launch that scores batch i-1 and selects batch i-2).  Everything is enqueue-only (no host synchronisation) and capturable in a CUDA graph as long as a
`flush()` is captured last (it joins the internal streams).  Results are bit-identical to `ClusterStore.score_topk`
(tests/test_gpu_pipeline.py).  There is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _cabi
from .store import ClusterStore


class Ticket:
    """One submitted batch.  `scores` [B, k] fp32 / `docids` [B, k] int32 are valid after `wait()`."""
    __slots__ = ("index", "scores", "docids", "event", "keep", "alpha", "host_out", "which", "ev_inv")

    def __init__(self, index, scores, docids, alpha, keep):
        self.index, self.scores, self.docids, self.alpha, self.keep = index, scores, docids, alpha, keep
        self.event: Optional[torch.cuda.Event] = None      # recorded once the launch that produces the outputs is enqueued
        self.host_out = None

    def wait(self, stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.event is None:
            raise RuntimeError("this batch's top-k has not been issued yet: submit the next batch or call flush() first")
        (stream or torch.cuda.current_stream()).wait_event(self.event)
        return self.scores, self.docids


class SmPartition:
    """Two disjoint SM sets of one device with streams of their own (include/gdr_b200.h gdr_partition_*; csrc/partition.cu).
    `big` / `small` are lists of torch streams (torch.cuda.ExternalStream over the library's green-context streams)."""

    def __init__(self, device, small_sms: int, n_big: int = 2, n_small: int = 5):
        import ctypes
        self._handle = None
        self.device = torch.device(device)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().gdr_partition_create(ctypes.byref(h), int(small_sms), int(n_big), int(n_small)))
        self._handle = h
        sms = (ctypes.c_int32 * 2)()
        _cabi.check(_cabi.lib().gdr_partition_sms(h, sms))
        self.sms_big, self.sms_small = int(sms[0]), int(sms[1])
        self.big = [torch.cuda.ExternalStream(int(_cabi.lib().gdr_partition_stream(h, 0, i)), device=self.device) for i in range(n_big)]
        self.small = [torch.cuda.ExternalStream(int(_cabi.lib().gdr_partition_stream(h, 1, i)), device=self.device) for i in range(n_small)]

    def close(self) -> None:
        if self._handle is not None:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                _cabi.lib().gdr_partition_destroy(self._handle)
            self._handle, self.big, self.small = None, [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PipelinedRetriever:
    def __init__(self, store, schedule: str = "auto", depth: int = 0, fused_ctas: int = 0, fused_groups: int = 0,
                 launch_priorities: bool = False, small_sms: int = 56, big_streams: int = 2, scoring_ctas_per_sm: int = 0):
        """store: the resident ClusterStore — or a list of stores of identical shape (several indexes served by one pipeline;
        bench.py cycles copies of the corpus so that every step streams embeddings that are not in L2), chosen per batch with
        `submit(..., which=i)`.  schedule: 'auto' (= fused where eligible, else batches) | 'fused' | 'batches' | 'partitioned'
        (SM partition: `small_sms` SMs for inversion + top-k, the rest for scoring on `big_streams` streams, with
        `scoring_ctas_per_sm` = 2 persistent scoring CTAs per SM by default — the 4-stage kernel of csrc/score_umma_x2.cu; falls back to
        'batches' for shapes that do not take the tcgen05 path).  depth: batches in flight for 'batches' / 'partitioned' (default 5; the fused
        schedule always uses three scratch sets)."""
        if schedule not in ("auto", "fused", "batches", "partitioned"):
            raise ValueError("schedule must be 'auto', 'fused', 'batches' or 'partitioned'")
        self.stores = list(store) if isinstance(store, (list, tuple)) else [store]
        store = self.stores[0]
        self.store, self.schedule_request = store, schedule
        self.depth = depth if depth > 0 else 5
        self.fused_ctas, self.fused_groups, self.launch_priorities = fused_ctas, fused_groups, launch_priorities
        self.small_sms, self.big_streams = small_sms, big_streams
        self.scoring_ctas_per_sm = int(scoring_ctas_per_sm) if scoring_ctas_per_sm else (2 if schedule == "partitioned" else 1)
        self.partition: Optional[SmPartition] = None     # 'partitioned' schedule: created on first use
        self.dev = store.emb.device
        self._fused_handles: Optional[List[List[ClusterStore]]] = None    # [scratch set][store]
        self._batch_handles: Optional[List[List[ClusterStore]]] = None    # [stream][store]
        self._streams: List[torch.cuda.Stream] = []
        self._s_inv: Optional[torch.cuda.Stream] = None
        self._n = 0                      # batches submitted so far
        self._inverted: Optional[Ticket] = None   # fused: inverted, its scoring launch not yet issued
        self._scored: Optional[Ticket] = None     # fused: scored, its top-k (part of the next launch) not yet issued
        self._launch_events = {}         # fused: index -> event recorded after launch i (scratch reuse ordering)
        self._open: List[Ticket] = []    # batches schedule: tickets not yet joined by flush()
        self._dirty = {}                 # internal streams that have work since the last flush() (only those are joined: a stream that took
                                         # no part in a CUDA graph capture must not be waited for inside it)
        self.last_schedule = None
        # host-buffer front end (submit_host): staging slots, one copy stream per direction
        self._slots: List[dict] = []
        self._s_h2d: Optional[torch.cuda.Stream] = None
        self._s_d2h: Optional[torch.cuda.Stream] = None
        self._host_jobs: List[Tuple[Ticket, dict, object]] = []

    # ------------------------------------------------------------------------------------------------------------------
    def fused_eligible(self, B: int, K: int, k: int, flags: int = 0) -> bool:
        """Mirror of gdr_score_fused's requirements (include/gdr_b200.h): the batch takes the tcgen05 path alone and the
        previous batch's top-k is the small-footprint select."""
        s = self.store
        if s.emb.dtype != torch.bfloat16 or s.dim % 64 != 0 or (flags & _cabi.FORCE_SIMT) or (flags & _cabi.Q_PER_BEAM):
            return False
        if not ((flags & _cabi.FORCE_UMMA) or B * K >= 3 * s.n_clusters):
            return False
        stride = (K * s.max_cluster + 3) // 4 * 4
        return k <= 128 and stride <= 65535

    def partition_eligible(self, B: int, K: int, flags: int = 0) -> bool:
        """The batch takes the tcgen05 path alone (one scoring kernel whose persistent CTAs can be sized to the big SM set)."""
        s = self.store
        if s.emb.dtype != torch.bfloat16 or s.dim % 64 != 0 or (flags & _cabi.FORCE_SIMT) or getattr(s, "p2p", None):
            return False
        return bool((flags & _cabi.FORCE_UMMA) or B * K >= 3 * s.n_clusters)

    def _handles_partitioned(self) -> List[List[ClusterStore]]:
        hs = self._handles_batches()
        if self.partition is None:
            self.partition = SmPartition(self.dev, self.small_sms, self.big_streams, self.depth)
            for h in (h for row in hs for h in row):
                h.set_option("umma_ctas_per_sm", self.scoring_ctas_per_sm)
                h.set_option("umma_ctas", self.scoring_ctas_per_sm * self.partition.sms_big)
        return hs

    def _handles_fused(self) -> List[List[ClusterStore]]:
        if self._fused_handles is None:
            self._fused_handles = [[s.clone_handle() for s in self.stores] for _ in range(3)]
            for h in (h for hs in self._fused_handles for h in hs):
                if self.fused_ctas:
                    h.set_option("umma_ctas", self.fused_ctas)
                if self.fused_groups:
                    h.set_option("fused_groups", self.fused_groups)
                if self.launch_priorities:
                    h.set_option("launch_priorities", 1)
            self._s_inv = torch.cuda.Stream(device=self.dev)
        return self._fused_handles

    def _handles_batches(self) -> List[List[ClusterStore]]:
        if self._batch_handles is None:
            self._batch_handles = [[s.clone_handle() for s in self.stores] for _ in range(self.depth)]
            for h in (h for hs in self._batch_handles for h in hs):
                if self.launch_priorities:
                    h.set_option("launch_priorities", 1)
                if self.scoring_ctas_per_sm > 1:
                    h.set_option("umma_ctas_per_sm", self.scoring_ctas_per_sm)
                    h.set_option("umma_ctas", self.scoring_ctas_per_sm * torch.cuda.get_device_properties(self.dev).multi_processor_count)
            self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.depth)]
        return self._batch_handles

    def reserve(self, B: int, K: int, k: int, flags: int = 0) -> "PipelinedRetriever":
        """Allocate every scratch set for this batch shape now (no allocation / synchronisation on the query path later)."""
        fused = self.schedule_request in ("auto", "fused") and self.fused_eligible(B, K, k, flags)
        for hs in (self._handles_fused() if fused else self._handles_batches()):
            for h in hs:
                h.reserve(B, K, k, flags)
        return self

    # ------------------------------------------------------------------------------------------------------------------
    def submit(self, q: torch.Tensor, beams: torch.Tensor, k: int, prob: Optional[torch.Tensor] = None, alpha: float = 1.0,
               act: Optional[str] = "none", flags: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               which: int = 0) -> Ticket:
        """Enqueue one batch (same arguments as ClusterStore.score_topk with a single alpha) against store number `which`.
        q / beams / prob must not be overwritten until the ticket's outputs are complete (the ticket keeps references)."""
        B, K = int(beams.shape[0]), int(beams.shape[1])
        fused = self.schedule_request in ("auto", "fused") and self.fused_eligible(B, K, k, flags)
        if self.schedule_request == "fused" and not fused:
            raise ValueError("this batch shape is not eligible for the fused schedule (see gdr_score_fused in include/gdr_b200.h)")
        sched = "fused" if fused else ("partitioned" if self.schedule_request == "partitioned" and self.partition_eligible(B, K, flags) else "batches")
        if self.last_schedule is not None and self.last_schedule != sched:
            self.flush()                                       # the schedule changes with the shape: drain first
        self.last_schedule = sched
        if out is None:
            B_out = self.store.p2p[2] if getattr(self.store, "p2p", None) else B
            out = (torch.empty((B_out, k), dtype=torch.float32, device=self.dev), torch.empty((B_out, k), dtype=torch.int32, device=self.dev))
        t = Ticket(self._n, out[0], out[1], float(alpha), (q, beams, prob))
        t.which = which
        cur = torch.cuda.current_stream(self.dev)
        if fused:
            self._submit_fused(t, q, beams, k, prob, act, flags, cur)
        elif sched == "partitioned":
            self._submit_partitioned(t, q, beams, k, prob, act, flags, cur)
        else:
            self._submit_batch(t, q, beams, k, prob, act, flags, cur)
        self._n += 1
        return t

    def _submit_fused(self, t, q, beams, k, prob, act, flags, cur):
        """Batch i: its inversion goes to the side stream NOW; the fused launch that scores batch i-1 (and selects batch i-2) is
        enqueued right behind it, so the inversion of batch i runs beside the scoring of batch i-1 — one batch ahead."""
        hs = self._handles_fused()
        i = t.index
        h = hs[i % 3][t.which]
        ev_in = torch.cuda.Event()
        ev_in.record(cur)                                      # the inputs are ready in stream order here
        self._dirty[id(self._s_inv)] = self._s_inv
        with torch.cuda.stream(self._s_inv):
            self._s_inv.wait_event(ev_in)
            if i - 2 in self._launch_events:                   # the batch that last used this scratch set has had its top-k
                self._s_inv.wait_event(self._launch_events.pop(i - 2))
            h.invert(q, beams, k, prob=prob, act=act, flags=flags)
            t.ev_inv = torch.cuda.Event()
            t.ev_inv.record(self._s_inv)
        if self._inverted is not None:
            self._launch_fused(self._inverted, cur)
        self._inverted = t

    def _launch_fused(self, t, cur):
        """ONE launch: score batch `t` (inverted earlier) and select the top-k of the batch scored by the previous launch."""
        hs = self._handles_fused()
        prev = self._scored
        cur.wait_event(t.ev_inv)
        with torch.cuda.stream(cur):
            hs[t.index % 3][t.which].score_fused(hs[prev.index % 3][prev.which] if prev is not None else None,
                                                 alpha=prev.alpha if prev is not None else 1.0,
                                                 out=(prev.scores, prev.docids) if prev is not None else None)
            ev = torch.cuda.Event()
            ev.record(cur)
        self._launch_events[t.index] = ev
        if prev is not None:
            prev.event, prev.keep = ev, None
        self._scored = t

    def _submit_batch(self, t, q, beams, k, prob, act, flags, cur):
        hs = self._handles_batches()
        i = t.index
        st = self._streams[i % self.depth]
        self._dirty[id(st)] = st
        ev_in = torch.cuda.Event()
        ev_in.record(cur)
        with torch.cuda.stream(st):
            st.wait_event(ev_in)
            hs[i % self.depth][t.which].score_topk(q, beams, k, prob=prob, alphas=[t.alpha], act=act, flags=flags,
                                          out=(t.scores.view(1, *t.scores.shape), t.docids.view(1, *t.docids.shape)))
            t.event = torch.cuda.Event()
            t.event.record(st)
        self._open.append(t)
        if len(self._open) > self.depth:                       # bounded bookkeeping; older tickets keep their own events
            self._open.pop(0)

    def _submit_partitioned(self, t, q, beams, k, prob, act, flags, cur):
        """The three phases of batch i as three calls on the same handle: inversion and top-k on small-set stream i % depth, scoring
        on big-set stream i % big_streams; events carry the order.  The handle's next batch (i + depth) starts on the same small-set
        stream, i.e. behind this batch's top-k, so a scratch set is never written while it is read."""
        hs = self._handles_partitioned()
        part, i = self.partition, t.index
        h = hs[i % self.depth][t.which]
        ss, bs = part.small[i % self.depth], part.big[i % len(part.big)]
        self._dirty[id(ss)], self._dirty[id(bs)] = ss, bs
        ev_in = torch.cuda.Event()
        ev_in.record(cur)
        with torch.cuda.stream(ss):
            ss.wait_event(ev_in)
            h.invert(q, beams, k, prob=prob, act=act, flags=flags)
            ev_inv = torch.cuda.Event()
            ev_inv.record(ss)
        with torch.cuda.stream(bs):
            bs.wait_event(ev_inv)
            h.score_topk(q, beams, k, prob=prob, alphas=[t.alpha], act=act, flags=flags | _cabi.SKIP_INVERT | _cabi.SKIP_TOPK,
                         out=h._fused_dummy, _unchecked_out=True)
            ev_sc = torch.cuda.Event()
            ev_sc.record(bs)
        with torch.cuda.stream(ss):
            ss.wait_event(ev_sc)
            h.score_topk(q, beams, k, prob=prob, alphas=[t.alpha], act=act, flags=flags | _cabi.SKIP_INVERT | _cabi.SKIP_SCORE,
                         out=(t.scores.view(1, *t.scores.shape), t.docids.view(1, *t.docids.shape)))
            t.event = torch.cuda.Event()
            t.event.record(ss)
        self._open.append(t)
        if len(self._open) > self.depth:
            self._open.pop(0)

    # ---- host-buffer front end: what a caller with inputs in (pinned) host memory uses ---------------------------------------
    def submit_host(self, in_host: torch.Tensor, B: int, K: int, k: int, out_host: torch.Tensor, alpha: float = 1.0,
                    act: Optional[str] = "none", flags: int = 0, which: int = 0) -> Ticket:
        """One batch whose inputs live in ONE pinned host buffer `in_host` (uint8: fp32 q [B, D] followed by int32 beams
        [B, K]) and whose results go to ONE pinned host buffer `out_host` (uint8: fp32 scores [B, k] then int32 docids [B, k]):
        a single H2D and a single D2H copy per batch, each direction on a copy stream of its own so that batch i+1's upload,
        batch i's kernels and batch i-1's download overlap.  `ticket.host_out` is the event after which `out_host` is valid."""
        D = self.store.dim
        q_bytes, b_bytes, r_bytes = B * D * 4, B * K * 4, B * k * 4
        if in_host.numel() != q_bytes + b_bytes or out_host.numel() != 2 * r_bytes or in_host.dtype != torch.uint8 or out_host.dtype != torch.uint8:
            raise ValueError("in_host must hold q then beams, out_host scores then docids, both as uint8 buffers")
        if self._s_h2d is None:
            self._s_h2d, self._s_d2h = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)
        n_slots = 5 if self.schedule_request in ("auto", "fused") and self.fused_eligible(B, K, k, flags) else self.depth + 1
        if len(self._slots) != n_slots or self._slots[0]["in"].numel() != q_bytes + b_bytes or self._slots[0]["out"].numel() != 2 * r_bytes:
            self.flush()
            self._slots = [dict(**{"in": torch.empty(q_bytes + b_bytes, dtype=torch.uint8, device=self.dev),
                                   "out": torch.empty(2 * r_bytes, dtype=torch.uint8, device=self.dev)}, free=None) for _ in range(n_slots)]
        slot = self._slots[self._n % n_slots]
        self._dirty[id(self._s_h2d)], self._dirty[id(self._s_d2h)] = self._s_h2d, self._s_d2h
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self._s_h2d):
            if slot["free"] is not None:
                self._s_h2d.wait_event(slot["free"])           # the batch that last used this slot has been downloaded
            else:
                self._s_h2d.wait_stream(cur)
            slot["in"].copy_(in_host, non_blocking=True)
            ev_up = torch.cuda.Event()
            ev_up.record(self._s_h2d)
        cur.wait_event(ev_up)
        q = slot["in"][:q_bytes].view(torch.float32).view(B, D)
        beams = slot["in"][q_bytes:].view(torch.int32).view(B, K)
        o_s = slot["out"][:r_bytes].view(torch.float32).view(B, k)
        o_d = slot["out"][r_bytes:].view(torch.int32).view(B, k)
        t = self.submit(q, beams, k, alpha=alpha, act=act, flags=flags, out=(o_s, o_d), which=which)
        self._host_jobs.append((t, slot, out_host))
        self._drain_host_jobs()
        return t

    def _drain_host_jobs(self) -> None:
        """Enqueue the download of every batch whose top-k has been issued."""
        while self._host_jobs and self._host_jobs[0][0].event is not None:
            t, slot, out_host = self._host_jobs.pop(0)
            with torch.cuda.stream(self._s_d2h):
                self._s_d2h.wait_event(t.event)
                out_host.copy_(slot["out"], non_blocking=True)
                t.host_out = torch.cuda.Event()
                t.host_out.record(self._s_d2h)
            slot["free"] = t.host_out

    def flush_scoring(self) -> None:
        """First half of `flush()`: issue the scoring launch that is still outstanding.  Only callers that drive SEVERAL pipelines
        from one host thread on one stream need it separately (the one-process tests of the peer-to-peer exchange: every rank's last
        scoring launch must be enqueued before any rank's final top-k, which waits for all of them)."""
        if self._inverted is not None:
            self._launch_fused(self._inverted, torch.cuda.current_stream(self.dev))
            self._inverted = None

    def flush(self) -> None:
        """Issue what is still outstanding (fused: the last scoring launch and the last batch's top-k) and join the internal streams
        into the current one."""
        cur = torch.cuda.current_stream(self.dev)
        self.flush_scoring()
        if self._scored is not None:
            t, self._scored = self._scored, None
            hs = self._handles_fused()
            with torch.cuda.stream(cur):
                ClusterStore.flush_fused(hs[t.index % 3][t.which], alpha=t.alpha, out=(t.scores, t.docids))
                t.event = torch.cuda.Event()
                t.event.record(cur)
            t.keep = None
        self._drain_host_jobs()
        host = self._s_h2d is not None and id(self._s_h2d) in self._dirty
        for st in self._dirty.values():
            cur.wait_stream(st)
        self._dirty.clear()
        if host:
            for slot in self._slots:
                slot["free"] = None                            # everything is joined into `cur`: the next upload waits for `cur`
        self._launch_events.clear()
        self._open.clear()

    def launches(self) -> int:
        """Kernels launched per pipelined step (synchronises): inversion kernels + the fused launch, or the whole call."""
        hs = self._fused_handles if self.last_schedule == "fused" else self._batch_handles
        n = int(hs[0][0].last_stats()["launches"])
        return n + 1 if self.last_schedule == "fused" else n
