"""Cluster-sharded multi-GPU execution (one process per GPU, torch.distributed for the plumbing).

New design — the reference keeps a full replica of the pickle on every DDP rank
(GDR_model/main_models.py:806-814) and has no cross-rank step in retrieval.
Clusters are the independent units: each rank owns a subset of the clusters (balanced by document
count), scores only the beams that land in its clusters, and produces a local sorted top-k padded
with (-inf, -1).  The single exchange step is an all-gather of the packed (score, docid) candidates
(Q*k*8 bytes per rank — NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests),
followed by the merge top-k kernel (gdr_merge_topk).  Ordering is by (score desc, docid asc), which
does not depend on candidate order, so the merged result equals the single-GPU result exactly.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def partition_clusters(sizes: Sequence[int], world_size: int) -> np.ndarray:
    """Assign clusters to ranks, balanced by document count: largest cluster first onto the currently
    lightest rank (LPT).  Deterministic; returns owner[c] in [0, world_size)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(sizes.size, dtype=np.int32)
    if world_size == 1:
        owner[:] = 0
        return owner
    # LPT with a vectorised fast path: after the first pass loads are near-equal, so round-robin over
    # ranks sorted by load per block of `world_size` clusters is within one cluster of true LPT.
    for start in range(0, sizes.size, world_size):
        blk = order[start:start + world_size]
        ranks = np.argsort(load, kind="stable")[:blk.size]
        owner[blk] = ranks
        load[ranks] += sizes[blk]
    return owner


def partition_contiguous(sizes: Sequence[int], world_size: int) -> np.ndarray:
    """Split the cluster sequence into `world_size` CONTIGUOUS ranges balanced by document count (the form the peer-to-peer
    exchange needs: a shard is one slab of rows).  Returns bounds[world_size + 1]: rank r owns clusters [bounds[r], bounds[r+1])."""
    sizes = np.asarray(sizes, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(sizes)])
    targets = csum[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(csum, targets, side="left")
    bounds = np.concatenate([[0], cuts, [sizes.size]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def global_to_local(owner: np.ndarray, rank: int) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (g2l [C] int32: local cluster index or -1, local_clusters: the global ids this rank owns,
    ascending)."""
    mine = np.nonzero(owner == rank)[0]
    g2l = np.full(owner.size, -1, dtype=np.int32)
    g2l[mine] = np.arange(mine.size, dtype=np.int32)
    return g2l, mine


def pack_candidates(scores: torch.Tensor, docids: torch.Tensor) -> torch.Tensor:
    """[B, k] fp32 + [B, k] int32 -> one int32 buffer [2, B, k] (score bits, docids) so the exchange is a
    single collective."""
    return torch.stack([scores.contiguous().view(torch.int32), docids.to(torch.int32)], dim=0).contiguous()


class ShardedRetriever:
    """One rank's view of a cluster-sharded corpus.

    local_topk(q, local_beams, k, prob, alphas, act) -> (scores [n_alpha, B, k], docids [n_alpha, B, k])
    merge(gathered int32 [G, 2, B, k], k) -> (scores [B, k], docids [B, k])
    default to the CUDA implementations (ClusterStore.score_topk / gdr_merge_topk); the CPU (gloo)
    tests inject checkers for them — the product path never leaves the device.
    """

    def __init__(self, store, g2l: torch.Tensor, group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.store = store
        self.g2l = g2l
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_topk = local_topk if local_topk is not None else self._cuda_local_topk
        self._merge = merge if merge is not None else self._cuda_merge

    # ---- CUDA implementations ---------------------------------------------------------------------
    def _cuda_local_topk(self, q, local_beams, k, prob, alphas, act, flags=0):
        return self.store.score_topk(q, local_beams, k, prob=prob, alphas=alphas if alphas is not None else [1.0], act=act, flags=flags)

    @staticmethod
    def _cuda_merge(gathered: torch.Tensor, k: int):
        import ctypes  # noqa: F401
        from . import _cabi
        G, _, B, k_in = gathered.shape
        out_s = torch.empty((B, k), dtype=torch.float32, device=gathered.device)
        out_d = torch.empty((B, k), dtype=torch.int32, device=gathered.device)
        base = gathered.data_ptr()
        with torch.cuda.device(gathered.device):
            _cabi.check(_cabi.lib().gdr_merge_topk(base, base + B * k_in * 4, G, B, k_in, 2 * B * k_in, k,
                                                   out_s.data_ptr(), out_d.data_ptr(), _cabi.stream_ptr()))
        return out_s, out_d

    # ---- one batch ----------------------------------------------------------------------------------
    def localize(self, beams: torch.Tensor) -> torch.Tensor:
        """Global cluster ids [B, K] (-1 = absent) -> this rank's local ids, -1 where another rank owns it."""
        idx = beams.long()
        loc = self.g2l[idx.clamp(min=0)]
        return torch.where(idx >= 0, loc, torch.full_like(loc, -1)).to(torch.int32)

    def score_topk(self, q: torch.Tensor, beams: torch.Tensor, k: int, prob: Optional[torch.Tensor] = None,
                   alpha: Optional[float] = None, act: str = "none", flags: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """q [B, D] and beams [B, K] (GLOBAL cluster ids) are replicated on every rank.
        Returns the merged (scores [B, k], docids [B, k]) on every rank.  `flags` (GDR_FORCE_SIMT / GDR_FORCE_UMMA) go to the local call."""
        if flags:
            s, d = self._local_topk(q, self.localize(beams), k, prob, None if alpha is None else [alpha], act, flags)
        else:
            s, d = self._local_topk(q, self.localize(beams), k, prob, None if alpha is None else [alpha], act)
        s, d = s[0], d[0]
        if self.world_size == 1:
            return s, d
        packed = pack_candidates(s, d)
        # output is the dim-0 concatenation of every rank's [2, B, k] block (the layout both NCCL and gloo accept)
        gathered = torch.empty((self.world_size * 2,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)
        return self._merge(gathered.view((self.world_size,) + tuple(packed.shape)), k)


class ShardedPipeline:
    """The north-star multi-GPU path on a cluster-sharded corpus, pipelined: one process per GPU, `stores` = this rank's shard
    (`ClusterStore.shard`; a list = several indexes / the bench's L2-defeating copies).  The global batch (world x b_own queries,
    identical q / beams on every rank, beams with GLOBAL cluster ids) is submitted on every rank; each rank gets the top-k of
    the b_own queries it owns.

    exchange = "p2p": candidates never become lists — every rank's scoring epilogue stores its scores straight into the owner's
    score buffer over NVLink and raises a flag; the owner's top-k waits for the flags (include/gdr_b200.h gdr_store_p2p_*).
    Several batches are in flight (PipelinedRetriever's schedules, one exchange buffer per handle); whatever waits for a peer is
    built so that it cannot keep a scoring kernel off the SMs: one spinning warp in front of the stand-alone top-k, or the top-k
    groups inside the scoring CTA of the fused launch.  All ranks must submit the same sequence of batches.
    exchange = "nccl" is `ShardedRetriever` (local top-k, all-gather of (score, docid) lists, merge): see there."""

    def __init__(self, stores, rank: int, world: int, b_own: int, K: int, k: int, group=None, flags: int = 0,
                 schedule: str = "auto", depth: int = 5, fused_ctas: int = 0, fused_groups: int = 0, local_peers=None):
        """schedule: 'batches' = whole calls round-robin on `depth` streams (each handle has its own exchange buffer; the stand-alone
        top-k is preceded by a one-warp wait kernel, so nothing that spins can keep a scoring kernel off the SMs), 'fused' = scoring of
        batch i + top-k of batch i-1 in one launch (the top-k groups wait for the flags inside the scoring CTA), 'auto' = batches."""
        from .pipeline import PipelinedRetriever
        self.rank, self.world, self.b_own, self.K, self.k, self.group = rank, world, b_own, K, k, group
        stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
        B = world * b_own
        want_fused = schedule == "fused"
        self.pr = PipelinedRetriever(stores, schedule="fused" if want_fused else "batches", depth=depth, fused_ctas=fused_ctas, fused_groups=fused_groups)
        fused = want_fused and self.pr.fused_eligible(B, K, k, flags)
        if want_fused and not fused:
            raise ValueError("this batch shape is not eligible for the fused schedule")
        self.schedule = "fused" if fused else f"batches x{depth}"
        self.handles = [h for hs in (self.pr._handles_fused() if fused else self.pr._handles_batches()) for h in hs]
        mine = [h.p2p_init(world, rank, b_own, K) for h in self.handles]
        if local_peers is not None:             # every rank's ShardedPipeline lives in this process (tests): wired by connect_local
            self._ipc = None
        else:
            gathered: List[Optional[list]] = [None] * world
            dist.all_gather_object(gathered, mine, group=group)
            for j, h in enumerate(self.handles):
                h.p2p_attach([gathered[r][j] for r in range(world)])
        for h in self.handles:
            h.reserve(B, K, k, flags)

    @staticmethod
    def connect_local(pipelines: Sequence["ShardedPipeline"]) -> None:
        """Wire the exchange buffers of pipelines that live in ONE process (pipelines[r] = rank r)."""
        for p in pipelines:
            for j, h in enumerate(p.handles):
                h.p2p_attach_local([q.handles[j] for q in pipelines])

    def submit(self, q: torch.Tensor, beams: torch.Tensor, prob: Optional[torch.Tensor] = None, alpha: float = 1.0,
               act: Optional[str] = "none", flags: int = 0, out=None, which: int = 0):
        """q [world*b_own, D], beams [world*b_own, K] global cluster ids (prob likewise) -> Ticket with [b_own, k] outputs."""
        if out is None:
            dev = q.device
            out = (torch.empty((self.b_own, self.k), dtype=torch.float32, device=dev), torch.empty((self.b_own, self.k), dtype=torch.int32, device=dev))
        return self.pr.submit(q, beams, self.k, prob=prob, alpha=alpha, act=act, flags=flags, out=out, which=which)

    def flush(self) -> None:
        self.pr.flush()


class PeerAllGather:
    """All-gather of a batch's per-rank inputs (queries, beams) over NVLink by the copy engines (include/gdr_b200.h gdr_xchg_*,
    csrc/xchg.cu): every rank copies its parts into slot `slot` of every rank's buffer, raises an arrival flag, and one warp waits
    for the peers' flags — no collective kernel that would compete with the persistent scoring CTAs for SMs.
    parts: [(rows_per_rank, cols, dtype), ...]; `gathered(slot, p)` is the [world * rows, cols] tensor the consumers read."""

    def __init__(self, rank: int, world: int, parts, n_slots: int, device, group=None, local: bool = False):
        import ctypes
        from . import _cabi
        self.rank, self.world, self.parts, self.n_slots = rank, world, list(parts), n_slots
        self.part_bytes = [int(r) * int(c) * torch.empty((), dtype=dt).element_size() for r, c, dt in self.parts]
        arr = (ctypes.c_int64 * len(self.parts))(*self.part_bytes)
        total = int(_cabi.lib().gdr_xchg_bytes(world, n_slots, arr, len(self.parts)))
        if total <= 0:
            raise ValueError("every part must be a positive multiple of 16 bytes (at most 4 parts, 8 ranks)")
        self.buf = torch.zeros(total, dtype=torch.uint8, device=device)
        self._handle = ctypes.c_void_p()
        blob = (ctypes.c_ubyte * 72)()
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().gdr_xchg_create(ctypes.byref(self._handle), self.buf.data_ptr(), world, rank, n_slots, arr, len(self.parts), blob))
        self.own_bytes = sum(self.part_bytes)
        if not local and world > 1:
            gathered: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(gathered, bytes(blob), group=group)
            with torch.cuda.device(device):
                _cabi.check(_cabi.lib().gdr_xchg_attach(self._handle, ctypes.c_char_p(b"".join(gathered))))

    @staticmethod
    def connect_local(objs: Sequence["PeerAllGather"]) -> None:
        import ctypes
        from . import _cabi
        arr = (ctypes.c_void_p * len(objs))(*[o._handle.value for o in objs])
        for o in objs:
            _cabi.check(_cabi.lib().gdr_xchg_attach_local(o._handle, arr))

    def gathered(self, slot: int, part: int) -> torch.Tensor:
        from . import _cabi
        rows, cols, dt = self.parts[part]
        off = int(_cabi.lib().gdr_xchg_part_offset(self._handle, slot, part))
        return self.buf[off:off + self.world * self.part_bytes[part]].view(dt).view(self.world * rows, cols)

    def all_gather(self, slot: int, own: torch.Tensor, stream=None) -> None:
        """own: this rank's parts back to back (uint8 [own_bytes], device).  Enqueue-only; work enqueued afterwards on the same
        stream sees the complete slot."""
        from . import _cabi
        if own.dtype != torch.uint8 or own.numel() != self.own_bytes or not own.is_cuda or not own.is_contiguous():
            raise ValueError(f"own must be a contiguous CUDA uint8 tensor of {self.own_bytes} bytes")
        with torch.cuda.device(self.buf.device):
            _cabi.check(_cabi.lib().gdr_xchg_all_gather(self._handle, int(slot), own.data_ptr(), _cabi.stream_ptr(stream)))

    def close(self):
        from . import _cabi
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _cabi.lib().gdr_xchg_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
