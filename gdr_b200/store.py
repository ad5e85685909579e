"""Cluster-contiguous CSR document-embedding store resident in HBM.

Replaces the reference's two host-side index structures:
  * `self.doc_embed`  — int-indexable container of 1-D fp32 [D] tensors, unpickled at start-up
                        (GDR_model/main_models.py:806-814, loader :182-187) and copied to the GPU one
                        document at a time at query time (:1458-1462);
  * `self.id_mapping` — Dict[str cluster id -> List[int doc index]] (main_models.py:874-889).
Here the rows are permuted once so that a cluster is one contiguous slab of `emb[N, D]` (bf16 or
fp32) with `offsets[C+1]` and `docid[N]` (the reference's doc index of each row) beside it.
All compute goes through libgdr_b200.so (gdr_b200/_cabi.py); there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi


def csr_from_reference(doc_embed, id_mapping: Dict[str, List[int]]
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[str]]:
    """(doc_embed, id_mapping) -> host CSR (emb [N, D] fp32, offsets [C+1], docid [N], keys).
    `doc_embed[i]` -> [D] tensor (or `doc_embed` is an [N, D] tensor).  Clusters keep the dict's
    insertion order, documents inside a cluster keep their list order (= the reference's candidate
    order, main_models.py:1441-1443).  A document listed by several clusters gets one row per cluster."""
    keys = list(id_mapping.keys())
    rows: List[int] = []
    offsets = [0]
    for key in keys:
        rows.extend(int(i) for i in id_mapping[key])
        offsets.append(len(rows))
    idx = torch.tensor(rows, dtype=torch.int64)
    if isinstance(doc_embed, torch.Tensor):
        table = doc_embed.detach().to("cpu").reshape(len(doc_embed), -1)
    else:
        table = torch.stack([torch.as_tensor(doc_embed[i]).detach().to("cpu").reshape(-1) for i in range(len(doc_embed))])
    return table.float()[idx].contiguous(), torch.tensor(offsets, dtype=torch.int64), idx, keys


class ClusterStore:
    def __init__(self, emb: torch.Tensor, offsets: torch.Tensor, docid: torch.Tensor,
                 keys: Optional[Sequence[str]] = None):
        """emb [N, D] cuda bf16/fp32 (cluster-contiguous), offsets [C+1] / docid [N] integer tensors."""
        if not emb.is_cuda:
            raise ValueError("ClusterStore lives in HBM: emb must be a CUDA tensor (no CPU fallback)")
        if emb.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("emb must be float32 or bfloat16")
        self.emb = emb.contiguous()
        off_host = offsets.detach().to("cpu", torch.int64)
        if off_host.numel() < 2 or int(off_host[0]) != 0 or int(off_host[-1]) != emb.shape[0] or \
                bool((off_host[1:] < off_host[:-1]).any()):
            raise ValueError("offsets must be a non-decreasing [C+1] array spanning [0, N]")
        self.offsets_host = off_host.numpy()
        self.sizes_host = np.diff(self.offsets_host)
        self.offsets = off_host.to(torch.int32).to(emb.device)
        self.docid = docid.detach().to(torch.int32).to(emb.device).contiguous()
        if self.docid.numel() != emb.shape[0]:
            raise ValueError("docid must have one entry per row of emb")
        self.n_docs, self.dim = int(emb.shape[0]), int(emb.shape[1])
        self.n_clusters = int(off_host.numel() - 1)
        self.max_cluster = int(self.sizes_host.max())
        self.keys = list(keys) if keys is not None else None
        self.cluster_index: Dict[str, int] = {k: i for i, k in enumerate(self.keys)} if self.keys is not None else {}
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(emb.device):
            _cabi.check(_cabi.lib().gdr_store_create(
                ctypes.byref(self._handle), self.emb.data_ptr(), self.n_docs, self.dim,
                _cabi.DTYPE_BF16 if emb.dtype == torch.bfloat16 else _cabi.DTYPE_F32,
                self.offsets.data_ptr(), self.n_clusters, self.docid.data_ptr(), self.max_cluster))

    # ---- one shard of a corpus split across GPUs by cluster (include/gdr_b200.h gdr_store_create_shard) ---------------------------
    @classmethod
    def shard(cls, emb_local: torch.Tensor, offsets_global, docid_global: torch.Tensor, c_lo: int, c_hi: int,
              keys: Optional[Sequence[str]] = None) -> "ClusterStore":
        """emb_local [n_local, D] cuda: the rows of the GLOBAL clusters [c_lo, c_hi) only; offsets_global [C+1] and docid_global
        [N] describe the whole corpus and are replicated on every GPU.  Beams given to this store carry global cluster ids."""
        self = cls.__new__(cls)
        if not emb_local.is_cuda or emb_local.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("emb_local must be a float32 / bfloat16 CUDA tensor")
        off_host = torch.as_tensor(np.asarray(offsets_global)).to("cpu", torch.int64)
        row_lo, row_hi = int(off_host[c_lo]), int(off_host[c_hi])
        if row_hi - row_lo != emb_local.shape[0]:
            raise ValueError(f"emb_local has {emb_local.shape[0]} rows, clusters [{c_lo}, {c_hi}) hold {row_hi - row_lo}")
        self.emb = emb_local.contiguous()
        self.offsets_host = off_host.numpy()
        self.sizes_host = np.diff(self.offsets_host)
        self.offsets = off_host.to(torch.int32).to(emb_local.device)
        self.docid = docid_global.detach().to(torch.int32).to(emb_local.device).contiguous()
        self.n_docs, self.dim = int(off_host[-1]), int(emb_local.shape[1])
        if self.docid.numel() != self.n_docs:
            raise ValueError("docid_global must have one entry per document of the whole corpus")
        self.n_clusters = int(off_host.numel() - 1)
        self.max_cluster = int(self.sizes_host.max())
        self.keys = list(keys) if keys is not None else None
        self.cluster_index = {k: i for i, k in enumerate(self.keys)} if self.keys is not None else {}
        self.shard_range = (int(c_lo), int(c_hi), row_lo)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(emb_local.device):
            _cabi.check(_cabi.lib().gdr_store_create_shard(
                ctypes.byref(self._handle), self.emb.data_ptr(), int(emb_local.shape[0]), self.dim,
                _cabi.DTYPE_BF16 if emb_local.dtype == torch.bfloat16 else _cabi.DTYPE_F32, self.offsets.data_ptr(), self.n_clusters,
                self.docid.data_ptr(), self.n_docs, self.max_cluster, int(c_lo), int(c_hi), row_lo))
        return self

    def p2p_init(self, n_ranks: int, my_rank: int, b_own: int, K: int) -> bytes:
        """Allocate this handle's peer-to-peer exchange buffer; returns its 64-byte CUDA IPC handle (gdr_store_p2p_init)."""
        buf = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_p2p_init(self._handle, int(n_ranks), int(my_rank), int(b_own), int(K), buf))
        self.p2p = (int(n_ranks), int(my_rank), int(b_own))
        return bytes(buf)

    def p2p_attach(self, all_handles: Sequence[bytes]) -> None:
        """Map every rank's exchange buffer (IPC handles in rank order, one process per GPU)."""
        blob = b"".join(all_handles)
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_p2p_attach(self._handle, ctypes.c_char_p(blob)))

    def p2p_attach_local(self, peers: Sequence["ClusterStore"]) -> None:
        """The same for handles that live in this process (peers[r] = rank r's handle)."""
        arr = (ctypes.c_void_p * len(peers))(*[p._handle.value for p in peers])
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_p2p_attach_local(self._handle, arr))

    # ---- construction from the reference's objects ------------------------------------------
    @classmethod
    def from_reference(cls, doc_embed, id_mapping: Dict[str, List[int]], dtype=torch.bfloat16,
                       device="cuda") -> "ClusterStore":
        """Build from exactly the two objects the reference holds (main_models.py:806-814, 874-889)."""
        emb, offsets, docid, keys = csr_from_reference(doc_embed, id_mapping)
        return cls(emb.to(dtype).to(device), offsets, docid, keys)

    @classmethod
    def from_csr(cls, emb: torch.Tensor, offsets, docid, keys=None, dtype=None, device="cuda") -> "ClusterStore":
        emb = torch.as_tensor(emb)
        if dtype is not None:
            emb = emb.to(dtype)
        return cls(emb.to(device), torch.as_tensor(np.asarray(offsets)), torch.as_tensor(np.asarray(docid)), keys)

    # ---- lookups ---------------------------------------------------------------------------------
    def beams_from_ids(self, dec: Sequence[Sequence[str]]) -> torch.Tensor:
        """B x K cluster-id strings -> int32 [B, K] cluster indices.  KeyError for an unknown id,
        like `self.id_mapping[cluster_id]` at main_models.py:1442."""
        return torch.tensor([[self.cluster_index[c] for c in row] for row in dec], dtype=torch.int32)

    def rows_of(self, doc_indices) -> torch.Tensor:
        """Store row of each document index (the reference indexes `doc_embed[index]`, main_models.py:983-996); a
        document listed by several clusters has several rows holding the same embedding — the first is returned.
        KeyError for a document that is in no cluster."""
        if getattr(self, "_row_of_doc", None) is None:
            d = self.docid.cpu().numpy().astype(np.int64)
            table = np.full(int(d.max()) + 1 if d.size else 0, -1, dtype=np.int32)
            table[d[::-1]] = np.arange(d.size - 1, -1, -1, dtype=np.int32)      # reversed: the first occurrence wins
            self._row_of_doc = table
        idx = np.asarray(doc_indices, dtype=np.int64).reshape(-1)
        if idx.size and (idx.min() < 0 or idx.max() >= self._row_of_doc.size or (self._row_of_doc[idx] < 0).any()):
            raise KeyError("document index not in the store")
        return torch.from_numpy(self._row_of_doc[idx])

    def candidate_counts(self, beams_host: torch.Tensor) -> torch.Tensor:
        sizes = torch.from_numpy(np.append(self.sizes_host, 0))     # index -1 -> 0
        return sizes[beams_host.long()].sum(dim=1)

    # ---- the hot path ----------------------------------------------------------------------------
    def score_topk(self, q: torch.Tensor, beams: torch.Tensor, k: int, prob: Optional[torch.Tensor] = None,
                   alphas: Optional[Sequence[float]] = None, act: Optional[str] = "none", per_beam: bool = False,
                   flags: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, stream=None,
                   _unchecked_out: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Cluster-restricted scoring + top-k for one batch (gdr_score_topk in include/gdr_b200.h).
        q [B, D] fp32 cuda ([B*K, D] with per_beam), beams [B, K] int32 cuda (-1 = absent),
        prob [B, K] fp32 cuda or None, alphas list of floats or None.
        Returns (scores [n_alpha, B, k] fp32, docids [n_alpha, B, k] int32) — [B, k] when alphas is None."""
        dev = self.emb.device
        B, K = int(beams.shape[0]), int(beams.shape[1])
        if q.device != dev or beams.device != dev:
            raise ValueError("q and beams must live on the store's device")
        if q.dtype != torch.float32 or beams.dtype != torch.int32:
            raise ValueError("q must be float32 and beams int32")
        if q.shape != ((B * K if per_beam else B), self.dim):
            raise ValueError(f"q has shape {tuple(q.shape)}, expected {(B * K if per_beam else B, self.dim)}")
        q = q.contiguous()
        beams = beams.contiguous()
        if prob is not None:
            if prob.shape != (B, K) or prob.dtype != torch.float32 or prob.device != dev:
                raise ValueError("prob must be float32 [B, K] on the store's device")
            prob = prob.contiguous()
        n_alpha = 1 if alphas is None else len(alphas)
        alpha_arr = None if alphas is None else (ctypes.c_float * n_alpha)(*[float(a) for a in alphas])
        B_out = self.p2p[2] if getattr(self, "p2p", None) else B       # a p2p shard handle returns its own b_own queries of the global batch
        if out is None:
            out_s = torch.empty((n_alpha, B_out, k), dtype=torch.float32, device=dev)
            out_d = torch.empty((n_alpha, B_out, k), dtype=torch.int32, device=dev)
        else:
            out_s, out_d = out
            if not _unchecked_out:      # (the inversion-only call passes placeholders: no output is written)
                _check_out(out_s, out_d, n_alpha * B_out * int(k), dev)
        if per_beam:
            flags |= _cabi.Q_PER_BEAM
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gdr_score_topk(
                self._handle, q.data_ptr(), beams.data_ptr(), None if prob is None else prob.data_ptr(), alpha_arr,
                n_alpha, B, K, _cabi.ACT[act], int(k), flags, out_s.data_ptr(), out_d.data_ptr(), _cabi.stream_ptr(stream)))
        self._last_shape = (B_out, int(k))
        if alphas is None and out is None:
            return out_s[0], out_d[0]
        return out_s, out_d

    # ---- handle options / scratch ---------------------------------------------------------------------------------
    def set_option(self, name: str, value: int) -> "ClusterStore":
        """Launch options of this handle (gdr_store_set_option; names in _cabi.OPTIONS).  None changes results."""
        _cabi.check(_cabi.lib().gdr_store_set_option(self._handle, _cabi.OPTIONS[name], int(value)))
        return self

    def reserve(self, B: int, K: int, k: int, flags: int = 0, per_beam: bool = False, stream=None) -> "ClusterStore":
        """Pre-size the scratch for batches of up to this shape: no allocation on the query path afterwards."""
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_reserve(self._handle, int(B), int(K), int(k), flags | (_cabi.Q_PER_BEAM if per_beam else 0),
                                                     _cabi.stream_ptr(stream)))
        return self

    def clone_handle(self) -> "ClusterStore":
        """Another handle (= another scratch set) over the SAME device arrays: nothing is copied.  One per batch in flight."""
        if getattr(self, "shard_range", None) is not None:
            return ClusterStore.shard(self.emb, self.offsets_host, self.docid, self.shard_range[0], self.shard_range[1], self.keys)
        return ClusterStore(self.emb, torch.as_tensor(self.offsets_host), self.docid, self.keys)

    # ---- pipelined schedule: fused scoring + top-k, one batch behind (include/gdr_b200.h gdr_score_fused; pipeline.py drives it) ----
    def invert(self, q: torch.Tensor, beams: torch.Tensor, k: int, prob: Optional[torch.Tensor] = None, act: Optional[str] = "none",
               flags: int = 0, stream=None) -> None:
        """Inversion phase of a batch alone (gdr_score_topk with GDR_SKIP_SCORE | GDR_SKIP_TOPK): leaves the batch's work lists
        in this handle's scratch for `score_fused`.  q / beams / prob must stay alive until the batch's top-k has run."""
        if getattr(self, "_fused_dummy", None) is None or self._fused_dummy[0].numel() < k:
            self._fused_dummy = (torch.empty(max(k, 128), dtype=torch.float32, device=self.emb.device),
                                 torch.empty(max(k, 128), dtype=torch.int32, device=self.emb.device))
        self.score_topk(q, beams, k, prob=prob, act=act, flags=flags | _cabi.SKIP_SCORE | _cabi.SKIP_TOPK,
                        out=self._fused_dummy, stream=stream, _unchecked_out=True)

    def score_fused(self, prev: Optional["ClusterStore"], alpha: float = 1.0,
                    out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, stream=None):
        """ONE launch that scores the batch last given to `self.invert` and, in the same CTAs, selects the top-k of the batch
        that `prev` (another handle = another scratch set) scored in the previous launch.  Returns prev's (scores [B, k],
        docids [B, k]) or None when prev is None (first batch).  `ClusterStore.flush_fused(prev, ...)` ends a sequence."""
        return _score_fused(self, prev, alpha, out, stream)

    @staticmethod
    def flush_fused(prev: "ClusterStore", alpha: float = 1.0, out=None, stream=None):
        """Top-k of the last batch of a `score_fused` sequence (the stand-alone top-k kernel)."""
        return _score_fused(None, prev, alpha, out, stream)

    def last_stats(self) -> Dict[str, int]:
        arr = (ctypes.c_int64 * 4)()
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_last_stats(self._handle, arr, _cabi.stream_ptr()))
        return {"simt_items": arr[0], "umma_tiles": arr[1], "launches": arr[2], "clusters_touched": arr[3]}

    def centroids(self, stream=None) -> torch.Tensor:
        """Leaf-cluster centroids [C, D] fp32: the `embedding` tree_embedding_calculate stores on every leaf-cluster node
        (reference main_models.py:154-158)."""
        out = torch.empty((self.n_clusters, self.dim), dtype=torch.float32, device=self.emb.device)
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_cluster_centroids(self._handle, out.data_ptr(), _cabi.stream_ptr(stream)))
        return out

    def set_profiling(self, enable: bool) -> None:
        _cabi.check(_cabi.lib().gdr_store_set_profiling(self._handle, int(enable)))

    def last_phase_ms(self) -> Dict[str, float]:
        arr = (ctypes.c_float * 4)()
        with torch.cuda.device(self.emb.device):
            _cabi.check(_cabi.lib().gdr_store_last_phase_ms(self._handle, arr))
        return {"invert": arr[0], "score_umma": arr[1], "score_simt": arr[2], "topk": arr[3]}

    def bytes_touched(self, beams_host: torch.Tensor) -> int:
        """Algorithmic HBM bytes of the embeddings one batch touches: every touched cluster once
        (SURVEY.md §8d)."""
        u = torch.unique(beams_host[beams_host >= 0]).long().numpy()
        return int(self.sizes_host[u].sum()) * self.dim * self.emb.element_size()

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _cabi.lib().gdr_store_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_out(out_s: torch.Tensor, out_d: torch.Tensor, numel: int, dev) -> None:
    """The kernels write `numel` consecutive elements through raw pointers: refuse anything else before it reaches the C ABI."""
    if out_s.dtype != torch.float32 or out_d.dtype != torch.int32:
        raise ValueError("out must be (float32 scores, int32 docids)")
    if out_s.device != dev or out_d.device != dev:
        raise ValueError("out tensors must live on the store's device")
    if not out_s.is_contiguous() or not out_d.is_contiguous():
        raise ValueError("out tensors must be contiguous")
    if out_s.numel() != numel or out_d.numel() != numel:
        raise ValueError(f"out tensors must hold exactly n_alpha * B * k = {numel} elements each, got {out_s.numel()} / {out_d.numel()}")


def _score_fused(cur: Optional[ClusterStore], prev: Optional[ClusterStore], alpha, out, stream):
    out_s = out_d = None
    if prev is not None:
        B, k = prev._last_shape
        out_s, out_d = out if out is not None else (torch.empty((B, k), dtype=torch.float32, device=prev.emb.device),
                                                    torch.empty((B, k), dtype=torch.int32, device=prev.emb.device))
        _check_out(out_s, out_d, B * k, prev.emb.device)
    dev = (cur if cur is not None else prev).emb.device
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gdr_score_fused(
            None if cur is None else cur._handle, None if prev is None else prev._handle, float(alpha),
            None if out_s is None else out_s.data_ptr(), None if out_d is None else out_d.data_ptr(), _cabi.stream_ptr(stream)))
    return None if prev is None else (out_s, out_d)
