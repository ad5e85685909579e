"""Training-time gather + contrastive loss on the cluster store.

Mirrors the reference's candidate gather in `T5FineTuner.forward` (GDR_model/main_models.py:983-996: `doc_embed[index]`
rows concatenated one at a time) and `encoder_cal` (main_models.py:1184-1221, called at :1275 with
`all_doc = cat([positive docs, in-cluster candidates])`).  One CUDA kernel reads the rows straight from the store and
returns the loss and d loss / d query; the document table gets no gradient, as in the reference (it is a fixed pickle,
only the query encoder trains).  No CPU path.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _cabi
from .store import ClusterStore


class _ContrastiveLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, store, pos_rows, cand_rows, cand_off, act, tau, intra_rate):
        dev = store.emb.device
        q = query.detach().to(dev, torch.float32).contiguous()
        B, S = q.shape[0], int(cand_rows.numel())
        per_query = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(q) if query.requires_grad else None
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gdr_contrastive_loss(
                store._handle, q.data_ptr(), pos_rows.data_ptr(), cand_rows.data_ptr() if S else None, cand_off.data_ptr(), B, S,
                _cabi.ACT[act], float(tau), float(intra_rate), per_query.data_ptr(), loss.data_ptr(),
                grad.data_ptr() if grad is not None else None, _cabi.stream_ptr()))
        ctx.grad_q = grad
        ctx.per_query = per_query
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.grad_q * grad_out if ctx.grad_q is not None else None
        return g, None, None, None, None, None, None, None


def encoder_cal(store: ClusterStore, query: torch.Tensor, positive_doc: Sequence[int], candidates_doc: Sequence[Sequence[int]],
                loss_func: str = "tanh", tau: float = 0.05, intra_rate: float = 1.0) -> torch.Tensor:
    """reference `encoder_cal(query, all_doc, valid_num)` (main_models.py:1184-1221) with the gather of main_models.py:983-996
    folded in: `positive_doc[i]` is query i's positive document index, `candidates_doc[i]` the document indices of its
    in-cluster candidates (`valid_num[i] = len(candidates_doc[i])`).  `query` [B, D] cuda fp32, may require grad.
    Returns the scalar loss (differentiable w.r.t. `query`).  KeyError for a document that is not in the store."""
    if not query.is_cuda:
        raise ValueError("the contrastive loss runs on the device: query must be a CUDA tensor (no CPU fallback)")
    if loss_func not in ("tanh", "sigmoid"):
        raise ValueError(f"loss_func {loss_func!r}: the reference defines only tanh and sigmoid (main_models.py:1187-1190)")
    B = query.shape[0]
    if len(positive_doc) != B or len(candidates_doc) != B:
        raise ValueError("one positive document and one candidate list per query")
    dev = store.emb.device
    valid_num = [len(c) for c in candidates_doc]
    flat = [d for c in candidates_doc for d in c]
    pos_rows = store.rows_of(positive_doc).to(dev)
    cand_rows = store.rows_of(flat).to(dev) if flat else torch.zeros(0, dtype=torch.int32, device=dev)
    cand_off = torch.from_numpy(np.concatenate([[0], np.cumsum(valid_num)]).astype(np.int32)).to(dev)
    return _ContrastiveLoss.apply(query, store, pos_rows, cand_rows, cand_off, loss_func, tau, intra_rate)
