"""Drop-in for the reference's GDR_model/dense.py API surface that touches the hot path.

    DenseModel.compute_similarity(q_reps, p_reps)   dense.py:53-54 (base encoder.py:128-129):
                                                    `torch.matmul(q_reps, p_reps.transpose(0, 1))`
    DensePooler                                     dense.py:10-27 (CLS pooling + linear projection; stock
                                                    PyTorch, kept only so code that builds one keeps working)

`compute_similarity` runs on libgdr_b200.so (gdr_similarity).  The cluster-restricted form of the
same product — the one the fine stage actually needs — is `ClusterStore.score_topk`.  Encoding
(`encode_passage` / `encode_query`, dense.py:31-51) is a HuggingFace forward and stays in stock PyTorch.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import _cabi


def compute_similarity(q_reps: Tensor, p_reps: Tensor, stream=None) -> Tensor:
    """[Q, D] x [P, D]^T -> [Q, P] fp32.  q_reps fp32; p_reps fp32 or bf16; both CUDA."""
    if not (q_reps.is_cuda and p_reps.is_cuda):
        raise ValueError("compute_similarity runs on the device: inputs must be CUDA tensors (no CPU fallback)")
    if q_reps.dim() != 2 or p_reps.dim() != 2 or q_reps.shape[1] != p_reps.shape[1]:
        raise ValueError("q_reps [Q, D] and p_reps [P, D] must agree on D")
    q = q_reps.to(torch.float32).contiguous()
    p = p_reps.contiguous()
    if p.dtype not in (torch.float32, torch.bfloat16):
        p = p.to(torch.float32)
    out = torch.empty((q.shape[0], p.shape[0]), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _cabi.check(_cabi.lib().gdr_similarity(
            q.data_ptr(), q.shape[0], p.data_ptr(), p.shape[0], q.shape[1],
            _cabi.DTYPE_BF16 if p.dtype == torch.bfloat16 else _cabi.DTYPE_F32, out.data_ptr(), _cabi.stream_ptr(stream)))
    return out


class DensePooler(nn.Module):
    """dense.py:10-27: separate linear heads for query and passage on the CLS state."""

    def __init__(self, input_dim: int = 768, output_dim: int = 768, normalize=False):
        super().__init__()
        self.normalize = normalize
        self.linear_q = nn.Linear(input_dim, output_dim)
        self.linear_p = nn.Linear(input_dim, output_dim)
        self._config = {"input_dim": input_dim, "output_dim": output_dim, "normalize": normalize}

    def forward(self, q: Tensor = None, p: Tensor = None, **kwargs):
        if q is not None:
            rep = self.linear_q(q[:, 0])
        elif p is not None:
            rep = self.linear_p(p[:, 0])
        else:
            raise ValueError
        return nn.functional.normalize(rep, dim=-1) if self.normalize else rep


class DenseModel(nn.Module):
    """Only the part of dense.py's DenseModel that is on the hot path.  `lm_q` / `lm_p` / `pooler` are
    whatever stock-PyTorch encoders the caller has; they are not touched here."""

    def __init__(self, lm_q: nn.Module = None, lm_p: nn.Module = None, pooler: nn.Module = None):
        super().__init__()
        self.lm_q, self.lm_p, self.pooler = lm_q, lm_p, pooler

    def encode_passage(self, psg):          # dense.py:31-40 — stock PyTorch forward
        if psg is None:
            return None
        hidden = self.lm_p(**psg, return_dict=True).last_hidden_state
        return self.pooler(p=hidden) if self.pooler is not None else hidden[:, 0]

    def encode_query(self, qry):            # dense.py:42-51 — stock PyTorch forward
        if qry is None:
            return None
        hidden = self.lm_q(**qry, return_dict=True).last_hidden_state
        return self.pooler(q=hidden) if self.pooler is not None else hidden[:, 0]

    def compute_similarity(self, q_reps, p_reps):   # dense.py:53-54
        return compute_similarity(q_reps, p_reps)
