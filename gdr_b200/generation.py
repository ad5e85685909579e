"""Docid logit masks for the constrained beam search, on the device.

    DeviceTrie / tree_mask_      the live prefix-tree mask of the reference's beam search,
                                 GDR_model/transformers/generation_utils_previous.py:712-730
                                 (a Python loop over B*K rows with .tolist() syncs, dict walks and
                                 one index_put per row) as ONE kernel over a CSR trie in HBM.
    position_mask_ /             the positional docid mask of GDR_model/transformers/modeling_t5.py:
    select_valid_embedding       1546-1571 (eval, applied :1646) and 1279-1301 (training buffer, :1644).

The beam search itself (T5 forward, log_softmax, topk(2K), hypothesis bookkeeping) stays in stock
PyTorch; `TreeMask` is the hook a maintainer drops where the reference's block sits.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _cabi


def flatten_trie(root) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """`Node` trie (anything with a `.children: Dict[int, node]`) -> CSR arrays
    (first_child [n_nodes+1], child_tok [n_edges], child_node [n_edges]); node 0 is the root,
    nodes are numbered breadth-first, each node's edges are sorted by token."""
    first_child, child_tok, child_node = [0], [], []
    queue = [root]
    head = 0
    while head < len(queue):
        node = queue[head]
        head += 1
        for tok in sorted(node.children):
            child_tok.append(int(tok))
            child_node.append(len(queue))
            queue.append(node.children[tok])
        first_child.append(len(child_tok))
    return (np.asarray(first_child, dtype=np.int32), np.asarray(child_tok, dtype=np.int32),
            np.asarray(child_node, dtype=np.int32))


def child_insertion_order(root) -> np.ndarray:
    """For the numbering of `flatten_trie`: every node's edge indices in the order its `children` dict was filled (the
    order the reference iterates `cur.children.keys()` in, main_models.py:164,241), as one [n_edges] array."""
    order, queue, head, n_edges = [], [root], 0, 0
    while head < len(queue):
        node = queue[head]
        head += 1
        toks = sorted(node.children)
        pos = {t: i for i, t in enumerate(toks)}
        order.extend(n_edges + pos[t] for t in node.children)          # dict order = insertion order
        for t in toks:
            queue.append(node.children[t])
        n_edges += len(toks)
    return np.asarray(order, dtype=np.int32)


class DeviceTrie:
    """Device-resident CSR form of the reference's `Node` trie (main_models.py:112-151)."""

    def __init__(self, first_child: np.ndarray, child_tok: np.ndarray, child_node: np.ndarray, device="cuda",
                 child_order: Optional[np.ndarray] = None):
        self.first_child = np.ascontiguousarray(first_child, dtype=np.int32)
        self.child_tok = np.ascontiguousarray(child_tok, dtype=np.int32)
        self.child_node = np.ascontiguousarray(child_node, dtype=np.int32)
        self.n_nodes = int(self.first_child.size - 1)
        self.n_edges = int(self.child_tok.size)
        self.device = torch.device(device)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().gdr_trie_create(
                ctypes.byref(self._handle), self.first_child.ctypes.data, self.child_tok.ctypes.data,
                self.child_node.ctypes.data, self.n_nodes, self.n_edges))
            if child_order is not None:
                self.child_order = np.ascontiguousarray(child_order, dtype=np.int32)
                _cabi.check(_cabi.lib().gdr_trie_set_child_order(self._handle, self.first_child.ctypes.data, self.child_order.ctypes.data))

    @classmethod
    def from_root(cls, root, device="cuda") -> "DeviceTrie":
        return cls(*flatten_trie(root), device=device, child_order=child_insertion_order(root))

    def find(self, tokens) -> int:
        """Node reached from the root along `tokens` (host walk over the CSR arrays), -1 if the path leaves the tree."""
        cur = 0
        for t in tokens:
            lo, hi = int(self.first_child[cur]), int(self.first_child[cur + 1])
            j = lo + int(np.searchsorted(self.child_tok[lo:hi], int(t)))
            if j >= hi or int(self.child_tok[j]) != int(t):
                return -1
            cur = int(self.child_node[j])
        return cur

    def mask_(self, scores: torch.Tensor, input_ids: torch.Tensor, eos_token_id: int = 1, strict: bool = False,
              stream=None) -> torch.Tensor:
        """In place on `scores` [R, V] fp32 (cuda, last dim contiguous) given `input_ids` [R, cur_len]
        int64: what generation_utils_previous.py:714-729 does to `scores`.  Returns `scores`."""
        if not (scores.is_cuda and input_ids.is_cuda):
            raise ValueError("tree mask runs on the device: scores and input_ids must be CUDA tensors (no CPU fallback)")
        if scores.dtype != torch.float32 or input_ids.dtype != torch.int64:
            raise ValueError("scores must be float32 and input_ids int64")
        if scores.dim() != 2 or input_ids.dim() != 2 or scores.shape[0] != input_ids.shape[0]:
            raise ValueError("scores [R, V] and input_ids [R, cur_len] must agree on R")
        if scores.stride(1) != 1 or input_ids.stride(1) != 1:
            raise ValueError("last dimension must be contiguous")
        R, V = scores.shape
        with torch.cuda.device(scores.device):
            _cabi.check(_cabi.lib().gdr_tree_mask(
                self._handle, input_ids.data_ptr(), input_ids.stride(0) if R > 1 else input_ids.shape[1], R,
                input_ids.shape[1], scores.data_ptr(), scores.stride(0) if R > 1 else V, V, int(eos_token_id),
                int(strict), _cabi.stream_ptr(stream)))
        return scores

    def beam_step(self, next_token_logits: torch.Tensor, input_ids: torch.Tensor, beam_scores: torch.Tensor,
                  num_beams: int, eos_token_id: int = 1, stream=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fused log_softmax -> tree mask -> + beam score -> view [B, K*V] -> topk(2K) of the reference's beam search
        (generation_utils_previous.py:694, 714-729, 757-771) with one read of the logits.
        next_token_logits [B*K, V] fp32, input_ids [B*K, cur_len] int64, beam_scores [B*K] fp32 (all CUDA).
        Returns (next_scores [B, 2K] fp32, next_tokens [B, 2K] int64 = beam * V + token).  When fewer than 2K entries survive the
        mask (routine at the leaf level, where every beam only allows EOS) the tail has score -inf and — like the in-row indices
        `torch.topk` returns for the reference's -inf entries — a VALID flat index (beam 0, token `eos_token_id`), so the
        Hugging Face bookkeeping that follows (`beam_id = idx // V`, `token_id = idx % V`, generation_utils_previous.py:783-834)
        never indexes outside its own batch element.
        The fused step stands in for `postprocess_next_token_scores` (:696-708) only when its options are at their no-op defaults
        (repetition_penalty 1.0, no_repeat_ngram_size 0, bad_words_ids None, min_length 0), which is how the reference calls
        `generate` (main_models.py:1380-1397); pass non-default options through the unfused path (`TreeMask.__call__`)."""
        if not (next_token_logits.is_cuda and input_ids.is_cuda and beam_scores.is_cuda):
            raise ValueError("beam_step runs on the device (no CPU fallback)")
        if next_token_logits.dtype != torch.float32 or input_ids.dtype != torch.int64 or next_token_logits.stride(1) != 1:
            raise ValueError("next_token_logits must be float32 with a contiguous last dimension, input_ids int64")
        R, V = next_token_logits.shape
        if R % num_beams or input_ids.shape[0] != R or beam_scores.numel() != R:
            raise ValueError("row counts must equal batch_size * num_beams")
        B = R // num_beams
        bs = beam_scores.to(torch.float32).contiguous().view(-1)
        ids = input_ids if input_ids.stride(1) == 1 else input_ids.contiguous()
        out_s = torch.empty((B, 2 * num_beams), dtype=torch.float32, device=next_token_logits.device)
        out_t = torch.empty((B, 2 * num_beams), dtype=torch.int32, device=next_token_logits.device)
        with torch.cuda.device(next_token_logits.device):
            _cabi.check(_cabi.lib().gdr_beam_step(
                self._handle, next_token_logits.data_ptr(), next_token_logits.stride(0) if R > 1 else V, ids.data_ptr(),
                ids.stride(0) if R > 1 else ids.shape[1], bs.data_ptr(), B, num_beams, ids.shape[1], V, int(eos_token_id),
                out_s.data_ptr(), out_t.data_ptr(), _cabi.stream_ptr(stream)))
        out_t = out_t.long()
        return out_s, torch.where(out_t < 0, torch.full_like(out_t, int(eos_token_id)), out_t)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _cabi.lib().gdr_trie_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TreeMask:
    """Callable hook for the beam-search loop: `scores = tree_mask(input_ids, scores)` replaces the
    `if decode_tree:` block of generation_utils_previous.py:714-729 (`decode_tree` = the `Node` root
    passed to `generate(..., decode_tree=self.root)`, main_models.py:1393)."""

    def __init__(self, decode_tree, eos_token_id: int = 1, strict: bool = False, device="cuda"):
        self.trie = decode_tree if isinstance(decode_tree, DeviceTrie) else DeviceTrie.from_root(decode_tree, device)
        self.eos_token_id = eos_token_id
        self.strict = strict

    def __call__(self, input_ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        return self.trie.mask_(scores, input_ids, self.eos_token_id, self.strict)

    def beam_step(self, next_token_logits, input_ids, beam_scores, num_beams):
        """Fused replacement of log_softmax + mask + beam-score add + topk(2K): see DeviceTrie.beam_step."""
        return self.trie.beam_step(next_token_logits, input_ids, beam_scores, num_beams, self.eos_token_id)


def position_mask_(logits: torch.Tensor, output_vocab_size: int, last_eos_only: bool = False, stream=None) -> torch.Tensor:
    """In place on logits [bz, seq_length, vocab] fp32: position t keeps tokens
    {t*V_out+2 .. t*V_out+V_out+1} and 1, everything else gets -1e9 added (modeling_t5.py:1566-1569).
    `last_eos_only` = the training-time `logit_mask` variant (modeling_t5.py:1296)."""
    if not logits.is_cuda:
        raise ValueError("position mask runs on the device (no CPU fallback)")
    if logits.dtype != torch.float32 or logits.dim() != 3 or not logits.is_contiguous():
        raise ValueError("logits must be a contiguous float32 [bz, seq_length, vocab] tensor")
    bz, sl, V = logits.shape
    with torch.cuda.device(logits.device):
        _cabi.check(_cabi.lib().gdr_position_mask(logits.data_ptr(), bz, sl, V, int(output_vocab_size),
                                                  int(last_eos_only), _cabi.stream_ptr(stream)))
    return logits


def select_valid_embedding(sequence: torch.Tensor, output_vocab_size: int) -> torch.Tensor:
    """Functional form with the reference's signature (nested function at modeling_t5.py:1546-1571,
    `self.output_vocab_size` passed explicitly): returns `sequence + mask` as a new tensor."""
    return position_mask_(sequence.contiguous().clone(), output_vocab_size, last_eos_only=False)


def build_logit_mask(max_output_length: int, decode_vocab_size: int, output_vocab_size: int, device="cuda") -> torch.Tensor:
    """The `self.logit_mask` buffer of modeling_t5.py:1279-1301: [1, max_output_length, decode_vocab_size],
    0 at valid tokens, -1e9 elsewhere, last position EOS-only."""
    z = torch.zeros(1, max_output_length, decode_vocab_size, dtype=torch.float32, device=device)
    return position_mask_(z, output_vocab_size, last_eos_only=True)
