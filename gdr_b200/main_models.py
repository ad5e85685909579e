"""Host-side mirror of the pieces of the reference's GDR_model/main_models.py that sit on the
hot path — same names, argument meaning and error behaviour — delegating all device work to
libgdr_b200.so.  What stays in stock PyTorch and out of this package: the T5 / DPR forward, the
Lightning harness, data loading and metrics.

    Node, TreeBuilder            main_models.py:112-151   (picklable trie the beam search is constrained by)
    encode_single_newid          main_models.py:297-319
    decode_token                 main_models.py:322-346
    dec_2d                       main_utils.py:70-76
    EncoderModel, encode_query   main_models.py:62-109    (query embedding = T5 encoder token-0 state)
    FineStage                    main_models.py:1434-1637 (the fine-grained stage of validation_step_i)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .store import ClusterStore


class Node(object):
    """Trie node (reference main_models.py:112-127): `children` maps token -> Node,
    `embedding_index` lists the docs of a leaf cluster."""

    __slots__ = ("token_id", "children", "embedding_index", "embedding", "all_leaf_num")

    def __init__(self, token_id) -> None:
        self.token_id = token_id
        self.children: Dict[int, "Node"] = {}
        self.embedding_index: List[int] = []
        self.embedding = None
        self.all_leaf_num = 0

    def __repr__(self):
        return "<tree node representation>"

    def __getstate__(self):
        return {s: getattr(self, s) for s in self.__slots__}

    def __setstate__(self, state):
        for s in self.__slots__:
            setattr(self, s, state.get(s))


class TreeBuilder(object):
    """reference main_models.py:130-151.  `add(seq, embedding_index)`: seq is a token path without the
    leading pad, with trailing EOS (=1) and possibly pads (=0).  A pad ends the walk; the doc index
    is recorded on the parent of the last node created/visited (the leaf-cluster node, parent of EOS)."""

    def __init__(self) -> None:
        self.root = Node(0)

    def build(self) -> Node:
        return self.root

    def add(self, seq: Sequence[int], embedding_index: int) -> None:
        node, parent = self.root, None
        for tok in seq:
            tok = int(tok)
            if tok == 0:
                return
            child = node.children.get(tok)
            if child is None:
                child = Node(tok)
                node.children[tok] = child
            parent, node = node, child
        parent.embedding_index.append(embedding_index)


def encode_single_newid(args, seq: str) -> List[int]:
    """reference main_models.py:297-319: "3-17-22" -> [5, 49, 84, 1] (token = i*kary + digit + 2 when
    args.position, EOS = 1 appended).  With args.kary == 0 the id is a string of decimal digits and
    the positional stride is 10 (:312-318)."""
    position = bool(getattr(args, "position", 0))
    if args.kary:
        digits = [int(c) for c in seq.split("-")]
        stride = args.kary
    else:
        digits = [int(c) for c in seq]
        stride = 10
    return [(i * stride if position else 0) + d + 2 for i, d in enumerate(digits)] + [1]


def decode_token(args, seqs) -> List[str]:
    """reference main_models.py:322-346.  Each row: drop the leading pad and everything from the first
    EOS (=1) on, subtract `pos*output_vocab_size + 2` (positional) or 2, join with '-' (kary) or ''.
    A row without EOS is decoded whole, leading token included (the reference's bare except, :331-335)."""
    position = bool(getattr(args, "position", 0))
    sep = "-" if args.kary else ""
    out = []
    for seq in seqs:
        seq = np.asarray(seq)
        hits = np.nonzero(seq == 1)[0]
        if hits.size:
            seq = seq[1:hits[0]]
        offset = (np.arange(len(seq)) * args.output_vocab_size + 2) if position else 2
        out.append(sep.join(str(c) for c in (seq - offset)))
    return out


def dec_2d(dec: list, size: int) -> List[list]:
    """reference main_utils.py:70-76."""
    return [dec[i:i + size] for i in range(0, len(dec), size)]


def encode_query(qry_hidden: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """reference main_models.py:102-109 with `self.output = None`: the query embedding is the T5
    encoder's token-0 hidden state."""
    if qry_hidden is None:
        return None
    return qry_hidden[:, 0]


class EncoderModel(torch.nn.Module):
    """reference main_models.py:62-109.  The reference wraps a DPR context encoder (stock PyTorch, loaded from local
    checkpoints); here the wrapped module is whatever the caller supplies.  `forward(query_enc=h)` is the part on the
    hot path: the query embedding is `h[:, 0]` (or the optional pooler `self.output`)."""

    def __init__(self, model: Optional[torch.nn.Module] = None, output: Optional[torch.nn.Module] = None, args=None):
        super().__init__()
        self.model, self.output, self.args = model, output, args

    def forward(self, passage=None, query_enc=None):
        if passage is not None:                                      # main_models.py:80-86 — stock PyTorch forward
            passage = {k: v.view(-1, v.size(-1)) for k, v in passage.items()}
            return self.model(**passage, return_dict=True).pooler_output
        if query_enc is not None:
            return self.encode_query(query_enc)

    def encode_query(self, qry_hidden):                              # main_models.py:102-109
        if qry_hidden is None:
            return None
        return self.output(q=qry_hidden) if self.output is not None else qry_hidden[:, 0]


class FineStage:
    """The fine-grained stage of `T5FineTuner.validation_step_i` (reference main_models.py:1434-1637):

        top_k( f(D_cand · q_b) + alpha * softmax(beam_scores_b)[cluster_of(candidate)] )   per query b, per alpha

    Construct once from the two objects the reference holds (`doc_embed`, `id_mapping`) and the
    flags it reads: `args.num_return_sequences` (beam width AND k, main_models.py:1364-1366,1625),
    `args.score_rate` (main.py:389), `args.loss_func` (main.py:393).  `k` may be given separately:
    the BASELINE configs use beam 10/20/100 with top-100/1000.
    """

    def __init__(self, args, doc_embed=None, id_mapping: Optional[Dict[str, List[int]]] = None,
                 store: Optional[ClusterStore] = None, dtype=torch.bfloat16, device="cuda", k: Optional[int] = None):
        self.args = args
        self.store = store if store is not None else ClusterStore.from_reference(doc_embed, id_mapping, dtype, device)
        self.k = int(k if k is not None else args.num_return_sequences)
        self.score_rate = list(getattr(args, "score_rate", [0]))
        self.loss_func = getattr(args, "loss_func", "tanh")
        if self.loss_func not in ("tanh", "sigmoid", "none"):
            raise ValueError(f"loss_func {self.loss_func!r}: the reference defines only tanh and sigmoid (main_models.py:1578-1581)")

    def retrieve(self, dec: Sequence[Sequence[str]], scores: Sequence[float], query_embeds: torch.Tensor,
                 per_beam: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """dec: B lists of K cluster-id strings (output of decode_token + dec_2d, :1398,1422);
        scores: the B*K beam scores returned by generate (:1380-1397); query_embeds [B, D]
        (`self.encoder(query_enc=...)`, :1466) or [B*K, D] with per_beam (:1467-1571).
        Returns (values [B, n_rate, k] fp32, docids [B, n_rate, k] int64) on the store's device.
        KeyError for an unknown cluster id (:1442); RuntimeError if a query has fewer than k candidates (:1625)."""
        st = self.store
        dev = st.emb.device
        B = len(dec)
        beams_host = st.beams_from_ids(dec)                                   # KeyError like :1442
        counts = st.candidate_counts(beams_host)
        if B and int(counts.min()) < self.k:
            # torch.topk's error at main_models.py:1625
            raise RuntimeError(f"selected index k out of range: a query has {int(counts.min())} candidates, k = {self.k}")
        # :1598-1601 — softmax over each query's K beam scores, on the host in fp32 exactly as the reference
        prob = torch.softmax(torch.tensor(list(scores), dtype=torch.float32).view(B, -1), dim=-1)
        q = query_embeds.to(dev, torch.float32)
        vals, ids = st.score_topk(q, beams_host.to(dev), self.k, prob=prob.to(dev), alphas=self.score_rate,
                                  act=self.loss_func, per_beam=per_beam)
        return vals.permute(1, 0, 2), ids.permute(1, 0, 2).long()

    def __call__(self, dec, scores, query_embeds, texts: Optional[Sequence[str]] = None,
                 gt_answers: Optional[Sequence[str]] = None, per_beam: bool = False):
        """Returns `inf_index_batch_all` in the reference's format (:1614-1637):
        out[b][rate_idx] = [[text, "docid,docid,...", ground_truth]]."""
        _, ids = self.retrieve(dec, scores, query_embeds, per_beam)
        ids = ids.cpu().tolist()
        out = []
        for b, per_rate in enumerate(ids):
            text = texts[b] if texts is not None else ""
            gt = gt_answers[b] if gt_answers is not None else ""
            out.append([[[text, ",".join(str(d) for d in row), gt]] for row in per_rate])
        return out
