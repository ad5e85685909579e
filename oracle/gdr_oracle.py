"""CPU oracle for the GDR fine-grained stage and the docid logit masks.

*** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the CPU baseline.  The product
package `gdr_b200` never imports anything under `oracle/`; it fails loudly when its CUDA
library is missing instead of falling back to this code.

This is a plain torch-CPU/numpy restatement of the reference's algorithm (ypw0102/GDR,
pure Python + torch; the arithmetic itself lives in third-party `torch`, reference pin
pytorch=1.10.0 at environment.yml:82, this image 2.11.0).  Every function cites the
reference file:line it follows.  Paths are relative to /root/reference/.

PINNING.  The reference ships no tests, fixtures or golden vectors for this path
(SURVEY.md §4/§8c): parity is "unpinned" by the reference's own test-suite.  It is pinned
here instead against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE DEV CONTAINER:
`oracle/make_golden.py` imports the unmodified reference (via `oracle/ref_shims.py`),
executes dense.py's `DenseModel.compute_similarity`, main_models.py's
`T5FineTuner.validation_step_i`, `TreeBuilder`, the codecs, the live tree-mask block of
generation_utils_previous.py:714-729 and modeling_t5.py's `select_valid_embedding`, and
commits their inputs/outputs under tests/golden/.  tests/test_oracle_golden.py checks this
file against those fixtures (and against the live reference when /root/reference exists).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a1. similarity  — GDR_model/dense.py:53-54 (base: GDR_model/encoder.py:128-129)
# --------------------------------------------------------------------------------------


def compute_similarity(q_reps: torch.Tensor, p_reps: torch.Tensor) -> torch.Tensor:
    """`torch.matmul(q_reps, p_reps.transpose(0, 1))` -> [Q, P] (dense.py:54)."""
    return torch.matmul(q_reps, p_reps.transpose(0, 1))


# --------------------------------------------------------------------------------------
# a10. prefix tree + codecs — GDR_model/main_models.py:112-151, 297-346
# --------------------------------------------------------------------------------------


class Node:
    """main_models.py:112-127."""

    def __init__(self, token_id):
        self.token_id = token_id
        self.children: Dict[int, "Node"] = {}
        self.embedding_index: List[int] = []
        self.embedding = None
        self.all_leaf_num = 0


class TreeBuilder:
    """main_models.py:130-151.  The doc index is appended to the PARENT of the last
    visited node (i.e. the parent of the EOS node = the leaf-cluster node); a pad token
    (0) stops the walk without recording anything (main_models.py:144-145)."""

    def __init__(self):
        self.root = Node(0)

    def build(self) -> Node:
        return self.root

    def add(self, seq: Sequence[int], embedding_index: int) -> None:
        cur, cur_top = self.root, None
        for tok in seq:
            if tok == 0:
                return
            nxt = cur.children.get(tok)
            if nxt is None:
                nxt = cur.children[tok] = Node(tok)
            cur_top, cur = cur, nxt
        cur_top.embedding_index.append(embedding_index)


def encode_single_newid(seq: str, kary: int = 30, position: bool = True) -> List[int]:
    """main_models.py:297-319.  "3-17-22" -> [5, 49, 84, 1]: token = i*kary + c + 2, EOS = 1."""
    out = []
    if kary:
        for i, c in enumerate(seq.split("-")):
            out.append(i * kary + int(c) + 2 if position else int(c) + 2)
    else:  # main_models.py:312-318 (decimal digits, hard-coded vocab 10)
        for i, c in enumerate(seq):
            out.append(i * 10 + int(c) + 2 if position else int(c) + 2)
    return out + [1]


def decode_token(seqs: np.ndarray, output_vocab_size: int = 30, position: bool = True,
                 kary: int = 30) -> List[str]:
    """main_models.py:322-346.  Drops the leading pad, cuts at the first EOS (=1); when no
    EOS exists the reference swallows the ValueError and decodes the WHOLE row including the
    leading token (main_models.py:331-335)."""
    result = []
    for seq in seqs:
        seq = np.asarray(seq)
        lst = seq.tolist()
        if 1 in lst:
            seq = seq[1:lst.index(1)]
        offset = np.arange(len(seq)) * output_vocab_size + 2 if position else 2
        res = seq - offset
        result.append(("-" if kary else "").join(str(c) for c in res))
    return result


def dec_2d(dec: list, size: int) -> List[list]:
    """main_utils.py:70-76."""
    return [dec[i:i + size] for i in range(0, len(dec), size)]


# --------------------------------------------------------------------------------------
# a4-a9. fine stage — GDR_model/main_models.py:1434-1637
# --------------------------------------------------------------------------------------

_ACT = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "none": lambda x: x}


def gather_candidates(id_mapping: Dict[str, List[int]], dec: List[List[str]]
                      ) -> Tuple[List[List[int]], List[List[int]]]:
    """main_models.py:1435-1443.  Returns per query: candidate doc indices (beam-major,
    in id_mapping order) and the K segment lengths.  KeyError for an unknown cluster id."""
    cand, seg = [], []
    for clusters in dec:
        ids, lens = [], []
        for cid in clusters:
            docs = id_mapping[cid]
            ids.extend(docs)
            lens.append(len(docs))
        cand.append(ids)
        seg.append(lens)
    return cand, seg


def fine_stage(doc_embed, id_mapping: Dict[str, List[int]], dec: List[List[str]],
               beam_scores: Sequence[float], query_embeds: torch.Tensor,
               score_rate: Sequence[float], loss_func: str, k: int
               ) -> List[List[Tuple[torch.Tensor, torch.Tensor, List[int]]]]:
    """Restatement of main_models.py:1434-1637 for `use_query_embed_encoder`:

        topk_k( f(D_cand · q_b) + alpha * softmax(beam_scores_b)[cluster_of(j)] )

    gather :1441-1462, score :1577-1582 (row b sliced to its own candidates :1606-1611),
    softmax over the K beam scores :1598-1601, per-alpha segment bias :1619-1624,
    topk(largest, sorted) :1625, index -> doc index :1628-1631.
    The reference evaluates every query against every query's candidates and then slices;
    only the sliced part is observable, so only that is computed here.
    Returns out[b][rate_idx] = (values[k], indices[k] into the query's candidate list, docids[k]).
    RuntimeError (from torch.topk) if a query has fewer than k candidates, as the reference.
    """
    f = _ACT[loss_func]
    B = len(dec)
    cand, seg = gather_candidates(id_mapping, dec)
    prob = torch.softmax(torch.tensor(list(beam_scores), dtype=torch.float32).view(B, -1), dim=-1)
    out = []
    for b in range(B):
        D = torch.stack([doc_embed[i].float() for i in cand[b]]) if cand[b] else torch.zeros(0, query_embeds.shape[1])
        # main_models.py:1582 — torch.mul(q[:,None,:], D[None,:,:]).sum(-1); fp32 throughout
        sim = f(torch.mul(query_embeds[b].float().view(1, 1, -1), D.unsqueeze(0)).sum(-1))[0]
        per_rate = []
        for alpha in score_rate:
            s = sim.clone()
            lo = 0
            for i, n in enumerate(seg[b]):
                s[lo:lo + n] = s[lo:lo + n] + alpha * prob[b][i]
                lo += n
            vals, idx = s.topk(k, dim=0, largest=True, sorted=True)
            per_rate.append((vals, idx, [cand[b][j] for j in idx.tolist()]))
        out.append(per_rate)
    return out


def dense_topk(q: torch.Tensor, emb: torch.Tensor, offsets: np.ndarray, docid: np.ndarray,
               beams: np.ndarray, k: int, bias: Optional[torch.Tensor] = None,
               act: str = "none") -> Tuple[torch.Tensor, torch.Tensor]:
    """The `dense.py` path the north star pins: per query, gather the rows of
    its K beam clusters (CSR slices, beam order), `compute_similarity(q[b:b+1], rows)`
    (dense.py:53-54), optional activation/bias (main_models.py:1582,1623-1624), then
    `Tensor.topk(k, largest=True, sorted=True)` (main_models.py:1625).
    `emb` is the cluster-contiguous [N, D] table, `offsets` [C+1], `docid` [N] global ids,
    `beams` [B, K] cluster indices (-1 = absent).  Queries with fewer than k candidates are
    padded with (-inf, -1) — the reference raises there (main_models.py:1625); the padded form
    is what the C ABI defines.
    Returns (scores [B, k] fp32, docids [B, k] int64)."""
    B = q.shape[0]
    f = _ACT[act]
    out_s = torch.full((B, k), float("-inf"), dtype=torch.float32)
    out_i = torch.full((B, k), -1, dtype=torch.int64)
    emb32 = emb.float()
    q32 = q.float()
    docid_t = torch.from_numpy(np.asarray(docid).astype(np.int64))
    for b in range(B):
        rows, seg_bias = [], []
        for i, c in enumerate(beams[b].tolist()):
            if c < 0:
                continue
            lo, hi = int(offsets[c]), int(offsets[c + 1])
            rows.append(torch.arange(lo, hi))
            if bias is not None:
                seg_bias.append(bias[b, i].float().expand(hi - lo))
        if not rows:
            continue
        rows = torch.cat(rows)
        if rows.numel() == 0:
            continue
        s = f(compute_similarity(q32[b:b + 1], emb32[rows])[0])
        if bias is not None:
            s = s + torch.cat(seg_bias)
        kk = min(k, s.numel())
        v, i = s.topk(kk, largest=True, sorted=True)
        out_s[b, :kk] = v
        out_i[b, :kk] = docid_t[rows[i]]
    return out_s, out_i


def leaf_centroids(doc_embed, id_mapping: Dict[str, List[int]]) -> torch.Tensor:
    """tree_embedding_calculate, leaf part (main_models.py:154-158): `sum([embedding[i] for i in idx]) / len(idx)`,
    one row per cluster in id_mapping order."""
    return torch.stack([sum([doc_embed[i] for i in idx]) / len(idx) for idx in id_mapping.values()])


def tree_embedding_insert(cluster_embedding: torch.Tensor, keys: List[str], id_mapping: Dict[str, List[int]], insert_doc,
                          docnum: int) -> Dict[str, List[int]]:
    """main_models.py:282-294: for index >= docnum: `sim = torch.mul(insert_doc[index], cluster_embedding).sum(-1)`,
    `np.argmax(sim)`, append the index to that cluster's list, `list(set(...))`.  (The reference walks its tree to get
    `cluster_embedding` (:268-281); here the [C, D] centroid matrix and the key of each row are passed in.)"""
    for index in range(len(insert_doc)):
        if index < docnum:
            continue
        sim = torch.mul(insert_doc[index], cluster_embedding).sum(dim=-1)
        target = keys[int(np.argmax(sim))]
        id_mapping[target].append(index)
        id_mapping[target] = list(set(id_mapping[target]))
    return id_mapping


def tree_embedding_calculate(root: Node, embedding) -> None:
    """main_models.py:154-179.  Depth-first: a node that lists documents (a leaf cluster) gets the mean of their
    embeddings (`sum([...]) / len`) and is not descended further; any other node gets the leaf-count-weighted mean of
    its children, accumulated in the children's insertion order (`embed * leaf_num` summed, then `/ sum(leaf_num)`)."""
    def dfs(cur):
        if len(cur.embedding_index) > 0:
            cur.embedding = sum([embedding[i] for i in cur.embedding_index]) / len(cur.embedding_index)
            cur.all_leaf_num = len(cur.embedding_index)
            return cur.embedding, cur.all_leaf_num
        embeds, leaf_nums = [], []
        for key in cur.children.keys():
            e, n = dfs(cur.children[key])
            embeds.append(e)
            leaf_nums.append(n)
        acc = None
        for e, n in zip(embeds, leaf_nums):
            acc = e.clone() * n if acc is None else acc + e.clone() * n
        cur.embedding = acc / sum(leaf_nums)
        cur.all_leaf_num = sum(leaf_nums)
        return cur.embedding, cur.all_leaf_num

    dfs(root)


def tree_match(root: Node, doc_embed: torch.Tensor) -> np.ndarray:
    """main_models.py:232-252.  Greedy descent from the root: at every node take the child whose embedding has the
    largest `torch.mul(doc, child).sum(-1)` (np.argmax: first maximum in the children's insertion order); stop at a
    node whose only child carries no embedding (the EOS child of a leaf cluster).  Returns [0, tok..., 1]."""
    cur, out = root, [0]
    while True:
        kids = list(cur.children.keys())
        if len(kids) == 1 and cur.children[kids[0]].embedding is None:
            break
        cand = torch.stack([cur.children[k].embedding for k in kids])
        sim = torch.mul(doc_embed, cand).sum(dim=-1)
        target = kids[int(np.argmax(sim))]
        out.append(target)
        cur = cur.children[target]
    out.append(1)
    return np.array(out)


def encoder_cal(query: torch.Tensor, all_doc: torch.Tensor, valid_num: Sequence[int], loss_func: str, tau: float,
                intra_rate: float) -> torch.Tensor:
    """main_models.py:1184-1221, the training-time contrastive loss.  all_doc = cat([positives (one per query), in-cluster
    candidates of query 0, of query 1, ...]) with valid_num[i] candidates for query i (main_models.py:1259-1273).
    dot_sim = f(q . d) for every (query, doc); per query  -log exp(s_ii / tau) + log( intra_rate * sum over its OWN
    candidates of exp(s / tau) + sum over the OTHER queries' candidates of exp(s / tau) ); mean over queries.  The
    reference's intra_rate == 1 branch (:1206-1219) is the same expression with intra_rate = 1, evaluated batched."""
    B = len(valid_num)
    func = torch.sigmoid if loss_func == "sigmoid" else torch.tanh
    dot_sim = func(torch.mul(query.unsqueeze(1), all_doc.unsqueeze(0)).sum(-1))
    if intra_rate != 1:
        loss = 0
        for i in range(B):
            lo, hi = B + sum(valid_num[:i]), B + sum(valid_num[:i + 1])
            nominator = torch.exp(dot_sim[i][i] / tau)
            intra = torch.sum(torch.exp(dot_sim[i][lo:hi] / tau), dim=-1)
            inter = torch.sum(torch.exp(torch.cat([dot_sim[i][B:lo], dot_sim[i][hi:]], dim=0) / tau), dim=-1)
            loss = loss + (-torch.log(nominator) + torch.log(intra_rate * intra + inter))
    else:
        nominator = torch.exp(torch.gather(dot_sim, index=torch.arange(B).unsqueeze(1), dim=1) / tau)
        denominator = torch.sum(torch.exp(dot_sim[:, B:] / tau), dim=-1)
        loss = -torch.sum(torch.log(nominator), dim=0) + torch.sum(torch.log(denominator), dim=0)
    return loss / B


def merge_topk(scores: torch.Tensor, docids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """New in the sharded design (no reference counterpart): merge G per-rank
    sorted candidate lists [G, B, k'] into one [B, k].  Equivalent to topk over the union."""
    G, B, kp = scores.shape
    s = scores.permute(1, 0, 2).reshape(B, G * kp)
    d = docids.permute(1, 0, 2).reshape(B, G * kp)
    v, i = s.topk(k, dim=1, largest=True, sorted=True)
    dd = torch.gather(d, 1, i)
    dd = torch.where(torch.isinf(v) & (v < 0), torch.full_like(dd, -1), dd)
    return v, dd


# --------------------------------------------------------------------------------------
# a11. prefix-tree mask — GDR_model/transformers/generation_utils_previous.py:712-730
# --------------------------------------------------------------------------------------


def tree_mask_allowed(root, input_ids_row: Sequence[int]) -> List[int]:
    """generation_utils_previous.py:716-727.  Walk the trie along input_ids[1:] (the leading
    pad/decoder-start token is ignored); allowed = children of the reached node, or [1] (EOS)
    if the path leaves the tree; a reached node with no children allows nothing."""
    cur = root
    for value in list(input_ids_row)[1:]:
        nxt = cur.children.get(value)
        if nxt is None:
            return [1]
        cur = nxt
    return list(cur.children.keys())


def tree_mask(scores: torch.Tensor, input_ids: torch.Tensor, root) -> torch.Tensor:
    """generation_utils_previous.py:714-729: mask = -inf everywhere, 0 at allowed tokens;
    `scores += mask` (so kept entries are `s + 0.0`, masked entries `s + (-inf)`)."""
    mask = torch.ones_like(scores) * float("-inf")
    ids = input_ids.tolist()
    for i in range(scores.shape[0]):
        mask[i, tree_mask_allowed(root, ids[i])] = 0
    return scores + mask


def beam_step(next_token_logits: torch.Tensor, input_ids: torch.Tensor, beam_scores: torch.Tensor, root, num_beams: int
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """generation_utils_previous.py:694 (log_softmax) -> :714-729 (tree mask) -> :757-771:
    `next_scores = scores + beam_scores[:, None]`, `.view(batch_size, num_beams * vocab_size)`,
    `torch.topk(next_scores, 2 * num_beams, dim=1, largest=True, sorted=True)`.
    (postprocess_next_token_scores, :696-708, is a no-op under the reference's generate() defaults.)"""
    scores = torch.log_softmax(next_token_logits, dim=-1)
    scores = tree_mask(scores, input_ids, root)
    next_scores = scores + beam_scores[:, None].expand_as(scores)
    B = next_token_logits.shape[0] // num_beams
    next_scores = next_scores.view(B, num_beams * next_token_logits.shape[1])
    return torch.topk(next_scores, 2 * num_beams, dim=1, largest=True, sorted=True)


# --------------------------------------------------------------------------------------
# a12. positional mask — GDR_model/transformers/modeling_t5.py:1546-1571 (eval) / 1279-1301 (train)
# --------------------------------------------------------------------------------------


def position_valid_indices(seq_length: int, output_vocab_size: int, last_eos_only: bool) -> torch.Tensor:
    """modeling_t5.py:1554-1559 (eval) / 1289-1296 (train: the last position allows only
    token 1, line 1296).  Row t = {t*V_out+2 .. t*V_out+V_out+1} ∪ {1}."""
    valid = torch.arange(output_vocab_size).view(1, -1) + torch.arange(seq_length).view(-1, 1) * output_vocab_size + 2
    valid = torch.cat((valid, torch.ones(seq_length, 1, dtype=valid.dtype)), dim=-1).long()
    if last_eos_only:
        valid[-1, :] = 1
    return valid


def position_mask(logits: torch.Tensor, output_vocab_size: int, last_eos_only: bool = False) -> torch.Tensor:
    """modeling_t5.py:1566-1569: `mask = zeros_like(x) - 1e9; mask.scatter_(-1, valid, 0); x + mask`
    on logits [bz, seq_length, vocab].  With `last_eos_only` it is the training-time
    `logit_mask` buffer (modeling_t5.py:1279-1301, applied :1644) for seq_length = max_output_length."""
    bz, sl, _ = logits.shape
    valid = position_valid_indices(sl, output_vocab_size, last_eos_only).unsqueeze(0).repeat(bz, 1, 1)
    mask = torch.zeros_like(logits) - 1e9
    mask = mask.scatter_(-1, valid, torch.zeros_like(logits))
    return logits + mask


# --------------------------------------------------------------------------------------
# synthetic workloads — shared by tests and bench so both sides see the same inputs
# --------------------------------------------------------------------------------------


def synth_corpus(N: int, C: int, D: int = 768, seed: int = 1234, zipf: float = 0.0):
    """emb = randn(N,D)*D^-0.5; assign = randint(0,C) (or Zipf-skewed); CSR by stable argsort.
    Returns (emb_sorted [N,D] fp32, offsets [C+1] int64, docid [N] int64 = original doc index)."""
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(N, D, generator=g) * D ** -0.5
    if zipf > 0:
        w = 1.0 / torch.arange(1, C + 1, dtype=torch.float64) ** zipf
        assign = torch.multinomial(w / w.sum(), N, replacement=True, generator=g)
    else:
        assign = torch.randint(0, C, (N,), generator=g)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=C)
    offsets = torch.zeros(C + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    return emb[order].contiguous(), offsets.numpy(), order.numpy()


def synth_queries(Q: int, C: int, K: int, D: int = 768, seed: int = 4321):
    """q = randn(Q,D); beams = randperm(C)[:K] per query (distinct clusters, like distinct beam
    hypotheses); beam scores = -cumsum(rand(K)) (descending log-prob-like)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(Q, D, generator=g)
    if K <= C:
        beams = torch.stack([torch.randperm(C, generator=g)[:K] for _ in range(Q)]) if Q * C <= 5e7 else \
            torch.argsort(torch.rand(Q, C, generator=g), dim=1)[:, :K]
    else:
        beams = torch.randint(0, C, (Q, K), generator=g)
    beam_scores = -torch.cumsum(torch.rand(Q, K, generator=g), dim=1)
    return q, beams.to(torch.int32).numpy(), beam_scores
