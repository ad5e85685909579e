"""Index expansion: assign new documents to the existing leaf clusters.

Mirrors the reference's `tree_embedding_calculate` (leaf part, GDR_model/main_models.py:154-158: a leaf cluster's
embedding is the mean of its members) and `tree_embedding_insert` (main_models.py:268-295: every document with index
>= args.docnum goes to `argmax_c doc . centroid_c` and is appended to that cluster's list in `id_mapping`), used by
`--expand` (main.py:396).  The arg-max over all clusters is the same gather-score-select primitive as the fine stage:
the centroids form a one-cluster store and every new document asks for its top-1.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .store import ClusterStore


def assign_to_clusters(centroids: torch.Tensor, docs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """centroids [C, D] fp32 cuda, docs [M, D] fp32 cuda -> (cluster index [M] int64, score [M] fp32):
    argmax_c (docs[m] * centroids[c]).sum(-1), ties to the lowest cluster index (as np.argmax, main_models.py:288)."""
    C = centroids.shape[0]
    dev = centroids.device
    store = ClusterStore(centroids.contiguous(), torch.tensor([0, C]), torch.arange(C))
    docs = docs.to(dev, torch.float32)
    # The call's scratch holds M x C scores (and as many keys): documents go in chunks that keep it within ~256 MB and the
    # score-buffer index below 2^31 — the reference's per-document loop (main_models.py:283-290) has no limit on M either.
    chunk = max(1, min(docs.shape[0], (64 << 20) // max(C, 1)))
    idx = torch.empty(docs.shape[0], dtype=torch.int64, device=dev)
    val = torch.empty(docs.shape[0], dtype=torch.float32, device=dev)
    for lo in range(0, docs.shape[0], chunk):
        part = docs[lo:lo + chunk].contiguous()
        beams = torch.zeros((part.shape[0], 1), dtype=torch.int32, device=dev)
        s, d = store.score_topk(part, beams, 1)
        idx[lo:lo + chunk], val[lo:lo + chunk] = d[:, 0].long(), s[:, 0]
    return idx, val


def tree_embedding_insert(store: ClusterStore, id_mapping: Dict[str, List[int]], insert_doc, docnum: int
                          ) -> Dict[str, List[int]]:
    """reference main_models.py:268-295 on top of a ClusterStore built from (doc_embed, id_mapping): documents
    `insert_doc[docnum:]` are appended to the `id_mapping` list of their nearest leaf cluster (centroids from the
    store: an fp32 store reproduces the reference's fp32 means bit for bit; a bf16 store averages the ROUNDED embeddings, so a
    document whose two best clusters are within bf16 rounding of each other can land in the other one — build the store with
    dtype=torch.float32 for index expansion when that matters).  Returns `id_mapping` (modified in place; each list de-duplicated like the reference's list(set(...)),
    here keeping first-seen order)."""
    if store.keys is None:
        raise ValueError("the store must carry its cluster keys")
    new = list(range(docnum, len(insert_doc)))
    if not new:
        return id_mapping
    docs = torch.stack([torch.as_tensor(insert_doc[i]).reshape(-1).float() for i in new]).to(store.emb.device)
    idx, _ = assign_to_clusters(store.centroids(), docs)
    for doc_index, c in zip(new, idx.cpu().tolist()):
        lst = id_mapping[store.keys[c]]
        if doc_index not in lst:
            lst.append(doc_index)
    return id_mapping


def node_embeddings(trie, store: ClusterStore, args) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference `tree_embedding_calculate` (main_models.py:154-179) on the device trie: every node's embedding
    ([n_nodes, D] fp32) and leaf count ([n_nodes] int32, 0 = the node carries no embedding).  Leaf clusters are the
    store's clusters (their keys are cluster-id strings, encoded with `encode_single_newid(args, key)` to find the
    node); inner nodes are the leaf-count-weighted means of their children, accumulated in the children's insertion
    order with the reference's operations."""
    from . import _cabi
    from .main_models import encode_single_newid
    if store.keys is None:
        raise ValueError("the store must carry its cluster keys")
    dev = store.emb.device
    node_cluster = np.full(trie.n_nodes, -1, dtype=np.int32)
    for c, key in enumerate(store.keys):
        n = trie.find(encode_single_newid(args, key)[:-1])                # the leaf-cluster node = parent of the EOS node
        if n < 0:
            raise KeyError(f"cluster {key!r} is not a path of the tree")
        node_cluster[n] = c
    leaf_emb = store.centroids()
    leaf_num = torch.as_tensor(np.diff(store.offsets_host).astype(np.int32), device=dev)
    node_emb = torch.zeros((trie.n_nodes, store.dim), dtype=torch.float32, device=dev)
    node_leaf_num = torch.zeros(trie.n_nodes, dtype=torch.int32, device=dev)
    nc = torch.as_tensor(node_cluster, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gdr_trie_node_embeddings(trie._handle, nc.data_ptr(), leaf_emb.data_ptr(), leaf_num.data_ptr(), store.dim,
                                                         node_emb.data_ptr(), node_leaf_num.data_ptr(), _cabi.stream_ptr()))
    return node_emb, node_leaf_num


def tree_match(trie, node_emb: torch.Tensor, node_leaf_num: torch.Tensor, docs: torch.Tensor, max_len: int = 16) -> List[np.ndarray]:
    """reference `tree_match` (main_models.py:232-252) for a batch of documents [M, D]: greedy descent by
    `doc . child embedding`; one int array `[0, tok, ..., 1]` per document, as the reference returns."""
    from . import _cabi
    dev = node_emb.device
    docs = docs.to(dev, torch.float32).contiguous()
    M = docs.shape[0]
    out = torch.zeros((M, max_len), dtype=torch.int32, device=dev)
    out_len = torch.zeros(M, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gdr_tree_match(trie._handle, node_emb.data_ptr(), node_leaf_num.data_ptr(), node_emb.shape[1], docs.data_ptr(), M,
                                               max_len, out.data_ptr(), out_len.data_ptr(), _cabi.stream_ptr()))
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    return [out[m, :out_len[m]].astype(np.int64) for m in range(M)]
