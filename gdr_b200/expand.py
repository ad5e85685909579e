"""Index expansion: assign new documents to the existing leaf clusters.

Mirrors the reference's `tree_embedding_calculate` (leaf part, GDR_model/main_models.py:154-158: a leaf cluster's
embedding is the mean of its members) and `tree_embedding_insert` (main_models.py:268-295: every document with index
>= args.docnum goes to `argmax_c doc . centroid_c` and is appended to that cluster's list in `id_mapping`), used by
`--expand` (main.py:396).  The arg-max over all clusters is the same gather-score-select primitive as the fine stage:
the centroids form a one-cluster store and every new document asks for its top-1.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from .store import ClusterStore


def assign_to_clusters(centroids: torch.Tensor, docs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """centroids [C, D] fp32 cuda, docs [M, D] fp32 cuda -> (cluster index [M] int64, score [M] fp32):
    argmax_c (docs[m] * centroids[c]).sum(-1), ties to the lowest cluster index (as np.argmax, main_models.py:288)."""
    C = centroids.shape[0]
    dev = centroids.device
    store = ClusterStore(centroids.contiguous(), torch.tensor([0, C]), torch.arange(C))
    beams = torch.zeros((docs.shape[0], 1), dtype=torch.int32, device=dev)
    s, d = store.score_topk(docs.to(dev, torch.float32), beams, 1)
    return d[:, 0].long(), s[:, 0]


def tree_embedding_insert(store: ClusterStore, id_mapping: Dict[str, List[int]], insert_doc, docnum: int
                          ) -> Dict[str, List[int]]:
    """reference main_models.py:268-295 on top of a ClusterStore built from (doc_embed, id_mapping): documents
    `insert_doc[docnum:]` are appended to the `id_mapping` list of their nearest leaf cluster (centroids from the
    store).  Returns `id_mapping` (modified in place; each list de-duplicated like the reference's list(set(...)),
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
