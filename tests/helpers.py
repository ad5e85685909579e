"""Shared test helpers (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NEAR_TIE = 1e-5     # SURVEY.md §8d parity rule: a docid mismatch is excused only for near-ties
SCORE_RTOL = 1e-3   # north star: scores within 1e-3 relative (fp32 accumulate)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def bf16_bits_to_float(bits: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(bits.astype(np.int16)).view(torch.bfloat16).float()


def tol(s):
    return NEAR_TIE * max(1.0, abs(float(s)))


def assert_topk_parity(got_s, got_d, ref_s, ref_d, what=""):
    """Position-by-position docid equality; a mismatch at position i is excused only when it is a
    near-tie under fp32 reassociation: the doc we put at i sits in the oracle's list at a position j
    whose oracle score is within 1e-5*max(1,|s|) of the oracle score at i (or, for a swap across the
    k boundary, our score is within that distance of the oracle's last score).  Scores must agree within
    1e-3 relative.  Returns the number of excused positions."""
    got_s = torch.as_tensor(got_s).float().cpu().numpy()
    ref_s = torch.as_tensor(ref_s).float().cpu().numpy()
    got_d = torch.as_tensor(got_d).cpu().numpy().astype(np.int64)
    ref_d = torch.as_tensor(ref_d).cpu().numpy().astype(np.int64)
    assert got_s.shape == ref_s.shape and got_d.shape == ref_d.shape, (what, got_s.shape, ref_s.shape)
    excused = 0
    finite = np.isfinite(ref_s)
    assert np.array_equal(np.isfinite(got_s), finite), f"{what}: padding pattern differs"
    np.testing.assert_allclose(got_s[finite], ref_s[finite], rtol=SCORE_RTOL, atol=SCORE_RTOL, err_msg=what)
    assert np.all(got_d[~finite] == -1), f"{what}: padded docids must be -1"
    for b in range(got_d.shape[0]):
        if np.array_equal(got_d[b], ref_d[b]):
            continue
        for i in np.nonzero(got_d[b] != ref_d[b])[0]:
            js = np.nonzero(ref_d[b] == got_d[b, i])[0]
            if js.size:
                ok = np.min(np.abs(ref_s[b, js] - ref_s[b, i])) <= tol(ref_s[b, i])
            else:
                last = ref_s[b][finite[b]][-1]
                ok = abs(got_s[b, i] - last) <= 2 * tol(last) and abs(ref_s[b, i] - last) <= 2 * tol(last)
            assert ok, f"{what}: query {b} position {i}: docid {got_d[b, i]} vs oracle {ref_d[b, i]} is not a near-tie"
            excused += 1
    return excused


def fine_stage_inputs(name):
    g = load_golden(name)
    emb = torch.from_numpy(g["emb"])
    id_mapping = json.loads(str(g["id_mapping_json"]))
    dec = [[str(x) for x in row] for row in g["dec"]]
    return dict(emb=emb, doc_embed=[emb[i] for i in range(emb.shape[0])], id_mapping=id_mapping, dec=dec,
                beam_scores=g["beam_scores"].tolist(), q=torch.from_numpy(g["q"]),
                score_rate=g["score_rate"].tolist(), loss_func=str(g["loss_func"]), docids=g["docids"])


def rebuild_tree(edges, leaves, node_cls):
    """Rebuild a trie from the fixture's (parent, token, child) edge list with the given Node class."""
    nodes = {0: node_cls(0)}
    for parent, tok, child in edges.tolist():
        nodes[child] = node_cls(tok)
    for parent, tok, child in edges.tolist():
        nodes[parent].children[tok] = nodes[child]
    for node, doc in leaves.tolist():
        nodes[node].embedding_index.append(doc)
    return nodes[0]
