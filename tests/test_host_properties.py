"""Property tests (hypothesis) of the host-side logic: codecs, trie flattening, cluster partitioning, candidate packing.
CPU only; the oracle is used as the checker for the codecs."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gdr_oracle as orc  # noqa: E402  (checker)
from gdr_b200.generation import child_insertion_order, flatten_trie  # noqa: E402
from gdr_b200.main_models import TreeBuilder, decode_token, encode_single_newid  # noqa: E402
from gdr_b200.sharded import global_to_local, pack_candidates, partition_clusters  # noqa: E402

paths_st = st.lists(st.lists(st.integers(0, 29), min_size=1, max_size=4), min_size=1, max_size=40)


@settings(max_examples=60, deadline=None)
@given(paths_st, st.booleans())
def test_codec_round_trip_and_oracle_agreement(paths, position):
    """encode_single_newid -> decode_token is the identity on cluster ids, and both agree with the oracle's restatement
    of main_models.py:297-346."""
    args = SimpleNamespace(kary=30, position=int(position), output_vocab_size=30)
    ids = ["-".join(str(d) for d in p) for p in paths]
    L = 7
    rows = np.zeros((len(ids), L), dtype=np.int64)
    for i, s in enumerate(ids):
        toks = encode_single_newid(args, s)
        assert toks == orc.encode_single_newid(s, kary=30, position=position)
        assert toks[-1] == 1 and len(toks) == len(paths[i]) + 1
        rows[i, 1:1 + len(toks)] = toks
    assert decode_token(args, rows) == ids
    assert decode_token(args, rows) == orc.decode_token(rows, output_vocab_size=30, position=position, kary=30)


@settings(max_examples=60, deadline=None)
@given(paths_st)
def test_flattened_trie_invariants(paths):
    """Breadth-first numbering (every depth a contiguous id range, children after parents), edges sorted by token inside a
    node, child_insertion_order a permutation of every node's own edges in dict order, and the CSR walk equals the dict walk."""
    tb = TreeBuilder()
    for i, p in enumerate(paths):
        tb.add([d + 2 + 30 * lvl for lvl, d in enumerate(p)] + [1], i)
    root = tb.build()
    fc, tok, node = flatten_trie(root)
    order = child_insertion_order(root)
    n_nodes = fc.size - 1
    assert fc[0] == 0 and fc[-1] == tok.size == node.size == order.size == n_nodes - 1      # a tree: one edge per non-root node
    depth = np.full(n_nodes, -1)
    depth[0] = 0
    objs = {0: root}
    for n in range(n_nodes):
        lo, hi = int(fc[n]), int(fc[n + 1])
        assert list(tok[lo:hi]) == sorted(tok[lo:hi])
        assert sorted(order[lo:hi].tolist()) == list(range(lo, hi))
        assert [int(tok[e]) for e in order[lo:hi]] == [int(t) for t in objs[n].children]     # dict (insertion) order
        for e in range(lo, hi):
            c = int(node[e])
            assert c > n and depth[c] == -1
            depth[c] = depth[n] + 1
            objs[c] = objs[n].children[int(tok[e])]
    assert (depth >= 0).all() and (np.diff(depth) >= 0).all()                                # levels are contiguous id ranges


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 5000), min_size=1, max_size=300), st.integers(1, 8))
def test_partition_covers_and_balances(sizes, world):
    owner = partition_clusters(sizes, world)
    assert owner.shape == (len(sizes),) and owner.min() >= 0 and owner.max() < world
    load = np.bincount(owner, weights=np.asarray(sizes, dtype=np.float64), minlength=world)
    assert load.max() - load.min() <= max(sizes) * 2 + 1                                    # LPT-like: within two largest clusters
    seen = np.zeros(len(sizes), dtype=int)
    for r in range(world):
        g2l, mine = global_to_local(owner, r)
        seen[mine] += 1
        assert (g2l[mine] == np.arange(mine.size)).all() and (np.delete(g2l, mine) == -1).all()
    assert (seen == 1).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(1, 40), st.integers(0, 2 ** 31 - 2))
def test_pack_candidates_is_lossless(B, k, seed):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(B, k, generator=g)
    s[0, 0] = float("-inf")
    d = torch.randint(-1, 2 ** 31 - 1, (B, k), generator=g, dtype=torch.int64).int()
    packed = pack_candidates(s, d)
    assert packed.dtype == torch.int32 and packed.shape == (2, B, k)
    assert torch.equal(packed[0].view(torch.float32), s) and torch.equal(packed[1], d)
