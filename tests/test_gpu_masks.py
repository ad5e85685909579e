"""Bit-exact parity of the mask kernels with the reference's own outputs (tests/golden) and the oracle."""
import numpy as np
import pytest
import torch

import gdr_oracle as orc
from helpers import load_golden, rebuild_tree

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


@pytest.mark.parametrize("strict", [False, True])
def test_tree_mask_golden(strict):
    from gdr_b200 import DeviceTrie, Node
    g = load_golden("tree")
    root = rebuild_tree(g["edges"], g["leaves"], Node)
    trie = DeviceTrie.from_root(root)
    for cur_len in (1, 2, 3, 4, 5):
        scores = torch.from_numpy(g[f"in_{cur_len}"]).cuda()
        ids = torch.from_numpy(g[f"ids_{cur_len}"]).cuda()
        out = trie.mask_(scores, ids, strict=strict)
        assert out.data_ptr() == scores.data_ptr()
        assert np.array_equal(_bits(out), g[f"out_{cur_len}"].view(np.uint32)), cur_len


def test_tree_mask_cfg4_shape_vs_oracle():
    """BASELINE.json configs[3]: V = 32,128, 3-level 30-ary tree with 1,024 leaf clusters; rows sampled at
    cur_len 1..4 with off-tree and finished rows (row count reduced so the Python oracle finishes fast;
    the full 25,600 rows are checked through the allowed-set property below)."""
    from gdr_b200 import DeviceTrie, Node, TreeBuilder, TreeMask
    rng = np.random.RandomState(4)
    paths = set()
    while len(paths) < 1024:
        paths.add(tuple(rng.randint(0, 30, 3)))
    paths = sorted(paths)
    tb, otb = TreeBuilder(), orc.TreeBuilder()
    toks = []
    for di, p in enumerate(paths):
        t = [i * 30 + int(c) + 2 for i, c in enumerate(p)] + [1]
        toks.append(t)
        tb.add(t, di); otb.add(t, di)
    hook = TreeMask(tb.build())
    V = 32128
    for cur_len in (1, 2, 3, 4, 5):
        R = 256
        ids = torch.zeros(R, cur_len, dtype=torch.int64)
        for r in range(R):
            t = toks[rng.randint(len(toks))]
            n = min(cur_len - 1, len(t))
            ids[r, 1:1 + n] = torch.tensor(t[:n])
            if r % 97 == 5 and cur_len > 1:
                ids[r, rng.randint(1, cur_len)] = 91 + rng.randint(1000)
        scores = torch.log_softmax(torch.randn(R, V, generator=torch.Generator().manual_seed(cur_len)), -1)
        ref = orc.tree_mask(scores, ids, otb.build())
        out = hook(ids.cuda(), scores.clone().cuda())
        assert np.array_equal(_bits(out), _bits(ref)), cur_len
    # full cfg4 row count: 256 x beam 100 = 25,600 rows, property check (allowed set per row)
    R, cur_len = 25600, 3
    pick = rng.randint(len(toks), size=R)
    ids = torch.zeros(R, cur_len, dtype=torch.int64)
    ids[:, 1:] = torch.tensor([toks[i][:2] for i in pick])
    scores = torch.full((R, V), -1.0, device="cuda")
    out = hook(ids.cuda(), scores)
    kept = (out == -1.0)
    assert torch.all((out == -1.0) | (out == float("-inf")))
    n_kept = kept.sum(1).cpu()
    root = otb.build()
    for r in range(0, R, 997):
        allowed = orc.tree_mask_allowed(root, ids[r].tolist())
        assert n_kept[r] == len(allowed) and bool(kept[r, allowed].all())


def test_tree_mask_strict_propagates_nan_like_reference():
    from gdr_b200 import DeviceTrie, Node, TreeBuilder
    tb = TreeBuilder(); tb.add([5, 40, 1], 0)
    trie = DeviceTrie.from_root(tb.build())
    scores = torch.zeros(2, 64); scores[0, 9] = float("nan"); scores[1, 5] = float("nan")
    ids = torch.zeros(2, 1, dtype=torch.int64)
    ref = orc.tree_mask(scores, ids, tb.build())
    out = trie.mask_(scores.clone().cuda(), ids.cuda(), strict=True)
    # NaN stays NaN at the same positions (the GPU returns the canonical NaN 0x7fffffff where the CPU keeps
    # the input payload, so NaNs are compared as NaNs, everything else bit for bit)
    nan_ref = torch.isnan(ref)
    assert torch.equal(torch.isnan(out).cpu(), nan_ref) and int(nan_ref.sum()) == 2
    assert np.array_equal(_bits(out)[~nan_ref.numpy()], _bits(ref)[~nan_ref.numpy()])
    # default (write-only) mode turns a masked NaN into -inf, an allowed NaN stays NaN
    out = trie.mask_(scores.clone().cuda(), ids.cuda(), strict=False).cpu()
    assert out[0, 9] == float("-inf") and torch.isnan(out[1, 5])
    # odd V and unaligned rows take the scalar path
    s = torch.randn(3, 61)
    ids = torch.tensor([[0, 5], [0, 6], [0, 5]])
    out = trie.mask_(s.clone().cuda(), ids.cuda())
    assert np.array_equal(_bits(out), _bits(orc.tree_mask(s, ids, tb.build())))


def test_position_mask_golden():
    from gdr_b200 import build_logit_mask, position_mask_, select_valid_embedding
    g = load_golden("position_mask")
    for key in [k[3:] for k in g.files if k.startswith("in_")]:
        v_out = int(key.split("_")[0])
        x = torch.from_numpy(g["in_" + key]).cuda()
        y = select_valid_embedding(x, v_out)
        assert np.array_equal(_bits(y), g["out_" + key].view(np.uint32)), key
        assert np.array_equal(_bits(x), g["in_" + key].view(np.uint32)), "functional form must not modify its input"
        assert np.array_equal(_bits(position_mask_(x, v_out)), g["out_" + key].view(np.uint32))
    assert np.array_equal(build_logit_mask(10, 302, 30).cpu().numpy(), g["train_logit_mask_30_10"])
    with pytest.raises(RuntimeError):
        position_mask_(torch.zeros(1, 10, 100, device="cuda"), 30)      # vocabulary too small (scatter_ would raise)
