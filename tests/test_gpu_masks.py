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


def _check_beam_step(got_s, got_t, ref_s, ref_t, what):
    """values within 1e-5; tokens equal wherever the reference value is finite and not within 1e-5 of a neighbour;
    past the survivors ours is (-inf, a valid in-row index: beam 0, EOS) — the reference's topk returns -inf with unspecified
    (but in-row) indices there, and the Hugging Face bookkeeping derives beam / token ids from them."""
    got_s, got_t = got_s.cpu().numpy(), got_t.cpu().numpy()
    finite = np.isfinite(ref_s)
    assert np.array_equal(np.isfinite(got_s), finite), what
    np.testing.assert_allclose(got_s[finite], ref_s[finite], rtol=1e-5, atol=1e-5, err_msg=what)
    assert np.all(got_t[~finite] == 1), what          # flat index 0 * V + eos (= 1)
    assert got_t.min() >= 0, what
    for b in range(ref_s.shape[0]):
        for i in np.nonzero(finite[b] & (got_t[b] != ref_t[b]))[0]:
            js = np.nonzero(ref_t[b] == got_t[b, i])[0]
            assert js.size and abs(ref_s[b, js[0]] - ref_s[b, i]) <= 1e-5 * max(1.0, abs(ref_s[b, i])), (what, b, i)


def test_beam_step_golden():
    from gdr_b200 import Node, TreeMask
    g = load_golden("beam_step")
    hook = TreeMask(rebuild_tree(g["edges"], g["leaves"], Node))
    for case in ("a", "b", "c"):
        K = int(g[f"K_{case}"])
        logits = torch.from_numpy(g[f"logits_{case}"]).cuda()
        before = logits.clone()
        s, t = hook.beam_step(logits, torch.from_numpy(g[f"ids_{case}"]).cuda(), torch.from_numpy(g[f"beam_{case}"]).cuda(), K)
        assert torch.equal(logits, before), "beam_step must not modify the logits"
        _check_beam_step(s, t, g[f"scores_{case}"], g[f"tokens_{case}"], f"beam_step golden {case}")


def test_beam_step_cfg4_shape_vs_oracle():
    """V = 32,128, 3-level 30-ary tree with 1,024 leaf clusters, beam 100 (rows reduced so the CPU oracle finishes fast)."""
    from gdr_b200 import TreeBuilder, TreeMask
    rng = np.random.RandomState(7)
    paths = set()
    while len(paths) < 1024:
        paths.add(tuple(rng.randint(0, 30, 3)))
    tb, otb = TreeBuilder(), orc.TreeBuilder()
    toks = []
    for di, p in enumerate(sorted(paths)):
        t = [i * 30 + int(c) + 2 for i, c in enumerate(p)] + [1]
        toks.append(t)
        tb.add(t, di); otb.add(t, di)
    hook = TreeMask(tb.build())
    V, B, K = 32128, 4, 100
    for cur_len in (1, 2, 3, 4):
        R = B * K
        ids = torch.zeros(R, cur_len, dtype=torch.int64)
        for r in range(R):
            t = toks[rng.randint(len(toks))]
            n = min(cur_len - 1, len(t))
            ids[r, 1:1 + n] = torch.tensor(t[:n])
            if r % 37 == 5 and cur_len > 1:
                ids[r, rng.randint(1, cur_len)] = 91 + rng.randint(1000)
        gen = torch.Generator().manual_seed(cur_len)
        logits = torch.randn(R, V, generator=gen) * 2
        beam = -torch.rand(R, generator=gen) * 8
        if cur_len == 1:
            beam = beam.view(B, K); beam[:, 1:] = -1e9; beam = beam.reshape(-1)
        ref_s, ref_t = orc.beam_step(logits, ids, beam, otb.build(), K)
        s, t = hook.beam_step(logits.cuda(), ids.cuda(), beam.cuda(), K)
        _check_beam_step(s, t, ref_s.numpy(), ref_t.numpy(), f"beam_step cfg4 cur_len={cur_len}")
