"""Parity of the CUDA path (through the C ABI) with the oracle: cluster-restricted scoring + top-k.
Bar (north star / SURVEY.md §8d): docid lists equal position by position, a mismatch excused only
for oracle near-ties (<= 1e-5*max(1,|s|)); scores within 1e-3 relative."""
import numpy as np
import pytest
import torch

import gdr_oracle as orc
from helpers import assert_topk_parity, bf16_bits_to_float, fine_stage_inputs, load_golden

pytestmark = pytest.mark.gpu

FLAG_SIMT, FLAG_UMMA = 2, 4


def _store(emb, offsets, docid, dtype):
    from gdr_b200 import ClusterStore
    return ClusterStore.from_csr(emb, offsets, docid, dtype=dtype)


@pytest.mark.parametrize("name", ["dense_topk_d768", "dense_topk_d128", "dense_topk_zipf"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_golden_dense_topk(name, dtype):
    g = load_golden(name)
    emb = bf16_bits_to_float(g["emb_bf16_bits"])       # bf16-representable: both store dtypes hold the same values
    st = _store(emb, g["offsets"], g["docid"], dtype)
    q = torch.from_numpy(g["q"]).cuda()
    beams = torch.from_numpy(g["beams"]).cuda()
    k = int(g["k"])
    s, d = st.score_topk(q, beams, k)
    assert_topk_parity(s, d, g["scores_plain"], g["docids_plain"], name)
    bias = torch.from_numpy(g["bias"]).cuda()
    s, d = st.score_topk(q, beams, k, prob=bias, alphas=[1.0], act="tanh")
    assert_topk_parity(s[0], d[0], g["scores_tanh_bias"], g["docids_tanh_bias"], name + "/tanh+bias")


@pytest.mark.parametrize("name", ["fine_stage_tanh", "fine_stage_sigmoid"])
@pytest.mark.parametrize("dtype", [torch.float32])
def test_golden_fine_stage_drop_in(name, dtype):
    """FineStage (mirror of main_models.py:1434-1637) reproduces the docid strings the real
    validation_step_i returned."""
    from types import SimpleNamespace
    from gdr_b200 import FineStage
    f = fine_stage_inputs(name)
    K = len(f["dec"][0])
    args = SimpleNamespace(num_return_sequences=K, score_rate=f["score_rate"], loss_func=f["loss_func"])
    fs = FineStage(args, f["doc_embed"], f["id_mapping"], dtype=dtype)
    out = fs(f["dec"], f["beam_scores"], f["q"], texts=["q"] * len(f["dec"]), gt_answers=["gt"] * len(f["dec"]))
    ref = orc.fine_stage(f["doc_embed"], f["id_mapping"], f["dec"], f["beam_scores"], f["q"], f["score_rate"], f["loss_func"], K)
    vals, ids = fs.retrieve(f["dec"], f["beam_scores"], f["q"])
    for b in range(len(f["dec"])):
        for r in range(len(f["score_rate"])):
            assert out[b][r][0][0] == "q" and out[b][r][0][2] == "gt"
            got = [int(x) for x in out[b][r][0][1].split(",")]
            assert_topk_parity(vals[b, r][None], torch.tensor(got)[None], ref[b][r][0][None],
                               torch.tensor(f["docids"][b, r])[None], f"{name} b{b} r{r}")
    with pytest.raises(KeyError):
        fs([["not-a-cluster"] * K] * len(f["dec"]), f["beam_scores"], f["q"])
    big = FineStage(args, store=fs.store, k=10 ** 3)
    with pytest.raises(RuntimeError):
        big(f["dec"], f["beam_scores"], f["q"])


CASES = [
    # N, C, D, Q, K, k, zipf, dtype, flags
    (20000, 256, 768, 64, 10, 100, 0.0, torch.float32, 0),
    (20000, 256, 768, 64, 10, 100, 0.0, torch.bfloat16, FLAG_SIMT),
    (20000, 128, 768, 256, 20, 100, 0.0, torch.bfloat16, 0),          # dense groups (~40 pairs per cluster)
    (20000, 128, 768, 256, 20, 100, 0.0, torch.bfloat16, FLAG_UMMA),
    (30000, 200, 256, 100, 16, 1000, 0.0, torch.bfloat16, 0),        # k = 1000
    (30000, 300, 128, 50, 8, 64, 1.2, torch.bfloat16, 0),            # Zipf-skewed clusters (global-keys top-k)
    (30000, 300, 128, 50, 8, 64, 1.2, torch.float32, 0),
    (5000, 64, 1024, 33, 7, 50, 0.0, torch.bfloat16, 0),             # max dim
    (5000, 64, 8, 33, 7, 50, 0.0, torch.float32, 0),                 # min dim
    (5000, 64, 200, 33, 7, 50, 0.0, torch.bfloat16, 0),              # dim not a multiple of 64 (no TMA path)
    (40000, 256, 64, 96, 24, 100, 0.0, torch.bfloat16, 0),            # ~3,700 candidates per query: top-100 streams the scores twice (more than fit in registers)
    (30000, 300, 64, 40, 16, 64, 1.2, torch.bfloat16, 0),            # K x largest cluster > 65,535: 256-thread top-k with 32-bit histogram bins
    (3000, 200, 64, 500, 3, 10, 0.0, torch.bfloat16, FLAG_UMMA),      # fewer tiles than CTAs in the tile queue, one-K-block tiles
    (20000, 128, 768, 256, 20, 100, 0.0, torch.float32, 0),          # fp32 store, dense groups (~40 pairs per cluster): the tiled fp32 kernel, chunked groups
    (9000, 96, 96, 200, 8, 64, 1.2, torch.float32, 0),               # ... Zipf-skewed clusters (multi-tile slabs), three K chunks
    (9000, 96, 96, 200, 8, 64, 0.0, torch.float32, FLAG_SIMT),       # ... and the same store forced onto the GEMV
]


@pytest.mark.parametrize("N,C,D,Q,K,k,zipf,dtype,flags", CASES)
def test_seeded_synthetic(N, C, D, Q, K, k, zipf, dtype, flags):
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=N + C, zipf=zipf)
    if dtype == torch.bfloat16:
        emb = emb.bfloat16().float()                 # the oracle consumes the rounded values
    q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=Q + K)
    prob = torch.softmax(beam_scores, -1)
    st = _store(emb, offsets, docid, dtype)
    if flags == FLAG_UMMA and D % 64 != 0:
        pytest.skip("tcgen05 path needs dim % 64 == 0")
    s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, flags=flags)
    ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k)
    n_exc = assert_topk_parity(s, d, ref_s, ref_d, "plain")
    s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, prob=prob.cuda(), alphas=[0.0, 1.0, 3.0], act="tanh", flags=flags)
    for r, alpha in enumerate([0.0, 1.0, 3.0]):
        ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k, bias=prob * alpha, act="tanh")
        n_exc += assert_topk_parity(s[r], d[r], ref_s, ref_d, f"tanh alpha={alpha}")
    assert n_exc <= 0.002 * 4 * Q * k + 2, f"too many near-tie excuses: {n_exc}"
    stats = st.last_stats()
    if flags == FLAG_SIMT:
        assert stats["umma_tiles"] == 0
    if flags == FLAG_UMMA:
        assert stats["simt_items"] == 0 and stats["umma_tiles"] > 0


def test_edge_cases_ragged_empty_absent_and_padding():
    D = 64
    sizes = [0, 1, 3, 0, 130, 257, 2, 0]                 # empty clusters, multi-tile clusters
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    N = int(offsets[-1])
    g = torch.Generator().manual_seed(8)
    emb = (torch.randn(N, D, generator=g) * D ** -0.5).bfloat16().float()
    docid = torch.randperm(N, generator=g).numpy()
    q = torch.randn(5, D, generator=g)
    beams = np.array([[0, 3, 7], [1, -1, 2], [4, 5, 6], [-1, -1, -1], [5, 5, 1]], dtype=np.int32)   # dup cluster, all absent
    for dtype, flags in ((torch.float32, 0), (torch.bfloat16, 0), (torch.bfloat16, FLAG_UMMA)):
        st = _store(emb, offsets, docid, dtype)
        for k in (1, 4, 300, 1000):
            s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, flags=flags)
            ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k)
            assert_topk_parity(s, d, ref_s, ref_d, f"edge k={k}")
            assert torch.all(s[0] == float("-inf")) and torch.all(d[3] == -1)
    # B = 0 is a no-op
    s, d = st.score_topk(q[:0].cuda(), torch.zeros(0, 3, dtype=torch.int32).cuda(), 4)
    assert s.shape == (0, 4)


def test_large_groups_are_chunked_on_the_tensor_path():
    """More pairs per cluster than one tcgen05 tile holds (32): groups are split into chunks, clusters into 128-row tiles."""
    N, C, D, Q, K, k = 6000, 12, 128, 200, 6, 100         # 100 pairs per cluster, clusters of ~500 rows
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=21)
    emb = emb.bfloat16().float()
    q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=22)
    st = _store(emb, offsets, docid, torch.bfloat16)
    ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k)
    for flags in (0, FLAG_UMMA, FLAG_SIMT):
        s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, flags=flags)
        assert_topk_parity(s, d, ref_s, ref_d, f"large groups flags={flags}")
    assert st.last_stats()["umma_tiles"] == 0
    st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k)
    assert st.last_stats()["umma_tiles"] >= 150           # default policy: dense batch -> tensor path, ~4 row tiles x 4 chunks x 12 clusters


def test_mass_ties_are_deterministic_and_docid_ordered():
    """tanh saturates to exactly 1.0 for large |q.d|: thousands of exact ties.  Ours
    orders ties by ascending docid, so the result is the k smallest docids among the tied maximum."""
    D, N, C = 64, 4000, 8
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=1)
    emb = (emb * 40).bfloat16().float()
    q = torch.randn(6, D, generator=torch.Generator().manual_seed(2)) * 10
    beams = np.stack([np.arange(C, dtype=np.int32)] * 6)
    st = _store(emb, offsets, docid, torch.bfloat16)
    k = 100
    s1, d1 = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, act="tanh")
    s2, d2 = st.score_topk(q.cuda(), torch.from_numpy(beams[:, ::-1].copy()).cuda(), k, act="tanh")
    assert torch.equal(d1, d2) and torch.equal(s1, s2), "result must not depend on candidate order"
    full = torch.tanh(q @ emb.T)
    for b in range(6):
        ones = np.sort(docid[(full[b] == 1.0).numpy()])
        if ones.size >= k:
            assert d1[b].cpu().tolist() == ones[:k].tolist()
            assert torch.all(s1[b] == 1.0)


@pytest.mark.parametrize("flags", [0, FLAG_SIMT, FLAG_UMMA])
def test_per_beam_queries(flags):
    """One query vector per (query, beam): main_models.py:1467-1571,1583-1594."""
    N, C, D, Q, K, k = 8000, 64, 128, 20, 5, 30
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=3)
    emb = emb.bfloat16().float()
    q, beams, _ = orc.synth_queries(Q * K, C, K, D, seed=4)
    beams = beams[:Q]
    st = _store(emb, offsets, docid, torch.bfloat16)
    s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, per_beam=True, flags=flags)
    # oracle: each beam segment scored with its own query vector
    ref_s = torch.full((Q, k), float("-inf"))
    ref_d = torch.full((Q, k), -1, dtype=torch.int64)
    for b in range(Q):
        sc, ids = [], []
        for i, c in enumerate(beams[b]):
            rows = torch.arange(int(offsets[c]), int(offsets[c + 1]))
            sc.append(orc.compute_similarity(q[b * K + i:b * K + i + 1], emb[rows])[0])
            ids.append(torch.from_numpy(docid)[rows])
        sc, ids = torch.cat(sc), torch.cat(ids)
        v, i = sc.topk(k)
        ref_s[b], ref_d[b] = v, ids[i]
    assert_topk_parity(s, d, ref_s, ref_d, "per-beam")


def test_full_size_cfg2_properties_and_oracle():
    """BASELINE.json configs[1] at full size: 109,739 x 768 bf16, C = 1,024, 1,024 queries, beam 20, top-100."""
    N, C, D, Q, K, k = 109739, 1024, 768, 1024, 20, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=1234)
    emb = emb.bfloat16().float()
    q, beams, _ = orc.synth_queries(Q, C, K, D, seed=4321)
    st = _store(emb, offsets, docid, torch.bfloat16)
    s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k)
    torch.cuda.synchronize()
    # size-independent properties
    assert torch.all(s[:, :-1] >= s[:, 1:]), "sorted descending"
    row_of = torch.empty(N, dtype=torch.int64); row_of[torch.from_numpy(docid)] = torch.arange(N)
    rows = row_of[d.cpu().long()]
    owner = torch.from_numpy(np.searchsorted(offsets, rows.numpy(), side="right") - 1)
    assert all(set(owner[b].tolist()) <= set(beams[b].tolist()) for b in range(Q)), "docids come from the beam clusters"
    recomputed = torch.einsum("bkd,bd->bk", emb[rows].double(), q.double())
    np.testing.assert_allclose(s.cpu().double().numpy(), recomputed.numpy(), rtol=1e-3, atol=1e-3)
    # and the oracle itself on a 128-query slice
    ref_s, ref_d = orc.dense_topk(q[:128], emb, offsets, docid, beams[:128], k)
    assert_topk_parity(s[:128], d[:128], ref_s, ref_d, "cfg2")
    # idempotence: same call, same bits
    s2, d2 = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k)
    assert torch.equal(s, s2) and torch.equal(d, d2)


def test_full_size_cfg1_fp32_store():
    """BASELINE.json configs[0] at its own shape: 109,739 x 768 fp32 store (what the reference holds, main_models.py:806-814),
    7,830 queries, beam 10, top-100 — in batches of 1,024 like the bench, oracle on a query slice of the first and the last
    (ragged) batch, size-independent properties on all of them."""
    N, C, D, Q, K, k = 109739, 1024, 768, 7830, 10, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=1234)
    q, beams, _ = orc.synth_queries(Q, C, K, D, seed=4322)
    st = _store(emb, offsets, docid, torch.float32)
    row_of = torch.empty(N, dtype=torch.int64); row_of[torch.from_numpy(docid)] = torch.arange(N)
    for lo in range(0, Q, 1024):
        hi = min(Q, lo + 1024)
        s, d = st.score_topk(q[lo:hi].cuda(), torch.from_numpy(beams[lo:hi]).cuda(), k)
        torch.cuda.synchronize()
        assert torch.all(s[:, :-1] >= s[:, 1:]), "sorted descending"
        rows = row_of[d.cpu().long()]
        recomputed = torch.einsum("bkd,bd->bk", emb[rows].double(), q[lo:hi].double())
        np.testing.assert_allclose(s.cpu().double().numpy(), recomputed.numpy(), rtol=1e-3, atol=1e-3)
        if lo == 0 or hi == Q:
            n = min(96, hi - lo)
            ref_s, ref_d = orc.dense_topk(q[hi - n:hi], emb, offsets, docid, beams[hi - n:hi], k)
            assert_topk_parity(s[-n:], d[-n:], ref_s, ref_d, f"cfg1 queries {hi - n}..{hi}")


def test_cfg3_shape_top1000():
    """BASELINE.json configs[2] at full batch: 73,970 x 768, 1,024 queries, beam 100, top-1000 (C = 1,024 assumed);
    properties on every query, the oracle on a 48-query slice."""
    N, C, D, Q, K, k = 73970, 1024, 768, 1024, 100, 1000
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=77)
    emb = emb.bfloat16().float()
    q, beams, _ = orc.synth_queries(Q, C, K, D, seed=78)
    st = _store(emb, offsets, docid, torch.bfloat16)
    s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k)
    torch.cuda.synchronize()
    assert torch.all(s[:, :-1] >= s[:, 1:]), "sorted descending"
    row_of = torch.empty(N, dtype=torch.int64); row_of[torch.from_numpy(docid)] = torch.arange(N)
    for lo in range(0, Q, 128):                       # recompute every returned score in fp64 (chunked: 1,024 x 1,000 x 768 doubles)
        rows = row_of[d[lo:lo + 128].cpu().long()]
        recomputed = torch.einsum("bkd,bd->bk", emb[rows].double(), q[lo:lo + 128].double())
        np.testing.assert_allclose(s[lo:lo + 128].cpu().double().numpy(), recomputed.numpy(), rtol=1e-3, atol=1e-3)
    ref_s, ref_d = orc.dense_topk(q[:48], emb, offsets, docid, beams[:48], k)
    assert_topk_parity(s[:48], d[:48], ref_s, ref_d, "cfg3")


def test_compute_similarity_drop_in():
    from gdr_b200 import DenseModel
    g = torch.Generator().manual_seed(5)
    # fp32 / small shapes: the GEMV; bf16 passages with dim % 64 == 0: the tcgen05 grouped GEMM (ragged query chunk, ragged last row tile)
    for Q, P, D, dt in ((7, 33, 96, torch.float32), (130, 1000, 768, torch.float32), (64, 513, 768, torch.bfloat16),
                        (1000, 3001, 768, torch.bfloat16), (33, 200, 128, torch.bfloat16), (8, 64, 64, torch.bfloat16)):
        q = torch.randn(Q, D, generator=g)
        p = (torch.randn(P, D, generator=g) * D ** -0.5).to(dt)
        out = DenseModel().compute_similarity(q.cuda(), p.cuda())
        ref = orc.compute_similarity(q, p.float())
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-3, atol=1e-4)


def test_merge_topk_kernel():
    from gdr_b200.sharded import ShardedRetriever, pack_candidates
    g = torch.Generator().manual_seed(6)
    G, B, k = 8, 50, 100
    s = torch.randn(G, B, k, generator=g).sort(dim=-1, descending=True).values
    d = torch.randint(0, 10 ** 8, (G, B, k), generator=g, dtype=torch.int64)
    s[3, :, 40:] = float("-inf"); d[3, :, 40:] = -1
    s[5] = float("-inf"); d[5] = -1
    packed = torch.stack([pack_candidates(s[r], d[r].int()) for r in range(G)]).cuda()
    ms, md = ShardedRetriever._cuda_merge(packed, k)
    ref_s, ref_d = orc.merge_topk(s, d, k)
    assert_topk_parity(ms, md, ref_s, ref_d, "merge")


def test_index_file_loads_into_an_identical_store(tmp_path):
    """load_store(file) == ClusterStore.from_reference: same results, whole corpus and a cluster shard."""
    from types import SimpleNamespace
    from gdr_b200 import ClusterStore
    from gdr_b200.index_io import load_store, write_index
    from gdr_b200.store import csr_from_reference
    f = fine_stage_inputs("fine_stage_tanh")
    emb, offsets, docid, keys = csr_from_reference(f["doc_embed"], f["id_mapping"])
    p = str(tmp_path / "idx.gdr")
    write_index(p, emb.bfloat16(), offsets.numpy(), docid.numpy(), keys)
    a = ClusterStore.from_reference(f["doc_embed"], f["id_mapping"], dtype=torch.bfloat16)
    b = load_store(p, chunk_rows=37)
    assert b.keys == a.keys and torch.equal(b.emb, a.emb) and torch.equal(b.docid, a.docid) and torch.equal(b.offsets, a.offsets)
    beams = a.beams_from_ids(f["dec"]).cuda()
    q = f["q"].cuda()
    sa, da = a.score_topk(q, beams, 5)
    sb, db = b.score_topk(q, beams, 5)
    assert torch.equal(sa, sb) and torch.equal(da, db)
    shard = load_store(p, clusters=np.array([1, 4, 7]), chunk_rows=16)
    assert shard.keys == [keys[1], keys[4], keys[7]] and shard.n_docs == int(sum(len(f["id_mapping"][keys[c]]) for c in (1, 4, 7)))


def test_index_expansion_matches_reference():
    """Centroid kernel + top-1 assignment (gdr_b200.expand) against the reference's tree_embedding_calculate /
    tree_embedding_insert run (tests/golden/expand.npz).  fp32 store: centroids bit-exact, assignments identical."""
    import json
    from gdr_b200 import ClusterStore
    from gdr_b200.expand import assign_to_clusters, tree_embedding_insert
    g = load_golden("expand")
    emb = torch.from_numpy(g["embedding"])
    docs = [emb[i] for i in range(emb.shape[0])]
    before = json.loads(str(g["before_json"]))
    after = json.loads(str(g["after_json"]))
    docnum = int(g["docnum"])
    store = ClusterStore.from_reference(docs[:docnum], before, dtype=torch.float32)
    cent = store.centroids()
    assert torch.equal(cent.cpu(), torch.from_numpy(g["centroids"])), "centroids must equal the reference's leaf embeddings"
    idx, score = assign_to_clusters(cent, emb[docnum:].cuda())
    ref = torch.from_numpy(g["centroids"]) @ emb[docnum:].T
    assert torch.equal(idx.cpu(), ref.argmax(0))
    out = tree_embedding_insert(store, {k: list(v) for k, v in before.items()}, docs, docnum)
    assert {k: sorted(v) for k, v in out.items()} == after


def test_many_clusters_multi_block_scan():
    """More than 8,192 clusters: the inversion's scan runs in several CTAs (block-local offsets + bases)."""
    N, C, D, Q, K, k = 60000, 20000, 64, 300, 50, 20
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=31)
    emb = emb.bfloat16().float()
    q, beams, _ = orc.synth_queries(Q, C, K, D, seed=32)
    beams[:, :3] = np.array([8191, 8192, 16383])[None, :]          # clusters on both sides of the block boundaries, shared by all queries
    beams[:, 3] = C - 1
    st = _store(emb, offsets, docid, torch.bfloat16)
    ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k)
    for flags in (FLAG_SIMT, FLAG_UMMA, 0):
        s, d = st.score_topk(q.cuda(), torch.from_numpy(beams).cuda(), k, flags=flags)
        assert_topk_parity(s, d, ref_s, ref_d, f"many clusters flags={flags}")


def test_phases_can_be_reissued_separately():
    """GDR_SKIP_* phase flags: scoring re-launched on an existing inversion (the tile queue resets itself), then top-k alone,
    must reproduce the fused call (bench.py times the scoring kernel this way)."""
    SK_I, SK_S, SK_T = 256, 512, 1024
    N, C, D, Q, K, k = 20000, 128, 768, 256, 20, 100
    emb, offsets, docid = orc.synth_corpus(N, C, D, seed=77)
    emb = emb.bfloat16().float()
    q, beams, _ = orc.synth_queries(Q, C, K, D, seed=78)
    st = _store(emb, offsets, docid, torch.bfloat16)
    qd, bd = q.cuda(), torch.from_numpy(beams).cuda()
    for flags in (0, FLAG_SIMT):
        ref_s, ref_d = st.score_topk(qd, bd, k, flags=flags)
        out = (torch.zeros((1, Q, k), device="cuda"), torch.zeros((1, Q, k), dtype=torch.int32, device="cuda"))
        st.score_topk(qd, bd, k, out=out, flags=flags | SK_S | SK_T)           # inversion only
        for _ in range(3):
            st.score_topk(qd, bd, k, out=out, flags=flags | SK_I | SK_T)       # scoring only, three times over
        st.score_topk(qd, bd, k, out=out, flags=flags | SK_I | SK_S)           # top-k only
        torch.cuda.synchronize()
        assert torch.equal(out[1][0], ref_d) and torch.equal(out[0][0], ref_s)


def test_node_embeddings_and_tree_match_match_reference():
    """gdr_trie_node_embeddings / gdr_tree_match against the reference's tree_embedding_calculate (all nodes) and tree_match
    (tests/golden/tree_match.npz).  fp32 store: node embeddings bit-identical; descents identical (a near-tie between two
    children — sims within 1e-5 — would be excused and counted; the fixture has none)."""
    import json
    from types import SimpleNamespace
    from gdr_b200 import ClusterStore, DeviceTrie, TreeBuilder
    from gdr_b200.expand import node_embeddings, tree_match
    from gdr_b200.main_models import encode_single_newid
    f = load_golden("tree_match")
    args = SimpleNamespace(kary=30, position=1, output_vocab_size=30)
    emb = torch.from_numpy(f["embedding"])
    docs = [emb[i] for i in range(emb.shape[0])]
    id_map = json.loads(str(f["id_map_json"]))
    order = [str(x) for x in f["cluster_order"]]
    id_map = {k: id_map[k] for k in order}                              # the reference's insertion order
    tb = TreeBuilder()
    for key in order:
        for d in id_map[key]:
            tb.add(encode_single_newid(args, key), d)
    trie = DeviceTrie.from_root(tb.build())
    store = ClusterStore.from_reference(docs, id_map, dtype=torch.float32)
    node_emb, node_leaf = node_embeddings(trie, store, args)
    paths = json.loads(str(f["node_paths_json"]))
    ne, nl = node_emb.cpu(), node_leaf.cpu()
    seen = set()
    for path, want, leaves in zip(paths, f["node_emb"], f["node_leaves"]):
        n = trie.find(path)
        seen.add(n)
        assert n >= 0 and int(nl[n]) == int(leaves)
        assert torch.equal(ne[n], torch.from_numpy(want)), f"node {path}: embedding differs from the reference"
    assert all(int(nl[n]) == 0 for n in range(trie.n_nodes) if n not in seen)      # EOS children carry no embedding
    got = tree_match(trie, node_emb, node_leaf, torch.from_numpy(f["new_docs"]).cuda())
    want = json.loads(str(f["matches_json"]))
    assert [g.tolist() for g in got] == want


def test_contrastive_loss_and_gradient_match_reference():
    """gdr_contrastive_loss against the reference's encoder_cal + autograd (tests/golden/contrastive.npz): loss within 1e-5
    relative, d loss / d query within 1e-4 of its largest entry (expf / tanhf vs. the CPU libm)."""
    from gdr_b200 import ClusterStore
    from gdr_b200.contrastive import encoder_cal
    f = load_golden("contrastive")
    doc = torch.from_numpy(f["doc_embed"])
    N = doc.shape[0]
    id_map = {str(c): list(range(c * 30, (c + 1) * 30)) for c in range(N // 30)}
    store = ClusterStore.from_reference([doc[i] for i in range(N)], id_map, dtype=torch.float32)
    for c in range(int(f["n_cases"])):
        valid = f[f"c{c}_valid_num"].tolist()
        cand = f[f"c{c}_cand"].tolist()
        cands, at = [], 0
        for v in valid:
            cands.append(cand[at:at + v])
            at += v
        q = torch.from_numpy(f[f"c{c}_q"]).cuda().requires_grad_(True)
        loss = encoder_cal(store, q, f[f"c{c}_pos"].tolist(), cands, str(f[f"c{c}_loss_func"]), float(f[f"c{c}_tau"]), float(f[f"c{c}_intra_rate"]))
        loss.backward()
        want = float(f[f"c{c}_loss"])
        assert abs(loss.item() - want) <= 1e-5 * abs(want), (c, loss.item(), want)
        gw = torch.from_numpy(f[f"c{c}_grad_q"])
        assert (q.grad.cpu() - gw).abs().max().item() <= 1e-4 * gw.abs().max().item(), c
    with pytest.raises(KeyError):
        encoder_cal(store, q, [N + 5] * q.shape[0], cands)
