"""Host-side mirror of the reference API (gdr_b200.main_models / generation / store / sharded) on CPU."""
import os
import pickle
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import gdr_oracle as orc
from helpers import fine_stage_inputs, load_golden, rebuild_tree

import gdr_b200
from gdr_b200 import _cabi
from gdr_b200.generation import flatten_trie
from gdr_b200.main_models import Node, TreeBuilder, dec_2d, decode_token, encode_query, encode_single_newid
from gdr_b200.sharded import global_to_local, pack_candidates, partition_clusters
from gdr_b200.store import csr_from_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = SimpleNamespace(kary=30, position=1, output_vocab_size=30)


def test_codecs_match_reference_fixture():
    g = load_golden("tree")
    paths = [str(p) for p in g["paths"]]
    assert encode_single_newid(ARGS, "3-17-22") == [5, 49, 84, 1]
    assert decode_token(ARGS, g["tok_arr"]) == paths
    assert decode_token(ARGS, g["noeos"]) == [str(x) for x in g["decoded_noeos"]]
    for p in paths[:50]:
        assert encode_single_newid(ARGS, p) == orc.encode_single_newid(p)
    nk = SimpleNamespace(kary=0, position=1, output_vocab_size=10)
    assert encode_single_newid(nk, "305") == orc.encode_single_newid("305", kary=0) == [5, 12, 27, 1]
    assert dec_2d(list(range(6)), 2) == [[0, 1], [2, 3], [4, 5]]
    h = torch.randn(3, 4, 8)
    assert torch.equal(encode_query(h), h[:, 0]) and encode_query(None) is None
    assert torch.equal(gdr_b200.EncoderModel()(query_enc=h), h[:, 0])


def test_tree_builder_matches_reference_fixture_and_pickles():
    g = load_golden("tree")
    paths = [str(p) for p in g["paths"]]
    b = TreeBuilder()
    for di, p in enumerate(paths):
        toks = encode_single_newid(ARGS, p)
        b.add(toks, di)
        if di % 3 == 0:
            b.add(toks + [0, 0], 1000 + di)
    ref_root = rebuild_tree(g["edges"], g["leaves"], Node)

    def same(a, c):
        assert list(a.children.keys()) == list(c.children.keys()) and a.embedding_index == c.embedding_index
        for t in a.children:
            same(a.children[t], c.children[t])
    root = b.build()
    same(root, ref_root)
    same(pickle.loads(pickle.dumps(root)), ref_root)


def test_flatten_trie_walk_equals_dict_walk():
    g = load_golden("tree")
    root = rebuild_tree(g["edges"], g["leaves"], Node)
    fc, tok, child = flatten_trie(root)
    assert fc[0] == 0 and fc[-1] == tok.size == child.size
    for n in range(fc.size - 1):
        seg = tok[fc[n]:fc[n + 1]]
        assert np.all(np.diff(seg) > 0)
    for cur_len in (1, 3, 5):
        for row in g[f"ids_{cur_len}"]:
            node = 0
            for t in row[1:]:
                seg = tok[fc[node]:fc[node + 1]]
                hit = np.nonzero(seg == t)[0]
                node = child[fc[node] + hit[0]] if hit.size else -1
                if node < 0:
                    break
            allowed = [1] if node < 0 else tok[fc[node]:fc[node + 1]].tolist()
            assert sorted(allowed) == sorted(orc.tree_mask_allowed(root, row.tolist()))


def test_csr_from_reference_layout():
    f = fine_stage_inputs("fine_stage_tanh")
    emb, offsets, docid, keys = csr_from_reference(f["doc_embed"], f["id_mapping"])
    assert keys == list(f["id_mapping"].keys())
    for c, key in enumerate(keys):
        lo, hi = int(offsets[c]), int(offsets[c + 1])
        assert docid[lo:hi].tolist() == f["id_mapping"][key]
        assert torch.equal(emb[lo:hi], f["emb"][docid[lo:hi]])
    emb2, *_ = csr_from_reference(f["emb"], f["id_mapping"])
    assert torch.equal(emb, emb2)


def test_partition_is_balanced_and_complete():
    rng = np.random.RandomState(1)
    sizes = rng.randint(1, 400, 999)
    for G in (1, 2, 4, 8):
        owner = partition_clusters(sizes, G)
        loads = [int(sizes[owner == r].sum()) for r in range(G)]
        assert sum(loads) == int(sizes.sum()) and max(loads) - min(loads) <= 400
        seen = np.zeros(sizes.size, dtype=int)
        for r in range(G):
            g2l, mine = global_to_local(owner, r)
            seen[mine] += 1
            assert np.array_equal(g2l[mine], np.arange(mine.size)) and np.all(g2l[owner != r] == -1)
        assert np.all(seen == 1)


def test_pack_candidates_roundtrip():
    s = torch.randn(4, 7)
    d = torch.randint(0, 1000, (4, 7), dtype=torch.int32)
    p = pack_candidates(s, d)
    assert torch.equal(p[0].view(torch.float32), s) and torch.equal(p[1], d)


def test_cabi_library_loads_and_exports_every_declared_symbol():
    """No compute calls here (no GPU): the .so must load and export what include/gdr_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "gdr_b200.h")).read()
    declared = set(re.findall(r"\b(gdr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gdr_abi_version() == 1
    assert lib.gdr_last_error() is not None


def test_cabi_argument_checks_need_no_gpu():
    """Bad arguments are refused with a status code and a message before any CUDA call (include/gdr_b200.h: errors are integer
    status codes, never aborts): the SM partition, the handle options and the batch call on a null handle."""
    import ctypes
    lib = _cabi.lib()
    h = ctypes.c_void_p()
    assert lib.gdr_partition_create(ctypes.byref(h), 4, 1, 1) == -1 and b"small_sms" in lib.gdr_last_error()
    assert lib.gdr_partition_create(ctypes.byref(h), 56, 0, 5) == -1 and h.value is None
    assert lib.gdr_partition_create(None, 56, 2, 5) == -1
    assert lib.gdr_partition_stream(None, 0, 0) is None and lib.gdr_partition_destroy(None) == 0
    assert lib.gdr_store_set_option(None, _cabi.OPTIONS["umma_ctas_per_sm"], 2) == -1
    assert lib.gdr_score_topk(None, None, None, None, None, 1, 1, 1, 0, 1, 0, None, None, None) == -1 and b"null" in lib.gdr_last_error()


def test_no_cpu_fallback_in_product_path():
    """The product refuses CPU tensors instead of silently computing elsewhere."""
    with pytest.raises(ValueError):
        gdr_b200.compute_similarity(torch.randn(2, 8), torch.randn(3, 8))
    with pytest.raises(ValueError):
        gdr_b200.ClusterStore(torch.randn(4, 8), torch.tensor([0, 4]), torch.arange(4))
    with pytest.raises(ValueError):
        gdr_b200.position_mask_(torch.zeros(1, 2, 70), 30)
    # and nothing under gdr_b200/ imports the oracle
    pkg = os.path.join(ROOT, "gdr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "gdr_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


def test_index_file_roundtrip_and_pickle_converter(tmp_path):
    """On-disk CSR format: write -> mmap round trip, and the converter from the reference's pickles."""
    import pickle
    from gdr_b200.index_io import convert_pickles, map_index, read_header, write_index
    f = fine_stage_inputs("fine_stage_tanh")
    emb, offsets, docid, keys = csr_from_reference(f["doc_embed"], f["id_mapping"])
    for dt in (torch.float32, torch.bfloat16):
        p = str(tmp_path / f"idx_{dt}.gdr")
        write_index(p, emb.to(dt), offsets.numpy(), docid.numpy(), keys)
        h = read_header(p)
        assert h["n_rows"] == emb.shape[0] and h["dim"] == emb.shape[1] and h["n_clusters"] == len(keys)
        m_emb, m_off, m_doc, m_keys, m_dt = map_index(p)
        assert m_dt == dt and m_keys == keys
        assert np.array_equal(m_off, offsets.numpy()) and np.array_equal(m_doc, docid.numpy())
        back = torch.from_numpy(np.array(m_emb))
        back = back.view(torch.int16).view(torch.bfloat16) if dt == torch.bfloat16 else back
        assert torch.equal(back, emb.to(dt))
    pe, pm = str(tmp_path / "doc_embedding.pkl"), str(tmp_path / "indexmap.pkl")
    pickle.dump(f["doc_embed"], open(pe, "wb"))
    pickle.dump(f["id_mapping"], open(pm, "wb"))
    h = convert_pickles(pe, pm, str(tmp_path / "conv.gdr"), dtype=torch.float32)
    assert h["n_clusters"] == len(f["id_mapping"]) and h["n_rows"] == sum(len(v) for v in f["id_mapping"].values())
    m_emb, m_off, m_doc, m_keys, _ = map_index(str(tmp_path / "conv.gdr"))
    assert m_keys == list(f["id_mapping"].keys()) and torch.equal(torch.from_numpy(np.array(m_emb)), emb)
    with pytest.raises(ValueError):
        read_header(pe)
    # the command-line converter a maintainer runs once per corpus (`python -m gdr_b200.index_io convert ...` / `info ...`)
    import json
    import subprocess
    import sys
    cli = str(tmp_path / "cli.gdr")
    out = subprocess.run([sys.executable, "-m", "gdr_b200.index_io", "convert", pe, pm, cli, "--dtype", "f32"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    assert json.loads(out.stdout.strip().splitlines()[-1])["n_rows"] == h["n_rows"]
    assert open(cli, "rb").read() == open(str(tmp_path / "conv.gdr"), "rb").read()
    out = subprocess.run([sys.executable, "-m", "gdr_b200.index_io", "info", cli], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and json.loads(out.stdout.strip().splitlines()[-1])["n_clusters"] == h["n_clusters"]


def test_child_insertion_order_matches_the_dict_order():
    """flatten_trie sorts every node's edges by token (the mask kernels binary-search them); child_insertion_order keeps the
    order the children were inserted in, which the reference's tree_embedding_calculate / tree_match iterate in."""
    from gdr_b200.generation import child_insertion_order, flatten_trie
    from gdr_b200.main_models import TreeBuilder
    tb = TreeBuilder()
    for i, seq in enumerate([[9, 40, 1], [3, 35, 1], [9, 33, 1], [5, 61, 1], [3, 34, 1]]):
        tb.add(seq, i)
    root = tb.build()
    fc, tok, node = flatten_trie(root)
    order = child_insertion_order(root)
    assert sorted(order.tolist()) == list(range(len(tok)))
    assert [int(tok[e]) for e in order[fc[0]:fc[1]]] == [9, 3, 5]            # root: inserted 9, 3, 5; stored 3, 5, 9
    n9 = int(node[fc[0]:fc[1]][list(tok[fc[0]:fc[1]]).index(9)])
    assert [int(tok[e]) for e in order[fc[n9]:fc[n9 + 1]]] == [40, 33]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host alone (it is the CPU arm the driver times beside ours): one JSON line
    with the contract's keys, the same metric / unit / config.workload as the GPU arm, and a zero-copy e2e object."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["value"] > 0 and line["steps"] == 2
    assert line["config"]["workload"] == "cfg2" and line["vs_baseline"] is None and line["higher_is_better"] is True
    # the unmodified reference dense.py when the tree (or its shipped copy baseline/_ref/) is present, else the oracle port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert line["metric"] == bench.METRIC            # ONE metric string for both arms (the driver refuses to divide otherwise)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_launch_autotune_decision(monkeypatch):
    """bench.py's launch autotune (child processes verify and time the device-resident loop under candidate schedules): a
    candidate replaces the `batches` default only when it is > 3 % faster, a failed, hung or garbled child leaves the default,
    and a schedule whose child reports differing results is never chosen."""
    import argparse
    import json
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    import bench

    args = argparse.Namespace(workload="cfg2", path="auto", schedule="auto", replicas=0)
    names = ("batches", "partitioned_56x2", "batches_priorities", "partitioned_48x2", "partitioned_64x2", "fused_140", "partitioned_48", "fused64_140")
    expect = {"fused_140": ("fused", "5", "140"), "fused64_140": ("fused", "9", "140")}
    expect_part = {"partitioned_56x2": ("56", "2"), "partitioned_48x2": ("48", "2"), "partitioned_64x2": ("64", "2"), "partitioned_48": ("48", "1")}
    seen = []

    def fake_run(times, fail=(), bad_line=None):
        def run(cmd, env=None, capture_output=None, text=None, timeout=None):
            name = names[len(seen) % len(names)]
            seen.append((cmd, env))
            assert "--probe" in cmd and "RANK" not in env and env["LOCAL_RANK"] == "2"
            sched = cmd[cmd.index("--schedule") + 1]
            if name in expect:
                assert (sched, cmd[cmd.index("--fused-groups") + 1], cmd[cmd.index("--fused-ctas") + 1]) == expect[name]
            elif name in expect_part:
                assert (sched, cmd[cmd.index("--small-sms") + 1], cmd[cmd.index("--ctas-per-sm") + 1]) == ("partitioned",) + expect_part[name]
            else:
                assert sched == "batches" and ("--launch-priorities" in cmd) == (name == "batches_priorities")
            if name in fail:
                if fail[name] == "timeout":
                    raise subprocess.TimeoutExpired(cmd, timeout)
                return subprocess.CompletedProcess(cmd, 1, stdout="", stderr="CUDA error: invalid value")
            line = {"probe": True, "us_per_step": times.get(name, 99.0), "schedule": sched}
            if (name in expect or name in expect_part) and bad_line is not None:
                line = bad_line
            return subprocess.CompletedProcess(cmd, 0, stdout="noise\n" + json.dumps(line) + "\n", stderr="")
        return run

    def tune(times, **kw):
        assert len(seen) % len(names) == 0
        monkeypatch.setattr(bench.subprocess, "run", fake_run(times, **kw))
        return bench.autotune(args, 2)

    monkeypatch.setenv("RANK", "0")
    best, rep = tune({"batches": 50.0, "fused_140": 37.0, "fused64_140": 47.0})
    assert (best["schedule"], best["fused_groups"], best["fused_ctas"], rep["chosen"]) == ("fused", 5, 140, "fused_140")
    assert rep["fused_140"]["verified_identical_to_serial"] and rep["fused64_140"]["us_per_step"] == 47.0
    best, rep = tune({"batches": 50.0, "fused_140": 49.5, "fused64_140": 60.0, "batches_priorities": 49.0})
    assert best["schedule"] == "batches" and rep["chosen"] == "batches"                      # within 3 %: the default stays
    best, rep = tune({"batches": 50.0, "batches_priorities": 40.0})
    assert best.get("launch_priorities") == "on" and rep["chosen"] == "batches_priorities"
    best, rep = tune({"batches": 50.0, "partitioned_56x2": 39.0, "partitioned_48": 41.0})
    assert (best["schedule"], best["small_sms"], best["ctas_per_sm"], rep["chosen"]) == ("partitioned", 56, 2, "partitioned_56x2")
    best, rep = tune({"batches": 50.0}, fail={"fused_140": "rc", "fused64_140": "timeout", "batches_priorities": "timeout",
                                             "partitioned_56x2": "rc", "partitioned_48x2": "timeout", "partitioned_64x2": "rc", "partitioned_48": "rc"})
    assert best["schedule"] == "batches" and all("failed" in rep[n] for n in names[1:])
    # a candidate whose child found differing results (or printed no time) is never chosen
    best, rep = tune({"batches": 50.0}, bad_line={"probe": True, "us_per_step": None, "failed": "RuntimeError: batch 3 differs from gdr_score_topk"})
    assert best["schedule"] == "batches" and "differs" in rep["fused_140"]["failed"]
    # other workloads do not try the fused schedule at all
    seen.clear()
    args3 = argparse.Namespace(workload="cfg3", path="auto", schedule="auto", replicas=0)

    def run3(cmd, env=None, capture_output=None, text=None, timeout=None):
        seen.append(cmd)
        return subprocess.CompletedProcess(cmd, 0, stdout=json.dumps({"probe": True, "us_per_step": 100.0, "schedule": "batches"}), stderr="")
    monkeypatch.setattr(bench.subprocess, "run", run3)
    bench.autotune(args3, 2)
    assert len(seen) == 2 and all("fused" not in c for c in seen)
