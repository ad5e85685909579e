#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the dev container.

TEST INFRASTRUCTURE.  Run here (needs /root/reference):   python oracle/make_golden.py
The fixtures it writes are committed; nothing on the GPU box reads /root/reference.

Each fixture holds the inputs and the reference's outputs for one piece of the hot path:
  dense_topk_*.npz   reference dense.py `DenseModel.compute_similarity` (dense.py:53-54) + `Tensor.topk`
                     per query over its beam clusters (BASELINE.md §2 recipe)
  fine_stage_*.npz   reference `T5FineTuner.validation_step_i` (main_models.py:1337-1642) driven unbound
                     with a stub `self` (fake generate/tokenizer/encoder), docid strings per alpha
  tree_*.npz         reference `TreeBuilder/Node` (main_models.py:112-151), codecs (297-346) and the live
                     tree-mask block, whose SOURCE LINES generation_utils_previous.py:714-729 are read
                     from the mounted reference at run time and exec'd (the block is inline in a
                     300-line method and cannot be called on its own)
  tree_match.npz     reference `tree_embedding_calculate` over all nodes (main_models.py:154-179) and the greedy descent
                     `tree_match` (main_models.py:232-252); generated on its own: python oracle/make_golden.py tree_match
  contrastive.npz    reference `encoder_cal` (main_models.py:1184-1221, exec'd source lines) + autograd d loss / d query;
                     generated on its own: python oracle/make_golden.py contrastive
  position_mask.npz  reference `select_valid_embedding` (modeling_t5.py:1546-1571), same technique,
                     and the training `logit_mask` recipe (modeling_t5.py:1279-1301)
"""
import json
import os
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)


def _bf16_bits(t):
    import torch
    return t.to(torch.bfloat16).view(torch.int16).numpy().astype("uint16")


# ---------------------------------------------------------------------------------------------
def gen_dense():
    import numpy as np
    import torch
    import ref_shims
    import gdr_oracle as orc

    ref_dense = ref_shims.load_ref_dense()
    cases = {
        # name: (N, C, D, Q, K, k, zipf)
        "dense_topk_d768": (512, 8, 768, 6, 3, 20, 0.0),
        "dense_topk_d128": (4096, 64, 128, 32, 8, 100, 0.0),
        "dense_topk_zipf": (3000, 40, 64, 16, 6, 64, 1.1),
    }
    for name, (N, C, D, Q, K, k, zipf) in cases.items():
        emb, offsets, docid = orc.synth_corpus(N, C, D, seed=11 + N, zipf=zipf)
        emb = emb.bfloat16().float()              # bf16-representable so one fixture serves both store dtypes
        q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=7 + Q)
        beams = beams.copy()
        beams[0, -1] = -1                         # an absent beam
        bias = torch.softmax(beam_scores, dim=-1)
        out = {}
        for tag, act, use_bias in (("plain", None, False), ("tanh_bias", torch.tanh, True)):
            S = torch.full((Q, k), float("-inf"))
            I = torch.full((Q, k), -1, dtype=torch.int64)
            for b in range(Q):
                rows, sb = [], []
                for i, c in enumerate(beams[b].tolist()):
                    if c < 0:
                        continue
                    r = torch.arange(int(offsets[c]), int(offsets[c + 1]))
                    rows.append(r)
                    sb.append(bias[b, i].expand(r.numel()))
                rows = torch.cat(rows)
                s = ref_dense.DenseModel.compute_similarity(None, q[b:b + 1], emb[rows])[0]   # dense.py:53-54
                if act is not None:
                    s = act(s)
                if use_bias:
                    s = s + torch.cat(sb)
                kk = min(k, s.numel())
                v, i = s.topk(kk, largest=True, sorted=True)                                  # main_models.py:1625
                S[b, :kk] = v
                I[b, :kk] = torch.from_numpy(docid)[rows[i]]
            out["scores_" + tag] = S.numpy()
            out["docids_" + tag] = I.numpy()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), emb_bf16_bits=_bf16_bits(emb), offsets=offsets,
                            docid=docid, q=q.numpy(), beams=beams, bias=bias.numpy(), k=np.int64(k), **out)
        print("wrote", name)


# ---------------------------------------------------------------------------------------------
def _ref_lines(path, first, last):
    """Source lines [first, last] (1-based, inclusive) of a reference file, dedented."""
    with open(path) as f:
        lines = f.read().splitlines()
    return textwrap.dedent("\n".join(lines[first - 1:last])) + "\n"


def gen_main_models():
    import numpy as np
    import ref_shims

    mm, gp = ref_shims.load_ref_main_models()
    import torch
    from types import SimpleNamespace
    import gdr_oracle as orc

    REF = ref_shims.REF_MODEL_DIR

    # ---------------- tree + codecs + tree mask ----------------
    rng = np.random.RandomState(5)
    args = SimpleNamespace(kary=30, position=1, output_vocab_size=30)
    paths = set()
    while len(paths) < 200:
        depth = rng.choice([2, 3, 3, 3])
        paths.add("-".join(str(rng.randint(0, 30)) for _ in range(depth)))
    paths = sorted(paths)
    builder = mm.TreeBuilder()
    tok_paths = []
    for di, p in enumerate(paths):
        toks = mm.encode_single_newid(args, p)                       # main_models.py:297-319
        tok_paths.append(toks)
        builder.add(toks, di)                                         # main_models.py:137-151
        if di % 3 == 0:
            builder.add(toks + [0, 0], 1000 + di)                     # trailing pads are ignored
    root = builder.build()
    L = 6
    arr = np.zeros((len(paths), L), dtype=np.int64)
    for i, t in enumerate(tok_paths):
        arr[i, 1:1 + len(t)] = t
    decoded = mm.decode_token(args, arr)                              # main_models.py:322-346
    assert decoded == paths, "reference codec round trip failed"
    noeos = np.array([[0, 5, 40, 70, 0, 0], [0, 9, 33, 62, 95, 0]], dtype=np.int64)
    decoded_noeos = mm.decode_token(args, noeos)

    # flatten the reference trie for the fixture: (parent, token, child) edges in insertion order + leaf docs
    edges, leaves, ids, stack = [], [], {id(root): 0}, [root]
    while stack:
        n = stack.pop()
        for tok, ch in n.children.items():
            ids[id(ch)] = len(ids)
            edges.append((ids[id(n)], tok, ids[id(ch)]))
            stack.append(ch)
        for d in n.embedding_index:
            leaves.append((ids[id(n)], d))

    block = _ref_lines(os.path.join(REF, "transformers", "generation_utils_previous.py"), 714, 729)
    assert block.startswith("if decode_tree:"), block[:40]
    V = 128
    out = {}
    for cur_len in (1, 2, 3, 4, 5):
        R = 48
        ids_t = torch.zeros(R, cur_len, dtype=torch.int64)
        for r in range(R):
            t = tok_paths[rng.randint(len(tok_paths))]
            n = min(cur_len - 1, len(t))
            ids_t[r, 1:1 + n] = torch.tensor(t[:n])
            if r % 7 == 3 and cur_len > 1:
                ids_t[r, rng.randint(1, cur_len)] = 99 + rng.randint(20)   # off-tree prefix
            if r % 11 == 5 and cur_len > 2:
                ids_t[r, cur_len - 1] = 0                                   # finished (padded) row
        g = torch.Generator().manual_seed(100 + cur_len)
        scores = torch.log_softmax(torch.randn(R, V, generator=g), dim=-1)
        scores[0, 3] = -0.0
        scores[1, 7] = float("-inf")
        ns = {"torch": torch, "decode_tree": root, "scores": scores.clone(), "input_ids": ids_t,
              "num_beams": 4, "batch_size": R // 4}
        exec(block, ns)                                                # generation_utils_previous.py:714-729
        out[f"ids_{cur_len}"] = ids_t.numpy()
        out[f"in_{cur_len}"] = scores.numpy()
        out[f"out_{cur_len}"] = ns["scores"].numpy()
    np.savez_compressed(os.path.join(GOLD, "tree.npz"), paths=np.array(paths), tok_arr=arr,
                        decoded_noeos=np.array(decoded_noeos), noeos=noeos,
                        edges=np.array(edges, dtype=np.int64), leaves=np.array(leaves, dtype=np.int64), **out)
    print("wrote tree")

    # ---------------- fused beam step: log_softmax (:694) -> mask block (:714-729) -> + beam scores, view, topk(2K) (:757-771) -----
    import torch.nn.functional as F
    bs_out = {}
    for case, (Bq, Kb, cur_len) in {"a": (3, 4, 3), "b": (2, 6, 1), "c": (4, 5, 4)}.items():
        R = Bq * Kb
        ids_t = torch.zeros(R, cur_len, dtype=torch.int64)
        for r in range(R):
            t = tok_paths[rng.randint(len(tok_paths))]
            n = min(cur_len - 1, len(t))
            ids_t[r, 1:1 + n] = torch.tensor(t[:n])
            if r % 5 == 2 and cur_len > 1:
                ids_t[r, rng.randint(1, cur_len)] = 99 + rng.randint(20)
        g = torch.Generator().manual_seed(500 + R)
        logits = torch.randn(R, V, generator=g) * 3
        beam_scores = -torch.rand(R, generator=g) * 5
        if cur_len == 1:
            beam_scores = beam_scores.view(Bq, Kb)
            beam_scores[:, 1:] = -1e9                                   # generation_utils_previous.py:668
            beam_scores = beam_scores.reshape(-1)
        scores = F.log_softmax(logits, dim=-1)                          # :694
        ns = {"torch": torch, "decode_tree": root, "scores": scores, "input_ids": ids_t, "num_beams": Kb, "batch_size": Bq}
        exec(block, ns)                                                 # :714-729
        next_scores = ns["scores"] + beam_scores[:, None].expand_as(ns["scores"])      # :757
        next_scores = next_scores.view(Bq, Kb * V)                      # :760-762
        top_s, top_t = torch.topk(next_scores, 2 * Kb, dim=1, largest=True, sorted=True)   # :771
        bs_out.update({f"logits_{case}": logits.numpy(), f"ids_{case}": ids_t.numpy(), f"beam_{case}": beam_scores.numpy(),
                       f"K_{case}": np.int64(Kb), f"scores_{case}": top_s.numpy(), f"tokens_{case}": top_t.numpy()})
    np.savez_compressed(os.path.join(GOLD, "beam_step.npz"), edges=np.array(edges, dtype=np.int64),
                        leaves=np.array(leaves, dtype=np.int64), **bs_out)
    print("wrote beam_step")

    # ---------------- index expansion: tree_embedding_calculate (:154-179) + tree_embedding_insert (:268-295) ----------------
    g = torch.Generator().manual_seed(321)
    n_cl, per, D_e, n_new = 10, 6, 32, 15
    cl_tok_paths = [t for t in tok_paths if len(t) == 4][:n_cl]         # token paths of 10 three-level clusters ([t0, t1, t2, 1])
    exp_builder = mm.TreeBuilder()
    docnum = n_cl * per
    embedding = [torch.randn(D_e, generator=g) for _ in range(docnum + n_new)]
    id_map = {}
    perm = torch.randperm(docnum, generator=g).tolist()
    for c, toks in enumerate(cl_tok_paths):
        key = mm.decode_token(args, [np.array([0] + toks)])[0]
        id_map[key] = perm[c * per:(c + 1) * per][: per - (c % 2)]
        for d in id_map[key]:
            exp_builder.add(toks, d)
    exp_root = exp_builder.build()
    mm.tree_embedding_calculate(exp_root, embedding)                    # main_models.py:154-179
    cluster_set = set("-".join(str(t) for t in toks[:-1]) for toks in cl_tok_paths)
    centroids = []
    for toks in cl_tok_paths:
        cur = exp_root
        for t in toks[:-1]:
            cur = cur.children[t]
        centroids.append(cur.embedding)
    before = {k: list(v) for k, v in id_map.items()}
    after = mm.tree_embedding_insert(exp_root, id_map, embedding, cluster_set, SimpleNamespace(docnum=docnum, **vars(args)))   # :268-295
    np.savez_compressed(os.path.join(GOLD, "expand.npz"), embedding=torch.stack(embedding).numpy(), docnum=np.int64(docnum),
                        before_json=np.array(json.dumps(before)), after_json=np.array(json.dumps({k: sorted(v) for k, v in after.items()})),
                        centroids=torch.stack(centroids).numpy())
    print("wrote expand")

    # ---------------- positional mask ----------------
    sel_src = _ref_lines(os.path.join(REF, "transformers", "modeling_t5.py"), 1546, 1571)
    assert sel_src.startswith("def select_valid_embedding(sequence):"), sel_src[:60]
    pm = {}
    for V_out, Lmax, sl in ((30, 10, 1), (30, 10, 4), (30, 10, 10), (10, 5, 5)):
        Vdec = V_out * Lmax + 2
        g = torch.Generator().manual_seed(V_out + sl)
        x = torch.randn(3, sl, Vdec, generator=g) * 4
        x[0, 0, 1] = -0.0
        ns = {"torch": torch, "self": SimpleNamespace(output_vocab_size=V_out)}
        exec(sel_src, ns)
        y = ns["select_valid_embedding"](x.clone())                    # modeling_t5.py:1546-1571
        key = f"{V_out}_{Lmax}_{sl}"
        pm["in_" + key], pm["out_" + key] = x.numpy(), y.numpy()
    # training-time buffer (modeling_t5.py:1279-1301): executed from source with a stub config/self
    tr_src = _ref_lines(os.path.join(REF, "transformers", "modeling_t5.py"), 1279, 1301)
    assert tr_src.startswith("if decode_embedding:"), tr_src[:40]
    stub = SimpleNamespace()
    ns = {"torch": torch, "decode_embedding": 2, "self": stub,
          "config": SimpleNamespace(max_output_length=10, decode_vocab_size=302, output_vocab_size=30)}
    exec(tr_src, ns)
    pm["train_logit_mask_30_10"] = stub.logit_mask.numpy()
    np.savez_compressed(os.path.join(GOLD, "position_mask.npz"), **pm)
    print("wrote position_mask")

    # ---------------- fine stage through the real validation_step_i ----------------
    for case, (loss_func, D, C, per, B, K) in {"fine_stage_tanh": ("tanh", 64, 12, 20, 3, 5),
                                               "fine_stage_sigmoid": ("sigmoid", 32, 9, 14, 2, 4)}.items():
        g = torch.Generator().manual_seed(77 + D)
        N = C * per
        emb = torch.randn(N, D, generator=g) * D ** -0.5
        doc_embed = [emb[i].clone() for i in range(N)]
        cl_paths = paths[:C]
        perm = torch.randperm(N, generator=g).tolist()
        id_mapping = {cl_paths[c]: perm[c * per:(c + 1) * per][: per - (c % 3)] for c in range(C)}
        beams = [torch.randperm(C, generator=g)[:K].tolist() for _ in range(B)]
        dec_flat = [cl_paths[c] for row in beams for c in row]
        outs = torch.zeros(B * K, L, dtype=torch.int64)
        for i, p in enumerate(dec_flat):
            t = mm.encode_single_newid(args, p)
            outs[i, 1:1 + len(t)] = torch.tensor(t)
        beam_scores = (-torch.cumsum(torch.rand(B, K, generator=g), dim=1)).flatten().tolist()
        q = torch.randn(B, D, generator=g)
        enc_hidden = torch.zeros(B * K, 4, D)
        enc_hidden[::K, 0] = q                                          # encoder token-0 state of each query's first beam row
        score_rate = [0, 0.5, 1, 3]
        a = SimpleNamespace(decode_embedding=2, position=1, max_output_length=L, hierarchic_decode=0,
                            output_vocab_size=30, softmax=0, gen_method="greedy", is_train_encoder=1,
                            multiple_decoder=0, num_return_sequences=K, length_penalty=0.8, kary=30,
                            label_length_cutoff=0, train_encoder_epoch=10 ** 9, use_query_embed_encoder=1,
                            use_query_embed_decoder_avg=0, use_query_embed_decoder_special=0,
                            loss_func=loss_func, score_rate=score_rate, eval_batch_size=B)
        model = SimpleNamespace(
            generate=lambda *x, **kw: ((outs, list(beam_scores)), SimpleNamespace(last_hidden_state=enc_hidden)),
            config=SimpleNamespace(hidden_size=D))
        stub_self = SimpleNamespace(args=a, epoch=0, model=model, root=None, cluster=set(cl_paths),
                                    tokenizer=SimpleNamespace(decode=lambda ids: "q"),
                                    id_mapping=id_mapping, doc_embed=doc_embed,
                                    encoder=lambda query_enc=None, passage=None: query_enc[:, 0],   # main_models.py:102-109
                                    softmax=torch.nn.Softmax(dim=-1))
        batch = {"source_ids": torch.zeros(B, 3, dtype=torch.int64), "source_mask": torch.ones(B, 3, dtype=torch.int64),
                 "target_mask": torch.ones(B, L, dtype=torch.int64), "rank": [], "oldid": [["gt"] * B]}
        res = mm.T5FineTuner.validation_step_i(stub_self, batch, -1)      # main_models.py:1337-1642
        docids = np.zeros((B, len(score_rate), K), dtype=np.int64)
        for b in range(B):
            for r in range(len(score_rate)):
                docids[b, r] = [int(x) for x in res["inf_index_batch"][b][r][0][1].split(",")]
        np.savez_compressed(os.path.join(GOLD, case + ".npz"), emb=emb.numpy(), q=q.numpy(),
                            id_mapping_json=np.array(json.dumps(id_mapping)), dec=np.array(dec_flat).reshape(B, K),
                            beam_scores=np.array(beam_scores, dtype=np.float64), score_rate=np.array(score_rate, dtype=np.float64),
                            loss_func=np.array(loss_func), docids=docids)
        # the restatement must agree with the reference right here, too
        mine = orc.fine_stage(doc_embed, id_mapping, [dec_flat[b * K:(b + 1) * K] for b in range(B)], beam_scores, q,
                              score_rate, loss_func, K)
        for b in range(B):
            for r in range(len(score_rate)):
                assert mine[b][r][2] == docids[b, r].tolist(), (case, b, r)
        print("wrote", case)


def gen_tree_match():
    """tree_match.npz: reference `tree_embedding_calculate` over ALL nodes (main_models.py:154-179: leaf clusters = mean of
    their documents, inner nodes = leaf-count-weighted mean of their children) and the greedy descent `tree_match`
    (main_models.py:232-252) for a set of new documents."""
    import numpy as np
    import torch
    import ref_shims
    from types import SimpleNamespace

    mm, _ = ref_shims.load_ref_main_models()
    args = SimpleNamespace(kary=30, position=1, output_vocab_size=30)
    rng = np.random.RandomState(17)
    g = torch.Generator().manual_seed(99)
    paths = set()
    while len(paths) < 60:                                              # 60 three-level clusters with shared prefixes
        paths.add("%d-%d-%d" % (rng.randint(0, 4), rng.randint(0, 5), rng.randint(0, 30)))
    paths = sorted(paths)
    rng.shuffle(paths)                                                  # insertion order != token order
    D, n_new = 48, 40
    builder = mm.TreeBuilder()
    id_map, embedding = {}, []
    for p in paths:
        toks = mm.encode_single_newid(args, p)                          # [t0, t1, t2, 1]
        n_docs = int(rng.randint(1, 6))
        id_map[p] = list(range(len(embedding), len(embedding) + n_docs))
        for d in id_map[p]:
            builder.add(toks, d)
            embedding.append(torch.randn(D, generator=g))
    root = builder.build()
    mm.tree_embedding_calculate(root, embedding)                        # main_models.py:154-179
    node_paths, node_emb, node_leaves, stack = [], [], [], [((), root)]
    while stack:
        path, n = stack.pop()
        if n.embedding is not None:
            node_paths.append(list(path))
            node_emb.append(n.embedding.clone())
            node_leaves.append(int(n.all_leaf_num))
        for tok, ch in n.children.items():
            stack.append((path + (int(tok),), ch))
    new_docs = torch.randn(n_new, D, generator=g)
    new_docs[:10] = torch.stack([embedding[i] for i in range(0, 50, 5)]) + 0.05 * torch.randn(10, D, generator=g)   # near existing docs
    matches = [mm.tree_match(root, new_docs[i]).tolist() for i in range(n_new)]          # main_models.py:232-252
    np.savez_compressed(os.path.join(GOLD, "tree_match.npz"), embedding=torch.stack(embedding).numpy(), id_map_json=np.array(json.dumps(id_map)),
                        cluster_order=np.array(paths), node_paths_json=np.array(json.dumps(node_paths)),
                        node_emb=torch.stack(node_emb).numpy(), node_leaves=np.array(node_leaves, dtype=np.int64),
                        new_docs=new_docs.numpy(), matches_json=np.array(json.dumps(matches)))
    print("wrote tree_match:", len(node_paths), "nodes with embeddings,", n_new, "matches, e.g.", matches[0])


def gen_contrastive():
    """contrastive.npz: the reference's training-time contrastive loss `encoder_cal` (a closure inside T5FineTuner.forward,
    main_models.py:1184-1221; its SOURCE LINES are read from the mounted reference and exec'd with a stub `self`) over
    `all_doc = cat([positives, in-cluster candidates])` gathered by document index (main_models.py:983-996, 1259-1275),
    plus d loss / d query from autograd.  `.cuda()` is shimmed to a no-op: this container has no GPU."""
    import numpy as np
    import torch
    import ref_shims
    from types import SimpleNamespace

    src = _ref_lines(os.path.join(ref_shims.REF_MODEL_DIR, "main_models.py"), 1184, 1221)
    assert src.lstrip().startswith("def encoder_cal(query, all_doc, valid_num):"), src[:80]
    torch.Tensor.cuda = lambda self, *a, **k: self
    g = torch.Generator().manual_seed(2024)
    N, D, B = 300, 64, 6
    doc_embed = torch.randn(N, D, generator=g) * D ** -0.5
    out = {"doc_embed": doc_embed.numpy()}
    cases = [("tanh", 1.0, 0.05), ("tanh", 0.5, 0.05), ("sigmoid", 2.0, 0.1)]
    for ci, (loss_func, intra_rate, tau) in enumerate(cases):
        valid_num = [int(v) for v in torch.randint(0 if ci else 1, 9, (B,), generator=g)]
        pos = torch.randint(0, N, (B,), generator=g)
        cand = torch.randint(0, N, (sum(valid_num),), generator=g)
        q = (torch.randn(B, D, generator=g) * 1.5).requires_grad_(True)
        stub = SimpleNamespace(args=SimpleNamespace(intra_rate=intra_rate, loss_func=loss_func), tau=tau)
        ns = {"torch": torch, "self": stub}
        exec(textwrap.dedent(src), ns)
        all_doc = torch.cat([doc_embed[pos], doc_embed[cand]], dim=0)          # main_models.py:983-996 gather + :1259-1273 concat
        loss = ns["encoder_cal"](q, all_doc, valid_num)                        # main_models.py:1184-1221
        loss = loss.reshape(())
        loss.backward()
        out.update({f"c{ci}_loss_func": np.array(loss_func), f"c{ci}_intra_rate": np.float64(intra_rate), f"c{ci}_tau": np.float64(tau),
                    f"c{ci}_valid_num": np.array(valid_num, dtype=np.int64), f"c{ci}_pos": pos.numpy(), f"c{ci}_cand": cand.numpy(),
                    f"c{ci}_q": q.detach().numpy(), f"c{ci}_loss": np.float64(loss.item()), f"c{ci}_loss_f32": loss.detach().numpy(),
                    f"c{ci}_grad_q": q.grad.numpy()})
        print("case", ci, loss_func, intra_rate, tau, valid_num, "loss", loss.item())
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(GOLD, "contrastive.npz"), **out)
    print("wrote contrastive")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1:
        {"dense": gen_dense, "main_models": gen_main_models, "tree_match": gen_tree_match, "contrastive": gen_contrastive}[sys.argv[1]]()
    else:
        for part in ("dense", "main_models"):      # separate processes: they need different `transformers`
            subprocess.check_call([sys.executable, os.path.abspath(__file__), part])
