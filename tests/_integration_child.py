"""Child process of tests/test_gpu_integration.py: the edits of INTEGRATION.md applied to the UNMODIFIED reference sources (the
git-ignored copy under baseline/_ref/ that __graft_entry__.build() ships, or /root/reference in the dev container) IN MEMORY —
the source text of the two reference modules is patched exactly as INTEGRATION.md §3(b) and §4(d) say and executed as new modules —
and run against the untouched reference on the same inputs:
  beam   vendored T5ForConditionalGeneration (tiny, random weights) through the reference's live `generate` + `_generate_beam_search`
         (generation_utils_previous.py:629-921): once with `decode_tree=root` (the reference's Python mask block :714-729) and once
         with the block replaced by `scores = decode_tree(input_ids, scores)` and `decode_tree=TreeMask(root)`: identical beams and scores
  fine   `T5FineTuner.validation_step_i` (main_models.py:1337-1642): the reference's own fine stage against the method with lines
         1434-1462 and 1573-1637 replaced by the `self.fine_stage(...)` call: identical `inf_index_batch`
A separate process because the reference needs its vendored transformers 3.4.0 first on sys.path.  `--cpu-dry-run` replaces the CUDA
kernels by the oracle (dev container, no GPU) to check the patching machinery itself."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shims                              # noqa: E402

DRY = "--cpu-dry-run" in sys.argv


def patched_module(path, name, package, edits):
    """Execute the source of `path` with `edits` = [(first_line, last_line, replacement_text)] (1-based, inclusive) applied."""
    lines = open(path).read().split("\n")
    for first, last, text in sorted(edits, reverse=True):
        lines[first - 1:last] = text.split("\n")
    mod = types.ModuleType(name)
    mod.__file__, mod.__package__ = path, package
    sys.modules[name] = mod
    exec(compile("\n".join(lines), path + " [gdr_b200 edits]", "exec"), mod.__dict__)
    return mod


def main():
    mm, gp = ref_shims.load_ref_main_models()          # puts the vendored transformers first on sys.path
    import numpy as np
    import torch
    from types import SimpleNamespace
    from transformers import T5Config, T5ForConditionalGeneration          # the VENDORED 3.4.0
    import gdr_oracle as orc
    dev = "cpu" if DRY else "cuda"
    REF = ref_shims.REF_MODEL_DIR
    out = {"reference": os.path.relpath(REF, ROOT)}

    # ---- a docid tree (30-ary, 3 levels) as the reference builds it
    rng = np.random.RandomState(7)
    args = SimpleNamespace(kary=30, position=1, output_vocab_size=30)
    paths = sorted({"-".join(str(rng.randint(0, 30)) for _ in range(3)) for _ in range(300)})
    builder = mm.TreeBuilder()
    for di, p in enumerate(paths):
        builder.add(mm.encode_single_newid(args, p), di)
    root = builder.build()

    if DRY:
        class TreeMask:                                 # stand-in with the product's call shape, oracle arithmetic
            def __init__(self, r): self.r = r
            def __call__(self, input_ids, scores): return orc.tree_mask(scores, input_ids, self.r)
    else:
        from gdr_b200 import TreeMask

    # ================= beam search: INTEGRATION.md §4(d) =================
    L, K, B = 6, 8, 5
    cfg = T5Config(vocab_size=400, d_model=32, d_kv=8, d_ff=64, num_layers=1, num_decoder_layers=1, num_heads=4, tie_word_embeddings=0,
                   decode_embedding=2, decode_vocab_size=30 * L + 2, output_vocab_size=30, max_output_length=L, Rdrop=0, adaptor_decode=0,
                   adaptor_efficient=0, multiple_decoder=0, embedding_distillation=0, weight_distillation=0, decoder_start_token_id=0,
                   pad_token_id=0, eos_token_id=1)
    torch.manual_seed(11)
    model = T5ForConditionalGeneration(cfg).to(dev).eval()
    ids = torch.randint(2, 400, (B, 9), generator=torch.Generator().manual_seed(3)).to(dev)
    gp_edit = patched_module(os.path.join(REF, "transformers", "generation_utils_previous.py"), "transformers.generation_utils_previous_gdr",
                             "transformers", [(714, 729, "            if decode_tree:\n                scores = decode_tree(input_ids, scores)")])

    def generate(mod, decode_tree):
        model._generate_beam_search = types.MethodType(mod.GenerationMixin._generate_beam_search, model)
        with torch.no_grad():
            (outs, scores), _ = mod.GenerationMixin.generate(
                model, ids, attention_mask=torch.ones_like(ids), num_beams=K, num_return_sequences=K, max_length=L, use_cache=False,
                early_stopping=False, length_penalty=0.8, decode_embedding=2, decode_vocab_size=30 * L + 2, decode_tree=decode_tree,
                decoder_index=-1, output_scores=True, output_encoder_embedding=True)
        return outs.cpu(), [float(s) for s in scores]

    ref_outs, ref_scores = generate(gp, root)
    our_outs, our_scores = generate(gp_edit, TreeMask(root))
    dec = mm.decode_token(args, ref_outs.numpy())
    out["beam"] = {"rows": int(ref_outs.shape[0]), "identical_beams": bool(torch.equal(ref_outs, our_outs)),
                   "max_score_diff": max(abs(a - b) for a, b in zip(ref_scores, our_scores)),
                   "valid_clusters": sum(d in set(paths) for d in dec)}
    ok_beam = out["beam"]["identical_beams"] and out["beam"]["max_score_diff"] <= 1e-6 and out["beam"]["valid_clusters"] > 0

    # ================= fine stage: INTEGRATION.md §3(b) =================
    mm_edit = patched_module(os.path.join(REF, "main_models.py"), "main_models_gdr", "", [
        (1434, 1462, "        if self.args.is_train_encoder:"),
        (1573, 1637, "            inf_index_batch_all = self.fine_stage(dec, scores, query_embeds, texts=texts, gt_answers=batch[\"oldid\"][0])")])
    ok_fine, fine = True, {}
    for loss_func, D, C, per, Bq, Kb in (("tanh", 64, 12, 20, 3, 5), ("sigmoid", 128, 20, 30, 4, 6)):
        g = torch.Generator().manual_seed(77 + D)
        N = C * per
        emb = (torch.randn(N, D, generator=g) * D ** -0.5).bfloat16().float()         # bf16-representable: the store holds the same values
        doc_embed = [emb[i].clone() for i in range(N)]
        cl_paths = paths[:C]
        perm = torch.randperm(N, generator=g).tolist()
        id_mapping = {cl_paths[c]: perm[c * per:(c + 1) * per][: per - (c % 3)] for c in range(C)}
        beams = [torch.randperm(C, generator=g)[:Kb].tolist() for _ in range(Bq)]
        dec_flat = [cl_paths[c] for row in beams for c in row]
        outs = torch.zeros(Bq * Kb, L, dtype=torch.int64)
        for i, p in enumerate(dec_flat):
            t = mm.encode_single_newid(args, p)
            outs[i, 1:1 + len(t)] = torch.tensor(t)
        beam_scores = (-torch.cumsum(torch.rand(Bq, Kb, generator=g), dim=1)).flatten().tolist()
        q = torch.randn(Bq, D, generator=g)
        enc_hidden = torch.zeros(Bq * Kb, 4, D)
        enc_hidden[::Kb, 0] = q
        score_rate = [0, 0.5, 1, 3]
        a = SimpleNamespace(decode_embedding=2, position=1, max_output_length=L, hierarchic_decode=0, output_vocab_size=30, softmax=0,
                            gen_method="greedy", is_train_encoder=1, multiple_decoder=0, num_return_sequences=Kb, length_penalty=0.8, kary=30,
                            label_length_cutoff=0, train_encoder_epoch=10 ** 9, use_query_embed_encoder=1, use_query_embed_decoder_avg=0,
                            use_query_embed_decoder_special=0, loss_func=loss_func, score_rate=score_rate, eval_batch_size=Bq)

        def stub():
            mdl = SimpleNamespace(generate=lambda *x, **kw: ((outs, list(beam_scores)), SimpleNamespace(last_hidden_state=enc_hidden.to(dev))),
                                  config=SimpleNamespace(hidden_size=D))
            return SimpleNamespace(args=a, epoch=0, model=mdl, root=None, cluster=set(cl_paths), tokenizer=SimpleNamespace(decode=lambda i_: "q"),
                                   id_mapping=id_mapping, doc_embed=doc_embed, encoder=lambda query_enc=None, passage=None: query_enc[:, 0],
                                   softmax=torch.nn.Softmax(dim=-1))
        batch = {"source_ids": torch.zeros(Bq, 3, dtype=torch.int64), "source_mask": torch.ones(Bq, 3, dtype=torch.int64),
                 "target_mask": torch.ones(Bq, L, dtype=torch.int64), "rank": [], "oldid": [["gt"] * Bq]}
        if DRY:
            torch.Tensor.cuda = lambda self, *x, **k: self
        ref = mm.T5FineTuner.validation_step_i(stub(), batch, -1)["inf_index_batch"]
        ours_self = stub()
        if DRY:
            def fine_stage(dec, scores, query_embeds, texts=None, gt_answers=None):
                r = orc.fine_stage(doc_embed, id_mapping, dec, scores, query_embeds, score_rate, loss_func, Kb)
                return [[[[texts[b], ",".join(str(x) for x in r[b][i][2]), gt_answers[b]]] for i in range(len(score_rate))] for b in range(len(dec))]
            ours_self.fine_stage = fine_stage
        else:
            from gdr_b200 import FineStage
            ours_self.fine_stage = FineStage(a, doc_embed, id_mapping, dtype=torch.bfloat16, k=Kb)
        ours = mm_edit.T5FineTuner.validation_step_i(ours_self, batch, -1)["inf_index_batch"]
        same = ref == ours
        fine[loss_func] = {"queries": Bq, "rates": len(score_rate), "identical_inf_index_batch": bool(same)}
        ok_fine = ok_fine and same
    out["fine"] = fine
    out["ok"] = bool(ok_beam and ok_fine)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
