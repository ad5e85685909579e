"""world_size-2 gloo test of the cluster-sharded path's host logic: partition, beam localisation,
candidate packing, all-gather, merge.  The local scorer / merge are the oracle here (CPU, checker
only); on the GPU box the same ShardedRetriever runs the CUDA kernels (tests/test_gpu_sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gdr_oracle as orc
    from gdr_b200.sharded import ShardedRetriever, global_to_local, partition_clusters

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, C, D, Q, K, k = 3000, 48, 32, 24, 6, 20
        emb, offsets, docid = orc.synth_corpus(N, C, D, seed=5)
        q, beams, beam_scores = orc.synth_queries(Q, C, K, D, seed=6)
        prob = torch.softmax(beam_scores, -1)
        sizes = np.diff(offsets)
        owner = partition_clusters(sizes, world)
        g2l, mine = global_to_local(owner, rank)
        # local CSR slab of this rank
        rows = np.concatenate([np.arange(offsets[c], offsets[c + 1]) for c in mine])
        loc_off = np.zeros(mine.size + 1, dtype=np.int64)
        loc_off[1:] = np.cumsum(sizes[mine])
        loc_emb, loc_docid = emb[rows], docid[rows]

        def local_topk(qq, local_beams, kk, pr, alphas, act):
            s, d = orc.dense_topk(qq, loc_emb, loc_off, loc_docid, local_beams.numpy(), kk,
                                  bias=None if pr is None else pr * alphas[0], act=act)
            return s[None], d.to(torch.int32)[None]

        def merge(gathered, kk):
            s = gathered[:, 0].contiguous().view(torch.float32)
            return orc.merge_topk(s, gathered[:, 1].long(), kk)

        r = ShardedRetriever(None, torch.from_numpy(g2l), local_topk=local_topk, merge=merge)
        s, d = r.score_topk(q, torch.from_numpy(beams), k, prob=prob, alpha=1.0, act="tanh")
        ref_s, ref_d = orc.dense_topk(q, emb, offsets, docid, beams, k, bias=prob, act="tanh")
        assert torch.equal(s, ref_s), "merged scores differ from the unsharded oracle"
        assert torch.equal(d.long(), ref_d), "merged docids differ from the unsharded oracle"
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_equals_unsharded_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
