"""Peer-to-peer candidate exchange of the cluster-sharded corpus (include/gdr_b200.h gdr_store_create_shard / gdr_store_p2p_*):
every rank's result for the queries it owns must equal the single-GPU call on the whole corpus BIT FOR BIT.
* one GPU: all "ranks" in one process, exchange buffers cross-wired with gdr_store_p2p_attach_local (tests/_p2p_child.py) —
  ownership, packed offsets, epochs and flags without NVLink;
* two or more GPUs: one process per GPU, CUDA IPC handles exchanged through torch.distributed, scores stored over NVLink,
  fused pipelined schedule (ShardedPipeline) — skipped on a one-GPU box."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(mode, world):
    out = subprocess.run([sys.executable, os.path.join(HERE, "_p2p_child.py"), mode, str(world)], capture_output=True, text=True, timeout=75)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"], line


@pytest.mark.parametrize("mode,world", [("umma", 2), ("simt", 3), ("fp32", 2)])
def test_p2p_exchange_one_process(mode, world):
    _run(mode, world)


@pytest.mark.parametrize("schedule,world", [("fused", 2), ("fused", 4)])
def test_p2p_pipeline_one_process(schedule, world):
    """(The `batches` schedule — several streams per rank, one-warp kernels that spin for a peer — is exercised with one process per GPU
    in test_p2p_sharded_over_nvlink and by bench.py --gpus N: with every "rank" on ONE GPU the spinning kernels of one rank and the
    persistent scoring grids of the others compete for the same device and can starve each other, which says nothing about the product.)"""
    _run(schedule, world)


def test_peer_all_gather_one_process():
    _run("gather", 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import gdr_oracle as orc
    from gdr_b200 import ClusterStore
    from gdr_b200.sharded import ShardedPipeline, partition_contiguous

    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        N, C, D, K, k, b_own = 40000, 256, 768, 20, 100, 192
        emb, offsets, docid = orc.synth_corpus(N, C, D, seed=9)
        emb = emb.bfloat16().float()
        bounds = partition_contiguous(np.diff(offsets), world)
        lo, hi = int(offsets[bounds[rank]]), int(offsets[bounds[rank + 1]])
        shard = ClusterStore.shard(emb[lo:hi].bfloat16().cuda(), offsets, torch.from_numpy(docid).cuda(), int(bounds[rank]), int(bounds[rank + 1]))
        full = ClusterStore.from_csr(emb, offsets, docid, dtype=torch.bfloat16)
        sp = ShardedPipeline(shard, rank, world, b_own, K, k, schedule=os.environ.get("GDR_TEST_SHARDED_SCHEDULE", "auto"))
        tickets, refs = [], []
        for i in range(6):
            q, beams, beam_scores = orc.synth_queries(world * b_own, C, K, D, seed=20 + i)
            qd, bd, pd = q.cuda(), torch.from_numpy(beams).cuda(), torch.softmax(beam_scores, -1).cuda()
            tickets.append(sp.submit(qd, bd, prob=pd, alpha=1.0, act="tanh"))
            s, d = full.score_topk(qd, bd, k, prob=pd, alphas=[1.0], act="tanh")
            refs.append((s[0], d[0]))
        sp.flush()
        torch.cuda.synchronize()
        # the inputs' all-gather over NVLink by the copy engines (PeerAllGather): every rank ends up with the rank-order concatenation
        from gdr_b200.sharded import PeerAllGather
        pag = PeerAllGather(rank, world, [(b_own, D, torch.float32), (b_own, K, torch.int32)], 2, torch.device("cuda", rank))
        for it in range(5):
            g = torch.Generator().manual_seed(500 + it)
            gq = torch.randn(world * b_own, D, generator=g)
            gb = torch.randint(0, C, (world * b_own, K), generator=g, dtype=torch.int32)
            own = torch.empty(pag.own_bytes, dtype=torch.uint8, device="cuda")
            own[:b_own * D * 4].view(torch.float32).view(b_own, D).copy_(gq[rank * b_own:(rank + 1) * b_own])
            own[b_own * D * 4:].view(torch.int32).view(b_own, K).copy_(gb[rank * b_own:(rank + 1) * b_own])
            pag.all_gather(it % 2, own)
            torch.cuda.synchronize()
            assert torch.equal(pag.gathered(it % 2, 0).cpu(), gq) and torch.equal(pag.gathered(it % 2, 1).cpu(), gb), "peer all-gather differs"
            dist.barrier()                 # (a slot is refilled only after every rank has consumed it)
        sl = slice(rank * b_own, (rank + 1) * b_own)
        for t, (rs, rd) in zip(tickets, refs):
            assert torch.equal(t.docids, rd[sl]) and torch.equal(t.scores, rs[sl]), "p2p-sharded result differs from the single-GPU result"
        dist.barrier()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_p2p_sharded_over_nvlink(tmp_path):
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    if world < 2:
        pytest.skip("needs at least two GPUs (the one-process tests above cover the protocol)")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert len(os.listdir(tmp_path)) == world
