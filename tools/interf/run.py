"""Interference probe: the scoring phase launched back to back on one stream while another stream loops a synthetic kernel
that uses ONE kind of SM resource, in CTAs shaped like the top-k's (1,024 x 128 threads).  Prints the scoring kernel's
average duration beside each co-runner (alone: ~35.5 us; beside the real top-k: ~41.8 us)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench
from gdr_b200 import ClusterStore, _cabi
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "libinterf.so"))
lib.interf_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda', 0)
emb, offsets, docid = bench.synth_shard(cfg, 1234, dev)
A = [ClusterStore(e, offsets, docid) for e in (emb, emb.clone())]
batches = bench.synth_batches(cfg, 2, cfg['C'], cfg['B'], 4321, dev)
out = (torch.empty((1, cfg['B'], 100), device=dev), torch.empty((1, cfg['B'], 100), dtype=torch.int32, device=dev))
for i in range(2): A[i].score_topk(*batches[i], 100, out=out)
src = torch.randn(2 * 1024 * 1024, device=dev)          # 8 MB, L2-resident
sink = torch.zeros(4, device=dev)
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run(kind, iters, n=60, reps=120):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(sa):
        torch.cuda._sleep(6_000_000); e[0].record(sa)
        for i in range(n): A[i % 2].score_topk(*batches[i % 2], 100, out=out, flags=256 | 1024)
        e[1].record(sa)
    if kind is not None:
        with torch.cuda.stream(sb):
            torch.cuda._sleep(6_000_000); e[2].record(sb)
            for i in range(reps): lib.interf_launch(kind, 1024, iters, src.data_ptr(), src.numel() // 4, sink.data_ptr(), _cabi.stream_ptr(sb))
            e[3].record(sb)
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) * 1000 / n, (e[2].elapsed_time(e[3]) * 1000 / reps if kind is not None else 0.0)
print("scoring alone: %.1f us" % run(None, 0)[0])
for name, kind, iters in (("ALU only", 0, 40), ("smem atomics", 1, 60), ("L2 reads", 2, 24), ("barriers", 3, 150), ("code 96 KB", 6, 1), ("ALU + 2 KB smem", 102, 40), ("ALU + 4 KB smem", 104, 40), ("ALU + 6 KB smem", 106, 40), ("ALU + 8 KB smem", 108, 40), ("ALU + 12 KB smem", 112, 40)):
    a, b = run(kind, iters)
    print("scoring || %-13s: scoring %.1f us, co-runner %.1f us per launch" % (name, a, b))
