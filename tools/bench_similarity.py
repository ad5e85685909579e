#!/usr/bin/env python
"""Dense similarity `q @ p.T` (GDR_model/dense.py:53-54) on the tensor cores: gdr_similarity against torch.matmul on the same GPU.

    python tools/bench_similarity.py [--Q 7830] [--P 109739] [--D 768]

Prints one JSON line: ms per call, effective TFLOP/s (2*Q*P*D), output GB/s, and the max deviation from an fp64 reference on a slice.
The product is exact to fp32 accumulation order (fp32 queries split into three bf16 terms), torch's bf16 matmul is not."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdr_b200 import compute_similarity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--Q", type=int, default=7830)
    ap.add_argument("--P", type=int, default=109739)
    ap.add_argument("--D", type=int, default=768)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn((args.Q, args.D), generator=g, device="cuda")
    p = (torch.randn((args.P, args.D), generator=g, device="cuda") * args.D ** -0.5).bfloat16()

    def timed(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(n):
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            del out
        return sorted(ts)[len(ts) // 2]

    ms = timed(lambda: compute_similarity(q, p))
    ms_t32 = timed(lambda: torch.matmul(q, p.float().T))
    ms_t16 = timed(lambda: torch.matmul(q.bfloat16(), p.T).float())
    out = compute_similarity(q[:64], p[:4096])
    ref = (q[:64].double() @ p[:4096].double().T)
    flops = 2.0 * args.Q * args.P * args.D
    print(json.dumps({"op": "gdr_similarity (tcgen05 grouped GEMM, exact 3-term split of fp32 queries)", "Q": args.Q, "P": args.P, "D": args.D,
                      "ms": ms, "tflops_effective": flops / ms / 1e9, "out_GBps": args.Q * args.P * 4 / ms / 1e6,
                      "torch_fp32_matmul_ms": ms_t32, "torch_bf16_matmul_ms": ms_t16,
                      "max_abs_err_vs_fp64": float((out.double() - ref).abs().max()),
                      "torch_bf16_max_abs_err_vs_fp64": float(((q[:64].bfloat16() @ p[:4096].T).double() - ref).abs().max())}))


if __name__ == "__main__":
    main()
