"""Per-kernel timing on one B200 (CUDA events, warm-up, L2-sized rotation of buffers is NOT done here: these are
hot-loop numbers for optimisation work; bench.py holds the judged measurement)."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--L", type=int, default=197)
    ap.add_argument("--D", type=int, default=768)
    ap.add_argument("--dtype", default="fp16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    M, D, H = args.batch * args.L, args.D, args.D // 64
    res = {}
    shapes = {"qkv": (M, 3 * D, D, _lib.EPI_BIAS_16), "out_proj": (M, D, D, _lib.EPI_BIAS_RES_32),
              "fc_gelu": (M, 4 * D, D, _lib.EPI_BIAS_GELU_16), "mlp_proj": (M, D, 4 * D, _lib.EPI_BIAS_RES_32)}
    for name, (m, n, k, epi) in shapes.items():
        a = torch.randn(m, k, device=dev).to(dt)
        w = (torch.randn(n, k, device=dev) * k ** -0.5).to(dt)
        bias = torch.randn(n, device=dev)
        o16 = torch.empty(m, n, device=dev, dtype=dt) if epi in (_lib.EPI_BIAS_16, _lib.EPI_BIAS_GELU_16) else None
        o32 = torch.zeros(m, n, device=dev) if o16 is None else None
        ms = timeit(lambda: ops.gemm16(a, w, epi, bias=bias, out16=o16, out32=o32))
        res[name] = {"ms": round(ms, 4), "tflops": round(2 * m * n * k / ms / 1e9, 1)}
    qkv = torch.randn(M, 3 * D, device=dev).to(dt)
    ms = timeit(lambda: ops.attention(qkv, args.batch, args.L, H))
    res["attention"] = {"ms": round(ms, 4), "tflops": round(4 * args.batch * args.L * args.L * D / ms / 1e9, 1)}
    x = torch.randn(M, D, device=dev)
    g = torch.ones(D, device=dev)
    ms = timeit(lambda: ops.layernorm(x, g, g, dt))
    res["layernorm"] = {"ms": round(ms, 4), "gbs": round(M * D * 6 / ms / 1e6, 1)}
    # torch (cuBLAS) reference point for the same GEMM shape
    a = torch.randn(M, D, device=dev).to(dt)
    w = torch.randn(3 * D, D, device=dev).to(dt)
    ms = timeit(lambda: torch.matmul(a, w.t()))
    res["cublas_qkv"] = {"ms": round(ms, 4), "tflops": round(2 * M * 3 * D * D / ms / 1e9, 1)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
