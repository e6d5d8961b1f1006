"""Launch one kernel class a few times (for ncu captures):  python tools/run_one.py attention|gemm_qkv|gemm_res|gemm_gelu|ln"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "attention"
B, L, D = 128, 197, 768
dev = torch.device("cuda:0")
M = B * L
dt = torch.float16
if which == "attention":
    qkv = torch.randn(M, 3 * D, device=dev).to(dt)
    for _ in range(4):
        ops.attention(qkv, B, L, D // 64)
elif which.startswith("gemm"):
    shapes = {"gemm_qkv": (M, 3 * D, D, _lib.EPI_BIAS_16), "gemm_res": (M, D, D, _lib.EPI_BIAS_RES_32),
              "gemm_gelu": (M, 4 * D, D, _lib.EPI_BIAS_GELU_16), "gemm_res4": (M, D, 4 * D, _lib.EPI_BIAS_RES_32)}
    m, n, k, epi = shapes[which]
    a = torch.randn(m, k, device=dev).to(dt)
    w = (torch.randn(n, k, device=dev) * k ** -0.5).to(dt)
    bias = torch.randn(n, device=dev)
    o16 = torch.empty(m, n, device=dev, dtype=dt) if epi in (_lib.EPI_BIAS_16, _lib.EPI_BIAS_GELU_16) else None
    o32 = torch.zeros(m, n, device=dev) if o16 is None else None
    for _ in range(4):
        ops.gemm16(a, w, epi, bias=bias, out16=o16, out32=o32)
elif which == "ln":
    x = torch.randn(M, D, device=dev)
    g = torch.ones(D, device=dev)
    for _ in range(4):
        ops.layernorm(x, g, g, dt)
torch.cuda.synchronize()
print("ok", which)
