"""Times the patch-embedding GEMM (EPI_PATCH_32) and a plain fp32-output GEMM (EPI_SCALE_32) at the ViT-B/16 batch-256 shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from aihab_clip_b200 import _lib, ops  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

n, g2, D, K = 256, 196, 768, 768
a = torch.randn(n * g2, K, device="cuda").half()
w = (torch.randn(D, K, device="cuda") * K ** -0.5).half()
pos = torch.randn(g2 + 1, D, device="cuda")
out = torch.zeros(n * (g2 + 1), D, device="cuda")
ms = timeit(lambda: ops.gemm16(a, w, _lib.EPI_PATCH_32, out32=out, pos=pos, g2=g2))
print("patch_embed", round(ms, 4), "ms", round(2 * n * g2 * D * K / ms / 1e9, 1), "TF")
o2 = torch.zeros(n * g2, D, device="cuda")
bias = torch.randn(D, device="cuda")
ms = timeit(lambda: ops.gemm16(a, w, _lib.EPI_SCALE_32, bias=bias, out32=o2, scale=2.0))
print("scale_32   ", round(ms, 4), "ms", round(2 * n * g2 * D * K / ms / 1e9, 1), "TF")
