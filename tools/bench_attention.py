"""Attention kernel timings for the sequence lengths of SURVEY.md section 8 (hot-loop numbers, CUDA events)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import ops  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

for (n, L, H) in [(64, 256, 16), (64, 257, 16), (32, 577, 16), (256, 197, 12), (512, 50, 12)]:
    qkv = torch.randn(n * L, 3 * H * 64, device="cuda").half()
    ms = timeit(lambda: ops.attention(qkv, n, L, H))
    print(n, L, H, round(ms, 4), "ms", round(4 * n * L * L * H * 64 / ms / 1e9, 1), "TF")
