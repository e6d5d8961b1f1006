"""One flash-attention launch per sequence length of ViT-L (for an ncu capture)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from aihab_clip_b200 import ops  # noqa: E402

for (n, L, H) in [(32, 577, 16), (64, 257, 16)]:
    qkv = torch.randn(n * L, 3 * H * 64, device="cuda").half()
    for _ in range(3):
        ops.attention(qkv, n, L, H)
torch.cuda.synchronize()
