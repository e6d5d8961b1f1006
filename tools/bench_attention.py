"""Hot-loop timing + fp64 error of the attention op for the library selected by AIHAB_CLIP_LIB.
    python tools/bench_attention.py [--batch 256] [--L 197] [--H 12]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--L", type=int, default=197)
ap.add_argument("--H", type=int, default=12)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(a.batch * a.L, 3 * a.H * 64, device=dev, generator=g) * 1.5).half()
out = ops.attention(qkv, a.batch, a.L, a.H)
nchk = 4
q, k, v = (t.reshape(a.batch, a.L, a.H, 64)[:nchk].permute(0, 2, 1, 3).double() for t in qkv.chunk(3, dim=1))
ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
err = (out.reshape(a.batch, a.L, a.H, 64)[:nchk].permute(0, 2, 1, 3).double() - ref).abs().max().item()
for _ in range(5):
    ops.attention(qkv, a.batch, a.L, a.H)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    ops.attention(qkv, a.batch, a.L, a.H)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 50
print(f"attention L={a.L} B={a.batch}: {ms:.4f} ms  {4.0 * a.batch * a.L * a.L * a.H * 64 / ms / 1e9:.1f} TFLOP/s  max |err| vs fp64 {err:.2e}")
