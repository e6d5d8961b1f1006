import sys, torch
sys.path.insert(0, "/root/repo")
from aihab_clip_b200 import ops
n, L, H = 64, 257, 16
qkv = torch.randn(n * L, 3 * H * 64, device="cuda").half()
for _ in range(4):
    ops.attention(qkv, n, L, H)
torch.cuda.synchronize()
