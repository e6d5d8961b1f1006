"""Times aihab_score16 (config 5: 1 M x 768 -> 512 -> 1000 classes, top-k) for a few k: python tools/probes/score16_time.py"""
import sys
import torch
sys.path.insert(0, ".")
from aihab_clip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
n = 1_000_000
feats = torch.randn(n, 768, device=dev, generator=g).half()
proj = (torch.randn(768, 512, device=dev, generator=g) * 768 ** -0.5).half()
tw = torch.nn.functional.normalize(torch.randn(1000, 512, device=dev, generator=g), dim=1).t().contiguous()
for k in (1, 5, 8):
    for _ in range(3):
        ops.score16(feats, proj, tw, 100.0, k)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.score16(feats, proj, tw, 100.0, k)
    e.record()
    torch.cuda.synchronize()
    print("k=%d  %.3f ms / 1M rows" % (k, s.elapsed_time(e) / 10), flush=True)
