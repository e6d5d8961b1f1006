"""One extraction batch through aihab_score (256 x 768 -> 512 -> C classes, top-1) for ncu launch lists / timing."""
import sys
import torch
sys.path.insert(0, ".")
from aihab_clip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 20
feats = torch.randn(256, 768, device=dev, generator=g)
proj = torch.randn(768, 512, device=dev, generator=g) * 768 ** -0.5
tw = torch.nn.functional.normalize(torch.randn(C, 512, device=dev, generator=g), dim=1).t().contiguous()
for _ in range(5):
    ops.score(feats, proj, tw, 100.0, 1)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    ops.score(feats, proj, tw, 100.0, 1)
e.record()
torch.cuda.synchronize()
print("C=%d  %.1f us per call" % (C, s.elapsed_time(e) * 20), flush=True)
