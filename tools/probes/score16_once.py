import torch, sys
sys.path.insert(0, ".")
from aihab_clip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
n=131072
feats = torch.randn(n, 768, device=dev, generator=g).half()
proj = (torch.randn(768, 512, device=dev, generator=g) * 768 ** -0.5).half()
tw = torch.nn.functional.normalize(torch.randn(1000, 512, device=dev, generator=g), dim=1).t().contiguous()
for _ in range(3): ops.score16(feats, proj, tw, 100.0, 5)
torch.cuda.synchronize()
