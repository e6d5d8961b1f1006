"""One 439 -> 224 resize of 256 uint8 images (the shipped crop_size case) for ncu / timing: python tools/probes/resize_once.py"""
import sys
import torch
sys.path.insert(0, ".")
from aihab_clip_b200 import ops
dev = torch.device("cuda:0")
u8 = torch.randint(0, 256, (256, 439, 439, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.preprocess_u8(u8, 224, torch.float16)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    ops.preprocess_u8(u8, 224, torch.float16)
e.record()
torch.cuda.synchronize()
print("439 -> 224, 256 images: %.3f ms" % (s.elapsed_time(e) / 20), flush=True)
