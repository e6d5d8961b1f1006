"""Phase timeline of the dual-stream attention kernel (attention_tcd.cu) from clock64() stamps of CTA 0, units 4..7.
Needs a library built with -DAIHAB_ATTN_TIMING:  bash tools/build_variant.sh timing -DAIHAB_ATTN_TIMING
    AIHAB_CLIP_LIB=$PWD/ab/lib_timing.so python tools/attn_timing.py [--batch 256]"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--L", type=int, default=197)
args = ap.parse_args()
dev = torch.device("cuda:0")
H = 12
qkv = torch.randn(args.batch * args.L, 3 * H * 64, device=dev).half()
for _ in range(3):
    ops.attention(qkv, args.batch, args.L, H)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (4 * 4 * 16))()
assert lib.aihab_debug_attn_timing(buf) == 0
ts = [[[buf[(r * 4 + u) * 16 + e] for e in range(16)] for u in range(4)] for r in range(4)]
t0 = min(v for r in ts for u in r for v in u if v > 0)
names_i = ["loop top", "Q/K landed", "buffer free (O read)", "S issued", "P ready", "PV issued"]
names_s = ["loop top", "S ready", "max pass done", "turn acquired", "exp pass done", "P stored (wait::st)", "P arrived",
           "O ready", "O read", "O stored"]
for r, nm in enumerate(["issuer A", "issuer B", "softmax A q0", "softmax B q0"]):
    print(f"--- {nm}")
    names = names_i if r < 2 else names_s
    for u in range(4):
        row = ts[r][u]
        print(f"  unit {u + 4}: " + "  ".join(f"{names[e]}={row[e] - t0}" for e in range(len(names)) if row[e] > 0))
per = [ts[2][u + 1][1] - ts[2][u][1] for u in range(3)]
print("cycles per unit (softmax A, S ready -> next S ready):", per)
