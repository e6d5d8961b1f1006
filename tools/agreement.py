"""Parity statistics of the B200 path against the reference logits of the 4096-image agreement set
(tests/golden/reference_outputs.npz: agree_logits).   python tools/agreement.py [--dtype fp16|bf16]"""
import argparse
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aihab_clip_b200.clip as clip  # noqa: E402
from aihab_clip_b200 import ops  # noqa: E402
from aihab_clip_b200.weights import make_state_dict, synthetic_images_u8  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp16")
args = ap.parse_args()
gold = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_outputs.npz")
dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as d:
    path = Path(d) / "b32.pt"
    torch.save(make_state_dict("ViT-B/32", 0), path)
    _, model, _ = clip.load(str(path), device=dev)
model.float()
model.visual.compute_dtype = args.dtype
n = 4096
u8 = np.concatenate([synthetic_images_u8(n // 2, 224, seed=777), synthetic_images_u8(n // 2, 224, seed=777, start=n // 2, smooth=True)])
feats = model.encode_image_u8(torch.from_numpy(u8).to(dev))
emb, logits, idx, _ = ops.score(feats, model.visual.proj, torch.from_numpy(gold["b32_text_w"]).to(dev), 100.0, 3)
got, ref = logits.cpu().numpy(), gold["agree_logits"]
err = float(np.abs(got - ref).max())
srt = np.sort(ref, axis=1)[:, ::-1]
margin = srt[:, 0] - srt[:, 1]
agree = idx[:, 0].cpu().numpy() == ref.argmax(1)
near = margin < 2 * err
print(f"dtype={args.dtype} n={n} max|dlogit|={err:.3e} rms={np.sqrt(((got - ref) ** 2).mean()):.3e} "
      f"strict_argmax_agreement={agree.mean() * 100:.3f}% near_ties={int(near.sum())} "
      f"untied_agreement={agree[~near].mean() * 100:.3f}% untied_disagreements={int((~agree[~near]).sum())}")
