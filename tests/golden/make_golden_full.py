"""Goldens at FULL depth for BASELINE.json configs 1, 3 and 4, from the UNMODIFIED reference (imported from
/root/reference, CPU fp32).  Run in the build container only:

    python tests/golden/make_golden_full.py      ->  tests/golden/reference_outputs_full.npz

* ``l14``    ViT-L/14 @224 (24 blocks, width 1024, 257 tokens), 8 images (4 noise + 4 smooth, 300 px so the resize
             runs), text head = the reference's ``clip_classifier`` over the 20 shipped class prompts with the model's
             own text tower (width 768, 12 heads)                                                       -> config 3
* ``l14336`` ViT-L/14 @336 (577 tokens), 4 images (400 px)                                              -> config 4
* ``c1``     ViT-B/32, N = 64 synthetic 224 x 224 images, 18 classes x 80 templates (the head is already in
             reference_outputs.npz as ``b32_text_w_18x80``)                                             -> config 1
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

from make_golden import import_reference, load_reference_model, ref_preprocess, sha  # noqa: E402
from aihab_clip_b200.weights import GEOMETRIES, synthetic_images_u8  # noqa: E402

FULL = {"l14": ("ViT-L/14", 0, 8, 300), "l14336": ("ViT-L/14@336px", 0, 4, 400)}


def mixed_images(n, side, seed=1234):
    return np.concatenate([synthetic_images_u8(n // 2, side, seed=seed),
                           synthetic_images_u8(n - n // 2, side, seed=seed, start=n // 2, smooth=True)])


def main():
    ref_clip = import_reference()
    import utils as ref_utils  # reference utils.py (clip_classifier)
    from data.templates import CS_CLASSNAMES, CS_TEMPLATES
    gold = {}
    for tag, (geom_name, seed, n, side) in FULL.items():
        geom = GEOMETRIES[geom_name]
        _, state, model, _ = load_reference_model(ref_clip, geom_name, seed)
        x = ref_preprocess(mixed_images(n, side), geom.image_resolution)       # data/clip_transforms.py:50-56
        with torch.no_grad():
            feats = model.encode_image(x)                                       # clip/model.py:335-336
            _, _, text_w = ref_utils.clip_classifier(CS_CLASSNAMES, CS_TEMPLATES, model)   # utils.py:31-57
            emb = F.normalize(feats @ state["visual.proj"], dim=-1)             # methods/ProLIP.py:40, methods/utils.py:184
            logits = 100. * emb @ text_w                                        # methods/utils.py:185
        gold[f"{tag}_pre_sha"] = np.frombuffer(bytes.fromhex(sha(x.numpy())), dtype=np.uint8)
        gold[f"{tag}_feats"], gold[f"{tag}_emb"] = feats.numpy(), emb.numpy()
        gold[f"{tag}_text_w"], gold[f"{tag}_logits"] = text_w.numpy(), logits.numpy()
        gold[f"{tag}_top3"] = logits.topk(3, 1, True, True)[1].numpy()
        print(tag, geom_name, "tokens", geom.tokens, "logits", tuple(logits.shape), "argmax", logits.argmax(1).tolist(),
              flush=True)
        del model, state

    # config 1 literally: ViT-B/32, batch 64, 224 px inputs, 18 x 80 head
    base = np.load(HERE / "reference_outputs.npz")
    w1880 = torch.from_numpy(base["b32_text_w_18x80"])
    _, state, model, _ = load_reference_model(ref_clip, "ViT-B/32", 0)
    x = ref_preprocess(mixed_images(64, 224, seed=4321), 224)
    with torch.no_grad():
        feats = model.encode_image(x)
        emb = F.normalize(feats @ state["visual.proj"], dim=-1)
        logits = 100. * emb @ w1880
    gold["c1_feats"], gold["c1_emb"], gold["c1_logits"] = feats.numpy(), emb.numpy(), logits.numpy()
    gold["c1_top3"] = logits.topk(3, 1, True, True)[1].numpy()
    print("c1 ViT-B/32 x 64, 18x80 head: logits", tuple(logits.shape), "classes hit", logits.argmax(1).unique().numel())
    np.savez_compressed(HERE / "reference_outputs_full.npz", **gold)
    print("wrote", HERE / "reference_outputs_full.npz", (HERE / "reference_outputs_full.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
