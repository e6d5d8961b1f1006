"""Supplementary goldens: the UNMODIFIED reference (imported from /root/reference, CPU fp32) at the ViT-L/14 widths
and sequence lengths (257 / 577 tokens, 16 heads of 64, 588-wide patches) of BASELINE.json configs 3 and 4, on the
2-block ViT-L-mini geometries and the exact inputs `tests/test_gpu_model.py::test_vit_l_geometries_match_oracle`
uses; and at the full headline geometry ViT-B/16 (configs[1]).  Run in the build container only:

    python tests/golden/make_golden_vitl.py      ->  tests/golden/reference_outputs_vitl.npz
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

CASES = {"lmini": ("ViT-L-mini/14", 3, 5, 300), "lmini336": ("ViT-L-mini/14@336px", 3, 3, 300)}
# the headline geometry itself (BASELINE.json configs[1]: ViT-B/16, 197 tokens, 12 blocks of width 768), with the
# shipped 20-class text head, on 16 images (half noise, half smooth, 300 px so the resize runs)
B16 = ("b16", "ViT-B/16", 0, 16, 300)


def main():
    ref_clip = import_reference()
    gold = {}
    for tag, (geom_name, seed, n, side) in CASES.items():
        geom = GEOMETRIES[geom_name]
        _, state, model, _ = load_reference_model(ref_clip, geom_name, seed)
        u8 = synthetic_images_u8(n, side, smooth=True)
        x = ref_preprocess(u8, geom.image_resolution)                       # data/clip_transforms.py:50-56
        with torch.no_grad():
            feats = model.encode_image(x)                                   # clip/model.py:335-336
            emb = F.normalize(feats @ state["visual.proj"], dim=-1)         # methods/ProLIP.py:40, methods/utils.py:184
        gold[f"{tag}_pre_sha"] = np.frombuffer(bytes.fromhex(sha(x.numpy())), dtype=np.uint8)
        gold[f"{tag}_feats"] = feats.numpy()
        gold[f"{tag}_emb"] = emb.numpy()
        print(tag, geom_name, "tokens", geom.tokens, "feats", tuple(feats.shape))
    import utils as ref_utils  # reference utils.py (clip_classifier)
    from data.templates import CS_CLASSNAMES, CS_TEMPLATES
    tag, geom_name, seed, n, side = B16
    geom = GEOMETRIES[geom_name]
    _, state, model, _ = load_reference_model(ref_clip, geom_name, seed)
    u8 = np.concatenate([synthetic_images_u8(n // 2, side, seed=1234),
                         synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])
    x = ref_preprocess(u8, geom.image_resolution)
    with torch.no_grad():
        feats = model.encode_image(x)
        _, _, text_w = ref_utils.clip_classifier(CS_CLASSNAMES, CS_TEMPLATES, model)   # utils.py:31-57
        emb = F.normalize(feats @ state["visual.proj"], dim=-1)
        logits = 100. * emb @ text_w                                                    # methods/utils.py:185
    gold[f"{tag}_pre_sha"] = np.frombuffer(bytes.fromhex(sha(x.numpy())), dtype=np.uint8)
    gold[f"{tag}_feats"], gold[f"{tag}_emb"] = feats.numpy(), emb.numpy()
    gold[f"{tag}_text_w"], gold[f"{tag}_logits"] = text_w.numpy(), logits.numpy()
    gold[f"{tag}_top3"] = logits.topk(3, 1, True, True)[1].numpy()
    print(tag, geom_name, "tokens", geom.tokens, "logits", tuple(logits.shape), "argmax", logits.argmax(1).tolist())
    np.savez_compressed(HERE / "reference_outputs_vitl.npz", **gold)
    print("wrote", HERE / "reference_outputs_vitl.npz", (HERE / "reference_outputs_vitl.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
