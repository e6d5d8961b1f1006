"""Parity at the benchmark's own geometry (BASELINE.json configs[1]): the full ViT-B/16 image tower (197 tokens, 12
blocks, width 768) + scoring on a B200 against outputs of the unmodified reference
(tests/golden/reference_outputs_vitl.npz, made by tests/golden/make_golden_vitl.py).  Gates from BASELINE.json:
embedding cosine >= 0.999, max |dlogit| <= 1e-2 on the x100 logits, argmax / top-3 exact where the reference scores are
untied.  (The file name sorts last on purpose: it is the slowest model-level case.)"""
import numpy as np
import pytest
import torch

from aihab_clip_b200.weights import GEOMETRIES, make_state_dict, synthetic_images_u8
from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu


def test_vit_b16_matches_reference_golden(tmp_path, cuda_device, gold_vitl, capsys):
    import aihab_clip_b200.clip as clip
    from aihab_clip_b200 import ops
    geom = GEOMETRIES["ViT-B/16"]
    path = tmp_path / "b16.pt"
    torch.save(make_state_dict(geom.name, 0), path)
    _, model, preprocess = clip.load(str(path), device=cuda_device)
    model.float()
    n, side = 16, 300
    u8 = np.concatenate([synthetic_images_u8(n // 2, side, seed=1234),
                         synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])
    x = preprocess.batch_u8(torch.from_numpy(u8).to(cuda_device))
    ref_x = np.stack([O.clip_preprocess(im, geom.image_resolution) for im in u8])
    np.testing.assert_array_equal(x.cpu().numpy(), ref_x)  # Pillow-exact resize + normalise, bit for bit
    feats = model.encode_image(x)
    assert tuple(feats.shape) == (n, geom.vision_width)
    text_w = torch.from_numpy(gold_vitl["b16_text_w"]).to(cuda_device)
    emb, logits, idx, _ = ops.score(feats, model.visual.proj, text_w, 100.0, 3)
    cos = (emb.cpu().numpy() * gold_vitl["b16_emb"]).sum(-1)
    ref_logits = gold_vitl["b16_logits"]
    err = float(np.abs(logits.cpu().numpy() - ref_logits).max())
    with capsys.disabled():
        print(f"\n[ViT-B/16 vs reference] min cosine {cos.min():.6f}  max |dlogit| {err:.2e}")
    assert cos.min() >= 0.999
    assert err <= 1e-2, f"max |dlogit| = {err}"
    srt = np.sort(ref_logits, axis=1)[:, ::-1]
    untied3 = np.abs(np.diff(srt[:, :4], axis=1)).min(axis=1) > 2 * err
    np.testing.assert_array_equal(idx.cpu().numpy()[untied3], gold_vitl["b16_top3"][untied3])
    # the fused uint8 path and a different batch composition give the same bits
    assert torch.equal(model.encode_image_u8(torch.from_numpy(u8).to(cuda_device)), feats)
    assert torch.equal(model.encode_image(x[5:9]), feats[5:9])
