"""The cache writers and the other extraction loops on the GPU path (aihab_clip_b200.feature_cache ->
extraction.extract_loader -> libaihab_clip.so), against files / values the UNMODIFIED reference functions produced on
the same checkpoint and dataset (tests/golden/make_golden_cache.py -> reference_cache.npz / .json):
cache_preprojection_features, cache_openclip_embeddings (aihab_utils/feature_cache.py:98-250),
compute_image_features(_test), build_cache_model (methods/utils.py:31-45, 142-189), pre_load_features (utils.py:60-82)."""
import io
import json
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLDEN))
import cache_case as CC  # noqa: E402

from aihab_clip_b200 import _lib, feature_cache as FC  # noqa: E402
from aihab_clip_b200.weights import make_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLDEN / "reference_cache.npz"), json.loads((GOLDEN / "reference_cache.json").read_text())


def load(tmp_path, device, fp32):
    import aihab_clip_b200.clip as clip
    path = tmp_path / "tiny16.pt"
    torch.save(make_state_dict("ViT-tiny/16", 0), path)
    state, model, preprocess = clip.load(str(path), device=device)
    if fp32:
        model.float()
    return state, model, preprocess


def test_cache_preprojection_features_files(tmp_path, cuda_device, ref, capsys):
    gold, meta = ref
    _, model, preprocess = load(tmp_path, cuda_device, fp32=True)
    cfg = dict(CC.CFG, root_path=str(tmp_path))
    n0 = _lib.kernel_launches()
    FC.cache_preprojection_features(cfg, {"clip_model": model}, CC.case_loader(preprocess, False), {"train_size": CC.N_IMAGES})
    assert _lib.kernel_launches() > n0                       # the CUDA engine did the work
    out = capsys.readouterr().out
    assert "'reload_shape_ok': True" in out and "'rows_match_expected': True" in out and "Feature caching complete." in out
    d = FC._feature_cache_dir(cfg)
    assert str(d.relative_to(tmp_path)) == meta["feature_dir_rel"]
    assert sorted(p.name for p in d.iterdir()) == meta["feature_files"]
    assert FC._feature_cache_exists(d, 2)
    for v in range(2):
        f = torch.load(d / f"f{v}.pth", weights_only=True)
        assert str(f.dtype) == meta[f"pre_f{v}_dtype"] and tuple(f.shape) == gold[f"pre_f{v}"].shape
        np.testing.assert_allclose(f.numpy(), gold[f"pre_f{v}"], atol=1e-2, rtol=0)
    lab = torch.load(d / "label.pth", weights_only=True)
    assert str(lab.dtype) == meta["pre_label_dtype"]
    np.testing.assert_array_equal(lab.numpy(), gold["pre_label"])
    # reference GPU behaviour: an fp16 model caches fp16 features (feature_cache.py:215 saves the model dtype)
    _, model16, _ = load(tmp_path, cuda_device, fp32=False)
    cfg16 = dict(cfg, seed=9, aug_views=1)
    FC.cache_preprojection_features(cfg16, {"clip_model": model16}, CC.case_loader(preprocess, False), {"train_size": CC.N_IMAGES})
    f16 = torch.load(FC._feature_cache_dir(cfg16) / "f0.pth", weights_only=True)
    assert f16.dtype == torch.float16
    np.testing.assert_allclose(f16.float().numpy(), gold["pre_f0"], atol=2e-2, rtol=0)
    # raw uint8 HWC batches take the fused GPU preprocessing and give the same bits as host PIL preprocessing
    x_u8, y_u8 = FC.compute_image_features(model, CC.case_loader(None, False, raw_u8=True), to_cpu=True)
    assert torch.equal(x_u8, torch.load(d / "f0.pth", weights_only=True)) and torch.equal(y_u8, lab)
    x_dev, y_dev = FC.compute_image_features(model, CC.case_loader(preprocess, False), to_cpu=False)
    assert x_dev.is_cuda and y_dev.is_cuda and torch.equal(x_dev.cpu(), x_u8)


def test_cache_openclip_embeddings_files(tmp_path, cuda_device, ref):
    gold, meta = ref
    _, model, preprocess = load(tmp_path, cuda_device, fp32=True)
    cfg = dict(CC.CFG, root_path=str(tmp_path))
    out = FC.cache_openclip_embeddings(cfg, model, CC.case_loader(preprocess, True), split="Test", checkpoint_path="ckpt/x.pt")
    assert str(out.relative_to(tmp_path)) == meta["embedding_dir_rel"]
    assert sorted(p.name for p in out.iterdir()) == meta["embedding_files"]
    e = torch.load(out / "embeddings.pt", weights_only=True)
    assert str(e.dtype) == meta["emb_dtype"] and tuple(e.shape) == gold["emb"].shape
    np.testing.assert_allclose(e.numpy(), gold["emb"], atol=2e-3, rtol=0)
    np.testing.assert_allclose(np.linalg.norm(e.numpy(), axis=1), 1.0, atol=1e-5)
    lab = torch.load(out / "labels.pt", weights_only=True)
    assert str(lab.dtype) == meta["emb_labels_dtype"]
    np.testing.assert_array_equal(lab.numpy(), gold["emb_labels"])
    assert (out / "metadata.csv").read_text() == meta["metadata_csv"]     # byte-identical CSV
    df = pd.read_csv(out / "metadata.csv")
    assert list(df.columns) == list(pd.read_csv(io.StringIO(meta["metadata_csv"])).columns)
    info = json.loads((out / "meta.json").read_text())
    assert list(info.keys()) == meta["meta_json_keys"]
    assert {k: v for k, v in info.items() if k not in ("timestamp", "cache_dir")} == meta["meta_json"]
    assert info["cache_dir"] == str(out)
    # not normalised, no metadata: defaults in the CSV, raw pre-projection features
    cfg2 = dict(cfg, seed=4, finetune=dict(cfg["finetune"], cache_embeddings_normalize=False))
    out2 = FC.cache_openclip_embeddings(cfg2, model, CC.case_loader(preprocess, False), split="val")
    np.testing.assert_allclose(torch.load(out2 / "embeddings.pt", weights_only=True).numpy(), gold["emb_raw"], atol=1e-2, rtol=0)
    assert (out2 / "metadata.csv").read_text() == meta["metadata_csv_default"]
    assert json.loads((out2 / "meta.json").read_text())["normalized"] is False
    with pytest.raises(ValueError):
        FC.cache_openclip_embeddings(cfg, model, [(torch.zeros(1, 3, 64, 64),)], split="x")


def test_other_extraction_loops_match_the_reference(tmp_path, cuda_device, ref):
    gold, meta = ref
    state, model, preprocess = load(tmp_path, cuda_device, fp32=True)
    tw = torch.from_numpy(gold["test_text_w"]).to(cuda_device)
    proj = state["visual.proj"].float()

    class VisProjViT(torch.nn.Module):               # methods/ProLIP.py:31-41
        def __init__(self, w):
            super().__init__()
            self.vit_proj = torch.nn.Parameter(w.clone())

        def forward(self, x):
            return x @ self.vit_proj
    for p in (proj, VisProjViT(proj).to(cuda_device), lambda x: x @ proj):   # tensor, VisProj module, any callable
        acc = FC.compute_image_features_test(model, CC.case_loader(preprocess, False), p, tw)
        assert acc == pytest.approx(meta["zero_shot_acc"])
    # pre_load_features: <split>_f.pt / _l.pt, the no-eps normalisation
    pcfg = {"load_pre_feat": False, "cache_dir": str(tmp_path)}
    f, l = FC.pre_load_features(pcfg, "val", model, CC.case_loader(preprocess, False))
    assert f.is_cuda and l.is_cuda and sorted(p.name for p in tmp_path.glob("val_*.pt")) == meta["preload_files"]
    np.testing.assert_allclose(f.cpu().numpy(), gold["preload_f"], atol=2e-3, rtol=0)
    np.testing.assert_array_equal(l.cpu().numpy(), gold["preload_l"])
    f2, l2 = FC.pre_load_features({"load_pre_feat": True, "cache_dir": str(tmp_path)}, "val", model, None)
    assert torch.equal(f2.cpu(), f.cpu()) and torch.equal(l2.cpu(), l.cpu())
    # build_cache_model: proj -> normalise -> mean over augment epochs -> renormalise -> [embed, N]; one-hot fp16 values
    ccfg = {"load_cache": False, "augment_epoch": 2, "cache_dir": str(tmp_path / "tip")}
    keys, values = FC.build_cache_model(ccfg, model, CC.case_loader(preprocess, False), 0, VisProjViT(proj).to(cuda_device))
    assert tuple(keys.shape) == gold["tip_keys"].shape and str(values.dtype) == meta["tip_values_dtype"]
    np.testing.assert_allclose(keys.float().cpu().numpy(), gold["tip_keys"], atol=2e-3, rtol=0)
    np.testing.assert_array_equal(values.float().cpu().numpy(), gold["tip_values"])
    assert (tmp_path / "tip").is_dir()


def test_l2_normalize_kernel(cuda_device):
    from aihab_clip_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(1)
    for cols in (768, 512, 100, 7):
        x = torch.randn(37, cols, generator=g).to(cuda_device)
        x[3] = 0
        for dt in (torch.float32, torch.float16, torch.bfloat16):
            xd = x.to(dt)
            ref = torch.nn.functional.normalize(xd.float(), dim=-1)
            y = ops.l2_normalize(xd, 1e-12)
            assert y.dtype == dt
            tol = 1e-6 if dt == torch.float32 else (1e-3 if dt == torch.float16 else 8e-3)
            assert (y.float() - ref).abs().max().item() <= tol
            y0 = ops.l2_normalize(xd, 0.0, torch.float32)        # f /= f.norm(): zero rows become NaN like the reference
            assert torch.isnan(y0[3]).all() and torch.allclose(y0[4], ref[4], atol=tol)
    with pytest.raises(RuntimeError):
        ops.l2_normalize(torch.zeros(2, 4))
