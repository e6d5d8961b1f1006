"""The actual drop-in proof: the reference's OWN callers of the hot path — imported UNMODIFIED from the oracle/_ref
archive (oracle/build_ref.py) — are handed the model that ``aihab_clip_b200.clip.load`` returns and must produce what
they produce with the reference's own model (goldens of tests/golden/make_golden_cache.py and make_golden.py):

  methods/utils.py   compute_image_features (:142-173), compute_image_features_test (:175-189)
  utils.py           clip_classifier (:31-57)
  aihab_utils/feature_cache.py   cache_preprojection_features (:189-250), cache_openclip_embeddings (:98-186)
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLDEN))
import cache_case as CC  # noqa: E402

from aihab_clip_b200 import _lib  # noqa: E402
from aihab_clip_b200.weights import make_state_dict  # noqa: E402
from oracle import build_ref  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reference_modules():
    if not build_ref.available():
        pytest.skip("oracle/_ref/reference_path.zip was not built (python -m oracle.build_ref where /root/reference exists)")
    build_ref.import_ref()
    import methods.utils as ref_methods_utils
    import utils as ref_utils
    from aihab_utils import feature_cache as ref_feature_cache
    from data.clip_transforms import build_clip_transforms
    from data.templates import CS_CLASSNAMES, CS_TEMPLATES
    return {"mu": ref_methods_utils, "u": ref_utils, "fc": ref_feature_cache, "tf": build_clip_transforms,
            "classes": CS_CLASSNAMES, "templates": CS_TEMPLATES}


def test_reference_callers_run_on_our_model(tmp_path, cuda_device, reference_modules, gold):
    R = reference_modules
    ref_gold = np.load(GOLDEN / "reference_cache.npz")
    ref_meta = json.loads((GOLDEN / "reference_cache.json").read_text())
    import aihab_clip_b200.clip as clip                       # instead of the reference's `import clip`
    path = tmp_path / "tiny16.pt"
    torch.save(make_state_dict("ViT-tiny/16", 0), path)
    state, model, preprocess = clip.load(str(path), device=cuda_device, jit=False)   # aihab_utils/model_init.py:145
    model.float()
    tf = R["tf"]({}, is_train=False, resolution=64)            # the reference's own eval transform
    n0 = _lib.kernel_launches()
    # methods/utils.py:142-173 with our model
    x, y = R["mu"].compute_image_features(model, CC.case_loader(tf, False), to_cpu=True)
    assert _lib.kernel_launches() > n0, "the reference loop must have driven libaihab_clip.so"
    assert x.dtype == torch.float32 and tuple(x.shape) == ref_gold["pre_f0"].shape
    np.testing.assert_allclose(x.numpy(), ref_gold["pre_f0"], atol=1e-2, rtol=0)
    np.testing.assert_array_equal(y.numpy(), ref_gold["pre_label"])
    # utils.py:31-57 with our model: the text head (PyTorch text tower, fp32 on this float() model)
    texts, w_before, text_w = R["u"].clip_classifier(R["classes"], R["templates"], model)
    assert tuple(text_w.shape) == gold["tiny16_text_w"].shape
    np.testing.assert_allclose(text_w.cpu().numpy(), gold["tiny16_text_w"], atol=2e-5, rtol=0)
    np.testing.assert_array_equal(texts.cpu().numpy(), gold["tiny16_texts"])
    # methods/utils.py:175-189 with our model and the reference's VisProj-style callable
    proj = state["visual.proj"].float()
    tw = torch.from_numpy(ref_gold["test_text_w"]).to(cuda_device)
    acc = R["mu"].compute_image_features_test(model, CC.case_loader(tf, False), lambda f: f @ proj, tw)
    assert acc == pytest.approx(ref_meta["zero_shot_acc"])
    # aihab_utils/feature_cache.py writers with our model: same files, same CSV, same values
    cfg = dict(CC.CFG, root_path=str(tmp_path))
    R["fc"].cache_preprojection_features(cfg, {"clip_model": model}, CC.case_loader(tf, False), {"train_size": CC.N_IMAGES})
    d = R["fc"]._feature_cache_dir(cfg)
    np.testing.assert_allclose(torch.load(d / "f1.pth", weights_only=True).numpy(), ref_gold["pre_f1"], atol=1e-2, rtol=0)
    out = R["fc"].cache_openclip_embeddings(cfg, model, CC.case_loader(tf, True), split="Test", checkpoint_path="ckpt/x.pt")
    np.testing.assert_allclose(torch.load(out / "embeddings.pt", weights_only=True).numpy(), ref_gold["emb"], atol=2e-3, rtol=0)
    assert (out / "metadata.csv").read_text() == ref_meta["metadata_csv"]
    # and the reference GPU dtype contract: fp16 model -> fp16 features from the reference loop
    _, model16, _ = clip.load(str(path), device=cuda_device)
    x16, _ = R["mu"].compute_image_features(model16, CC.case_loader(tf, False), to_cpu=False)
    assert x16.dtype == torch.float16 and x16.is_cuda
    np.testing.assert_allclose(x16.float().cpu().numpy(), ref_gold["pre_f0"], atol=2e-2, rtol=0)


def test_reference_pre_load_features_and_cache_model_on_our_model(tmp_path, cuda_device, reference_modules):
    """utils.pre_load_features (utils.py:60-82) and methods.utils.build_cache_model (methods/utils.py:31-45) — the
    reference's own code, which moves batches with `.cuda()` and normalises with torch — on the model our clip.load
    returns, against what they produce on the reference's model."""
    R = reference_modules
    ref_gold = np.load(GOLDEN / "reference_cache.npz")
    ref_meta = json.loads((GOLDEN / "reference_cache.json").read_text())
    import aihab_clip_b200.clip as clip
    path = tmp_path / "tiny16.pt"
    torch.save(make_state_dict("ViT-tiny/16", 0), path)
    state, model, _ = clip.load(str(path), device=cuda_device, jit=False)
    model.float()
    tf = R["tf"]({}, is_train=False, resolution=64)
    f, l = R["u"].pre_load_features({"load_pre_feat": False, "cache_dir": str(tmp_path)}, "val", model, CC.case_loader(tf, False))
    assert sorted(p.name for p in tmp_path.glob("val_*.pt")) == ref_meta["preload_files"]
    np.testing.assert_allclose(f.cpu().numpy(), ref_gold["preload_f"], atol=2e-3, rtol=0)
    np.testing.assert_array_equal(l.cpu().numpy(), ref_gold["preload_l"])
    proj = state["visual.proj"].float()
    keys, values = R["mu"].build_cache_model({"load_cache": False, "augment_epoch": 2, "cache_dir": str(tmp_path / "tip")},
                                             model, CC.case_loader(tf, False), 0, lambda x: x @ proj)
    np.testing.assert_allclose(keys.float().cpu().numpy(), ref_gold["tip_keys"], atol=2e-3, rtol=0)
    np.testing.assert_array_equal(values.float().cpu().numpy(), ref_gold["tip_values"])
