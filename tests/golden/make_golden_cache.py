"""Goldens for the cache writers and the other extraction loops, produced by the UNMODIFIED reference functions
(imported from /root/reference, CPU fp32) on the ViT-tiny/16 checkpoint and the dataset of tests/golden/cache_case.py:

  aihab_utils/feature_cache.py  cache_preprojection_features (:189-250), cache_openclip_embeddings (:98-186)
  methods/utils.py              compute_image_features (:142-173), compute_image_features_test (:175-189),
                                build_cache_model (:31-45)
  utils.py                      pre_load_features (:60-82)

    python tests/golden/make_golden_cache.py   ->  tests/golden/reference_cache.npz + reference_cache.json
"""
from __future__ import annotations

import contextlib
import io
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

from make_golden import import_reference, load_reference_model  # noqa: E402
import cache_case as CC  # noqa: E402


def main():
    ref_clip = import_reference()
    from aihab_utils import feature_cache as RFC      # the reference's writers
    from data.clip_transforms import build_clip_transforms
    import methods.utils as RMU
    import utils as RU
    _, state, model, _ = load_reference_model(ref_clip, "ViT-tiny/16", 0)
    tf = build_clip_transforms({}, is_train=False, resolution=64)
    gold, meta = {}, {}
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        cfg = dict(CC.CFG, root_path=tmp)
        # ---- cache_preprojection_features
        RFC.cache_preprojection_features(cfg, {"clip_model": model}, CC.case_loader(tf, False), {"train_size": CC.N_IMAGES})
        d = RFC._feature_cache_dir(cfg)
        meta["feature_files"] = sorted(p.name for p in d.iterdir())
        for v in range(2):
            f = torch.load(d / f"f{v}.pth", weights_only=True)
            gold[f"pre_f{v}"] = f.numpy()
            meta[f"pre_f{v}_dtype"] = str(f.dtype)
        lab = torch.load(d / "label.pth", weights_only=True)
        gold["pre_label"], meta["pre_label_dtype"] = lab.numpy(), str(lab.dtype)
        # ---- cache_openclip_embeddings (normalised, with metadata) and without metadata
        out = RFC.cache_openclip_embeddings(cfg, model, CC.case_loader(tf, True), split="Test", checkpoint_path="ckpt/x.pt")
        meta["embedding_files"] = sorted(p.name for p in out.iterdir())
        e = torch.load(out / "embeddings.pt", weights_only=True)
        gold["emb"], meta["emb_dtype"] = e.numpy(), str(e.dtype)
        l2 = torch.load(out / "labels.pt", weights_only=True)
        gold["emb_labels"], meta["emb_labels_dtype"] = l2.numpy(), str(l2.dtype)
        meta["metadata_csv"] = (out / "metadata.csv").read_text()
        info = json.loads((out / "meta.json").read_text())
        meta["meta_json_keys"] = list(info.keys())
        meta["meta_json"] = {k: v for k, v in info.items() if k not in ("timestamp", "cache_dir")}
        meta["embedding_dir_rel"] = str(out.relative_to(tmp))
        meta["feature_dir_rel"] = str(d.relative_to(tmp))
        cfg2 = dict(cfg, seed=4, finetune=dict(cfg["finetune"], cache_embeddings_normalize=False))
        out2 = RFC.cache_openclip_embeddings(cfg2, model, CC.case_loader(tf, False), split="val")
        gold["emb_raw"] = torch.load(out2 / "embeddings.pt", weights_only=True).numpy()
        meta["metadata_csv_default"] = (out2 / "metadata.csv").read_text()
        # ---- compute_image_features_test, pre_load_features, build_cache_model (they call .cuda(): run on CPU by
        # patching Tensor.cuda to the identity for the duration of the call — the arithmetic is untouched)
        tw = torch.nn.functional.normalize(torch.randn(64, 20, generator=torch.Generator().manual_seed(5)), dim=0)
        gold["test_text_w"] = tw.numpy()
        proj = state["visual.proj"].float()
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            acc = RMU.compute_image_features_test(model, CC.case_loader(tf, False), lambda x: x @ proj, tw)
            pcfg = {"load_pre_feat": False, "cache_dir": tmp}
            pf, pl = RU.pre_load_features(pcfg, "val", model, CC.case_loader(tf, False))
            ccfg = {"load_cache": False, "augment_epoch": 2, "cache_dir": tmp + "/tip"}
            ck, cv = RMU.build_cache_model(ccfg, model, CC.case_loader(tf, False), 0, lambda x: x @ proj)
        finally:
            torch.Tensor.cuda = orig_cuda
        meta["zero_shot_acc"] = float(acc)
        gold["preload_f"], gold["preload_l"] = pf.numpy(), pl.numpy()
        meta["preload_files"] = sorted(p.name for p in Path(tmp).glob("val_*.pt"))
        gold["tip_keys"], gold["tip_values"] = ck.numpy(), cv.float().numpy()
        meta["tip_values_dtype"] = str(cv.dtype)
    np.savez_compressed(HERE / "reference_cache.npz", **gold)
    (HERE / "reference_cache.json").write_text(json.dumps(meta, indent=1))
    print({k: v.shape for k, v in gold.items()})
    print(json.dumps({k: v for k, v in meta.items() if "csv" not in k}, indent=1))


if __name__ == "__main__":
    main()
