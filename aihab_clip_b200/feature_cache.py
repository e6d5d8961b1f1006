"""Feature-cache extraction — drop-in for ``aihab_utils/feature_cache.py`` and ``methods/utils.py:142-173``.

Same function names, arguments, return values, directory layout and file formats as the reference:

* ``<root>/features_<Backbone>_<dataset>/<shots>_shot/seed<seed>/f{v}.pth + label.pth``   (ref :35-43, :215-222)
* ``<cache_embeddings_dir>/<backbone>_<dataset>/<split>/seed<seed>/{embeddings.pt, labels.pt, metadata.csv,
  meta.json}``                                                                            (ref :53-65, :147-174)

What changes is where the work happens: batches go through the sm_100a engine (``model.encode_image`` for the
preprocessed float batches a reference DataLoader yields, ``model.encode_image_u8`` for raw uint8 HWC batches),
results stay in HBM for the whole pass and are copied to the host ONCE instead of per batch (the reference syncs
on ``.to('cpu')`` every batch, methods/utils.py:164), and the L2 normalisation runs in the scoring kernel.
"""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import ops


def _canonical_backbone_name(backbone: str) -> str:
    """ref :15-32 — folder-safe backbone name (ViT-B/16 -> ViTB16, ViT-B/32 -> ViTB32, others sanitised)."""
    if not backbone:
        return "unknown"
    short = {"ViT-B/16": "ViTB16", "ViT-B/32": "ViTB32"}
    if backbone in short:
        return short[backbone]
    name = backbone.replace("hf-hub:", "hf-hub_")
    for ch in "/ :":
        name = name.replace(ch, "_")
    return name


def _backbone_of(cfg) -> str:
    if str(cfg.get("clip_backend", "openai")).lower() == "openclip":
        return cfg.get("open_clip_model", cfg.get("backbone", "RN50"))
    return cfg.get("backbone", "RN50")


def _feature_cache_dir(cfg) -> Path:
    """ref :35-43."""
    root = Path(cfg.get("root_path", "./"))
    shots = int(cfg.get("shots", 0) or 0)
    seed = int(cfg.get("seed", 1) or 1)
    return (root / f"features_{_canonical_backbone_name(_backbone_of(cfg))}_{cfg.get('dataset', 'cs')}"
            / f"{shots}_shot" / f"seed{seed}")


def _embedding_cache_dir(cfg, split: str) -> Path:
    """ref :53-65."""
    root = Path(cfg.get("root_path", "./"))
    out_root = Path(cfg.get("finetune", {}).get("cache_embeddings_dir", "feat_cache_vis"))
    if not out_root.is_absolute():
        out_root = root / out_root
    seed = int(cfg.get("seed", 1) or 1)
    return (out_root / f"{_canonical_backbone_name(_backbone_of(cfg))}_{cfg.get('dataset', 'cs')}"
            / str(split).lower() / f"seed{seed}")


def _feature_cache_exists(cache_dir: Path, aug_views: int) -> bool:
    """ref :253-261."""
    cache_dir = Path(cache_dir)
    return (cache_dir.exists() and (cache_dir / "label.pth").is_file()
            and all((cache_dir / f"f{v}.pth").is_file() for v in range(aug_views)))


def _model_device(clip_model) -> torch.device:
    try:
        return next(clip_model.parameters()).device
    except StopIteration:
        return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _encode_batch(clip_model, images: torch.Tensor, device) -> torch.Tensor:
    images = images.to(device, non_blocking=True)
    if images.dtype == torch.uint8:  # raw HWC batch: GPU preprocessing fused in front of the tower
        return clip_model.encode_image_u8(images)
    return clip_model.encode_image(images)


def compute_image_features(clip_model, loader, to_cpu: bool = False):
    """ref methods/utils.py:142-173 — pre-projection features ``[N, width]`` (model dtype) and labels for a loader
    of ``(images, target)`` batches.  ``to_cpu=True`` returns CPU tensors (one device->host copy at the end)."""
    device = _model_device(clip_model)
    feats, labels = [], []
    with torch.no_grad():
        for images, target in loader:
            feats.append(_encode_batch(clip_model, images, device))
            labels.append(target.detach() if to_cpu else target.to(device, non_blocking=True))
    x = torch.cat(feats, dim=0)
    y = torch.cat(labels, dim=0)
    if to_cpu:
        x, y = x.detach().to("cpu"), y.to("cpu")
    return x, y


def cache_preprojection_features(cfg, clip_bundle: dict, dl_tr, info: dict):
    """ref :189-250 — writes f{v}.pth (features in model dtype) per augmentation view and label.pth (int64) once,
    then reloads each file to validate its shape."""
    clip_model = clip_bundle["clip_model"]
    cache_dir = _feature_cache_dir(cfg)
    num_views = int(cfg.get("aug_views", 1) or 1)
    expected_n = int(info.get("train_size", len(dl_tr.dataset)))
    print("\n==== Feature Caching (pre-projection) ====")
    print({"cache_dir": str(cache_dir), "backbone": cfg.get("backbone", "RN50"), "dataset": cfg.get("dataset", "cs"),
           "shots": int(cfg.get("shots", 0) or 0), "seed": int(cfg.get("seed", 1) or 1), "aug_views": num_views,
           "expected_train_size": expected_n})
    clip_model.eval()
    cache_dir.mkdir(parents=True, exist_ok=True)
    for v in range(num_views):
        feats_t, labels_t = compute_image_features(clip_model, dl_tr, to_cpu=True)
        fpath = cache_dir / f"f{v}.pth"
        torch.save(feats_t, fpath)
        print(f"[cache] view {v} -> {fpath}")
        print({"features.shape": tuple(feats_t.shape), "features.dtype": str(feats_t.dtype)})
        if v == 0:
            lpath = cache_dir / "label.pth"
            torch.save(labels_t, lpath)
            print(f"[cache] labels -> {lpath}")
            print({"labels.shape": tuple(labels_t.shape), "labels.dtype": str(labels_t.dtype),
                   "num_unique_labels": int(labels_t.unique().numel())})
        loaded = torch.load(fpath, map_location="cpu", weights_only=True)
        print({"reload_shape_ok": tuple(loaded.shape) == tuple(feats_t.shape),
               "rows_match_labels": feats_t.shape[0] == labels_t.shape[0],
               "rows_match_expected": feats_t.shape[0] == expected_n})
    print("\nFeature caching complete.")


def _to_py(v: Any) -> Any:
    if isinstance(v, torch.Tensor):
        return v.item() if v.numel() == 1 else v.detach().cpu().tolist()
    if isinstance(v, np.generic):
        return v.item()
    return v


def _metadata_rows(metadata: Any, batch_size: int) -> List[Dict[str, Any]]:
    """ref :84-95 — per-sample dicts from a collated metadata dict; defaults when missing."""
    if not isinstance(metadata, dict):
        print("[warn] metadata missing; writing default values in metadata.csv." if metadata is None
              else "[warn] metadata is not a dict; writing default values in metadata.csv.")
        return [{} for _ in range(batch_size)]
    rows = []
    for i in range(batch_size):
        rows.append({k: _to_py(v[i] if isinstance(v, (list, tuple, np.ndarray, torch.Tensor)) else v)
                     for k, v in metadata.items()})
    return rows


def cache_openclip_embeddings(cfg: dict, model: torch.nn.Module, loader, split: str = "test",
                              checkpoint_path: Optional[str] = None) -> Path:
    """ref :98-186 — embeddings.pt (optionally L2-normalised), labels.pt, metadata.csv, meta.json."""
    import pandas as pd
    normalize = bool(cfg.get("finetune", {}).get("cache_embeddings_normalize", True))
    cache_dir = _embedding_cache_dir(cfg, split)
    cache_dir.mkdir(parents=True, exist_ok=True)
    device = _model_device(model)
    model.eval()
    feats_list, labels_list, rows = [], [], []
    with torch.no_grad():
        for batch in loader:
            if isinstance(batch, (list, tuple)) and len(batch) == 3:
                images, targets, metadata = batch
            elif isinstance(batch, (list, tuple)) and len(batch) == 2:
                images, targets = batch
                metadata = None
            else:
                raise ValueError("Expected batch to be (images, targets) or (images, targets, metadata).")
            feats = _encode_batch(model, images, device)
            if normalize:  # F.normalize(feats, dim=-1) in the scoring kernel, cast back to the model dtype
                emb, _, _, _ = ops.score(feats, None, None, k=0)
                feats = emb.to(feats.dtype)
            feats_list.append(feats)
            targets_cpu = targets.detach().to("cpu")
            labels_list.append(targets_cpu)
            for i, row in enumerate(_metadata_rows(metadata, int(targets_cpu.shape[0]))):
                rows.append({"file_name": row.get("file_name", ""),
                             "ground_truth_num_label": int(targets_cpu[i].item()),
                             "ground_truth_word_label": row.get("plot_word_label", ""),
                             "ground_truth_L2_num_label": row.get("l2_label", -1)})
    feats_all = torch.cat(feats_list, dim=0).detach().to("cpu")
    labels_all = torch.cat(labels_list, dim=0)
    torch.save(feats_all, cache_dir / "embeddings.pt")
    torch.save(labels_all, cache_dir / "labels.pt")
    columns = ["file_name", "ground_truth_num_label", "ground_truth_word_label", "ground_truth_L2_num_label"]
    pd.DataFrame(rows).reindex(columns=columns).to_csv(cache_dir / "metadata.csv", index=False)
    info = {"timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S"), "split": str(split), "normalized": normalize,
            "num_samples": int(feats_all.shape[0]), "dim": int(feats_all.shape[1]) if feats_all.ndim == 2 else None,
            "checkpoint_path": str(checkpoint_path) if checkpoint_path is not None else None,
            "cache_dir": str(cache_dir)}
    (cache_dir / "meta.json").write_text(json.dumps(info, indent=2))
    print("\n==== OpenCLIP Embedding Cache ====")
    print({"cache_dir": str(cache_dir), "num_samples": info["num_samples"], "dim": info["dim"],
           "normalized": normalize})
    return cache_dir


def compute_image_features_test(clip_model, loader, proj, text_weights) -> float:
    """ref methods/utils.py:175-189 — zero-shot accuracy (%): encode -> proj -> normalise -> 100 * f @ W -> argmax.
    ``proj`` is a ``[width, embed]`` tensor or a module with a ``vit_proj`` parameter (methods/ProLIP.py:31-41)."""
    device = _model_device(clip_model)
    p = getattr(proj, "vit_proj", proj)
    correct, total = 0, 0
    with torch.no_grad():
        for images, target in loader:
            feats = _encode_batch(clip_model, images, device)
            _, _, idx, _ = ops.score(feats, p.to(device), text_weights.to(device), 100.0, 1, want_emb=False,
                                     want_logits=False)
            correct += int((idx[:, 0].cpu() == target.cpu()).sum())
            total += int(target.shape[0])
    return 100.0 * correct / max(total, 1)
