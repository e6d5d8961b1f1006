"""Feature-cache extraction — drop-in for ``aihab_utils/feature_cache.py`` and ``methods/utils.py:142-173``.

Same function names, arguments, return values, directory layout and file formats as the reference:

* ``<root>/features_<Backbone>_<dataset>/<shots>_shot/seed<seed>/f{v}.pth + label.pth``   (ref :35-43, :215-222)
* ``<cache_embeddings_dir>/<backbone>_<dataset>/<split>/seed<seed>/{embeddings.pt, labels.pt, metadata.csv,
  meta.json}``                                                                            (ref :53-65, :147-174)

plus the two remaining extraction loops of the reference, ``pre_load_features`` (utils.py:60-82, ``<split>_f.pt`` /
``<split>_l.pt``) and ``build_cache_model`` (methods/utils.py:31-45).

What changes is where the work happens.  Every function here is a thin caller of
``extraction.extract_loader``: batches go through the sm_100a engine (``model.encode_image`` for the preprocessed
float batches a reference DataLoader yields, ``model.encode_image_u8`` for raw uint8 HWC batches); the loader's batches
are sharded BY IMAGE BATCH across the ranks of an initialised ``torch.distributed`` group (1/2/4/8 B200s, one process
per GPU; a single process is the 1-GPU case); rows stay in HBM for the whole pass (the reference syncs on
``.to('cpu')`` every batch, methods/utils.py:164), meet in ONE all-gather, and reach the host in ONE copy;
normalisation runs in ``aihab_l2_normalize`` / the scoring kernel.  Rank 0 writes the files; every rank returns the
same tensors.
"""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import ops
from .extraction import extract_loader, metadata_rows as _metadata_rows  # noqa: F401


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


def _is_writer() -> bool:
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def _sync_ranks():
    """Files written by rank 0 are visible to every rank when a writer returns."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def compute_image_features(clip_model, loader, to_cpu: bool = False, _encode_fn=None):
    """ref methods/utils.py:142-173 — pre-projection features ``[N, width]`` (model dtype) and labels for a loader
    of ``(images, target)`` batches, in loader order.  ``to_cpu=True`` returns CPU tensors (ONE device->host copy at the
    end; very large passes spill to pinned memory in bounded blocks, ``AIHAB_CACHE_SPILL_MB``)."""
    x, y, _ = extract_loader(clip_model, loader, to_cpu=to_cpu, encode_fn=_encode_fn)
    return x, y


def cache_preprojection_features(cfg, clip_bundle: dict, dl_tr, info: dict, _encode_fn=None):
    """ref :189-250 — writes f{v}.pth (features in model dtype) per augmentation view and label.pth (int64) once,
    then reloads each file to validate its shape."""
    clip_model = clip_bundle["clip_model"]
    cache_dir = _feature_cache_dir(cfg)
    num_views = int(cfg.get("aug_views", 1) or 1)
    expected_n = int(info.get("train_size", len(dl_tr.dataset)))
    writer = _is_writer()
    if writer:
        print("\n==== Feature Caching (pre-projection) ====")
        print({"cache_dir": str(cache_dir), "backbone": cfg.get("backbone", "RN50"), "dataset": cfg.get("dataset", "cs"),
               "shots": int(cfg.get("shots", 0) or 0), "seed": int(cfg.get("seed", 1) or 1), "aug_views": num_views,
               "expected_train_size": expected_n})
    if hasattr(clip_model, "eval"):
        clip_model.eval()
    for v in range(num_views):
        feats_t, labels_t = compute_image_features(clip_model, dl_tr, to_cpu=True, _encode_fn=_encode_fn)
        if writer:
            fpath = cache_dir / f"f{v}.pth"
            fpath.parent.mkdir(parents=True, exist_ok=True)
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
        _sync_ranks()
    if writer:
        print("\nFeature caching complete.")


def cache_openclip_embeddings(cfg: dict, model: torch.nn.Module, loader, split: str = "test",
                              checkpoint_path: Optional[str] = None, _encode_fn=None, _normalize_fn=None) -> Path:
    """ref :98-186 — embeddings.pt (optionally L2-normalised, model dtype), labels.pt (int64), metadata.csv,
    meta.json."""
    import pandas as pd
    normalize = bool(cfg.get("finetune", {}).get("cache_embeddings_normalize", True))
    cache_dir = _embedding_cache_dir(cfg, split)
    if hasattr(model, "eval"):
        model.eval()
    post = None
    if normalize:  # F.normalize(feats, dim=-1), ref :126-127 — in the model dtype, fp32 statistics
        post = _normalize_fn or (lambda f: ops.l2_normalize(f, 1e-12))
    feats_all, labels_all, meta = extract_loader(model, loader, post_fn=post, to_cpu=True, want_metadata=True,
                                                 encode_fn=_encode_fn)
    if _is_writer():
        cache_dir.mkdir(parents=True, exist_ok=True)
        torch.save(feats_all, cache_dir / "embeddings.pt")
        torch.save(labels_all, cache_dir / "labels.pt")
        labels_py = labels_all.tolist()   # one conversion, no per-sample .item()
        rows = [{"file_name": row.get("file_name", ""), "ground_truth_num_label": int(lab),
                 "ground_truth_word_label": row.get("plot_word_label", ""),
                 "ground_truth_L2_num_label": row.get("l2_label", -1)} for row, lab in zip(meta, labels_py)]
        columns = ["file_name", "ground_truth_num_label", "ground_truth_word_label", "ground_truth_L2_num_label"]
        pd.DataFrame(rows).reindex(columns=columns).to_csv(cache_dir / "metadata.csv", index=False)
        info = {"timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S"), "split": str(split), "normalized": normalize,
                "num_samples": int(feats_all.shape[0]), "dim": int(feats_all.shape[1]) if feats_all.ndim == 2 else None,
                "checkpoint_path": str(checkpoint_path) if checkpoint_path is not None else None,
                "cache_dir": str(cache_dir)}
        (cache_dir / "meta.json").write_text(json.dumps(info, indent=2))
        print("\n==== OpenCLIP Embedding Cache ====")
        print({"cache_dir": str(cache_dir), "embeddings": str(cache_dir / "embeddings.pt"),
               "labels": str(cache_dir / "labels.pt"), "metadata": str(cache_dir / "metadata.csv"),
               "num_samples": info["num_samples"], "dim": info["dim"], "normalized": normalize})
    _sync_ranks()
    return cache_dir


def _project_fn(proj, device):
    """x -> proj(x): a ``[width, embed]`` tensor, a module with a ``vit_proj`` parameter (VisProjViT,
    methods/ProLIP.py:31-41), or any callable."""
    p = getattr(proj, "vit_proj", proj)
    if torch.is_tensor(p):
        return p.detach().to(device), None
    return None, proj


def compute_image_features_test(clip_model, loader, proj, text_weights) -> float:
    """ref methods/utils.py:175-189 — zero-shot accuracy (%): encode -> proj -> normalise -> 100 * f @ W -> argmax."""
    device = _model_device(clip_model)
    p, call = _project_fn(proj, device)
    tw = text_weights.detach().to(device)

    def post(f):
        if call is not None:   # a generic projector module: its forward, then normalise + logits + argmax in the kernel
            _, _, idx, _ = ops.score(call(f), None, tw, 100.0, 1, want_emb=False, want_logits=False)
        else:
            _, _, idx, _ = ops.score(f, p, tw, 100.0, 1, want_emb=False, want_logits=False)
        return idx.to(torch.float32)   # [n, 1]; exact for C < 2**24, rides in the single all-gather

    preds, labels, _ = extract_loader(clip_model, loader, post_fn=post, to_cpu=True)
    if preds.shape[0] == 0:
        return 0.0
    return 100.0 * float((preds[:, 0].to(torch.int64) == labels).double().mean())


def pre_load_features(cfg, split, clip_model, loader):
    """ref utils.py:60-82 — ``encode_image`` -> ``f /= f.norm(dim=-1, keepdim=True)`` (NO eps: a zero row becomes NaN
    as in the reference) -> ``<cache_dir>/<split>_f.pt`` / ``<split>_l.pt``; loads them when ``cfg['load_pre_feat']``.
    Returns (features, labels) on the model's device like the reference."""
    if cfg["load_pre_feat"] is False:
        features, labels, _ = extract_loader(clip_model, loader, post_fn=lambda f: ops.l2_normalize(f, 0.0))
        if _is_writer():
            torch.save(features, cfg["cache_dir"] + "/" + split + "_f.pt")
            torch.save(labels, cfg["cache_dir"] + "/" + split + "_l.pt")
        _sync_ranks()
    else:
        features = torch.load(cfg["cache_dir"] + "/" + split + "_f.pt")
        labels = torch.load(cfg["cache_dir"] + "/" + split + "_l.pt")
    return features, labels


def build_cache_model(cfg, clip_model, train_loader_cache, task, proj):
    """ref methods/utils.py:31-45 — Tip-Adapter style cache: per augment epoch encode -> proj -> F.normalize, mean over
    the epochs, renormalise (no eps), transpose to ``[embed, N]``; values = one-hot labels in fp16.  (The reference's
    ``load_cache`` branch is broken — it passes undefined tensors to torch.load, methods/utils.py:57-60 — so only the
    build branch exists here.)"""
    if cfg["load_cache"] is not False:
        raise NotImplementedError("build_cache_model(load_cache=True) is a dead path in the reference "
                                  "(methods/utils.py:57-60 calls torch.load on undefined tensors)")
    device = _model_device(clip_model)
    p, call = _project_fn(proj, device)

    def post(f):   # proj(x) -> F.normalize(dim=-1), ref :38-39; fp32 embedding from the scoring kernel, model dtype out
        if call is not None:
            return ops.l2_normalize(call(f), 1e-12)
        emb, _, _, _ = ops.score(f, p, None, k=0)
        return emb.to(f.dtype)

    keys, values = [], None
    for augment_idx in range(cfg["augment_epoch"]):
        print("Augment Epoch: {:} / {:}".format(augment_idx, cfg["augment_epoch"]))
        feats, labels, _ = extract_loader(clip_model, train_loader_cache, post_fn=post)
        keys.append(feats.unsqueeze(0))
        if augment_idx == 0:
            values = labels
    cache_keys = torch.cat(keys, dim=0).mean(dim=0)
    cache_keys = ops.l2_normalize(cache_keys, 0.0)           # cache_keys /= cache_keys.norm(dim=-1, keepdim=True)
    cache_keys = cache_keys.permute(1, 0)
    cache_values = torch.nn.functional.one_hot(values).half()
    Path(cfg["cache_dir"]).mkdir(parents=True, exist_ok=True)
    return cache_keys, cache_values
