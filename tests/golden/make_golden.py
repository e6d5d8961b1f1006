"""Generates tests/golden/*.npz / *.json by running the UNMODIFIED reference (imported from /root/reference) on
CPU fp32.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Weights come from aihab_clip_b200.weights (numpy PCG64, reproducible everywhere), are saved as a state_dict and
loaded through the reference's own ``clip.load(path, device='cpu')`` — the real API surface.
Inputs are the seeded synthetic uint8 images of aihab_clip_b200.weights.synthetic_images_u8 pushed through the
reference's ``build_clip_transforms(is_train=False)`` on PIL images.  Only outputs are stored; tests regenerate the
inputs from the same seeds.
"""
from __future__ import annotations

import hashlib
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from aihab_clip_b200.weights import GEOMETRIES, make_state_dict, state_dict_digest, synthetic_images_u8  # noqa: E402


def import_reference():
    """ftfy is the only missing import of the reference's clip package; fix_text is the identity on ASCII prompts
    (clip/simple_tokenizer.py:50-53).  evaluation.py additionally imports plotting / metric packages it does not
    need for the functions exercised here."""
    for name in ("ftfy",):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.fix_text = lambda s: s
            sys.modules[name] = m
    for name in ("seaborn", "torcheval", "torcheval.metrics", "timm", "timm.data", "open_clip"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.MulticlassF1Score = m.MulticlassConfusionMatrix = object
                m.create_transform = m.resolve_data_config = None
                sys.modules[name] = m
    try:
        import matplotlib  # noqa: F401
    except Exception:
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, str(REF))
    import clip  # noqa: E402  (the reference package)
    assert Path(clip.__file__).resolve().is_relative_to(REF), clip.__file__
    return clip


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_reference_model(ref_clip, geom_name: str, seed: int):
    sd = make_state_dict(geom_name, seed)
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save(sd, f.name)
        state, model, preprocess = ref_clip.load(f.name, device="cpu")
    return sd, state, model, preprocess


def ref_preprocess(u8: np.ndarray, R: int) -> torch.Tensor:
    from data.clip_transforms import build_clip_transforms
    tf = build_clip_transforms({}, is_train=False, resolution=R)
    return torch.stack([tf(Image.fromarray(im)) for im in u8])


def model_case(ref_clip, geom_name: str, seed: int, n: int, side: int, templates, classnames, out: dict, tag: str,
               layers_trace: bool = False):
    import utils as ref_utils  # reference utils.py (clip_classifier)
    geom = GEOMETRIES[geom_name]
    sd, state, model, _ = load_reference_model(ref_clip, geom_name, seed)
    R = geom.image_resolution
    u8 = np.concatenate([synthetic_images_u8(n // 2, side, seed=1234), synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])
    x = ref_preprocess(u8, R)
    with torch.no_grad():
        feats = model.encode_image(x)                                         # methods/utils.py:162
        texts, w_before, text_w = ref_utils.clip_classifier(classnames, templates, model)  # utils.py:31
        emb = F.normalize(feats @ state["visual.proj"], dim=-1)               # methods/ProLIP.py:40, methods/utils.py:184
        logits = 100. * emb @ text_w                                          # methods/utils.py:185
        top3 = logits.topk(3, 1, True, True)[1]
    out[f"{tag}_digest"] = np.frombuffer(bytes.fromhex(state_dict_digest(sd)), dtype=np.uint8)
    out[f"{tag}_pre_sha"] = np.frombuffer(bytes.fromhex(sha(x.numpy())), dtype=np.uint8)
    out[f"{tag}_feats"] = feats.numpy()
    out[f"{tag}_emb"] = emb.numpy()
    out[f"{tag}_text_w"] = text_w.numpy()
    out[f"{tag}_texts"] = texts.numpy()
    out[f"{tag}_logits"] = logits.numpy()
    out[f"{tag}_argmax"] = logits.argmax(dim=1).numpy()
    out[f"{tag}_top3"] = top3.numpy()
    if layers_trace:  # residual stream after ln_pre and after every block (hooks on the reference modules)
        acts = []
        hooks = [model.visual.ln_pre.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone()))]
        for blk in model.visual.transformer.resblocks:
            hooks.append(blk.register_forward_hook(lambda m, i, o: acts.append(o.permute(1, 0, 2).detach().clone())))
        with torch.no_grad():
            model.encode_image(x[:2])
        for h in hooks:
            h.remove()
        tr = torch.stack(acts)  # [layers+1, 2, L, D]
        out[f"{tag}_trace"] = (tr if tr[0].numel() < 20000 else tr[:, :1, :4]).numpy()  # big models: image 0, tokens 0..3
    # text tower fixtures: tokens and both encode_text outputs for the first 3 classes
    prompts = [templates[0].format(c.replace("_", " ")) for c in classnames[:3]]
    tok = ref_clip.tokenize(prompts)
    with torch.no_grad():
        tb, tx = model.encode_text(tok)
    out[f"{tag}_tok"] = tok.numpy()
    out[f"{tag}_text_before"] = tb.numpy()
    out[f"{tag}_text_emb"] = tx.numpy()
    print(tag, "feats", tuple(feats.shape), "logits", tuple(logits.shape), "argmax", logits.argmax(1).tolist())


def main():
    ref_clip = import_reference()
    from data.templates import CS_CLASSNAMES, CS_TEMPLATES, gen_prompts
    from data import build_l3_to_l2_map, NAME_LABEL_L2

    meta = {"classnames": list(CS_CLASSNAMES), "templates": list(CS_TEMPLATES)}
    l3_to_l2, l2_names = build_l3_to_l2_map()
    meta["l3_to_l2"], meta["l2_names"] = l3_to_l2, l2_names
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        meta["gen_prompts_flat"] = gen_prompts(use_hierarchy=False, use_descriptive=False)
        meta["gen_prompts_hier"] = gen_prompts(use_hierarchy=True, use_descriptive=False)
    bench_templates = [f"a habitat photo {i} of {{}}." for i in range(80)]
    meta["bench_templates"] = bench_templates

    # ---- tokenizer fixtures
    strings = [t.format(c.replace("_", " ")) for c in CS_CLASSNAMES for t in CS_TEMPLATES]
    strings += ["A photo of a CAT!!  with   spaces", "fen, marsh &amp; swamp's edge", "supra-littoral rock 42 m2", "",
                "é ü ß naïve café", "x" * 60]
    meta["tok_strings"] = strings
    gold = {"tok_tokens": ref_clip.tokenize(strings).numpy()}
    long_text = "habitat " * 100
    gold["tok_truncated"] = ref_clip.tokenize([long_text], truncate=True).numpy()
    meta["tok_long"] = long_text

    # ---- model fixtures
    model_case(ref_clip, "ViT-tiny/16", 0, 6, 64, CS_TEMPLATES, CS_CLASSNAMES, gold, "tiny16", layers_trace=True)
    model_case(ref_clip, "ViT-tiny/14", 1, 6, 111, CS_TEMPLATES, CS_CLASSNAMES, gold, "tiny14")
    model_case(ref_clip, "ViT-B/32", 0, 8, 439, CS_TEMPLATES, CS_CLASSNAMES, gold, "b32", layers_trace=True)

    # config 1 text head: 18 classes x 80 templates (BASELINE.json configs[0]); one-time ~35 s on CPU
    import utils as ref_utils
    _, _, model, _ = load_reference_model(ref_clip, "ViT-B/32", 0)
    with torch.no_grad():
        _, _, w1880 = ref_utils.clip_classifier(CS_CLASSNAMES[:18], bench_templates, model)
    gold["b32_text_w_18x80"] = w1880.numpy()

    # ---- preprocessing fixtures: sha256 of the reference transform output + a strided sample
    pre_cases = [(439, 439, 224), (256, 256, 224), (480, 480, 336), (300, 500, 224), (500, 300, 224), (64, 64, 64),
                 (111, 111, 56), (200, 200, 224), (224, 224, 224), (225, 338, 224)]
    meta["pre_cases"] = pre_cases
    for (h, w, R) in pre_cases:
        rng = np.random.Generator(np.random.PCG64([99, h, w, R]))
        u8 = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        u8[1] = synthetic_images_u8(1, max(h, w), seed=5, smooth=True)[0][:h, :w]
        y = ref_preprocess(u8, R).numpy()
        gold[f"pre_{h}x{w}_{R}_sha"] = np.frombuffer(bytes.fromhex(sha(y)), dtype=np.uint8)
        gold[f"pre_{h}x{w}_{R}_sample"] = y[:, :, ::37, ::41].copy()

    # ---- evaluation fixtures (aihab_utils/evaluation.py, methods/utils.py:16-21)
    from aihab_utils.evaluation import aggregate_logits_to_l2, map_l3_targets_to_l2, ClassificationTracker
    from methods.utils import cls_acc
    g = torch.Generator().manual_seed(3)
    ev_logits = torch.randn(64, 20, generator=g) * 3
    ev_logits[5, 7] = ev_logits[5, 2]          # exact tie -> lowest index first
    ev_labels = torch.randint(0, 20, (64,), generator=g)
    gold["ev_logits"], gold["ev_labels"] = ev_logits.numpy(), ev_labels.numpy()
    for red in ("sum", "mean", "logsumexp"):
        gold[f"ev_l2_{red}"] = aggregate_logits_to_l2(ev_logits, l3_to_l2, len(NAME_LABEL_L2), reduce=red).numpy()
    gold["ev_l2_targets"] = map_l3_targets_to_l2(ev_labels, l3_to_l2).numpy()
    c3, i3, p3 = ClassificationTracker().top3_metrics(ev_logits, ev_labels)
    gold["ev_top3_correct"], gold["ev_top3_idx"], gold["ev_top3_probs"] = np.asarray(int(c3)), i3.numpy(), p3.numpy()
    gold["ev_acc1"] = np.asarray(cls_acc(ev_logits, ev_labels, 1))
    gold["ev_acc3"] = np.asarray(cls_acc(ev_logits, ev_labels, 3))

    # ---- cache directory naming (aihab_utils/feature_cache.py:15-65)
    from aihab_utils.feature_cache import _canonical_backbone_name, _embedding_cache_dir, _feature_cache_dir
    cfgs = [
        {"root_path": "/data/x", "backbone": "ViT-B/16", "dataset": "cs", "shots": 16, "seed": 3},
        {"root_path": "./", "backbone": "ViT-B/32", "dataset": "cs", "shots": 0, "seed": 0},
        {"backbone": "ViT-L/14", "clip_backend": "openclip", "open_clip_model": "hf-hub:timm/ViT-SO400M-14-SigLIP",
         "finetune": {"cache_embeddings_dir": "feat_cache_vis"}, "seed": 5},
        {"root_path": "/r", "backbone": "ViT-B/16", "finetune": {"cache_embeddings_dir": "/abs/emb"}, "seed": 2},
    ]
    meta["cache_cfgs"] = cfgs
    meta["cache_feature_dirs"] = [str(_feature_cache_dir(c)) for c in cfgs]
    meta["cache_embedding_dirs"] = [str(_embedding_cache_dir(c, "Test")) for c in cfgs]
    meta["cache_backbone_names"] = {b: _canonical_backbone_name(b) for b in
                                    ["ViT-B/16", "ViT-B/32", "ViT-L/14", "RN50", "", "hf-hub:timm/x y:z"]}

    # ---- large agreement set (BASELINE.json gate: zero-shot argmax agreement >= 99.9 %): 4096 images, ViT-B/32,
    # 224 px inputs (no resize), 20 shipped classes; logits stored so that max |dlogit| is measured on the same set
    _, state32, model32, _ = load_reference_model(ref_clip, "ViT-B/32", 0)
    n_agree = 4096
    u8 = np.concatenate([synthetic_images_u8(n_agree // 2, 224, seed=777),
                         synthetic_images_u8(n_agree // 2, 224, seed=777, start=n_agree // 2, smooth=True)])
    tw = torch.from_numpy(gold["b32_text_w"])
    chunks = []
    with torch.no_grad():
        for i in range(0, n_agree, 64):
            x = ref_preprocess(u8[i:i + 64], 224)
            f = model32.encode_image(x)
            emb = F.normalize(f @ state32["visual.proj"], dim=-1)
            chunks.append((100. * emb @ tw).numpy())
    gold["agree_logits"] = np.concatenate(chunks).astype(np.float32)
    print("agreement set:", gold["agree_logits"].shape, "classes hit:", np.unique(gold["agree_logits"].argmax(1)).size)

    np.savez_compressed(HERE / "reference_outputs.npz", **gold)
    (HERE / "reference_meta.json").write_text(json.dumps(meta, indent=1))
    print("wrote", HERE / "reference_outputs.npz", (HERE / "reference_outputs.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
