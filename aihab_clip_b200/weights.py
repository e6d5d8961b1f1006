"""Deterministic random-init CLIP ViT checkpoints in the reference's state_dict layout.

No pretrained weights are reachable offline, so parity and benchmarks run on random-init models.
The reference builds them with ``torch.manual_seed(s); CLIP(...)`` (clip/model.py:238-321), which needs the
reference source; this module regenerates an equivalent checkpoint from numpy's PCG64 stream so that the GPU box
(which has no /root/reference) sees bit-identical weights to the ones the golden fixtures were made from.

Key names and shapes follow ``build_model`` (clip/model.py:396-433); the tensors that ``convert_weights``
(clip/model.py:372-393) casts to fp16 are rounded through fp16 here, so the dict is a fixed point of
``clip.load`` (clip/clip.py:134-137).
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class Geometry:
    """ctor arguments of the reference ``CLIP`` (clip/model.py:239-252)."""
    name: str
    embed_dim: int
    image_resolution: int
    vision_layers: int
    vision_width: int
    vision_patch_size: int
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_heads: int = 8
    transformer_layers: int = 12

    @property
    def grid(self) -> int:
        return self.image_resolution // self.vision_patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def vision_heads(self) -> int:
        return self.vision_width // 64


# SURVEY.md §8 geometry table (+ a tiny geometry for fast CPU tests)
GEOMETRIES = {
    "ViT-B/32": Geometry("ViT-B/32", 512, 224, 12, 768, 32),
    "ViT-B/16": Geometry("ViT-B/16", 512, 224, 12, 768, 16),
    "ViT-L/14": Geometry("ViT-L/14", 768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": Geometry("ViT-L/14@336px", 768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-tiny/16": Geometry("ViT-tiny/16", 64, 64, 2, 128, 16, 77, 49408, 64, 1, 2),
    "ViT-tiny/14": Geometry("ViT-tiny/14", 128, 56, 3, 256, 14, 77, 49408, 64, 1, 1),
    # ViT-L/14 widths and sequence lengths (257 / 577 tokens) with 2 blocks only: parity cases for configs 3 and 4
    "ViT-L-mini/14": Geometry("ViT-L-mini/14", 768, 224, 2, 1024, 14, 77, 49408, 64, 1, 1),
    "ViT-L-mini/14@336px": Geometry("ViT-L-mini/14@336px", 768, 336, 2, 1024, 14, 77, 49408, 64, 1, 1),
}


def _fp16_round(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float16).astype(np.float32)


def _block(rng: np.random.Generator, prefix: str, width: int, sd: "OrderedDict[str, np.ndarray]") -> None:
    """One ResidualAttentionBlock (clip/model.py:165-177): MHA in/out projections, two LayerNorms, the MLP."""
    d = width
    xav = float(np.sqrt(6.0 / (d + 3 * d)))
    lin = float(1.0 / np.sqrt(d))
    lin4 = float(1.0 / np.sqrt(4 * d))
    sd[prefix + "ln_1.weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    sd[prefix + "ln_1.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
    sd[prefix + "attn.in_proj_weight"] = _fp16_round(rng.uniform(-xav, xav, (3 * d, d)).astype(np.float32))
    sd[prefix + "attn.in_proj_bias"] = _fp16_round((0.02 * rng.standard_normal(3 * d)).astype(np.float32))
    sd[prefix + "attn.out_proj.weight"] = _fp16_round(rng.uniform(-lin, lin, (d, d)).astype(np.float32))
    sd[prefix + "attn.out_proj.bias"] = _fp16_round((0.02 * rng.standard_normal(d)).astype(np.float32))
    sd[prefix + "ln_2.weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    sd[prefix + "ln_2.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
    sd[prefix + "mlp.c_fc.weight"] = _fp16_round(rng.uniform(-lin, lin, (4 * d, d)).astype(np.float32))
    sd[prefix + "mlp.c_fc.bias"] = _fp16_round(rng.uniform(-lin, lin, 4 * d).astype(np.float32))
    sd[prefix + "mlp.c_proj.weight"] = _fp16_round(rng.uniform(-lin4, lin4, (d, 4 * d)).astype(np.float32))
    sd[prefix + "mlp.c_proj.bias"] = _fp16_round(rng.uniform(-lin4, lin4, d).astype(np.float32))


def make_state_dict_np(geom: Geometry | str, seed: int = 0, with_text: bool = True) -> "OrderedDict[str, np.ndarray]":
    """numpy state_dict (fp32) for ``geom``; visual.* always, text tower when ``with_text``."""
    if isinstance(geom, str):
        geom = GEOMETRIES[geom]
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    d, p, L = geom.vision_width, geom.vision_patch_size, geom.tokens
    scale = d ** -0.5
    kconv = float(1.0 / np.sqrt(3 * p * p))
    sd["visual.class_embedding"] = (scale * rng.standard_normal(d)).astype(np.float32)
    sd["visual.positional_embedding"] = (scale * rng.standard_normal((L, d))).astype(np.float32)
    sd["visual.proj"] = _fp16_round((scale * rng.standard_normal((d, geom.embed_dim))).astype(np.float32))
    sd["visual.conv1.weight"] = _fp16_round(rng.uniform(-kconv, kconv, (d, 3, p, p)).astype(np.float32))
    sd["visual.ln_pre.weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    sd["visual.ln_pre.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
    for i in range(geom.vision_layers):
        _block(rng, f"visual.transformer.resblocks.{i}.", d, sd)
    sd["visual.ln_post.weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    sd["visual.ln_post.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
    if with_text:
        w = geom.transformer_width
        trng = np.random.Generator(np.random.PCG64(seed + 7919))
        sd["positional_embedding"] = (0.01 * trng.standard_normal((geom.context_length, w))).astype(np.float32)
        sd["text_projection"] = _fp16_round((w ** -0.5 * trng.standard_normal((w, geom.embed_dim))).astype(np.float32))
        sd["logit_scale"] = np.asarray(np.log(1 / 0.07), dtype=np.float32)
        sd["token_embedding.weight"] = (0.02 * trng.standard_normal((geom.vocab_size, w), dtype=np.float32))
        for i in range(geom.transformer_layers):
            _block(trng, f"transformer.resblocks.{i}.", w, sd)
        sd["ln_final.weight"] = (1.0 + 0.1 * trng.standard_normal(w)).astype(np.float32)
        sd["ln_final.bias"] = (0.05 * trng.standard_normal(w)).astype(np.float32)
    return sd


def make_state_dict(geom: Geometry | str, seed: int = 0, with_text: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """torch (CPU fp32) version of :func:`make_state_dict_np`; what ``torch.save`` + ``clip.load(path)`` consume."""
    return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v))) for k, v in make_state_dict_np(geom, seed, with_text).items())


def state_dict_digest(sd) -> str:
    """sha256 over the visual.* tensors, to pin golden fixtures to the generator."""
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        if not k.startswith("visual."):
            continue
        v = sd[k]
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        h.update(k.encode())
        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return h.hexdigest()


def synthetic_images_u8(n: int, side: int, seed: int = 1234, start: int = 0, smooth: bool = False) -> np.ndarray:
    """uint8 [n, side, side, 3] synthetic images keyed by (seed, global image index) so that any sharding of an
    extraction job sees identical pixels.  ``smooth`` gives low-frequency images (7x7 noise
    upsampled) whose zero-shot logits are not near-tied."""
    out = np.empty((n, side, side, 3), dtype=np.uint8)
    for i in range(n):
        rng = np.random.Generator(np.random.PCG64([seed, start + i]))
        if smooth:
            coarse = rng.uniform(0, 255, (7, 7, 3))
            idx = np.linspace(0, 6, side)
            i0 = np.clip(np.floor(idx).astype(int), 0, 5)
            f = (idx - i0)[:, None]
            rows = coarse[i0] * (1 - f[:, :, None]) + coarse[i0 + 1] * f[:, :, None]  # [side, 7, 3]
            img = rows[:, i0] * (1 - f[None, :, :]) + rows[:, i0 + 1] * f[None, :, :]
            out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        else:
            out[i] = rng.integers(0, 256, (side, side, 3), dtype=np.uint8)
    return out
