"""CPU oracle, torch edition — TEST INFRASTRUCTURE ONLY (same rules as oracle/clip_oracle.py).

The reference's CPU path IS PyTorch: ``clip.load(path, device='cpu')`` builds fp32 ``nn.Module``s and
``encode_image`` runs torch's own CPU operators (oneDNN / MKL GEMMs, fused SDPA).  This module restates that path
with the *same torch operators the reference calls*, in the same order and layouts (LND inside the transformer), but
flat over a state_dict instead of through the reference's module classes, so it can travel to the GPU box where
/root/reference does not exist.  It serves two purposes:

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time THIS (all host threads): it is what a user of the
    reference would run on the box's host cores, unlike the numpy restatement, which pays for numpy's unfused
    elementwise passes and is several times slower;
  * a second, independent restatement: ``tests/test_oracle_golden.py`` holds it to the same reference goldens as the
    numpy oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.  Every function cites
the reference file:line it follows (paths relative to the reference checkout).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # data/clip_transforms.py:22
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)  # data/clip_transforms.py:23


def to_torch_state(sd: dict) -> dict:
    """fp32 CPU tensors of the visual tower + scoring weights (clip/clip.py:134-137: ``model.float()`` on CPU)."""
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in sd.items() if k.startswith("visual.")}


def preprocess_pil(images_u8: np.ndarray, resolution: int) -> torch.Tensor:
    """data/clip_transforms.py:50-56 (= clip/clip.py:74-81) on ``Image.fromarray(uint8 HWC)`` inputs
    (data/dataloader.py:415): Resize(R, BICUBIC) -> CenterCrop(R) -> ToTensor -> Normalize, one PIL image at a time in
    the calling thread, as the reference's DataLoader does with ``num_workers: 0`` (configs/cs.yaml:17)."""
    from PIL import Image
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF

    out = []
    for a in images_u8:
        im = Image.fromarray(np.ascontiguousarray(a))
        im = TF.resize(im, resolution, interpolation=InterpolationMode.BICUBIC)
        im = TF.center_crop(im, [resolution, resolution])
        t = TF.pil_to_tensor(im).to(torch.float32).div_(255.0)  # ToTensor: HWC u8 -> CHW f32 / 255
        out.append(TF.normalize(t, CLIP_MEAN, CLIP_STD))
    return torch.stack(out)


def _layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """clip/model.py:151-157 — nn.LayerNorm in fp32, eps 1e-5."""
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)


def _block(x: torch.Tensor, sd: dict, pre: str, heads: int) -> torch.Tensor:
    """clip/model.py:179-186 — x + attn(ln_1 x); x + c_proj(QuickGELU(c_fc(ln_2 x))).  x is [L, N, D] (LND)."""
    d = x.shape[-1]
    y = _layer_norm(x, sd[pre + "ln_1.weight"], sd[pre + "ln_1.bias"])
    # nn.MultiheadAttention(d, heads)(y, y, y, need_weights=False, attn_mask=None)[0]   (:169,181)
    a, _ = F.multi_head_attention_forward(
        y, y, y, d, heads, sd[pre + "attn.in_proj_weight"], sd[pre + "attn.in_proj_bias"], None, None, False, 0.0,
        sd[pre + "attn.out_proj.weight"], sd[pre + "attn.out_proj.bias"], training=False, need_weights=False)
    x = x + a
    y = _layer_norm(x, sd[pre + "ln_2.weight"], sd[pre + "ln_2.bias"])
    h = F.linear(y, sd[pre + "mlp.c_fc.weight"], sd[pre + "mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)                                           # QuickGELU, :160-162
    return x + F.linear(h, sd[pre + "mlp.c_proj.weight"], sd[pre + "mlp.c_proj.bias"])


@torch.no_grad()
def encode_image(sd: dict, images: torch.Tensor) -> torch.Tensor:
    """clip/model.py:216-235 (VisionTransformer.forward) through CLIP.encode_image (:335-336): PRE-projection features
    ``ln_post(x[:, 0, :])`` [N, width].  ``sd``: to_torch_state(...)."""
    w = sd["visual.conv1.weight"]
    x = F.conv2d(images.to(w.dtype), w, None, stride=w.shape[-1])              # :217
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                 # :218-219
    cls = sd["visual.class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]         # :220-221
    x = _layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])   # :222
    x = x.permute(1, 0, 2)                                                     # NLD -> LND, :224
    heads = x.shape[-1] // 64                                                  # :267
    layers = len([k for k in sd if k.endswith(".attn.in_proj_weight")])
    for i in range(layers):                                                    # :196-197
        x = _block(x, sd, f"visual.transformer.resblocks.{i}.", heads)
    x = x.permute(1, 0, 2)                                                     # :226
    return _layer_norm(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])  # :228


@torch.no_grad()
def score(feats: torch.Tensor, proj: torch.Tensor, text_w: torch.Tensor, scale: float = 100.0, k: int = 1):
    """methods/ProLIP.py:40 (x @ vit_proj) -> methods/utils.py:184 (F.normalize) -> :185 (100. * f @ text_weights) ->
    :17 topk / :186 argmax.  Returns (emb [N,E], logits [N,C], topk_idx [N,k])."""
    emb = F.normalize(feats.float() @ proj.float(), dim=-1)
    logits = scale * emb @ text_w.float()
    return emb, logits, logits.topk(k, 1, True, True)[1]


def reference_pass(sd: dict, text_w: torch.Tensor, images_u8: np.ndarray, resolution: int):
    """One batch of the benchmark workload on the CPU: uint8 HWC -> preprocess -> encode_image -> proj -> L2-norm ->
    x100 logits -> argmax (methods/utils.py:175-189, compute_image_features_test)."""
    x = preprocess_pil(images_u8, resolution)
    return score(encode_image(sd, x), sd["visual.proj"], text_w, 100.0, 1)
