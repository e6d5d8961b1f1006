"""CPU oracle for the aihab-clip hot path — TEST INFRASTRUCTURE ONLY.

A plain numpy fp32 restatement of the reference algorithm, one function per reference site, each citing the
reference file:line it follows (paths relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module, and only as the checker / the timed CPU baseline; the product (``aihab_clip_b200``) never does and has
no CPU fallback.

Pinning: the reference ships no tests and no golden vectors for this path, so the oracle is
pinned against outputs of the reference itself: ``tests/golden/make_golden.py`` imports the reference from
/root/reference in the build container and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those fixtures.

Third-party arithmetic restated here because its source is not under /root/reference (no version is pinned by
the reference; container versions: torch 2.11.0, torchvision 0.26.0, Pillow 12.2.0):
  * torch ``nn.LayerNorm`` / ``nn.MultiheadAttention`` / ``nn.Linear`` / ``nn.Conv2d`` / ``F.normalize`` / ``topk``
  * torchvision ``v2.Resize`` / ``v2.CenterCrop`` / ``v2.ToTensor`` / ``v2.Normalize``
  * Pillow ``ImagingResample`` (src/libImaging/Resample.c): 8-bit two-pass fixed-point bicubic
"""
from __future__ import annotations

import math

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # data/clip_transforms.py:22
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)  # data/clip_transforms.py:23


# ----------------------------------------------------------------------------------------------- model pieces
def layer_norm(x: np.ndarray, weight: np.ndarray, bias: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """clip/model.py:151-157 — nn.LayerNorm over the last dim, computed in fp32, eps 1e-5, affine."""
    x = x.astype(np.float32)
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    xc = x - mean
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=np.float32)
    return (xc / np.sqrt(var + np.float32(eps))) * weight + bias


def quick_gelu(x: np.ndarray) -> np.ndarray:
    """clip/model.py:160-162 — x * sigmoid(1.702 x)."""
    return x * (1.0 / (1.0 + np.exp(-1.702 * x, dtype=np.float32)))


def _softmax(s: np.ndarray) -> np.ndarray:
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s, dtype=np.float32)
    return e / e.sum(axis=-1, keepdims=True, dtype=np.float32)


def multihead_attention(x: np.ndarray, in_w, in_b, out_w, out_b, heads: int, causal: bool = False) -> np.ndarray:
    """clip/model.py:169,179-181 — nn.MultiheadAttention(x, x, x, need_weights=False): fused in_proj
    ([3D, D], rows q,k,v), heads of D/heads, softmax(q k^T / sqrt(hd)) v, out_proj.  x: [N, L, D].
    ``causal`` adds the text tower's -inf upper-triangular mask (clip/model.py:323-329)."""
    n, L, d = x.shape
    hd = d // heads
    qkv = x @ in_w.T + in_b
    q, k, v = np.split(qkv, 3, axis=-1)

    def heads_first(t):
        return t.reshape(n, L, heads, hd).transpose(0, 2, 1, 3)

    q, k, v = heads_first(q), heads_first(k), heads_first(v)
    s = (q @ k.transpose(0, 1, 3, 2)) * np.float32(1.0 / math.sqrt(hd))
    if causal:
        s = s + np.triu(np.full((L, L), -np.inf, dtype=np.float32), 1)
    o = _softmax(s) @ v
    o = o.transpose(0, 2, 1, 3).reshape(n, L, d)
    return o @ out_w.T + out_b


def residual_block(x: np.ndarray, sd: dict, prefix: str, heads: int, causal: bool = False) -> np.ndarray:
    """clip/model.py:183-186 — x + attn(ln_1 x); x + c_proj(quickgelu(c_fc(ln_2 x)))."""
    y = layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    x = x + multihead_attention(y, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"],
                                sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"], heads, causal)
    y = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    hdn = quick_gelu(y @ sd[prefix + "mlp.c_fc.weight"].T + sd[prefix + "mlp.c_fc.bias"])
    return x + hdn @ sd[prefix + "mlp.c_proj.weight"].T + sd[prefix + "mlp.c_proj.bias"]


def patch_embed(images: np.ndarray, conv_w: np.ndarray) -> np.ndarray:
    """clip/model.py:204,217-219 — stride-p p x p conv without bias == GEMM over im2col rows with column order
    (c, ky, kx).  images [N,3,R,R] -> [N, g*g, D]."""
    n, c, R, _ = images.shape
    d, _, p, _ = conv_w.shape
    g = R // p
    rows = images.reshape(n, c, g, p, g, p).transpose(0, 2, 4, 1, 3, 5).reshape(n, g * g, c * p * p)
    return rows @ conv_w.reshape(d, -1).T


def encode_image(sd: dict, images: np.ndarray, return_layers: bool = False):
    """clip/model.py:216-235 (VisionTransformer.forward) via CLIP.encode_image (:335-336): returns the
    PRE-projection features ln_post(x[:, 0, :]) of shape [N, width]; visual.proj is not applied."""
    sd = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items() if k.startswith("visual.")}
    images = np.asarray(images, dtype=np.float32)
    x = patch_embed(images, sd["visual.conv1.weight"])
    n, _, d = x.shape
    cls = np.broadcast_to(sd["visual.class_embedding"], (n, 1, d))           # :220
    x = np.concatenate([cls, x], axis=1) + sd["visual.positional_embedding"]  # :221
    x = layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])   # :222
    layers = len([k for k in sd if k.endswith(".attn.in_proj_weight")])
    heads = d // 64                                                           # :267
    trace = [x]
    for i in range(layers):                                                   # :196-197
        x = residual_block(x, sd, f"visual.transformer.resblocks.{i}.", heads)
        if return_layers:
            trace.append(x)
    out = layer_norm(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])  # :228
    return (out, trace) if return_layers else out


def encode_text(sd: dict, tokens: np.ndarray):
    """clip/model.py:338-353 — token + positional embedding, causal transformer, ln_final, EOT pooling at
    text.argmax(-1), @ text_projection.  Returns (x_before_proj [T, W], x [T, E])."""
    sd = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items() if not k.startswith("visual.")}
    tokens = np.asarray(tokens)
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]
    w = x.shape[-1]
    layers = len([k for k in sd if k.startswith("transformer.") and k.endswith(".attn.in_proj_weight")])
    heads = w // 64                                                           # build_model, :415
    for i in range(layers):
        x = residual_block(x, sd, f"transformer.resblocks.{i}.", heads, causal=True)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    before = x[np.arange(x.shape[0]), tokens.argmax(axis=-1)]
    return before, before @ sd["text_projection"]


def text_head(class_embeddings: list) -> np.ndarray:
    """utils.py:45-54 (clip_classifier) — per class: normalise the T prompt embeddings, mean over templates,
    renormalise; stack to [E, C]."""
    cols = []
    for e in class_embeddings:
        e = e / np.linalg.norm(e, axis=-1, keepdims=True)
        m = e.mean(axis=0)
        cols.append(m / np.linalg.norm(m))
    return np.stack(cols, axis=1).astype(np.float32)


# ----------------------------------------------------------------------------------------------- scoring
def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """methods/utils.py:184 — F.normalize(x, dim=-1): x / max(||x||_2, eps)."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True, dtype=np.float32))
    return x / np.maximum(n, np.float32(eps))


def topk_indices(logits: np.ndarray, k: int) -> np.ndarray:
    """methods/utils.py:17 / aihab_utils/evaluation.py:262 — topk(k, dim=1, largest, sorted); exact ties resolve
    to the lowest index first (what torch returns on CPU for these sizes)."""
    order = np.argsort(-logits, axis=1, kind="stable")
    return order[:, :k].astype(np.int64)


def score(feats: np.ndarray, proj, text_w, scale: float = 100.0, k: int = 1):
    """methods/ProLIP.py:40 (x @ vit_proj) -> methods/utils.py:184 (normalize) -> :185 (100. * f @ text_weights)
    -> :186 argmax / methods/utils.py:17 topk.  Returns (emb [N,E], logits [N,C], topk_idx [N,k])."""
    f = np.asarray(feats, dtype=np.float32)
    emb = f @ np.asarray(proj, dtype=np.float32) if proj is not None else f
    emb = l2_normalize(emb)
    logits = (np.float32(scale) * emb) @ np.asarray(text_w, dtype=np.float32)
    return emb, logits, topk_indices(logits, k)


def aggregate_logits_to_l2(logits_l3: np.ndarray, l3_to_l2, num_l2: int, reduce: str = "mean") -> np.ndarray:
    """aihab_utils/evaluation.py:92-142 — per-L2-group sum / mean / logsumexp of L3 logits."""
    l3_to_l2 = list(l3_to_l2)
    if logits_l3.shape[1] != len(l3_to_l2):
        raise ValueError("class count mismatch")
    if reduce not in {"sum", "mean", "logsumexp"}:
        raise ValueError(f"Unsupported reduce='{reduce}'")
    n = logits_l3.shape[0]
    if reduce == "logsumexp":
        out = np.full((n, num_l2), -np.inf, dtype=np.float32)
        for l3, l2 in enumerate(l3_to_l2):
            out[:, l2] = np.logaddexp(out[:, l2], logits_l3[:, l3]).astype(np.float32)
        return out
    out = np.zeros((n, num_l2), dtype=np.float32)
    counts = np.zeros(num_l2, dtype=np.float32)
    for l3, l2 in enumerate(l3_to_l2):
        out[:, l2] += logits_l3[:, l3]
        counts[l2] += 1
    if reduce == "mean":
        out = out / np.maximum(counts, 1)
    return out


def top3_metrics(outputs: np.ndarray, labels: np.ndarray):
    """aihab_utils/evaluation.py:261-273 — top-3 indices, softmax probabilities gathered at them, #correct."""
    idx = topk_indices(outputs, 3)
    probs = _softmax(outputs.astype(np.float32))
    top3 = np.take_along_axis(probs, idx, axis=1)
    correct = int((idx == labels[:, None]).any(axis=1).sum())
    return correct, idx, top3


def prototype_scores(emb: np.ndarray, labels: np.ndarray, prototypes: np.ndarray, owner: np.ndarray):
    """tools/outlier_cleaning.py:553-668 (MultiPrototypeScorer.score_prototype_distance, tensor part) — cosine
    similarity of L2-normalised embeddings to the nearest prototype of their own class (index inside the class block,
    first maximum), best similarity to any other class's prototype (NaN if there is none) and the margin.  With one
    prototype per class it is SingleCentroidScorer.score_centroid_distance (:296-337).
    prototypes [P, E] with class blocks contiguous, owner [P] = class id of each prototype."""
    emb = np.asarray(emb, dtype=np.float32)
    sim = emb @ np.asarray(prototypes, dtype=np.float32).T
    labels = np.asarray(labels)
    owner = np.asarray(owner)
    same = owner[None, :] == labels[:, None]
    own = np.where(same, sim, -np.inf)
    best = own.argmax(axis=1)                      # first maximum
    sim_own = own[np.arange(len(labels)), best]
    first = np.array([np.flatnonzero(owner == c)[0] for c in labels])
    other = np.where(same, -np.inf, sim).max(axis=1)
    other = np.where(np.isinf(other), np.nan, other).astype(np.float32)
    return sim_own.astype(np.float32), (best - first).astype(np.int64), other, (sim_own - other).astype(np.float32)


def cls_acc(output: np.ndarray, target: np.ndarray, topk: int = 1) -> float:
    """methods/utils.py:16-21 — top-k accuracy in percent."""
    pred = topk_indices(output, topk)
    return 100.0 * float((pred == target[:, None]).any(axis=1).sum()) / target.shape[0]


# ----------------------------------------------------------------------------------------------- preprocessing
def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2.0):
    per output index the first input index, the tap count and the 22-bit fixed-point taps."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            f = v * (1 << 22)
            kk[xx, x] = int(-0.5 + f) if f < 0 else int(0.5 + f)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One Pillow pass along axis 0 of a uint8 array: 2^21 + sum(px * k) >> 22, clamped to [0, 255]."""
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for i in range(out_size):
        lo, cnt = bounds[i]
        acc = np.tensordot(kk[i, :cnt].astype(np.int64), src[lo:lo + cnt], axes=(0, 0)) + (1 << 21)
        out[i] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return out


def pil_bicubic_resize(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow Image.resize(BICUBIC) on an HWC uint8 image: horizontal pass first, then vertical; a pass whose
    input and output sizes are equal is skipped (ImagingResample)."""
    h, w = img.shape[:2]
    if w != out_w:
        img = _resample_axis0(img.transpose(1, 0, 2), out_w).transpose(1, 0, 2)
    if h != out_h:
        img = _resample_axis0(img, out_h)
    return img


def resized_size(h: int, w: int, size: int):
    """torchvision _compute_resized_output_size for an int size: short side -> size, long -> int(size*long/short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def clip_preprocess(img_u8: np.ndarray, resolution: int) -> np.ndarray:
    """data/clip_transforms.py:50-56 (eval branch) == clip/clip.py:74-81: Resize(res, BICUBIC) on the PIL image,
    CenterCrop(res), ToTensor (HWC u8 -> CHW f32 / 255), Normalize(CLIP_MEAN, CLIP_STD).  [H,W,3] u8 -> [3,R,R]."""
    h, w = img_u8.shape[:2]
    nh, nw = resized_size(h, w, resolution)
    if (nh, nw) != (h, w):
        img_u8 = pil_bicubic_resize(img_u8, nh, nw)
    top = int(round((nh - resolution) / 2.0))
    left = int(round((nw - resolution) / 2.0))
    crop = img_u8[top:top + resolution, left:left + resolution]
    x = crop.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    mean = np.asarray(CLIP_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(CLIP_STD, dtype=np.float32)[:, None, None]
    return (x - mean) / std
