"""CPU emulation of the GPU path's roundings (analysis tool, not the product, not the oracle): fp16 tensor-core operands
with fp32 accumulation, fp32 residual stream / LayerNorm statistics / softmax, 16-bit qkv / P / attention output /
MLP hidden, LayerNorm folded into the consumer GEMM (A = fp16(gamma * x), statistics applied in the epilogue) and
QuickGELU through tanh with a 2^-11 relative error bound.  Prints the distance to the unmodified reference's goldens
at a geometry, i.e. the error budget the CUDA path should land in:

    python tools/emulate_fp16_tower.py            # ViT-B/16 headline geometry, 16 images
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from aihab_clip_b200.weights import GEOMETRIES, make_state_dict_np, synthetic_images_u8  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402


def r16(t):
    return t.to(torch.float16).to(torch.float32)


def mm16(a, w):  # fp16 operands, fp32 accumulate (products of fp16 values are exact in fp32)
    return r16(a) @ r16(w).T


def ln_fold_gemm(x, g, b, w, bias):
    """out = r * (fp16(g*x) @ W^T) - r*mu*s + b'   (EPI_LN_*)"""
    mu = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mu * mu
    r = torch.rsqrt(var.clamp_min(0) + 1e-5)
    w16 = r16(w).double()
    s = (w16 * g.double()).sum(-1).float()
    bp = (bias.double() + (w16 * b.double()).sum(-1)).float()
    return r * mm16(g * x, w) - r * mu * s + bp


def tower(sd, x, heads_dim=64):
    t = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in sd.items() if k.startswith("visual.")}
    w = t["visual.conv1.weight"]
    p = w.shape[-1]
    n, c, R, _ = x.shape
    g = R // p
    rows = x.reshape(n, c, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(n, g * g, c * p * p)
    h = mm16(rows, w.reshape(w.shape[0], -1))
    D = h.shape[-1]
    h = torch.cat([t["visual.class_embedding"].expand(n, 1, D), h], 1) + t["visual.positional_embedding"]
    h = torch.nn.functional.layer_norm(h, (D,), t["visual.ln_pre.weight"], t["visual.ln_pre.bias"], 1e-5)
    H = D // heads_dim
    L = h.shape[1]
    layers = len([k for k in t if k.endswith(".attn.in_proj_weight")])
    for i in range(layers):
        pre = f"visual.transformer.resblocks.{i}."
        qkv = r16(ln_fold_gemm(h, t[pre + "ln_1.weight"], t[pre + "ln_1.bias"], t[pre + "attn.in_proj_weight"],
                               t[pre + "attn.in_proj_bias"]))
        q, k, v = (u.reshape(n, L, H, heads_dim).permute(0, 2, 1, 3) for u in qkv.split(D, dim=-1))
        s = q @ k.transpose(-1, -2) / 8.0
        pr = torch.exp(s - s.amax(-1, keepdim=True))
        o = (r16(pr) @ v) / pr.sum(-1, keepdim=True)
        o = r16(o.permute(0, 2, 1, 3).reshape(n, L, D))
        h = h + mm16(o, t[pre + "attn.out_proj.weight"]) + t[pre + "attn.out_proj.bias"]
        u = ln_fold_gemm(h, t[pre + "ln_2.weight"], t[pre + "ln_2.bias"], t[pre + "mlp.c_fc.weight"], t[pre + "mlp.c_fc.bias"])
        th = torch.tanh(0.851 * u) * (1 + (torch.rand_like(u) - 0.5) * 2.0 ** -10)  # tanh.approx.f32: 2^-11 rel. error
        hid = r16(0.5 * u + 0.5 * u * th)
        h = h + mm16(hid, t[pre + "mlp.c_proj.weight"]) + t[pre + "mlp.c_proj.bias"]
    return torch.nn.functional.layer_norm(h[:, 0], (D,), t["visual.ln_post.weight"], t["visual.ln_post.bias"], 1e-5)


def main():
    torch.manual_seed(0)
    gold = np.load(REPO / "tests" / "golden" / "reference_outputs_vitl.npz")
    geom = GEOMETRIES["ViT-B/16"]
    sd = make_state_dict_np(geom, 0, with_text=False)
    n, side = 16, 300
    u8 = np.concatenate([synthetic_images_u8(n // 2, side, seed=1234),
                         synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])
    x = torch.from_numpy(np.stack([O.clip_preprocess(im, geom.image_resolution) for im in u8]))
    with torch.no_grad():
        feats = tower(sd, x)
    emb, logits, _ = O.score(feats.numpy(), sd["visual.proj"], gold["b16_text_w"], 100.0, 3)
    cos = (emb * gold["b16_emb"]).sum(-1)
    print("ViT-B/16, 16 images, emulated fp16-operand tower vs unmodified reference:")
    print(f"  min cosine {cos.min():.6f}   max |dfeat| {np.abs(feats.numpy() - gold['b16_feats']).max():.2e}"
          f"   max |dlogit| {np.abs(logits - gold['b16_logits']).max():.2e}")


if __name__ == "__main__":
    main()
