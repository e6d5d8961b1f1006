"""torch.Tensor front-ends of the C-ABI entry points.  PyTorch is used for device memory and streams only; every
computation below happens inside libaihab_clip.so.  All functions require CUDA tensors and raise otherwise."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}; expected float32, float16 or bfloat16") from None


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("aihab_clip_b200 ops run on CUDA tensors only (no CPU fallback)")


def gemm16(a: torch.Tensor, w: torch.Tensor, epilogue: int, bias=None, out16=None, out32=None, pos=None, g2: int = 0,
           scale: float = 1.0):
    """D = a @ w.T with the fused epilogue `epilogue` (see _lib.EPI_*).  a [M,K], w [N,K], fp16 or bf16."""
    _need_cuda(a, w, bias, out16, out32, pos)
    assert a.dtype == w.dtype and a.dtype in (torch.float16, torch.bfloat16)
    assert a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    out = out16 if out16 is not None else out32
    ldo = out.stride(0)
    rc = _lib.load().aihab_gemm16(_ptr(a), _ptr(w), M, N, K, dtype_code(a.dtype), epilogue, _ptr(bias), _ptr(out16),
                                  _ptr(out32), ldo, _ptr(pos), g2, C.c_float(scale), _stream(a.device))
    _lib.check(rc, "aihab_gemm16")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_dtype: torch.dtype = torch.float32):
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    rows, D = x.shape
    out = torch.empty(rows, D, dtype=out_dtype, device=x.device)
    o32 = out if out_dtype == torch.float32 else None
    o16 = None if out_dtype == torch.float32 else out
    rc = _lib.load().aihab_layernorm(_ptr(x), rows, D, _ptr(gamma), _ptr(beta), _ptr(o32), _ptr(o16),
                                     dtype_code(out_dtype), _stream(x.device))
    _lib.check(rc, "aihab_layernorm")
    return out


def attention(qkv: torch.Tensor, n: int, L: int, H: int, causal: bool = False):
    """qkv [n*L, 3*H*64] fp16/bf16 -> [n*L, H*64]; causal = the text tower's mask (64 < L <= 224)."""
    _need_cuda(qkv)
    assert qkv.is_contiguous() and qkv.shape == (n * L, 3 * H * 64)
    out = torch.empty(n * L, H * 64, dtype=qkv.dtype, device=qkv.device)
    fn = _lib.load().aihab_attention_causal if causal else _lib.load().aihab_attention
    rc = fn(_ptr(qkv), _ptr(out), n, L, H, dtype_code(qkv.dtype), _stream(qkv.device))
    _lib.check(rc, "aihab_attention_causal" if causal else "aihab_attention")
    return out


def preprocess_u8(images_u8: torch.Tensor, resolution: int, out_dtype: torch.dtype = torch.float32):
    """uint8 [N,H,W,3] (CUDA) -> normalised [N,3,R,R]; GPU twin of data/clip_transforms.py:50-56."""
    _need_cuda(images_u8)
    assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[-1] == 3
    images_u8 = images_u8.contiguous()
    n, sh, sw, _ = images_u8.shape
    out = torch.empty(n, 3, resolution, resolution, dtype=out_dtype, device=images_u8.device)
    rc = _lib.load().aihab_preprocess_u8(_ptr(images_u8), n, sh, sw, resolution, _ptr(out), dtype_code(out_dtype),
                                         _stream(images_u8.device))
    _lib.check(rc, "aihab_preprocess_u8")
    return out


def score(feats: torch.Tensor, proj, text_w, scale: float = 100.0, k: int = 1, want_emb: bool = True,
          want_logits: bool = True):
    """proj -> L2 normalise -> scale * emb @ text_w -> top-k.  Returns (emb, logits, topk_idx, topk_val); entries
    not requested are None.  fp32 throughout (methods/ProLIP.py:40, methods/utils.py:183-186)."""
    _need_cuda(feats, proj, text_w)
    f = feats.float().contiguous()
    n, D = f.shape
    p = proj.float().contiguous() if proj is not None else None
    E = p.shape[1] if p is not None else D
    w = text_w.float().contiguous() if text_w is not None else None
    Cn = w.shape[1] if w is not None else 0
    dev = f.device
    emb = torch.empty(n, E, dtype=torch.float32, device=dev) if want_emb else None
    logits = torch.empty(n, Cn, dtype=torch.float32, device=dev) if (want_logits and w is not None) else None
    idx = torch.empty(n, k, dtype=torch.int64, device=dev) if (k > 0 and w is not None) else None
    val = torch.empty(n, k, dtype=torch.float32, device=dev) if idx is not None else None
    rc = _lib.load().aihab_score(_ptr(f), n, D, _ptr(p), E, _ptr(w), Cn, C.c_float(scale), k if idx is not None else 0,
                                 _ptr(emb), _ptr(logits), _ptr(idx), _ptr(val), _stream(dev))
    _lib.check(rc, "aihab_score")
    return emb, logits, idx, val


def l2_normalize(x: torch.Tensor, eps: float = 1e-12, out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """Rows of x [n, D] (fp32 / fp16 / bf16) divided by max(||row||_2, eps), fp32 statistics, one 128-bit-load kernel.
    eps = 1e-12 is F.normalize (aihab_utils/feature_cache.py:126-127); eps = 0 is `f /= f.norm(dim=-1, keepdim=True)`
    (utils.py:69)."""
    _need_cuda(x)
    if x.dim() != 2:
        raise ValueError("l2_normalize expects a 2-D tensor [rows, cols]")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    rc = _lib.load().aihab_l2_normalize(_ptr(x), dtype_code(x.dtype), x.shape[0], x.shape[1], C.c_float(eps), _ptr(out),
                                        dtype_code(out.dtype), _stream(x.device))
    _lib.check(rc, "aihab_l2_normalize")
    return out


def score16(feats16: torch.Tensor, proj16: torch.Tensor, text_w: torch.Tensor, scale: float = 100.0, k: int = 1,
            want_emb: bool = False, want_logits: bool = False):
    """Tensor-core scoring over cached 16-bit features (config 5: ProLIP / linear-probe scoring).  feats16 [n, D]
    and proj16 [D, E] fp16 (or bf16), text_w [E, C] fp32.  Returns (emb, logits, topk_idx, topk_val) like score()."""
    _need_cuda(feats16, proj16, text_w)
    if feats16.dtype not in (torch.float16, torch.bfloat16) or proj16.dtype != feats16.dtype:
        raise TypeError("score16 expects fp16 or bf16 features and projection of the same dtype")
    f, p, w = feats16.contiguous(), proj16.contiguous(), text_w.float().contiguous()
    n, D = f.shape
    E, Cn = p.shape[1], w.shape[1]
    dev = f.device
    emb = torch.empty(n, E, dtype=torch.float32, device=dev) if want_emb else None
    logits = torch.empty(n, Cn, dtype=torch.float32, device=dev) if want_logits else None
    idx = torch.empty(n, k, dtype=torch.int64, device=dev) if k > 0 else None
    val = torch.empty(n, k, dtype=torch.float32, device=dev) if k > 0 else None
    rc = _lib.load().aihab_score16(_ptr(f), n, D, dtype_code(f.dtype), _ptr(p), E, _ptr(w), Cn, C.c_float(scale), k,
                                   _ptr(emb), _ptr(logits), _ptr(idx), _ptr(val), _stream(dev))
    _lib.check(rc, "aihab_score16")
    return emb, logits, idx, val


_REDUCE = {"sum": 0, "mean": 1, "logsumexp": 2}


def l2_metrics(logits_l3: torch.Tensor, l3_to_l2, num_l2: int, reduce: str = "mean", k: int = 1,
               want_logits: bool = True, want_top3: bool = True):
    """Fused metrics epilogue (aihab_utils/evaluation.py:92-142, 186-221, 261-273) on a CUDA logits tensor [n, C3].
    Returns (logits_l2 [n, num_l2] | None, topk_idx [n, k] int64, topk_val [n, k], top3_idx [n, 3] | None,
    top3_prob [n, 3] | None)."""
    _need_cuda(logits_l3)
    if reduce not in _REDUCE:
        raise ValueError(f"Unsupported reduce='{reduce}'. Expected one of: sum, mean, logsumexp.")
    x = logits_l3.float().contiguous()
    n, c3 = x.shape
    dev = x.device
    lut = (l3_to_l2.to(device=dev, dtype=torch.int32) if torch.is_tensor(l3_to_l2)
           else torch.tensor(list(l3_to_l2), device=dev, dtype=torch.int32)).contiguous()
    if lut.numel() != c3:
        raise ValueError(f"logits_l3 has {c3} classes, but l3_to_l2 has {lut.numel()} entries.")
    out = torch.empty(n, num_l2, dtype=torch.float32, device=dev) if want_logits else None
    idx = torch.empty(n, k, dtype=torch.int64, device=dev) if k > 0 else None
    val = torch.empty(n, k, dtype=torch.float32, device=dev) if k > 0 else None
    t3i = torch.zeros(n, 3, dtype=torch.int64, device=dev) if want_top3 else None
    t3p = torch.zeros(n, 3, dtype=torch.float32, device=dev) if want_top3 else None
    rc = _lib.load().aihab_l2_metrics(_ptr(x), n, c3, _ptr(lut), int(num_l2), _REDUCE[reduce], k, _ptr(out), _ptr(idx),
                                      _ptr(val), _ptr(t3i), _ptr(t3p), _stream(dev))
    _lib.check(rc, "aihab_l2_metrics")
    return out, idx, val, t3i, t3p


def prototype_scores(emb: torch.Tensor, labels: torch.Tensor, prototypes: torch.Tensor, owner: torch.Tensor):
    """tools/outlier_cleaning.py:553-668 on CUDA tensors: emb [n, E] L2-normalised, labels [n], prototypes [P, E]
    (class blocks contiguous), owner [P] class id per prototype.  Returns (sim_to_prototype [n], prototype_id [n]
    index inside the class block, sim_to_other_class_best [n] (NaN without another class), margin [n])."""
    _need_cuda(emb, labels, prototypes, owner)
    e = emb.float().contiguous()
    n, E = e.shape
    pt = prototypes.to(device=e.device, dtype=torch.float32).t().contiguous()  # [E, P]
    P = pt.shape[1]
    lab = labels.to(device=e.device, dtype=torch.int64).contiguous()
    own = owner.to(device=e.device, dtype=torch.int64).contiguous()
    sim = torch.empty(n, dtype=torch.float32, device=e.device)
    pid = torch.empty(n, dtype=torch.int64, device=e.device)
    oth = torch.empty(n, dtype=torch.float32, device=e.device)
    mar = torch.empty(n, dtype=torch.float32, device=e.device)
    rc = _lib.load().aihab_prototype_scores(_ptr(e), _ptr(lab), n, E, _ptr(pt), _ptr(own), P, _ptr(sim), _ptr(pid),
                                            _ptr(oth), _ptr(mar), _stream(e.device))
    _lib.check(rc, "aihab_prototype_scores")
    return sim, pid, oth, mar
