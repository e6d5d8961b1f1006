"""Per-kernel parity on a B200: every CUDA kernel is called through the C ABI (ctypes) and compared with the
numpy oracle on the same seeded inputs.  Integer work (preprocessing, top-k indices) must be bit-exact;
floating-point work is held to tolerances written next to each assert."""
import numpy as np
import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from aihab_clip_b200 import _lib, ops
    return _lib, ops


def _rand16(rng, shape, scale, dtype):
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32)).to(dtype)


GEMM_SHAPES = [(300, 2304, 768), (389, 768, 3072), (50, 128, 64), (1000, 200, 640), (128, 256, 768),
               (4100, 3072, 768), (7, 64, 128)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias_16(cuda_device, dtype, M, N, K):
    _lib, ops = _ops()
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a, w = _rand16(rng, (M, K), 1.0, dtype), _rand16(rng, (N, K), K ** -0.5, dtype)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    out = torch.full((M, N), float("nan"), dtype=dtype, device=cuda_device)
    ops.gemm16(a.to(cuda_device), w.to(cuda_device), _lib.EPI_BIAS_16, bias=bias.to(cuda_device), out16=out)
    torch.cuda.synchronize()
    ref = a.double().numpy() @ w.double().numpy().T + bias.double().numpy()
    got = out.float().cpu().numpy()
    # fp32 accumulation of exact 16-bit products + one rounding to the 16-bit output format
    eps = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, atol=eps * 4 + 1e-4, rtol=eps * 1.01)


@pytest.mark.parametrize("M,N,K", [(300, 3072, 768), (129, 256, 128)])
def test_gemm_bias_quickgelu(cuda_device, M, N, K):
    _lib, ops = _ops()
    rng = np.random.default_rng(11)
    a, w = _rand16(rng, (M, K), 1.0, torch.float16), _rand16(rng, (N, K), 2 * K ** -0.5, torch.float16)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    out = torch.empty((M, N), dtype=torch.float16, device=cuda_device)
    ops.gemm16(a.to(cuda_device), w.to(cuda_device), _lib.EPI_BIAS_GELU_16, bias=bias.to(cuda_device), out16=out)
    u = a.double().numpy() @ w.double().numpy().T + bias.double().numpy()
    ref = u / (1.0 + np.exp(-1.702 * u))
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=2e-3, rtol=2.0 ** -10)


@pytest.mark.parametrize("M,N,K", [(300, 768, 3072), (513, 1024, 1024), (5, 128, 64)])
def test_gemm_bias_residual_fp32(cuda_device, M, N, K):
    _lib, ops = _ops()
    rng = np.random.default_rng(12)
    a, w = _rand16(rng, (M, K), 1.0, torch.float16), _rand16(rng, (N, K), K ** -0.5, torch.float16)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    x0 = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32))
    x = x0.clone().to(cuda_device)
    ops.gemm16(a.to(cuda_device), w.to(cuda_device), _lib.EPI_BIAS_RES_32, bias=bias.to(cuda_device), out32=x)
    ref = x0.double().numpy() + a.double().numpy() @ w.double().numpy().T + bias.double().numpy()
    np.testing.assert_allclose(x.cpu().numpy(), ref, atol=2e-4, rtol=1e-5)  # fp32 accumulate + fp32 add


def test_gemm_patch_embed_epilogue(cuda_device):
    _lib, ops = _ops()
    rng = np.random.default_rng(13)
    n_img, g2, D, K = 5, 49, 256, 640
    a, w = _rand16(rng, (n_img * g2, K), 1.0, torch.float16), _rand16(rng, (D, K), K ** -0.5, torch.float16)
    pos = torch.from_numpy(rng.standard_normal((g2 + 1, D)).astype(np.float32))
    x = torch.zeros(n_img * (g2 + 1), D, dtype=torch.float32, device=cuda_device)
    ops.gemm16(a.to(cuda_device), w.to(cuda_device), _lib.EPI_PATCH_32, out32=x, pos=pos.to(cuda_device), g2=g2)
    acc = (a.double().numpy() @ w.double().numpy().T).reshape(n_img, g2, D) + pos.double().numpy()[None, 1:]
    got = x.cpu().numpy().reshape(n_img, g2 + 1, D)
    np.testing.assert_allclose(got[:, 1:], acc, atol=2e-4, rtol=1e-5)
    assert (got[:, 0] == 0).all()  # class-token rows are written by the ln_pre kernel, not the GEMM


def test_gemm_scale_fp32(cuda_device):
    _lib, ops = _ops()
    rng = np.random.default_rng(14)
    M, N, K = 777, 1000, 512
    a, w = _rand16(rng, (M, K), K ** -0.5, torch.float16), _rand16(rng, (N, K), K ** -0.5, torch.float16)
    out = torch.empty(M, N, dtype=torch.float32, device=cuda_device)
    ops.gemm16(a.to(cuda_device), w.to(cuda_device), _lib.EPI_SCALE_32, out32=out, scale=100.0)
    ref = 100.0 * (a.double().numpy() @ w.double().numpy().T)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("D", [128, 768, 1024])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_layernorm(cuda_device, D, out_dtype):
    _lib, ops = _ops()
    rng = np.random.default_rng(D)
    x = (rng.standard_normal((333, D)) * 3 + 0.5).astype(np.float32)
    g, b = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32), rng.standard_normal(D).astype(np.float32)
    y = ops.layernorm(torch.from_numpy(x).to(cuda_device), torch.from_numpy(g).to(cuda_device),
                      torch.from_numpy(b).to(cuda_device), out_dtype)
    ref = O.layer_norm(x, g, b)
    tol = {torch.float32: 1e-5, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}[out_dtype]
    np.testing.assert_allclose(y.float().cpu().numpy(), ref, atol=tol * 4, rtol=tol)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L,H,n", [(50, 12, 3), (197, 12, 2), (257, 16, 2), (577, 16, 1), (17, 2, 4), (64, 1, 1),
                                   (256, 4, 2), (100, 2, 3), (129, 2, 2), (65, 1, 5), (197, 12, 9), (225, 2, 3),
                                   (272, 3, 5), (273, 2, 2), (384, 2, 2), (600, 1, 2), (1024, 1, 1), (257, 16, 40),
                                   (80, 1, 2), (96, 1, 2), (160, 2, 2), (176, 1, 3), (192, 2, 2), (208, 1, 2),
                                   (50, 12, 5), (16, 1, 9), (33, 2, 7), (64, 2, 3), (43, 1, 2), (50, 2, 301), (24, 3, 11),
                                   (224, 2, 3), (210, 3, 2), (144, 1, 5), (130, 2, 1)])
def test_attention(cuda_device, dtype, L, H, n):
    _lib, ops = _ops()
    rng = np.random.default_rng(L + H)
    D = H * 64
    qkv = _rand16(rng, (n * L, 3 * D), 1.5, dtype)
    out = ops.attention(qkv.to(cuda_device), n, L, H).float().cpu().numpy()
    q, k, v = (t.reshape(n, L, H, 64).transpose(0, 2, 1, 3) for t in np.split(qkv.double().numpy(), 3, axis=-1))
    s = q @ k.transpose(0, 1, 3, 2) / 8.0
    p = np.exp(s - s.max(-1, keepdims=True))
    ref = ((p / p.sum(-1, keepdims=True)) @ v).transpose(0, 2, 1, 3).reshape(n * L, D)
    # P is rounded to 16 bits before PV and the output is 16-bit: error ~ eps * |v|
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    np.testing.assert_allclose(out, ref, atol=eps * 3, rtol=eps * 2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L,H,n", [(257, 2, 3), (577, 2, 2), (300, 1, 4)])
def test_attention_online_rescale(cuda_device, dtype, L, H, n):
    """Peaky scores whose maximum grows along the key axis: every key block of the flash kernel raises the running
    row maximum by much more than 2^8, so the O rescale in TMEM runs at every block (clip/model.py:179-181)."""
    _lib, ops = _ops()
    rng = np.random.default_rng(7 * L + H)
    D = H * 64
    q = rng.standard_normal((n, L, H, 64)) * 2.0
    k = rng.standard_normal((n, L, H, 64)) * (0.5 + 3.0 * np.arange(L)[None, :, None, None] / L)
    v = rng.standard_normal((n, L, H, 64))
    qkv = torch.from_numpy(np.concatenate([q, k, v], axis=2).reshape(n * L, 3 * D)).to(dtype)
    out = ops.attention(qkv.to(cuda_device), n, L, H).float().cpu().numpy()
    qd, kd, vd = (t.reshape(n, L, H, 64).transpose(0, 2, 1, 3) for t in np.split(qkv.double().numpy(), 3, axis=-1))
    s_ = qd @ kd.transpose(0, 1, 3, 2) / 8.0
    p_ = np.exp(s_ - s_.max(-1, keepdims=True))
    ref = ((p_ / p_.sum(-1, keepdims=True)) @ vd).transpose(0, 2, 1, 3).reshape(n * L, D)
    assert np.isfinite(out).all()
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    np.testing.assert_allclose(out, ref, atol=eps * 4, rtol=eps * 3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L,H,n", [(77, 8, 5), (65, 1, 3), (128, 2, 2), (224, 2, 2), (100, 12, 7)])
def test_attention_causal(cuda_device, dtype, L, H, n):
    """The text tower's mask (clip/model.py:323-329: -inf above the diagonal): key j visible to query i for j <= i."""
    _lib, ops = _ops()
    rng = np.random.default_rng(L * 3 + H)
    D = H * 64
    qkv = _rand16(rng, (n * L, 3 * D), 1.5, dtype)
    out = ops.attention(qkv.to(cuda_device), n, L, H, causal=True).float().cpu().numpy()
    q, k, v = (t.reshape(n, L, H, 64).transpose(0, 2, 1, 3) for t in np.split(qkv.double().numpy(), 3, axis=-1))
    s = q @ k.transpose(0, 1, 3, 2) / 8.0
    s = np.where(np.tril(np.ones((L, L), dtype=bool)), s, -np.inf)
    p = np.exp(s - s.max(-1, keepdims=True))
    ref = ((p / p.sum(-1, keepdims=True)) @ v).transpose(0, 2, 1, 3).reshape(n * L, D)
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    np.testing.assert_allclose(out, ref, atol=eps * 3, rtol=eps * 2)


def test_preprocess_bit_exact(cuda_device, gold, meta):
    _lib, ops = _ops()
    from aihab_clip_b200.weights import synthetic_images_u8
    import hashlib
    for (h, w, R) in meta["pre_cases"]:
        rng = np.random.Generator(np.random.PCG64([99, h, w, R]))
        u8 = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        u8[1] = synthetic_images_u8(1, max(h, w), seed=5, smooth=True)[0][:h, :w]
        y = ops.preprocess_u8(torch.from_numpy(u8).to(cuda_device), R).cpu().numpy()
        ref = np.stack([O.clip_preprocess(im, R) for im in u8])
        np.testing.assert_array_equal(y, ref, err_msg=f"{h}x{w}->{R}")
        assert hashlib.sha256(y.tobytes()).digest() == gold[f"pre_{h}x{w}_{R}_sha"].tobytes()
    # 16-bit outputs are the fp32 result rounded once
    u8 = synthetic_images_u8(3, 439)
    y32 = ops.preprocess_u8(torch.from_numpy(u8).to(cuda_device), 224)
    y16 = ops.preprocess_u8(torch.from_numpy(u8).to(cuda_device), 224, torch.float16)
    assert torch.equal(y16, y32.half())


@pytest.mark.parametrize("n,D,E,Cn,k", [(64, 768, 512, 20, 3), (1000, 768, 512, 1000, 5), (5, 128, 64, 18, 1),
                                        (9000, 768, 512, 100, 5)])
def test_score_matches_oracle(cuda_device, n, D, E, Cn, k):
    _lib, ops = _ops()
    rng = np.random.default_rng(n + Cn)
    feats = rng.standard_normal((n, D)).astype(np.float32)
    proj = (D ** -0.5 * rng.standard_normal((D, E))).astype(np.float32)
    tw = O.l2_normalize(rng.standard_normal((Cn, E)).astype(np.float32)).T.copy()
    emb, logits, idx, val = ops.score(torch.from_numpy(feats).to(cuda_device), torch.from_numpy(proj).to(cuda_device),
                                      torch.from_numpy(tw).to(cuda_device), 100.0, k)
    r_emb, r_logits, r_idx = O.score(feats, proj, tw, 100.0, k)
    np.testing.assert_allclose(emb.cpu().numpy(), r_emb, atol=2e-6, rtol=0)
    np.testing.assert_allclose(logits.cpu().numpy(), r_logits, atol=2e-4, rtol=0)
    # top-k of the kernel's own logits is exact (torch.topk semantics, lowest index first among ties)
    np.testing.assert_array_equal(idx.cpu().numpy(), O.topk_indices(logits.cpu().numpy(), k))
    srt = np.sort(r_logits, axis=1)[:, ::-1][:, :k + 1]
    untied = np.abs(np.diff(srt, axis=1)).min(axis=1) > 1e-3
    np.testing.assert_array_equal(idx.cpu().numpy()[untied], r_idx[untied])
    np.testing.assert_array_equal(val.cpu().numpy(), np.take_along_axis(logits.cpu().numpy(), idx.cpu().numpy(), 1))


def test_score_mid_path_equals_the_chunked_path_bit_for_bit(cuda_device):
    """65..8192 rows take the column-sliced kernels (score_proj_kernel / score_logits_kernel), more rows the chunked
    sgemm -> l2norm -> sgemm path: the same ascending fmaf chains and the same normalisation, so a row's embedding,
    logits and top-k must not depend on which path (or which batch) it went through."""
    _lib, ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(5)
    for (D, E, Cn, k) in ((768, 512, 1000, 5), (256, 128, 260, 3)):
        feats = torch.randn(9000, D, generator=g).to(cuda_device)
        proj = (torch.randn(D, E, generator=g) * D ** -0.5).to(cuda_device)
        tw = torch.nn.functional.normalize(torch.randn(Cn, E, generator=g), dim=1).t().contiguous().to(cuda_device)
        emb_c, lg_c, idx_c, val_c = ops.score(feats, proj, tw, 100.0, k)                 # chunked (n > 8192)
        for n in (300, 77, 8192):
            emb_m, lg_m, idx_m, val_m = ops.score(feats[:n].contiguous(), proj, tw, 100.0, k)   # mid path
            assert torch.equal(emb_m, emb_c[:n]) and torch.equal(lg_m, lg_c[:n])
            assert torch.equal(idx_m, idx_c[:n]) and torch.equal(val_m, val_c[:n])
            _, _, idx_n, _ = ops.score(feats[:n].contiguous(), proj, tw, 100.0, k, want_emb=False, want_logits=False)
            assert torch.equal(idx_n, idx_c[:n])                                         # scratch logits, same result
        # already projected features (proj = None, the call of the extraction step): normalise + logits + top-k only
        e_c, l_c, i_c, _ = ops.score(emb_c * 3.0, None, tw, 100.0, k)
        e_m, l_m, i_m, _ = ops.score((emb_c * 3.0)[:256].contiguous(), None, tw, 100.0, k)
        assert torch.equal(e_m, e_c[:256]) and torch.equal(l_m, l_c[:256]) and torch.equal(i_m, i_c[:256])


def test_topk_exact_ties(cuda_device):
    _lib, ops = _ops()
    lg = np.zeros((4, 8), dtype=np.float32)
    lg[0] = [1, 3, 3, 2, 3, 0, -1, 2]
    lg[1] = 5
    lg[2] = np.arange(8)
    lg[3] = -np.arange(8)
    # score() with identity projection-free path: feats are one-hot rows so logits = scale * text_w rows
    eye = torch.eye(4, dtype=torch.float32, device=cuda_device)
    emb, logits, idx, _ = ops.score(eye, None, torch.from_numpy(lg).to(cuda_device), 1.0, 4)
    np.testing.assert_array_equal(logits.cpu().numpy(), lg)
    np.testing.assert_array_equal(idx.cpu().numpy(), O.topk_indices(lg, 4))
    assert idx[0].tolist() == [1, 2, 4, 3] and idx[1].tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,D,E,Cn,k", [(5000, 768, 512, 1000, 5), (40000, 768, 512, 100, 3), (33, 128, 64, 20, 1)])
def test_score16_tensor_core_path(cuda_device, dtype, n, D, E, Cn, k):
    """Cached-feature scoring on the tensor cores: 16-bit features x 16-bit projection are exact products, the
    logits GEMM uses an fp16 hi/lo split -> results match the fp64 reference to fp32-accumulation noise."""
    _lib, ops = _ops()
    rng = np.random.default_rng(n + Cn)
    feats = torch.from_numpy(rng.standard_normal((n, D)).astype(np.float32)).to(dtype)
    proj = torch.from_numpy((D ** -0.5 * rng.standard_normal((D, E))).astype(np.float32)).to(dtype)
    tw = O.l2_normalize(rng.standard_normal((Cn, E)).astype(np.float32)).T.copy()
    emb, logits, idx, val = ops.score16(feats.to(cuda_device), proj.to(cuda_device), torch.from_numpy(tw).to(cuda_device),
                                        100.0, k, want_emb=True, want_logits=True)
    e64 = feats.double().numpy() @ proj.double().numpy()
    e64 /= np.maximum(np.linalg.norm(e64, axis=1, keepdims=True), 1e-12)
    l64 = 100.0 * e64 @ tw.astype(np.float64)
    np.testing.assert_allclose(emb.cpu().numpy(), e64, atol=3e-7, rtol=0)
    np.testing.assert_allclose(logits.cpu().numpy(), l64, atol=3e-4, rtol=0)   # fp32 accumulation + 2^-21 split error
    np.testing.assert_array_equal(idx.cpu().numpy(), O.topk_indices(logits.cpu().numpy(), k))
    srt = np.sort(l64, axis=1)[:, ::-1][:, :k + 1]
    untied = np.abs(np.diff(srt, axis=1)).min(axis=1) > 2e-3
    assert untied.mean() > 0.9
    np.testing.assert_array_equal(idx.cpu().numpy()[untied], O.topk_indices(l64, k)[untied])
    # and it agrees with the exact-fp32 CUDA-core path on the same inputs
    _, l32, i32, _ = ops.score(feats.to(cuda_device).float(), proj.to(cuda_device).float(),
                               torch.from_numpy(tw).to(cuda_device), 100.0, k)
    np.testing.assert_allclose(logits.cpu().numpy(), l32.cpu().numpy(), atol=5e-4, rtol=0)
    # without the embeddings the first GEMM writes the hi | hi | lo split of the RAW rows itself (EPI_SPLIT3_16) and the
    # logits GEMM normalises its accumulator rows: same gates against fp64, with and without the logits in HBM
    _, lg2, idx2, val2 = ops.score16(feats.to(cuda_device), proj.to(cuda_device), torch.from_numpy(tw).to(cuda_device),
                                     100.0, k, want_emb=False, want_logits=True)
    np.testing.assert_allclose(lg2.cpu().numpy(), l64, atol=3e-4, rtol=0)
    np.testing.assert_array_equal(idx2.cpu().numpy(), O.topk_indices(lg2.cpu().numpy(), k))
    np.testing.assert_array_equal(idx2.cpu().numpy()[untied], O.topk_indices(l64, k)[untied])
    if k <= 8:
        _, _, idx3, val3 = ops.score16(feats.to(cuda_device), proj.to(cuda_device), torch.from_numpy(tw).to(cuda_device),
                                       100.0, k)
        assert torch.equal(idx3, idx2) and torch.equal(val3, val2)        # top-k in the GEMM epilogue, same arithmetic


@pytest.mark.parametrize("n,C3,C2,k", [(257, 20, 11, 3), (5, 20, 11, 1), (4097, 37, 5, 5), (64, 1000, 100, 5)])
@pytest.mark.parametrize("reduce", ["sum", "mean", "logsumexp"])
def test_l2_metrics_match_oracle(cuda_device, n, C3, C2, k, reduce):
    """Fused L3 -> L2 aggregation + top-k + top-3 / softmax probabilities (aihab_utils/evaluation.py:92-142, 186-221,
    261-273) against the numpy restatement: sum / mean bit-exact (same accumulation order), logsumexp and the
    probabilities to 2e-6 relative, indices exact wherever the oracle's scores are untied."""
    _lib, ops = _ops()
    rng = np.random.default_rng(n * 31 + C3)
    logits = (100.0 * rng.uniform(-0.3, 0.3, (n, C3))).astype(np.float32)
    logits[0, :] = logits[0, 0]  # a fully tied row: lowest index first
    lut = rng.integers(0, C2, C3)
    lut[:min(C2, C3)] = np.arange(min(C2, C3))  # every group that can be populated is
    out, idx, val, t3i, t3p = ops.l2_metrics(torch.from_numpy(logits).to(cuda_device), lut.tolist(), C2, reduce, k=k)
    ref = O.aggregate_logits_to_l2(logits, lut, C2, reduce)
    got = out.cpu().numpy()
    if reduce == "logsumexp":
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=2e-6, atol=2e-6)
    else:
        np.testing.assert_array_equal(got, ref)
    ref_idx = O.topk_indices(got, k)  # ordering rule checked on the kernel's own L2 logits
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_array_equal(val.cpu().numpy(), np.take_along_axis(got, ref_idx, axis=1))
    _, r3i, r3p = O.top3_metrics(logits, np.zeros(n, dtype=np.int64))
    np.testing.assert_array_equal(t3i.cpu().numpy(), r3i)
    np.testing.assert_allclose(t3p.cpu().numpy(), r3p, rtol=3e-6, atol=1e-9)


def test_evaluation_mirrors_use_the_fused_kernel_on_cuda(cuda_device):
    """L2MetricsAccumulator / ClassificationTracker / aggregate_logits_to_l2 give the same numbers on CUDA (fused
    kernel) as on CPU (the reference's torch formulation)."""
    from aihab_clip_b200 import evaluation as E
    _lib, _ = _ops()
    rng = np.random.default_rng(5)
    logits = torch.from_numpy((100.0 * rng.uniform(-0.3, 0.3, (300, 20))).astype(np.float32))
    targets = torch.from_numpy(rng.integers(0, 20, 300))
    lut = [0, 0, 1, 2, 2, 3, 4, 5, 5, 5, 6, 7, 8, 8, 9, 10, 10, 3, 1, 0]
    n0 = _lib.kernel_launches()
    for reduce in ("sum", "mean", "logsumexp"):
        a = E.aggregate_logits_to_l2(logits, lut, 11, reduce)
        b = E.aggregate_logits_to_l2(logits.to(cuda_device), lut, 11, reduce).cpu()
        if reduce == "logsumexp":
            torch.testing.assert_close(b, a, rtol=2e-6, atol=2e-6)
        else:
            assert torch.equal(a, b)
        acc_c = E.L2MetricsAccumulator(lut, 11, reduce=reduce, topk=(1, 3), mode="logits")
        acc_g = E.L2MetricsAccumulator(lut, 11, reduce=reduce, topk=(1, 3), mode="logits")
        acc_c.update(logits, targets)
        acc_g.update(logits.to(cuda_device), targets.to(cuda_device))
        mc, mg = acc_c.compute(), acc_g.compute()
        assert mc["top1"] == mg["top1"] and mc["top3"] == mg["top3"] and mc["f1"] == mg["f1"]
    c_cpu, i_cpu, p_cpu = E.ClassificationTracker().top3_metrics(logits, targets)
    c_gpu, i_gpu, p_gpu = E.ClassificationTracker().top3_metrics(logits.to(cuda_device), targets.to(cuda_device))
    assert int(c_cpu) == int(c_gpu) and torch.equal(i_cpu, i_gpu.cpu())
    torch.testing.assert_close(p_gpu.cpu(), p_cpu, rtol=3e-6, atol=1e-9)
    assert _lib.kernel_launches() > n0  # the CUDA calls went through libaihab_clip.so


def test_prototype_scores_match_reference_golden(cuda_device):
    """aihab_prototype_scores against the unmodified reference's outputs and the oracle."""
    from pathlib import Path
    _lib, ops = _ops()
    g = np.load(Path(__file__).resolve().parent / "golden" / "prototype_scores.npz")
    t = lambda k: torch.from_numpy(g[k]).to(cuda_device)
    sim, pid, other, margin = ops.prototype_scores(t("emb"), t("labels"), t("prototypes"), t("owner"))
    np.testing.assert_allclose(sim.cpu().numpy(), g["sim_to_prototype"], atol=2e-6, rtol=0)
    np.testing.assert_array_equal(pid.cpu().numpy(), g["prototype_id"])
    np.testing.assert_allclose(other.cpu().numpy(), g["sim_to_other_class_best"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(margin.cpu().numpy(), g["margin_to_other_class"], atol=4e-6, rtol=0)
    sim_c, pid_c, _, _ = ops.prototype_scores(t("emb"), t("labels"), t("centroids"), t("centroid_owner"))
    np.testing.assert_allclose(sim_c.cpu().numpy(), g["sim_to_centroid"], atol=2e-6, rtol=0)
    assert int(pid_c.abs().sum()) == 0
    # larger than one chunk of rows, many prototypes, single-class NaN rule: against the oracle
    rng = np.random.default_rng(9)
    n, E, P = 70000, 512, 96
    emb = rng.standard_normal((n, E)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    protos = rng.standard_normal((P, E)).astype(np.float32)
    protos /= np.linalg.norm(protos, axis=1, keepdims=True)
    owner = np.sort(rng.integers(0, 20, P)).astype(np.int64)
    labels = rng.choice(np.unique(owner), n).astype(np.int64)
    got = ops.prototype_scores(*(torch.from_numpy(a).to(cuda_device) for a in (emb, labels, protos, owner)))
    ref = O.prototype_scores(emb, labels, protos, owner)
    np.testing.assert_allclose(got[0].cpu().numpy(), ref[0], atol=3e-6, rtol=0)
    tied = np.abs(got[0].cpu().numpy() - ref[0]) > 0  # index must agree unless two prototypes are within rounding
    assert (got[1].cpu().numpy() == ref[1]).mean() > 0.9999
    np.testing.assert_allclose(got[2].cpu().numpy(), ref[2], atol=3e-6, rtol=0)
    np.testing.assert_allclose(got[3].cpu().numpy(), ref[3], atol=6e-6, rtol=0)
    one = owner == owner[0]
    o1 = ops.prototype_scores(torch.from_numpy(emb[:100]).to(cuda_device), torch.full((100,), int(owner[0]), device=cuda_device),
                              torch.from_numpy(protos[one]).to(cuda_device), torch.from_numpy(owner[one]).to(cuda_device))
    assert torch.isnan(o1[2]).all() and torch.isnan(o1[3]).all()


def test_score16_fused_topk_equals_the_unfused_path(cuda_device):
    """EPI_TOPK_32: the logits GEMM keeps per-row top-k candidates in its epilogue and a merge pass finishes the selection,
    so the [rows, C] logits never reach HBM.  Same GEMM arithmetic, same (value desc, index asc) order: indices and
    values must equal those of the store + top-k kernel pair (requesting the logits selects that path) bit for bit,
    including exact ties and rows spanning several passes."""
    _lib, ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(3)
    for (n, D, E, Cn, k) in ((40000, 768, 512, 1000, 5), (777, 128, 64, 20, 3), (3000, 256, 128, 260, 8)):
        feats = torch.randn(n, D, generator=g).half().to(cuda_device)
        proj = (torch.randn(D, E, generator=g) * D ** -0.5).half().to(cuda_device)
        tw = torch.nn.functional.normalize(torch.randn(Cn, E, generator=g), dim=1).t().contiguous()
        tw[:, 7] = tw[:, 3]                                    # two identical classes: exact ties in every row
        tw = tw.to(cuda_device)
        _, _, idx_f, val_f = ops.score16(feats, proj, tw, 100.0, k)                      # fused (no logits requested)
        _, logits, idx_u, val_u = ops.score16(feats, proj, tw, 100.0, k, want_logits=True)  # logits stored + top-k kernel
        assert torch.equal(idx_f, idx_u) and torch.equal(val_f, val_u)
        ref = logits.topk(k, 1, True, True)
        assert torch.equal(val_f, ref.values)
        both = (idx_f == 3) | (idx_f == 7)
        assert both.any() and (idx_f[both.any(dim=1)].eq(3).any(dim=1)).all()   # tie: the lower index is always present first
