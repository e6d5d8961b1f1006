// HBM-bound kernels of the hot path: LayerNorm, im2col, weight cast.  128-bit loads/stores, one warp per row.
#include "kernels.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include "ptx.cuh"

namespace aihab {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp normalises one row of D = 128 * VPL floats held entirely in registers (two-pass statistics).
template <int VPL>
__global__ void __launch_bounds__(128) layernorm_kernel(const float* __restrict__ x, size_t ldx,
                                                        const float* __restrict__ cls0, int L,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* out32, void* out16,
                                                        int out_bf16, int rows) {
  constexpr int D = VPL * 128;
  ptx::griddep_launch();
  ptx::griddep_wait();  // x comes from the previous kernel of the stream
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = (cls0 != nullptr && (row % L) == 0) ? cls0 : x + static_cast<size_t>(row) * ldx;
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out32 != nullptr) *reinterpret_cast<float4*>(out32 + static_cast<size_t>(row) * D + col) = y;
    if (out16 != nullptr) {
      uint2 pk;
      pk.x = out_bf16 ? ptx::pack2<true>(y.x, y.y) : ptx::pack2<false>(y.x, y.y);
      pk.y = out_bf16 ? ptx::pack2<true>(y.z, y.w) : ptx::pack2<false>(y.z, y.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out16) + static_cast<size_t>(row) * D + col) = pk;
    }
  }
}

// ln_pre followed by ln_1 of the first block in one pass over the row: y = LN(x; g1, b1) is written as the fp32 residual
// stream and, still in registers, normalised again (LN(y; g2, b2)) into the 16-bit A operand of the first in_proj GEMM.
// The second stage is the arithmetic of layernorm_kernel on the values it would have re-read, so the result is
// bit-identical to the two launches; it saves the launch and one read of the residual stream per step.
template <int VPL>
__global__ void __launch_bounds__(128) layernorm2_kernel(const float* __restrict__ x, size_t ldx,
                                                         const float* __restrict__ cls0, int L,
                                                         const float* __restrict__ g1, const float* __restrict__ b1,
                                                         float* out32, const float* __restrict__ g2,
                                                         const float* __restrict__ b2, void* out16, int out_bf16, int rows) {
  constexpr int D = VPL * 128;
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = (cls0 != nullptr && (row % L) == 0) ? cls0 : x + static_cast<size_t>(row) * ldx;
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
#pragma unroll 1
  for (int stage = 0; stage < 2; ++stage) {
    const float* gamma = stage ? g2 : g1;
    const float* beta = stage ? b2 : b1;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (stage == 0) {
        *reinterpret_cast<float4*>(out32 + static_cast<size_t>(row) * D + col) = y;
        v[i] = y;
      } else {
        uint2 pk;
        pk.x = out_bf16 ? ptx::pack2<true>(y.x, y.y) : ptx::pack2<false>(y.x, y.y);
        pk.y = out_bf16 ? ptx::pack2<true>(y.z, y.w) : ptx::pack2<false>(y.z, y.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out16) + static_cast<size_t>(row) * D + col) = pk;
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Each thread writes 8 consecutive patch-row columns (one 16 B store).
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ img, int n, int R, int p, int g, int Kpad,
                                                     uint16_t* __restrict__ out, int out_bf16) {
  const int units_per_row = Kpad >> 3;
  const long total = static_cast<long>(n) * g * g * units_per_row;
  const int pp = p * p;
  const int K = 3 * pp;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long prow = idx / units_per_row;
    const int col0 = static_cast<int>(idx - prow * units_per_row) * 8;
    const int im = static_cast<int>(prow / (g * g));
    const int pi = static_cast<int>(prow - static_cast<long>(im) * g * g);
    const int gy = pi / g, gx = pi - gy * g;
    int c = col0 / pp;
    int rem = col0 - c * pp;
    int ky = rem / p;
    int kx = rem - ky * p;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (col0 + j < K) {
        const size_t off = ((static_cast<size_t>(im) * 3 + c) * R + (gy * p + ky)) * R + (gx * p + kx);
        f[j] = load_as_float<T>(img + off);
      } else {
        f[j] = 0.f;
      }
      if (++kx == p) {
        kx = 0;
        if (++ky == p) {
          ky = 0;
          ++c;
        }
      }
    }
    uint4 pk;
    if (out_bf16) {
      pk.x = ptx::pack2<true>(f[0], f[1]);
      pk.y = ptx::pack2<true>(f[2], f[3]);
      pk.z = ptx::pack2<true>(f[4], f[5]);
      pk.w = ptx::pack2<true>(f[6], f[7]);
    } else {
      pk.x = ptx::pack2<false>(f[0], f[1]);
      pk.y = ptx::pack2<false>(f[2], f[3]);
      pk.z = ptx::pack2<false>(f[4], f[5]);
      pk.w = ptx::pack2<false>(f[6], f[7]);
    }
    *reinterpret_cast<uint4*>(out + prow * Kpad + col0) = pk;
  }
}

__global__ void cast_pad_kernel(const float* __restrict__ src, int rows, int cols, uint16_t* __restrict__ dst,
                                int cols_pad, int out_bf16) {
  const long total = static_cast<long>(rows) * cols_pad;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    const float v = c < cols ? src[r * cols + c] : 0.f;
    if (out_bf16) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      dst[i] = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __half h = __float2half_rn(v);
      dst[i] = *reinterpret_cast<uint16_t*>(&h);
    }
  }
}

// Text tower input (clip/model.py:341-342): x[row, :] = token_embedding[tokens[row], :] + positional_embedding[row % L, :]
__global__ void __launch_bounds__(256) embed_tokens_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ emb,
                                                           const float* __restrict__ pos, float* __restrict__ x, long rows,
                                                           int L, int D, int vocab) {
  const int d4 = D >> 2;
  const long total = rows * d4;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    long tok = tokens[row];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb + tok * D) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (row % L) * D) + c);
    reinterpret_cast<float4*>(x + row * D)[c] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

// EOT pooling (clip/model.py:350): per sequence the row at tokens.argmax(-1) (first maximum, like torch) is copied
// out of the residual stream; ln_final then runs on those n rows only (LayerNorm is per row, so this equals
// ln_final(x)[arange(n), eot]).
__global__ void __launch_bounds__(128) eot_gather_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ x,
                                                         float* __restrict__ out, int n, int L, int D) {
  const int seq = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (seq >= n) return;
  long best = -0x7fffffffffffffffL - 1;
  int bi = 0x7fffffff;
  for (int t = lane; t < L; t += 32) {
    const long v = tokens[static_cast<long>(seq) * L + t];
    if (v > best || (v == best && t < bi)) {
      best = v;
      bi = t;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  const float4* src = reinterpret_cast<const float4*>(x + (static_cast<long>(seq) * L + bi) * D);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<long>(seq) * D);
  for (int c = lane; c < (D >> 2); c += 32) dst[c] = src[c];
}


template <int CODE> struct ElemOf;
template <> struct ElemOf<0> { using type = float; };
template <> struct ElemOf<1> { using type = __half; };
template <> struct ElemOf<2> { using type = __nv_bfloat16; };
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Row L2 normalisation for any of fp32 / fp16 / bf16 in and out: y = x / max(||x||_2, eps), statistics in fp32
// (F.normalize, methods/utils.py:184, aihab_utils/feature_cache.py:127; eps = 0 gives the `f /= f.norm()` variant of
// utils.py:69, NaN rows for zero vectors included).  One warp per row; 16-byte loads and stores when the row pitch
// allows it (cols % (16 / element size) == 0 and aligned bases), element-wise otherwise.
template <int IN, int OUT>  // 0 = fp32, 1 = fp16, 2 = bf16
__global__ void __launch_bounds__(128) l2norm_rows_kernel(const void* __restrict__ xin, void* __restrict__ yout, int rows,
                                                          int cols, float eps, int vec) {
  using TI = typename ElemOf<IN>::type;
  using TO = typename ElemOf<OUT>::type;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TI* src = static_cast<const TI*>(xin) + static_cast<size_t>(row) * cols;
  TO* dst = static_cast<TO*>(yout) + static_cast<size_t>(row) * cols;
  constexpr int VI = 16 / sizeof(TI);  // elements per 16-byte load
  float s = 0.f;
  if (vec) {
    for (int c = lane * VI; c < cols; c += 32 * VI) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + c);
      const TI* e = reinterpret_cast<const TI*>(&raw);
#pragma unroll
      for (int j = 0; j < VI; ++j) {
        const float f = to_float(e[j]);
        s = fmaf(f, f, s);
      }
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      const float f = to_float(src[c]);
      s = fmaf(f, f, s);
    }
  }
  const float inv_denom_src = fmaxf(sqrtf(warp_sum(s)), eps);
  if (vec) {
    for (int c = lane * VI; c < cols; c += 32 * VI) {  // second read of the row hits L1/L2
      const uint4 raw = *reinterpret_cast<const uint4*>(src + c);
      const TI* e = reinterpret_cast<const TI*>(&raw);
      TO o[VI];
#pragma unroll
      for (int j = 0; j < VI; ++j) o[j] = from_float<TO>(to_float(e[j]) / inv_denom_src);
      constexpr int NV = VI * sizeof(TO) / 16 > 0 ? VI * sizeof(TO) / 16 : 1;
      if constexpr (VI * sizeof(TO) >= 16) {
#pragma unroll
        for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(dst + c)[q] = reinterpret_cast<const uint4*>(o)[q];
      } else {  // fp32 in, 16-bit out: 4 elements = 8 bytes
        *reinterpret_cast<uint2*>(dst + c) = *reinterpret_cast<const uint2*>(o);
      }
    }
  } else {
    for (int c = lane; c < cols; c += 32) dst[c] = from_float<TO>(to_float(src[c]) / inv_denom_src);
  }
}

}  // namespace

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("AIHAB_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

cudaError_t launch_layernorm(const float* x, size_t ldx, const float* cls0, int L, const float* gamma,
                             const float* beta, float* out32, void* out16, int out_bf16, int rows, int D,
                             cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (D % 128 != 0 || D > 2048 || (ldx & 3) != 0) return cudaErrorInvalidValue;
  const int grid = (rows + 3) / 4;
  cudaError_t e = cudaSuccess;
#define AIHAB_LN(V)                                                                                            \
  case V:                                                                                                      \
    e = launch_kernel(layernorm_kernel<V>, grid, 128, 0, stream, 1, true, x, ldx, cls0, L > 0 ? L : 1, gamma, beta, \
                      out32, out16, out_bf16, rows);                                                             \
    break;
  switch (D / 128) {
    AIHAB_LN(1) AIHAB_LN(2) AIHAB_LN(3) AIHAB_LN(4) AIHAB_LN(5) AIHAB_LN(6) AIHAB_LN(7) AIHAB_LN(8)
    AIHAB_LN(9) AIHAB_LN(10) AIHAB_LN(11) AIHAB_LN(12) AIHAB_LN(13) AIHAB_LN(14) AIHAB_LN(15) AIHAB_LN(16)
    default: return cudaErrorInvalidValue;
  }
#undef AIHAB_LN
  return e;
}

cudaError_t launch_layernorm2(const float* x, size_t ldx, const float* cls0, int L, const float* g1, const float* b1,
                              float* out32, const float* g2, const float* b2, void* out16, int out_bf16, int rows, int D,
                              cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (D % 128 != 0 || D > 2048 || (ldx & 3) != 0 || out32 == nullptr || out16 == nullptr) return cudaErrorInvalidValue;
  const int grid = (rows + 3) / 4;
  cudaError_t e = cudaSuccess;
#define AIHAB_LN2(V)                                                                                              \
  case V:                                                                                                         \
    e = launch_kernel(layernorm2_kernel<V>, grid, 128, 0, stream, 1, true, x, ldx, cls0, L > 0 ? L : 1, g1, b1, out32, \
                      g2, b2, out16, out_bf16, rows);                                                               \
    break;
  switch (D / 128) {
    AIHAB_LN2(1) AIHAB_LN2(2) AIHAB_LN2(3) AIHAB_LN2(4) AIHAB_LN2(5) AIHAB_LN2(6) AIHAB_LN2(7) AIHAB_LN2(8)
    AIHAB_LN2(9) AIHAB_LN2(10) AIHAB_LN2(11) AIHAB_LN2(12) AIHAB_LN2(13) AIHAB_LN2(14) AIHAB_LN2(15) AIHAB_LN2(16)
    default: return cudaErrorInvalidValue;
  }
#undef AIHAB_LN2
  return e;
}

cudaError_t launch_im2col(const void* images, int in_dtype, int n, int R, int p, int Kpad, void* out, int out_bf16,
                          cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (R % p != 0 || (Kpad & 7) != 0 || Kpad < 3 * p * p) return cudaErrorInvalidValue;
  const int g = R / p;
  const long total = static_cast<long>(n) * g * g * (Kpad / 8);
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  switch (in_dtype) {
    case 0: im2col_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(images), n, R, p, g, Kpad, o, out_bf16); break;
    case 1: im2col_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(images), n, R, p, g, Kpad, o, out_bf16); break;
    case 2: im2col_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(images), n, R, p, g, Kpad, o, out_bf16); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_cast_pad(const float* src, int rows, int cols, void* dst, int cols_pad, int out_bf16,
                            cudaStream_t stream) {
  const long total = static_cast<long>(rows) * cols_pad;
  if (total <= 0) return cudaSuccess;
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 32));
  cast_pad_kernel<<<grid, 256, 0, stream>>>(src, rows, cols, reinterpret_cast<uint16_t*>(dst), cols_pad, out_bf16);
  return cudaGetLastError();
}

cudaError_t launch_embed_tokens(const int64_t* tokens, const float* emb, const float* pos, float* x, long rows, int L,
                                int D, int vocab, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if ((D & 3) != 0 || L <= 0 || vocab <= 0) return cudaErrorInvalidValue;
  const long total = rows * (D >> 2);
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  embed_tokens_kernel<<<grid, 256, 0, stream>>>(tokens, emb, pos, x, rows, L, D, vocab);
  return cudaGetLastError();
}

cudaError_t launch_eot_gather(const int64_t* tokens, const float* x, float* out, int n, int L, int D, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if ((D & 3) != 0) return cudaErrorInvalidValue;
  eot_gather_kernel<<<(n + 3) / 4, 128, 0, stream>>>(tokens, x, out, n, L, D);
  return cudaGetLastError();
}

cudaError_t launch_l2norm_rows(const void* x, int in_dtype, void* y, int out_dtype, int rows, int cols, float eps,
                               cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (cols <= 0 || in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2) return cudaErrorInvalidValue;
  const int isz = in_dtype == 0 ? 4 : 2, osz = out_dtype == 0 ? 4 : 2;
  const int vi = 16 / isz;
  const int vec = (cols % vi == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                  (reinterpret_cast<uintptr_t>(y) % (vi * osz >= 16 ? 16 : 8) == 0) &&
                  ((static_cast<size_t>(cols) * osz) % (vi * osz >= 16 ? 16 : 8) == 0);
  const dim3 grid((rows + 3) / 4);
#define AIHAB_L2N(I, O) l2norm_rows_kernel<I, O><<<grid, 128, 0, stream>>>(x, y, rows, cols, eps, vec)
  switch (in_dtype * 3 + out_dtype) {
    case 0: AIHAB_L2N(0, 0); break;
    case 1: AIHAB_L2N(0, 1); break;
    case 2: AIHAB_L2N(0, 2); break;
    case 3: AIHAB_L2N(1, 0); break;
    case 4: AIHAB_L2N(1, 1); break;
    case 5: AIHAB_L2N(1, 2); break;
    case 6: AIHAB_L2N(2, 0); break;
    case 7: AIHAB_L2N(2, 1); break;
    default: AIHAB_L2N(2, 2); break;
  }
#undef AIHAB_L2N
  return cudaGetLastError();
}

}  // namespace aihab

