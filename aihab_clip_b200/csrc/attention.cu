// Fused single-pass attention, head dim 64, sequence lengths 50/197/257/577 (any L <= 640).
// One CTA per (head, image): K and V of the whole sequence are staged once in shared memory (XOR-swizzled
// 128 B rows, cp.async), every warp owns 16-query tiles and streams over 64-key chunks with an online softmax
// kept in registers (quad shuffles for the row max / row sum), P never leaves registers: QK^T and PV are
// mma.sync m16n8k16 with fp32 accumulators.   Reference: clip/model.py:179-181 (nn.MultiheadAttention -> SDPA,
// scale 1/sqrt(64), no mask, dropout 0).
#include "kernels.cuh"
#include "ptx.cuh"

namespace aihab {

namespace {

constexpr int HD = 64;  // head dim (width // 64 heads, clip/model.py:267)

template <bool BF16>
__global__ void __launch_bounds__(256, 2)
attention_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int L, int H, int Lp, int q_tile0) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + static_cast<size_t>(Lp) * 128;

  const int h = blockIdx.x;
  const int img = blockIdx.y;
  const int D = H * HD;
  const size_t ld = static_cast<size_t>(3) * D;
  const uint16_t* base = qkv + static_cast<size_t>(img) * L * ld + h * HD;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;

  // ---- stage K and V: row = key, 8 x 16 B units per row, unit' = unit ^ (row & 7); rows >= L zero-filled
  for (int idx = threadIdx.x; idx < Lp * 8; idx += blockDim.x) {
    const int row = idx >> 3, u = idx & 7;
    const bool valid = row < L;
    const uint16_t* src = base + static_cast<size_t>(valid ? row : 0) * ld + u * 8;
    const int off = row * 128 + ((u ^ (row & 7)) << 4);
    ptx::cp_async16(sK + off, src + D, valid);
    ptx::cp_async16(sV + off, src + 2 * D, valid);
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const float sl2 = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
  const int q_tiles = (L + 15) >> 4;
  const uint32_t sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);
  const int lm = lane >> 3, lr = lane & 7;  // ldmatrix: matrix id and row inside it

  for (int qt = q_tile0 + warp; qt < q_tiles; qt += nwarps) {  // q_tile0 > 0: only the rows from 16 * q_tile0 on
    const int q0 = qt * 16;
    // Q fragments straight from global memory in the m16k16 A layout (4 k-steps x 4 regs)
    uint32_t qa[4][4];
    {
      const int r0 = q0 + g, r1 = q0 + g + 8;
      const uint16_t* p0 = base + static_cast<size_t>(r0 < L ? r0 : 0) * ld;
      const uint16_t* p1 = base + static_cast<size_t>(r1 < L ? r1 : 0) * ld;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c = ks * 16 + 2 * t;
        qa[ks][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c) : 0u;
        qa[ks][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c) : 0u;
        qa[ks][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c + 8) : 0u;
        qa[ks][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c + 8) : 0u;
      }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kc = 0; kc < Lp; kc += 64) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // S = Q K^T over this 64-key chunk
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
          const int key = kc + np * 16 + (lm >> 1) * 8 + lr;
          const int du = ks * 2 + (lm & 1);
          uint32_t b0, b1, b2, b3;
          ptx::ldmatrix_x4(sK_u + key * 128 + ((du ^ (key & 7)) << 4), b0, b1, b2, b3);
          ptx::mma_16816<BF16>(s[2 * np], qa[ks], b0, b1);
          ptx::mma_16816<BF16>(s[2 * np + 1], qa[ks], b2, b3);
        }
      }
      // mask keys >= L (only the last chunk has any)
      if (kc + 64 > L) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = kc + i * 8 + 2 * t;
          if (key >= L) s[i][0] = s[i][2] = -INFINITY;
          if (key + 1 >= L) s[i][1] = s[i][3] = -INFINITY;
        }
      }
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cm0 = fmaxf(cm0, fmaxf(s[i][0], s[i][1]));
        cm1 = fmaxf(cm1, fmaxf(s[i][2], s[i][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);  // finite: every chunk has >= 1 valid key
      const float corr0 = exp2f((m0 - mn0) * sl2), corr1 = exp2f((m1 - mn1) * sl2);
      m0 = mn0;
      m1 = mn1;
      const float ms0 = mn0 * sl2, ms1 = mn1 * sl2;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = exp2f(fmaf(s[i][0], sl2, -ms0));
        s[i][1] = exp2f(fmaf(s[i][1], sl2, -ms0));
        s[i][2] = exp2f(fmaf(s[i][2], sl2, -ms1));
        s[i][3] = exp2f(fmaf(s[i][3], sl2, -ms1));
        rs0 += s[i][0] + s[i][1];
        rs1 += s[i][2] + s[i][3];
      }
      l0 = l0 * corr0 + rs0;
      l1 = l1 * corr1 + rs1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i][0] *= corr0;
        o[i][1] *= corr0;
        o[i][2] *= corr1;
        o[i][3] *= corr1;
      }
      // O += P V : P re-used from the S accumulators as A fragments (16 keys per k-step)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = ptx::pack2<BF16>(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = ptx::pack2<BF16>(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = ptx::pack2<BF16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = ptx::pack2<BF16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d n-tiles
          const int key = kc + kk * 16 + (lm & 1) * 8 + lr;
          const int du = dp * 2 + (lm >> 1);
          uint32_t b0, b1, b2, b3;
          ptx::ldmatrix_x4_trans(sV_u + key * 128 + ((du ^ (key & 7)) << 4), b0, b1, b2, b3);
          ptx::mma_16816<BF16>(o[2 * dp], pa, b0, b1);
          ptx::mma_16816<BF16>(o[2 * dp + 1], pa, b2, b3);
        }
      }
    }
    // finalise: quad-reduce the row sums, normalise, store 16-bit
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    uint16_t* ob = out + static_cast<size_t>(img) * L * D + h * HD;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 8 + 2 * t;
      if (r0 < L) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * D + c) = ptx::pack2<BF16>(o[i][0] * inv0, o[i][1] * inv0);
      if (r1 < L) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * D + c) = ptx::pack2<BF16>(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

// A few query rows per (image, head): one 8-warp CTA per row, no tensor cores.  Eight consecutive lanes share a key
// (one 16-byte chunk = 8 head dims each), so a warp reads four complete 128-byte K (V) rows per load instruction and
// every thread has several independent loads in flight.  Scores pass through smem (L <= 1024 floats); the partial
// sum_k p_k V_k is folded over the four key slots of a warp with shuffles and over the warps through smem, all fp32.
// Used for the <= 16 rows the 128-row tiles of the flash kernel leave over (L = 257: row 256).
constexpr int ROWS_MAX_L = 1024;

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = ptx::unpack2<BF16>(v.x), b = ptx::unpack2<BF16>(v.y), c = ptx::unpack2<BF16>(v.z),
               d = ptx::unpack2<BF16>(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <bool BF16>
__global__ void __launch_bounds__(256)
attention_rows_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int L, int H, int q_row0, int n_rows) {
  __shared__ float s_p[ROWS_MAX_L];
  __shared__ float s_red[16];
  __shared__ float s_o[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = lane >> 3, c = lane & 7;  // key slot of the warp, 16-byte chunk (head dims 8c .. 8c + 7)
  const int r = blockIdx.x % n_rows;
  const int u = blockIdx.x / n_rows;
  const int h = u % H;
  const int img = u / H;
  const int D = H * HD;
  const size_t ld = static_cast<size_t>(3) * D;
  const uint16_t* base = qkv + static_cast<size_t>(img) * L * ld + h * HD + 8 * c;
  const int qrow = q_row0 + r;
  const float sl2 = 0.125f * 1.4426950408889634f;

  float qf[8];
  unpack8<BF16>(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(qrow) * ld)), qf);
  float m = -INFINITY;
#pragma unroll 4
  for (int kb = warp * 4; kb < L; kb += 32) {  // warp-uniform trip count
    const int key = kb + ks;
    const bool valid = key < L;
    float kf[8];
    unpack8<BF16>(valid ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(key) * ld + D))
                        : make_uint4(0u, 0u, 0u, 0u), kf);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], kf[e], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (valid) {
      acc *= sl2;
      m = fmaxf(m, acc);
      if (c == 0) s_p[key] = acc;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);

  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  float l = 0.f;
#pragma unroll 4
  for (int kb = warp * 4; kb < L; kb += 32) {
    const int key = kb + ks;
    if (key < L) {
      const float pk = exp2f(s_p[key] - m);
      if (c == 0) l += pk;
      float vf[8];
      unpack8<BF16>(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(key) * ld + 2 * D)), vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(pk, vf[e], o[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {  // fold the four key slots of the warp
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
  if (ks == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_o[warp][8 * c + e] = o[e];
  }
  if (lane == 0) s_red[8 + warp] = l;
  __syncthreads();
  if (warp == 0) {
    float lt = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      lt += s_red[8 + w];
      o0 += s_o[w][2 * lane];
      o1 += s_o[w][2 * lane + 1];
    }
    const float inv = 1.0f / lt;
    uint16_t* dst = out + (static_cast<size_t>(img) * L + qrow) * D + h * HD;
    *reinterpret_cast<uint32_t*>(dst + 2 * lane) = ptx::pack2<BF16>(o0 * inv, o1 * inv);
  }
}

int g_attn_max_smem[64] = {};  // per device: cudaFuncSetAttribute is a per-device setting

}  // namespace

cudaError_t attention_init(int max_L) {
  const int Lp = (max_L + 63) / 64 * 64;
  const int smem = Lp * 256;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem <= g_attn_max_smem[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  g_attn_max_smem[dev] = smem;
  return cudaSuccess;
}

cudaError_t launch_attention(const void* qkv, void* out, int n_img, int L, int H, int is_bf16, cudaStream_t stream) {
  if (n_img <= 0) return cudaSuccess;
  if (L <= 0 || H <= 0) return cudaErrorInvalidValue;
  const int Lp = (L + 63) / 64 * 64;
  const int smem = Lp * 256;
  {
    cudaError_t e = attention_init(L);  // no-op once this device's limit covers the request
    if (e != cudaSuccess) return e;
  }
  const int q_tiles = (L + 15) / 16;
  const int rounds = (q_tiles + 7) / 8;  // <= 8 warps per CTA (128 registers per thread, 2 CTAs / SM)
  const int nwarps = (q_tiles + rounds - 1) / rounds;
  dim3 grid(H, n_img);
  if (is_bf16)
    attention_kernel<true><<<grid, nwarps * 32, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, Lp, 0);
  else
    attention_kernel<false><<<grid, nwarps * 32, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, Lp, 0);
  return cudaGetLastError();
}

cudaError_t launch_attention_rows(const void* qkv, void* out, int n_img, int L, int H, int is_bf16, int q_row0,
                                  cudaStream_t stream) {
  if (n_img <= 0 || q_row0 >= L) return cudaSuccess;
  if (L <= 0 || H <= 0 || q_row0 < 0 || (q_row0 & 15)) return cudaErrorInvalidValue;
  if (L - q_row0 <= 16 && L <= ROWS_MAX_L) {  // a handful of rows: one small CTA each, no smem staging
    const int n_rows = L - q_row0;
    const int grid = n_img * H * n_rows;
    if (is_bf16)
      attention_rows_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, q_row0, n_rows);
    else
      attention_rows_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, q_row0, n_rows);
    return cudaGetLastError();
  }
  const int Lp = (L + 63) / 64 * 64;
  const int smem = Lp * 256;
  {
    cudaError_t e = attention_init(L);
    if (e != cudaSuccess) return e;
  }
  // all 8 warps stage K / V; only the first (L - q_row0 + 15) / 16 of them own a query tile
  dim3 grid(H, n_img);
  if (is_bf16)
    attention_kernel<true><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, Lp, q_row0 >> 4);
  else
    attention_kernel<false><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), L, H, Lp, q_row0 >> 4);
  return cudaGetLastError();
}

}  // namespace aihab
