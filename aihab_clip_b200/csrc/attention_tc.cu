// tcgen05 / TMEM attention for sequences of 65..256 tokens (ViT-B/16: 197), head dim 64.
// One CTA = one 128-query tile of one (image, head); two CTAs are resident per SM (96 KB smem, 256 TMEM columns
// each) so one CTA's TMA loads / epilogue overlap the other's softmax.
//   warp 0 (lane 0): TMA loads Q [128x64], K [Lk x 64], V [Lk x 64] (SWIZZLE_128B) and issues the MMAs
//       S[128 x Lk] = Q K^T        4 x tcgen05.mma  (M=128, N=Lk, K=16), A and B K-major
//       O[128 x 64] = P V      Lk/16 x tcgen05.mma  (M=128, N=64, K=16), A = P K-major, B = V MN-major (no transpose)
//   warps 1..8: softmax, two threads per query row (TMEM lane = row, each thread owns half of the key chunks):
//       pass 1 row max, pass 2 exp2 / row sum (tcgen05.ld of chunk c+1 in flight while chunk c is processed), P packed
//       to 16-bit and written in the UMMA K-major SWIZZLE_128B layout over the (dead) Q/K buffers; after the PV MMAs
//       the same threads read O from TMEM (it aliases S's first 64 columns), scale by 1/sum and store.
// Keys >= L are masked (P = 0); query rows >= L are computed on garbage and never stored (rows are independent).
// Reference: clip/model.py:179-181 (nn.MultiheadAttention core: softmax(q k^T / sqrt(64)) v, no mask).
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace aihab {

namespace {

constexpr int TC_THREADS = 288;  // warp 0: TMA + MMA issue; warps 1..8: softmax / epilogue
constexpr int OFF_Q = 0;               // 16 KB
constexpr int OFF_K = 16 * 1024;       // up to 32 KB
constexpr int OFF_P = 0;               // 4 blocks x 16 KB, aliases Q and K once S is complete
constexpr int OFF_V = 64 * 1024;       // up to 32 KB
constexpr int OFF_BAR = 96 * 1024;
constexpr int OFF_RED = OFF_BAR + 64;   // row max [2][128] + row sum [2][128] fp32
constexpr int TC_SMEM = OFF_RED + 2048 + 1024;  // + alignment slack

template <bool BF16>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    uint16_t* __restrict__ out, int L, int H, int Lk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024 B aligned, still a __shared__ pointer
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_max = reinterpret_cast<float*>(smem + OFF_RED);
  float* s_sum = s_max + 256;

  const int qt = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int D = H * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_kv);
      ptx::mbar_init(bar_qk, 1);
      ptx::mbar_init(bar_v, 1);
      ptx::mbar_init(bar_s, 1);
      ptx::mbar_init(bar_p, 256);
      ptx::mbar_init(bar_o, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = img * L;
      ptx::mbar_expect_tx(bar_qk, 128 * 128 + Lk * 128);
      ptx::tma_load_2d(smem + OFF_Q, &tmap_q, bar_qk, h * 64, row0 + qt * 128);
      ptx::tma_load_2d(smem + OFF_K, &tmap_kv, bar_qk, D + h * 64, row0);
      ptx::mbar_expect_tx(bar_v, Lk * 128);
      ptx::tma_load_2d(smem + OFF_V, &tmap_kv, bar_v, 2 * D + h * 64, row0);

      // S = Q K^T
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      const uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, Lk);
      const uint64_t qd = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + OFF_Q));
      const uint64_t kd = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + OFF_K));
#pragma unroll
      for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
      ptx::umma_commit(bar_s);

      // O = P V  (O overwrites S columns [0, 64): every softmax thread has finished reading S when bar_p completes)
      ptx::mbar_wait(bar_v, 0);
      ptx::mbar_wait(bar_p, 0);
      ptx::tc_fence_after();
      const uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, 64, /*b_mn_major=*/1);
      const uint32_t p_base = ptx::smem_u32(smem + OFF_P), v_base = ptx::smem_u32(smem + OFF_V);
      const int ksteps = Lk >> 4;
      for (int j = 0; j < ksteps; ++j) {
        const uint64_t pd = ptx::make_kmajor_sw128_desc(p_base + (j >> 2) * 16384 + (j & 3) * 32);
        const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + j * 2048);  // 16 keys x 128 B
        ptx::umma_f16(tmem, pd, vd, idesc_o, j != 0);
      }
      ptx::umma_commit(bar_o);
    }
  } else {
    // ---- softmax / epilogue: 8 warps, two threads per query row (each owns half of the key chunks / O columns)
    const int sw = warp - 1;                  // 0..7
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = sw >> 2;
    const int row = quad * 32 + lane;         // query row inside the tile
    const int grow = qt * 128 + row;          // token index inside the image
    const bool warp_has_rows = qt * 128 + quad * 32 < L;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;
    const int chunks = Lk >> 4;
    const int c_begin = half ? (chunks + 1) >> 1 : 0;
    const int c_end = half ? chunks : (chunks + 1) >> 1;

    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float m = -INFINITY;
    if (warp_has_rows) {
      // pass 1: row max over this thread's chunks; the load of chunk c+1 is in flight while chunk c is reduced
      uint32_t ra[16], rb[16];
      ptx::tmem_ld_32x16(t_row + c_begin * 16, ra);
      ptx::tmem_ld_wait();
      for (int c = c_begin; c < c_end; c += 2) {
        if (c + 1 < c_end) ptx::tmem_ld_32x16(t_row + (c + 1) * 16, rb);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c * 16 + j < L) m = fmaxf(m, __uint_as_float(ra[j]));
        ptx::tmem_ld_wait();
        if (c + 1 < c_end) {
          if (c + 2 < c_end) ptx::tmem_ld_32x16(t_row + (c + 2) * 16, ra);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((c + 1) * 16 + j < L) m = fmaxf(m, __uint_as_float(rb[j]));
          ptx::tmem_ld_wait();
        }
      }
      s_max[half * 128 + row] = m;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float l = 0.f;
    if (warp_has_rows) {
      m = fmaxf(s_max[row], s_max[128 + row]);
      const float ms = m * sl2;
      uint8_t* prow = smem + OFF_P + row * 128;
      auto emit = [&](int c, const uint32_t (&r)[16]) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c0 = c * 16 + 2 * j;
          float p0 = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * j]), sl2, -ms));
          float p1 = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -ms));
          if (c0 >= L) p0 = 0.f;
          if (c0 + 1 >= L) p1 = 0.f;
          l += p0 + p1;
          pk[j] = ptx::pack2<BF16>(p0, p1);
        }
        // 16 keys = two 16 B units of the 128 B row of key block c / 4
        uint8_t* blk = prow + (c >> 2) * 16384;
        const int u = (c & 3) * 2;
        *reinterpret_cast<uint4*>(blk + (((u) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(blk + (((u + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      uint32_t ra[16], rb[16];
      ptx::tmem_ld_32x16(t_row + c_begin * 16, ra);
      ptx::tmem_ld_wait();
      for (int c = c_begin; c < c_end; c += 2) {
        if (c + 1 < c_end) ptx::tmem_ld_32x16(t_row + (c + 1) * 16, rb);
        emit(c, ra);
        ptx::tmem_ld_wait();
        if (c + 1 < c_end) {
          if (c + 2 < c_end) ptx::tmem_ld_32x16(t_row + (c + 2) * 16, ra);
          emit(c + 1, rb);
          ptx::tmem_ld_wait();
        }
      }
      s_sum[half * 128 + row] = l;
    }
    ptx::tc_fence_before();      // all tcgen05.ld of S retired (wait::ld above) before O may overwrite it
    ptx::fence_proxy_async();    // P written through the generic proxy -> visible to the MMA (async proxy)
    ptx::mbar_arrive(bar_p);

    ptx::mbar_wait(bar_o, 0);    // also orders the s_sum writes of the partner thread (it arrived on bar_p first)
    ptx::tc_fence_after();
    if (warp_has_rows) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(t_row + half * 32, o);
      ptx::tmem_ld_wait();
      if (grow < L) {
        const float inv_l = 1.0f / (s_sum[row] + s_sum[128 + row]);
        uint16_t* dst = out + (static_cast<size_t>(img) * L + grow) * D + h * 64 + half * 32;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = ptx::pack2<BF16>(__uint_as_float(o[8 * u]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l);
          v.y = ptx::pack2<BF16>(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l);
          v.z = ptx::pack2<BF16>(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l);
          v.w = ptx::pack2<BF16>(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + 8 * u) = v;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

bool g_tc_init[64] = {};  // per device

}  // namespace

bool attention_tc_supported(int L) { return L > 64 && L <= 256; }

int attention_tc_key_rows(int L) { return (L + 15) / 16 * 16; }

cudaError_t launch_attention_tc(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, void* out, int n_img, int L,
                                int H, int is_bf16, cudaStream_t stream) {
  if (n_img <= 0) return cudaSuccess;
  if (!attention_tc_supported(L)) return cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!g_tc_init[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    g_tc_init[dev] = true;
  }
  const int Lk = attention_tc_key_rows(L);
  dim3 grid((L + 127) / 128, H, n_img);
  if (is_bf16)
    attention_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM, stream>>>(tmap_q, tmap_kv, static_cast<uint16_t*>(out), L, H, Lk);
  else
    attention_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, stream>>>(tmap_q, tmap_kv, static_cast<uint16_t*>(out), L, H, Lk);
  return cudaGetLastError();
}

}  // namespace aihab
