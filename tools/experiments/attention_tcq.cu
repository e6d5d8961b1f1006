// EXPERIMENT, NOT BUILT (round 2; measured slower, kept as a record - see profiles/attention_r2.txt).  To try it again: copy
// into aihab_clip_b200/csrc/, add it to the Makefile and dispatch to launch_attention_tcq() in launch_attention_tcp().
// Split-row variant of the dual-stream attention kernel (attention_tcd.cu) for 128 < L <= 224: the same two streams
// (the two 128-row query tiles of an (image, head) unit, sharing K and V in shared memory) but TWO threads per query
// row, each owning half of the key columns - 16 softmax warps, four per scheduler instead of two.
// Why: in attention_tcd the period of a unit is the length of ONE stream's dependent chain (maximum pass over 208
// columns, exp2 pass over 208 columns, O epilogue over 64 columns: 7 450 of 7 650 cycles, profiles/attention_r2.txt) while
// the MUFU pipe is busy 47 % and the issue slots 52 %.  Halving the columns per thread halves every phase of the chain;
// the price is two shared-memory exchanges per unit (row maximum, row sum) behind a 256-thread named barrier.
//   warp 16           TMA producer (as attention_tcd)
//   warps 17, 18      MMA issuers, one per stream
//   warps 0..15       softmax: warp = 8 g + 4 half + quad (warp % 4 = TMEM lane quadrant); half 0 owns the key groups
//                     [0, h0) of 16 columns, half 1 the groups [h0, Lk / 16), h0 = ceil(Lk / 32)
// TMEM plan of a stream (256 columns): S fp32 [0, Lk).  Every half packs its P INSIDE ITS OWN S range, behind its read
// pointer: group j -> 8 columns at 8 j (half 0) or 16 h0 + 8 (j - h0) (half 1), so no half ever overwrites S columns the
// other one still has to read.  O fp32 -> [8 (Lk / 16 + h0), + 64): behind both P regions, written by the PV MMAs only
// after the whole row's P is in TMEM.  The O epilogue is split the same way (32 columns per thread); the two warps of a
// (stream, quadrant) fill one swizzled 32-row x 128 B staging tile and one of them issues its TMA store.
// Reference: clip/model.py:179-181 (nn.MultiheadAttention core: softmax(q k^T / sqrt(64)) v, no mask).
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

// -DAIHAB_TCQ_NO_TURN: no exp2-phase token between the streams (A/B experiment)
namespace aihab {

namespace {

constexpr int Q_THREADS = 608;  // 16 softmax warps + TMA + 2 issuers
constexpr int NSTG = 2;
constexpr int ST_Q0 = 0;
constexpr int ST_Q1 = 16384;
constexpr int ST_K = 32768;
__host__ __device__ constexpr int st_v(int Lk) { return ST_K + Lk * 128; }
__host__ __device__ constexpr int st_bytes(int Lk) { return ST_K + 2 * Lk * 128; }
constexpr int BAR_BYTES = 256;
constexpr int STG_BYTES = 8 * 4096;        // O staging: one 32-row x 128 B SWIZZLE_128B tile per (stream, quadrant)
constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4;  // row maximum | row sum exchange: [kind][stream][half][row] fp32
__host__ __device__ constexpr int smem_bytes(int Lk) { return NSTG * st_bytes(Lk) + STG_BYTES + XCH_BYTES + BAR_BYTES + 1024; }
constexpr int SMEM_CAP = 227 * 1024;
constexpr int TM_STREAM = 256;
static_assert(smem_bytes(224) <= SMEM_CAP, "smem budget");

__device__ __forceinline__ void named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(Q_THREADS, 1)
attention_tcq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_out, int L, int H, int Lk, int total, int reverse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int ST_V = st_v(Lk), ST_BYTES = st_bytes(Lk);
  uint8_t* staging = smem + NSTG * ST_BYTES;  // [8 (stream, quadrant)][32 rows][128 B], 1024 B aligned
  float* xch = reinterpret_cast<float*>(smem + NSTG * ST_BYTES + STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTG * ST_BYTES + STG_BYTES + XCH_BYTES);
  uint64_t* bar_qk = bars + 0;      // [NSTG] Q0 + Q1 + K of a stage landed
  uint64_t* bar_v = bars + 2;       // [NSTG] V of a stage landed
  uint64_t* bar_kfree = bars + 4;   // [NSTG] both streams' S MMAs of the stage retired
  uint64_t* bar_sfull = bars + 6;   // [2] S of stream g written
  uint64_t* bar_p = bars + 8;       // [2] P of stream g written to TMEM (8 warp arrivals)
  uint64_t* bar_o = bars + 10;      // [2] O of stream g written
  uint64_t* bar_bfree = bars + 12;  // [2] O of stream g read (8 warp arrivals)
  uint64_t* bar_vfree = bars + 14;  // [NSTG] both streams' PV MMAs of the stage retired
  uint64_t* bar_turn = bars + 16;   // [2][4] exp2-phase token of stream g, quadrant q (2 arrivals: both halves)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int D = H * 64;
  const int n_units = blockIdx.x < total ? (total - static_cast<int>(blockIdx.x) + G - 1) / G : 0;
  auto decode = [&](int u, int& img, int& h) {
    int idx = static_cast<int>(blockIdx.x) + u * G;
    if (reverse) idx = total - 1 - idx;
    img = idx / H;
    h = idx - img * H;
  };
  const int n16 = Lk >> 4;          // key groups of 16 columns
  const int h0 = (n16 + 1) >> 1;    // groups of half 0
  const int o_col = 8 * (n16 + h0);  // O columns inside the stream's TMEM buffer
  auto p_col = [&](int j) { return j < h0 ? 8 * j : 16 * h0 + 8 * (j - h0); };

  ptx::griddep_launch();
  if (warp == 16 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    ptx::prefetch_tmap(&tmap_out);
    for (int i = 0; i < NSTG; ++i) {
      ptx::mbar_init(&bar_qk[i], 1);
      ptx::mbar_init(&bar_v[i], 1);
      ptx::mbar_init(&bar_kfree[i], 2);
      ptx::mbar_init(&bar_vfree[i], 2);
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar_turn[i], 2);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar_sfull[i], 1);
      ptx::mbar_init(&bar_p[i], 8);
      ptx::mbar_init(&bar_o[i], 1);
      ptx::mbar_init(&bar_bfree[i], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 17) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::griddep_wait();

  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int u = 0; u < n_units; ++u) {
        int img, h;
        decode(u, img, h);
        const int s = u % NSTG, k = u / NSTG;
        uint8_t* st = smem + s * ST_BYTES;
        if (u >= NSTG) ptx::mbar_wait(&bar_kfree[s], (k - 1) & 1);
        const int row0 = img * L;
        ptx::mbar_expect_tx(&bar_qk[s], 2 * 16384 + Lk * 128);
        ptx::tma_load_2d(st + ST_K, &tmap_kv, &bar_qk[s], D + h * 64, row0);
        ptx::tma_load_2d(st + ST_Q0, &tmap_q, &bar_qk[s], h * 64, row0);
        ptx::tma_load_2d(st + ST_Q1, &tmap_q, &bar_qk[s], h * 64, row0 + 128);
        if (u >= NSTG) ptx::mbar_wait(&bar_vfree[s], (k - 1) & 1);
        ptx::mbar_expect_tx(&bar_v[s], Lk * 128);
        ptx::tma_load_2d(st + ST_V, &tmap_kv, &bar_v[s], 2 * D + h * 64, row0);
      }
    }
  } else if (warp >= 17) {
    // ------------------------------------------------------------------ MMA issuer of stream g (whole warp, converged)
    const int g = warp - 17;
    const uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, Lk);
    const uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, 64, /*b_mn_major=*/1);
    const uint32_t tbuf = tmem + g * TM_STREAM;
    for (int u = 0; u < n_units; ++u) {
      const int s = u % NSTG, ks = u / NSTG;
      ptx::mbar_wait(&bar_qk[s], ks & 1);
      if (u > 0) ptx::mbar_wait(&bar_bfree[g], (u - 1) & 1);  // O(u-1) has been read out of this buffer
      ptx::tc_fence_after();
      const uint32_t st = ptx::smem_u32(smem + s * ST_BYTES);
      const uint64_t qd = ptx::make_kmajor_sw128_desc(st + (g ? ST_Q1 : ST_Q0));
      const uint64_t kd = ptx::make_kmajor_sw128_desc(st + ST_K);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_w(tbuf, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
      ptx::umma_commit_w(&bar_sfull[g]);
      ptx::umma_commit_w(&bar_kfree[s]);
      const uint32_t v_base = st + ST_V;
      ptx::mbar_wait(&bar_p[g], u & 1);
      ptx::mbar_wait(&bar_v[s], ks & 1);
      ptx::tc_fence_after();
      for (int j = 0; j < n16; ++j) {
        const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + j * 2048);
        ptx::umma_f16_ts_w(tbuf + o_col, tbuf + p_col(j), vd, idesc_o, j != 0);  // A = P from TMEM
      }
      ptx::umma_commit_w(&bar_o[g]);
      ptx::umma_commit_w(&bar_vfree[s]);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: two threads per query row
    const int g = warp >> 3;
    const int hf = (warp >> 2) & 1;
    const int quad = warp & 3;
    const bool has_rows = g * 128 + quad * 32 < L;  // warp-uniform (and the same for both halves)
    const uint32_t t_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + g * TM_STREAM;
    const float sl2 = 0.125f * 1.4426950408889634f;
    const int j_lo = hf ? h0 : 0, j_hi = hf ? n16 : h0;  // this thread's key groups
    const int row = quad * 32 + lane;
    float* x_mine = xch + (g * 2 + hf) * 128 + row;                 // row maximum of this half
    const float* x_other = xch + (g * 2 + (hf ^ 1)) * 128 + row;
    float* s_mine = x_mine + 512;                                   // partial row sum of this half (separate slots: a
    const float* s_other = x_other + 512;                           // half may still read the maximum when the other is done)
    const int sbar = 1 + g;                // named barrier of the stream's 8 softmax warps
    const int pbar = 3 + g * 4 + quad;     // named barrier of the two warps that share a staging tile

    for (int u = 0; u < n_units; ++u) {
      int img, h;
      decode(u, img, h);
      ptx::mbar_wait(&bar_sfull[g], u & 1);
      ptx::tc_fence_after();
      float l = 0.f;
      if (has_rows) {
        // ---- pass 1: maximum of this half of the row, two 16-column loads in flight (96 registers per thread at 608 threads)
        float m0 = -INFINITY, m1 = -INFINITY;
        auto max16 = [&](const uint32_t (&r)[16], int j) {
          if (j * 16 + 16 <= L) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(r[i]));
              m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (j * 16 + i < L) m0 = fmaxf(m0, __uint_as_float(r[i]));
          }
        };
#pragma unroll 1
        for (int j = j_lo; j < j_hi; j += 2) {
          uint32_t r0[16], r1[16];
          ptx::tmem_ld_32x16(t_row + j * 16, r0);
          if (j + 1 < j_hi) ptx::tmem_ld_32x16(t_row + (j + 1) * 16, r1);
          ptx::tmem_ld_wait();
          max16(r0, j);
          if (j + 1 < j_hi) max16(r1, j + 1);
        }
        *x_mine = fmaxf(m0, m1);
      }
      named_sync(sbar, 256);  // both halves' maxima are in shared memory
      if (has_rows) {
        const float ms = fmaxf(*x_mine, *x_other) * sl2;
        // ---- pass 2: P = exp2(S * scale - max) packed behind the read pointer, partial row sum in the thread
#ifndef AIHAB_TCQ_NO_TURN
        ptx::mbar_wait(&bar_turn[g * 4 + quad], (u & 1) ^ (g == 0 ? 1 : 0));
#endif
        float l0 = 0.f, l1 = 0.f;
        uint32_t ra[16], rb[16];
        auto exp16 = [&](uint32_t (&r)[16], int j) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(ptx::ex2_approx(fmaf(__uint_as_float(r[i]), sl2, -ms)));
          if (j * 16 + 16 > L) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (j * 16 + i >= L) r[i] = 0u;
          }
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            l0 += __uint_as_float(r[2 * i]);
            l1 += __uint_as_float(r[2 * i + 1]);
            pk[i] = ptx::pack2<BF16>(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          }
          ptx::tmem_st_32x8(t_row + p_col(j), pk);  // 16 keys -> 8 packed columns
        };
        ptx::tmem_ld_32x16(t_row + j_lo * 16, ra);
#pragma unroll 1
        for (int j = j_lo; j < j_hi; j += 2) {
          ptx::tmem_ld_wait_regs16(ra);
          if (j + 1 < j_hi) ptx::tmem_ld_32x16(t_row + (j + 1) * 16, rb);
          exp16(ra, j);
          if (j + 1 < j_hi) {
            ptx::tmem_ld_wait_regs16(rb);
            if (j + 2 < j_hi) ptx::tmem_ld_32x16(t_row + (j + 2) * 16, ra);
            exp16(rb, j + 1);
          }
        }
#ifndef AIHAB_TCQ_NO_TURN
        if (lane == 0) ptx::mbar_arrive(&bar_turn[(g ^ 1) * 4 + quad]);  // the other stream's turn (after both halves)
#endif
        l = l0 + l1;
        *s_mine = l;
        ptx::tmem_st_wait();
      } else {
#ifndef AIHAB_TCQ_NO_TURN
        ptx::mbar_wait(&bar_turn[g * 4 + quad], (u & 1) ^ (g == 0 ? 1 : 0));
        if (lane == 0) ptx::mbar_arrive(&bar_turn[(g ^ 1) * 4 + quad]);
#endif
      }
      ptx::tc_fence_before();  // P stored (wait::st) before the PV MMA may read it
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_p[g]);
      named_sync(sbar, 256);  // both halves' partial sums are in shared memory (and nobody still reads the maxima)
      if (has_rows) l += *s_other;

      ptx::mbar_wait(&bar_o[g], u & 1);
      ptx::tc_fence_after();
      uint32_t o[32];
      if (has_rows) {
        ptx::tmem_ld_32x32(t_row + o_col + 32 * hf, o);
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();  // O read (wait::ld) before the next S MMA may overwrite the buffer
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_bfree[g]);
      if (has_rows) {
        // x 1/sum -> this half's 64 B of the 16-bit row -> swizzled staging tile shared with the other half's warp ->
        // ONE TMA store of the 32-row x 64-column box (rows past the end of the sequence are clipped by the 3-D map)
        uint8_t* stg = staging + (g * 4 + quad) * 4096;
        if (u > 0) {  // the previous store (issued by the half-0 warp) has finished reading the tile
          if (hf == 0 && lane == 0) ptx::bulk_wait_read<0>();
          named_sync(pbar, 64);
        }
        const float inv_l = 1.0f / l;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = ptx::pack2<BF16>(__uint_as_float(o[8 * q]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
          v.y = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
          v.z = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
          v.w = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * hf + q) ^ (lane & 7)) << 4)) = v;
        }
        ptx::fence_proxy_async();
        named_sync(pbar, 64);  // both halves of the tile are written
        if (hf == 0 && lane == 0) {
          ptx::tma_store_3d(&tmap_out, stg, h * 64, g * 128 + quad * 32, img);
          ptx::bulk_commit();
        }
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

bool attention_tcq_supported(int L) {
  static const bool enabled = [] {
    const char* e = getenv("AIHAB_ATTN_SPLIT");
    return e != nullptr && e[0] == '1';
  }();
  return enabled && L > 128 && L <= 224;
}

cudaError_t launch_attention_tcq(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const CUtensorMap& tmap_out,
                                 int n_img, int L, int H, int is_bf16, int num_sms, cudaStream_t stream, int reverse) {
  if (n_img <= 0) return cudaSuccess;
  if (L <= 128 || L > 224) return cudaErrorInvalidValue;
  const int Lk = (L + 15) / 16 * 16;
  const int total = n_img * H;
  const int grid = total < num_sms ? total : num_sms;
  static bool attr_set[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev][is_bf16 ? 1 : 0]) {
    cudaError_t e = is_bf16 ? cudaFuncSetAttribute(attention_tcq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP)
                            : cudaFuncSetAttribute(attention_tcq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP);
    if (e != cudaSuccess) return e;
    attr_set[dev][is_bf16 ? 1 : 0] = true;
  }
  if (is_bf16)
    return launch_kernel(attention_tcq_kernel<true>, grid, Q_THREADS, smem_bytes(Lk), stream, 1, true, tmap_q, tmap_kv, tmap_out, L,
                         H, Lk, total, reverse);
  return launch_kernel(attention_tcq_kernel<false>, grid, Q_THREADS, smem_bytes(Lk), stream, 1, true, tmap_q, tmap_kv, tmap_out, L,
                       H, Lk, total, reverse);
}

}  // namespace aihab
