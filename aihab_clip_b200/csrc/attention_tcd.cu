// Dual-stream persistent tcgen05 / TMEM attention for 128 < L <= 224 (ViT-B/16: L = 197, Lk = 208, two 128-row query
// tiles per (image, head)).  One CTA per SM walks (image, head) UNITS; the two query tiles of a unit are two independent
// STREAMS that share the unit's K and V tiles in shared memory (loaded once, not once per tile):
//   warp 8            TMA producer : Q0, Q1 [128 x 64], K, V [Lk x 64] of unit u+1 land in the other stage
//   warps 9, 10       MMA issuers  : one per stream (converged warp, one elected lane issues).  S = Q K^T into the
//                                    stream's TMEM buffer, later O = P V (A = P from TMEM, V as an MN-major smem operand)
//   warps 0..3, 4..7  softmax      : ONE THREAD PER QUERY ROW.  Pass 1 reads the row's S from TMEM for its maximum, pass 2
//                                    reads it again, exp2, packs P as 16-bit pairs back INTO the S columns
//                                    (tcgen05.st); row max and row sum stay in the thread - no shared-memory exchange,
//                                    no named barrier; the O epilogue is done by the same thread: x 1/sum -> 16-bit row ->
//                                    swizzled smem tile -> one TMA store per warp through a 3-D map [image][row][col]
//                                    that clips the partial last tile at the end of its image.
// Why two streams: the softmax is bound by the MUFU pipe (16 exp2 / clock / SM) in its exp2 phase and does not touch
// it in the other phases (S wait, maximum, P hand-off, PV wait, O read).  In the single-stream kernel
// (attention_tcp.cu) all eight softmax warps are in the same phase at the same time, so MUFU idles ~55 % of an item.
// Here the two streams take TURNS in the exp2 phase (a per-SM-sub-partition token, bar_turn: the softmax warp of stream
// A and the one of stream B that share a scheduler ping-pong), so one stream's exp2 phase fills the other's MUFU-free
// phases instead of both competing for the pipe at the same time and idling together afterwards.
// TMEM: stream g owns columns [256 g, 256 g + 224): S fp32 [0, Lk) -> P 16-bit pairs [0, Lk / 2) -> O fp32 [128, 192)
// (S is dead once the row's P is written, so P and O overlay it).  smem: 2 stages x (Q0 16 KB | Q1 16 KB | K | V).
// Reference: clip/model.py:179-181 (nn.MultiheadAttention core: softmax(q k^T / sqrt(64)) v, no mask).
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#ifdef AIHAB_ATTN_TIMING
// -DAIHAB_ATTN_TIMING (tools/attn_timing.py): CTA 0 stamps clock64() at the phase boundaries of units 4..7 for every
// role; read back with aihab_debug_attn_timing().  [role 0..3][unit][event 0..15]; roles: 0/1 issuer A/B, 2/3 softmax A/B quad 0
__device__ long long g_attn_ts[4 * 4 * 16];
#define TSTAMP(role, ev)                                                                                        \
  do {                                                                                                          \
    if (blockIdx.x == 0 && lane == 0 && u >= 4 && u < 8) g_attn_ts[((role) * 4 + (u - 4)) * 16 + (ev)] = clock64(); \
  } while (0)
extern "C" __attribute__((visibility("default"))) int aihab_debug_attn_timing(long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, g_attn_ts, sizeof(g_attn_ts)));
}
#else
#define TSTAMP(role, ev) do {} while (0)
#endif

// exp2 of every element j of a 32-column chunk with (AIHAB_ATTN_POLYMASK >> (j % 8)) & 1 set is evaluated on the FMA
// pipe instead of the MUFU pipe (measured on B200: one warp-wide MUFU.EX2 per ~16 cycles and scheduler, which bounds the
// softmax): Cody-Waite split x = floor(x) + f, 2^f by a degree-3 minimax polynomial (max relative error 8.8e-5, below
// the 4.9e-4 half-ulp of the 16-bit P it is rounded to), 2^floor(x) by an integer add into the exponent field.
// Default 0: measured on B200 (profiles/attention_r2.txt) the polynomial is SLOWER for every fraction tried
// (0.0872 ms at 0, 0.0905 at 1/4, 0.0956 at 1/2, 0.1135 at 3/4): the exp2 pass is not MUFU-bound in this kernel.
#ifndef AIHAB_ATTN_POLYMASK
#define AIHAB_ATTN_POLYMASK 0x00
#endif
// timing experiments only (results are wrong): bit 0 no P store, bit 1 no S load in the exp pass, bit 2 no exp2,
// bit 3 no turn token, bit 4 no max pass
#ifndef AIHAB_ATTN_EXPERIMENT
#define AIHAB_ATTN_EXPERIMENT 0
#endif

// waits of the softmax / issuer warps: -DAIHAB_TCD_BACKOFF=<ns> sleeps between polls (experiment: do polling warps take
// issue slots from the working warp of their scheduler?)
#ifndef AIHAB_TCD_BACKOFF
#define AIHAB_TCD_BACKOFF 0
#endif

namespace aihab {

namespace {

__device__ __forceinline__ void tcd_wait(uint64_t* bar, uint32_t parity) {
#if AIHAB_TCD_BACKOFF > 0
  ptx::mbar_wait_backoff(bar, parity, AIHAB_TCD_BACKOFF);
#else
  ptx::mbar_wait(bar, parity);
#endif
}

__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = __fadd_rd(x, 12582912.0f);  // 1.5 * 2^23: floor(x) sits in the low mantissa bits
  const float f = x - (t - 12582912.0f);      // [0, 1)
  float p = fmaf(0.077119089663028717041015625f, f, 0.227564394474029541015625f);
  p = fmaf(p, f, 0.695146143436431884765625f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

constexpr int D_THREADS = 352;  // warps 0..3 / 4..7 softmax of stream A / B (warp % 4 = TMEM lane quadrant), 8 TMA, 9 / 10 MMA issuers
constexpr int NSTG = 2;
constexpr int ST_Q0 = 0;
constexpr int ST_Q1 = 16384;
constexpr int ST_K = 32768;
__host__ __device__ constexpr int st_v(int Lk) { return ST_K + Lk * 128; }
__host__ __device__ constexpr int st_bytes(int Lk) { return ST_K + 2 * Lk * 128; }  // Lk = 224: 90112
constexpr int BAR_BYTES = 256;
constexpr int STG_BYTES = 8 * 4096;  // O staging: one 32-row x 128 B SWIZZLE_128B tile per softmax warp
__host__ __device__ constexpr int smem_bytes(int Lk) { return NSTG * st_bytes(Lk) + STG_BYTES + BAR_BYTES + 1024; }
constexpr int SMEM_CAP = 227 * 1024;
constexpr int TM_STREAM = 256;  // TMEM columns per stream
constexpr int P_SPLIT = 4;      // 32-key chunks of P handed to the PV MMA early (while the exp2 pass does the rest)
// Column plan inside a stream's 256 TMEM columns.  Everything that is written while the exp2 pass still reads S must
// land on columns whose S values are already consumed (chunks are read in ascending order) or on the spare columns
// behind S.  EARLY plan (Lk <= 208, more than P_SPLIT chunks): P of chunks 0..2 -> spare [208, 256), P of chunk c >= 3
// -> [16 (c - 3), ...) inside S chunks 0, 1, O -> [64, 128) = S chunks 2, 3: after chunk 3 the PV MMAs over the first
// four chunks may run while chunks 4.. are still being read.  PLAIN plan: P of chunk c -> [16 c, ...), O -> [128, 192),
// PV only after the whole row is done.
__host__ __device__ constexpr bool plan_early(int Lk) { return Lk <= 208 && (Lk >> 5) > P_SPLIT; }
__device__ __forceinline__ int p_col(bool early, int c) { return early ? (c < 3 ? 208 + 16 * c : 16 * (c - 3)) : 16 * c; }
__device__ __forceinline__ int o_col(bool early) { return early ? 64 : 128; }
static_assert(smem_bytes(224) <= SMEM_CAP, "smem budget");
static_assert(st_v(144) % 1024 == 0 && st_bytes(144) % 1024 == 0, "SWIZZLE_128B tiles need 1024 B alignment");

template <bool BF16>
__global__ void __launch_bounds__(D_THREADS, 1)
attention_tcd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_out, int L, int H, int Lk, int total, int reverse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int ST_V = st_v(Lk), ST_BYTES = st_bytes(Lk);
  uint8_t* staging = smem + NSTG * ST_BYTES;  // [8 warps][32 rows][128 B], 1024 B aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTG * ST_BYTES + STG_BYTES);
  uint64_t* bar_qk = bars + 0;       // [NSTG] Q0 + Q1 + K of a stage landed
  uint64_t* bar_v = bars + 2;        // [NSTG] V of a stage landed
  uint64_t* bar_kfree = bars + 4;    // [NSTG] both streams' S MMAs of the stage retired: Q0, Q1, K may be overwritten
  uint64_t* bar_vfree = bars + 16;   // [NSTG] both streams' PV MMAs of the stage retired: V may be overwritten
  uint64_t* bar_turn = bars + 18;    // [2][4] exp2-phase token of stream g, quadrant q
  uint64_t* bar_p1 = bars + 26;      // [2] first P_SPLIT key chunks of P written: the PV MMAs over them may start
  uint64_t* bar_sfull = bars + 6;    // [2] S of stream g written
  uint64_t* bar_p = bars + 8;        // [2] P of stream g written to TMEM (4 warp arrivals)
  uint64_t* bar_o = bars + 10;       // [2] O of stream g written
  uint64_t* bar_bfree = bars + 12;   // [2] O of stream g read: its TMEM buffer is free for the next S (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int D = H * 64;
  const int n_units = blockIdx.x < total ? (total - static_cast<int>(blockIdx.x) + G - 1) / G : 0;
  auto decode = [&](int u, int& img, int& h) {
    int idx = static_cast<int>(blockIdx.x) + u * G;
    if (reverse) idx = total - 1 - idx;  // walk the units from the end: the producer's freshest rows first
    img = idx / H;
    h = idx - img * H;
  };

  ptx::griddep_launch();
  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    ptx::prefetch_tmap(&tmap_out);
    for (int i = 0; i < NSTG; ++i) {
      ptx::mbar_init(&bar_qk[i], 1);
      ptx::mbar_init(&bar_v[i], 1);
      ptx::mbar_init(&bar_kfree[i], 2);
      ptx::mbar_init(&bar_vfree[i], 2);
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar_turn[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar_sfull[i], 1);
      ptx::mbar_init(&bar_p[i], 4);
      ptx::mbar_init(&bar_o[i], 1);
      ptx::mbar_init(&bar_bfree[i], 4);
      ptx::mbar_init(&bar_p1[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::griddep_wait();  // qkv comes from the previous kernel of the stream

  if (warp == 8) {
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
        if (u >= NSTG) ptx::mbar_wait(&bar_vfree[s], (k - 1) & 1);  // V lives until the stage's PV MMAs retire
        ptx::mbar_expect_tx(&bar_v[s], Lk * 128);
        ptx::tma_load_2d(st + ST_V, &tmap_kv, &bar_v[s], 2 * D + h * 64, row0);
      }
    }
  } else if (warp >= 9) {
    // ------------------------------------------------------------------ MMA issuer of stream g (whole warp, converged)
    const int g = warp - 9;
    const uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, Lk);
    const uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, 64, /*b_mn_major=*/1);
    const int ksteps = Lk >> 4;
    const uint32_t tbuf = tmem + g * TM_STREAM;
    for (int u = 0; u < n_units; ++u) {
      const int s = u % NSTG, ks = u / NSTG;
      TSTAMP(g, 0);
      tcd_wait(&bar_qk[s], ks & 1);
      TSTAMP(g, 1);
      if (u > 0) tcd_wait(&bar_bfree[g], (u - 1) & 1);  // O(u-1) has been read out of this buffer
      ptx::tc_fence_after();
      TSTAMP(g, 2);
      const uint32_t st = ptx::smem_u32(smem + s * ST_BYTES);
      const uint64_t qd = ptx::make_kmajor_sw128_desc(st + (g ? ST_Q1 : ST_Q0));
      const uint64_t kd = ptx::make_kmajor_sw128_desc(st + ST_K);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_w(tbuf, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
      ptx::umma_commit_w(&bar_sfull[g]);
      ptx::umma_commit_w(&bar_kfree[s]);
      TSTAMP(g, 3);
      // O = P V in two instalments: the k-steps of the first P_SPLIT chunks as soon as they are in TMEM (the softmax
      // is still in its exp2 pass over the remaining keys), the rest after the full hand-off
      const uint32_t v_base = st + ST_V;
      const bool early = plan_early(Lk);
      const int k_early = early ? 2 * P_SPLIT : 0;
      const uint32_t t_o = tbuf + o_col(early);
      if (k_early) {
        tcd_wait(&bar_p1[g], u & 1);
        tcd_wait(&bar_v[s], ks & 1);
        ptx::tc_fence_after();
        for (int j = 0; j < k_early; ++j) {
          const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + j * 2048);
          ptx::umma_f16_ts_w(t_o, tbuf + p_col(early, j >> 1) + (j & 1) * 8, vd, idesc_o, j != 0);  // A = P from TMEM
        }
      }
      tcd_wait(&bar_p[g], u & 1);
      TSTAMP(g, 4);
      tcd_wait(&bar_v[s], ks & 1);
      ptx::tc_fence_after();
      for (int j = k_early; j < ksteps; ++j) {
        const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + j * 2048);
        ptx::umma_f16_ts_w(t_o, tbuf + p_col(early, j >> 1) + (j & 1) * 8, vd, idesc_o, j != 0);  // A = P from TMEM
      }
      ptx::umma_commit_w(&bar_o[g]);
      ptx::umma_commit_w(&bar_vfree[s]);
      TSTAMP(g, 5);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue of stream g: a thread per row
    const int g = warp >> 2;
    const int quad = warp & 3;
    const bool has_rows = g * 128 + quad * 32 < L;       // warp-uniform
    const uint32_t t_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + g * TM_STREAM;
    const float sl2 = 0.125f * 1.4426950408889634f;
    const int n32 = Lk >> 5;
    const bool tail16 = (Lk & 31) != 0;
    const bool early = plan_early(Lk);

    for (int u = 0; u < n_units; ++u) {
      int img, h;
      decode(u, img, h);
      if (quad == 0) TSTAMP(2 + g, 0);
      tcd_wait(&bar_sfull[g], u & 1);
      ptx::tc_fence_after();
      if (quad == 0) TSTAMP(2 + g, 1);
      float l = 0.f;
      uint32_t buf[96];
      if (has_rows) {
        // ---- pass 1: row maximum over the valid key columns [0, L)
        float m0 = -INFINITY, m1 = -INFINITY;
        // ONE register buffer for every phase of the item (3 chunks in the max pass, 2 in the exp2 pass, the 64 O values
        // in the epilogue), so that the phases do not add up in the allocator
        uint32_t (&ra)[32] = reinterpret_cast<uint32_t(&)[32]>(buf[0]);
        uint32_t (&rb)[32] = reinterpret_cast<uint32_t(&)[32]>(buf[32]);
        auto max32 = [&](const uint32_t (&r)[32], int c) {
          if (c * 32 + 32 <= L) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              m0 = fmaxf(m0, __uint_as_float(r[j]));
              m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < L) m0 = fmaxf(m0, __uint_as_float(r[j]));
          }
        };
        // three tcgen05.ld.x32 in flight per wait: the pass is bound by the TMEM load latency, not by its 208 FMNMX
#pragma unroll 1
        for (int c = 0; c < ((AIHAB_ATTN_EXPERIMENT & 16) ? 0 : n32); c += 3) {
          uint32_t (&rc)[32] = reinterpret_cast<uint32_t(&)[32]>(buf[64]);
          ptx::tmem_ld_32x32(t_row + c * 32, ra);
          if (c + 1 < n32) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
          if (c + 2 < n32) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, rc);
          ptx::tmem_ld_wait();
          max32(ra, c);
          if (c + 1 < n32) max32(rb, c + 1);
          if (c + 2 < n32) max32(rc, c + 2);
        }
        if (tail16) {
          uint32_t r[16];
          ptx::tmem_ld_32x16(t_row + n32 * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n32 * 32 + j < L) m0 = fmaxf(m0, __uint_as_float(r[j]));
        }
        const float ms = fmaxf(m0, m1) * sl2;
        // ---- pass 2: P = exp2(S * scale - max) packed over the S columns, row sum in the thread.  The exp2 phase is
        // MUFU-bound: the two streams' warps of this scheduler take turns in it
        if (quad == 0) TSTAMP(2 + g, 2);
        if (!(AIHAB_ATTN_EXPERIMENT & 8)) tcd_wait(&bar_turn[g * 4 + quad], (u & 1) ^ (g == 0 ? 1 : 0));
        if (quad == 0) TSTAMP(2 + g, 3);
        float l0 = 0.f, l1 = 0.f;
        float l2 = 0.f, l3 = 0.f;
        auto exp32 = [&](uint32_t (&r)[32], int c) {
          // all 32 exp2 first, their consumers afterwards: one warp per scheduler owns the MUFU pipe during its turn,
          // so the MUFU latency has to be covered by independent MUFU work of the SAME warp
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = fmaf(__uint_as_float(r[j]), sl2, -ms);
            if (AIHAB_ATTN_EXPERIMENT & 4) r[j] = __float_as_uint(x);
            else r[j] = __float_as_uint(((AIHAB_ATTN_POLYMASK >> (j & 7)) & 1) ? exp2_poly(x) : ptx::ex2_approx(x));
          }
          if (c * 32 + 32 > L) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j >= L) r[j] = 0u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            l0 += __uint_as_float(r[2 * j]);
            l1 += __uint_as_float(r[2 * j + 1]);
            l2 += __uint_as_float(r[2 * j + 2]);
            l3 += __uint_as_float(r[2 * j + 3]);
            pk[j] = ptx::pack2<BF16>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            pk[j + 1] = ptx::pack2<BF16>(__uint_as_float(r[2 * j + 2]), __uint_as_float(r[2 * j + 3]));
          }
          if (!(AIHAB_ATTN_EXPERIMENT & 1)) ptx::tmem_st_32x16(t_row + p_col(early, c), pk);  // 32 keys -> 16 packed columns
          else l0 += __uint_as_float(pk[0] ^ pk[5] ^ pk[10] ^ pk[15]);
        };
        if (!(AIHAB_ATTN_EXPERIMENT & 2)) ptx::tmem_ld_32x32(t_row, ra);
#pragma unroll 1
        for (int c = 0; c < n32; c += 2) {
          if (!(AIHAB_ATTN_EXPERIMENT & 2)) {
            ptx::tmem_ld_wait_regs(ra);
            if (c + 1 < n32) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
          }
          exp32(ra, c);
          if (c + 1 < n32) {
            if (!(AIHAB_ATTN_EXPERIMENT & 2)) {
              ptx::tmem_ld_wait_regs(rb);
              if (c + 2 < n32) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, ra);
            }
            exp32(rb, c + 1);
          }
          if (c + 2 == P_SPLIT && early) {  // first instalment of P is complete: let its PV MMAs start
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar_p1[g]);
          }
        }
        if (tail16) {
          uint32_t r[16];
          ptx::tmem_ld_32x16(t_row + n32 * 32, r);
          ptx::tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x0 = fmaf(__uint_as_float(r[2 * j]), sl2, -ms), x1 = fmaf(__uint_as_float(r[2 * j + 1]), sl2, -ms);
            float p0 = ((AIHAB_ATTN_POLYMASK >> ((2 * j) & 7)) & 1) ? exp2_poly(x0) : ptx::ex2_approx(x0);
            float p1 = ((AIHAB_ATTN_POLYMASK >> ((2 * j + 1) & 7)) & 1) ? exp2_poly(x1) : ptx::ex2_approx(x1);
            if (n32 * 32 + 2 * j >= L) p0 = 0.f;
            if (n32 * 32 + 2 * j + 1 >= L) p1 = 0.f;
            l0 += p0;
            l1 += p1;
            pk[j] = ptx::pack2<BF16>(p0, p1);
          }
          ptx::tmem_st_32x8(t_row + p_col(early, n32), pk);
        }
        if (lane == 0) ptx::mbar_arrive(&bar_turn[(g ^ 1) * 4 + quad]);  // the other stream's turn
        l = (l0 + l1) + (l2 + l3);
        if (quad == 0) TSTAMP(2 + g, 4);
        ptx::tmem_st_wait();
        if (quad == 0) TSTAMP(2 + g, 5);
      } else {  // a quadrant without rows (second tile, rows >= L) only passes the token on
        if (early && lane == 0) ptx::mbar_arrive(&bar_p1[g]);
        if (!(AIHAB_ATTN_EXPERIMENT & 8)) tcd_wait(&bar_turn[g * 4 + quad], (u & 1) ^ (g == 0 ? 1 : 0));
        if (lane == 0) ptx::mbar_arrive(&bar_turn[(g ^ 1) * 4 + quad]);
      }
      ptx::tc_fence_before();  // P stored (wait::st) before the PV MMA may read it
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_p[g]);

      if (quad == 0) TSTAMP(2 + g, 6);
      tcd_wait(&bar_o[g], u & 1);
      ptx::tc_fence_after();
      if (quad == 0) TSTAMP(2 + g, 7);
      uint32_t (&o)[64] = reinterpret_cast<uint32_t(&)[64]>(buf[0]);
      if (has_rows) {
        ptx::tmem_ld_32x32(t_row + o_col(early), reinterpret_cast<uint32_t(&)[32]>(o[0]));
        ptx::tmem_ld_32x32(t_row + o_col(early) + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();  // O read (wait::ld) before the next S MMA may overwrite the buffer
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_bfree[g]);
      if (quad == 0) TSTAMP(2 + g, 8);
      if (has_rows) {
        // x 1/sum -> 16-bit row (128 B) -> swizzled staging tile -> ONE TMA store of the warp's 32-row x 64-column box;
        // rows past the end of the sequence are clipped by the 3-D tensor map (dim 1 = rows of ONE image)
        uint8_t* stg = staging + warp * 4096;
        if (u > 0) {  // the previous store of this warp has finished reading the tile
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
        }
        const float inv_l = 1.0f / l;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 v;
          v.x = ptx::pack2<BF16>(__uint_as_float(o[8 * q]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
          v.y = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
          v.z = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
          v.w = ptx::pack2<BF16>(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = v;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_3d(&tmap_out, stg, h * 64, g * 128 + quad * 32, img);
          ptx::bulk_commit();
        }
      }
      if (quad == 0) TSTAMP(2 + g, 9);
    }
    if (lane == 0) ptx::bulk_wait_all();  // the last stores have completed before the CTA (and its smem) goes away
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

bool attention_tcd_supported(int L) {
  static const bool enabled = [] {
    const char* e = getenv("AIHAB_ATTN_DUAL");
    return !(e != nullptr && e[0] == '0');
  }();
  return enabled && L > 128 && L <= 224;
}

cudaError_t launch_attention_tcd(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const CUtensorMap& tmap_out,
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
    cudaError_t e = is_bf16 ? cudaFuncSetAttribute(attention_tcd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP)
                            : cudaFuncSetAttribute(attention_tcd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP);
    if (e != cudaSuccess) return e;
    attr_set[dev][is_bf16 ? 1 : 0] = true;
  }
  if (is_bf16)
    return launch_kernel(attention_tcd_kernel<true>, grid, D_THREADS, smem_bytes(Lk), stream, 1, true, tmap_q, tmap_kv, tmap_out, L,
                         H, Lk, total, reverse);
  return launch_kernel(attention_tcd_kernel<false>, grid, D_THREADS, smem_bytes(Lk), stream, 1, true, tmap_q, tmap_kv, tmap_out, L,
                       H, Lk, total, reverse);
}

}  // namespace aihab
