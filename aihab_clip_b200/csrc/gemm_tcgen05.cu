// tcgen05 / TMEM / TMA GEMM for sm_100a.  See gemm_tcgen05.cuh for the contract.
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace aihab {

namespace {

constexpr int BM = 128;      // UMMA M (one CTA, cta_group::1): accumulator row i <-> TMEM lane i
constexpr int BK = 64;       // 64 x 16-bit = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;   // fixed for 16-bit operands
constexpr int EPI_WARP0 = 4;  // warps 4.. are the epilogue (warp % 4 selects the TMEM lane quadrant)
// Two warps per TMEM lane quadrant (8 epilogue warps, 384 threads) so the epilogue keeps up with the MMA.  The fp32
// residual epilogue streams 4 KB TMA boxes through per-warp rings, one warp per quadrant (two per quadrant were
// measured no faster in round 1: in the step these GEMMs wait for DRAM, not for the epilogue warps).
// EPI_RES_WIDE (internal, CTA pairs only): EPI_BIAS_RES_32 with TWO warps per quadrant (8 rings of 3 boxes) and 3 operand
// stages, for the residual GEMM with a SHORT main loop (out_proj, K = D): there the tile time is the latency chain of
// the residual epilogue (one warp per scheduler issues 25 % of the time), so a second chain per scheduler pays
// (-10 % on out_proj in the step), while c_proj (K = 4 D) is main-loop bound and keeps 4 stages + 4 deep rings.
constexpr int EPI_RES_WIDE = 100;
// EPI_TOPK5 (internal): EPI_TOPK_32 keeping 5 instead of TOPK_SLOTS candidates per (row, warp) - enough for k <= 5 and
// 40 % fewer instructions in the branch-free insertion chain (the epilogue of the K = 3E logits GEMM must stay hidden)
constexpr int EPI_TOPK5 = 101;
__host__ __device__ constexpr bool is_res(int epi) { return epi == EPI_BIAS_RES_32 || epi == EPI_RES_WIDE; }
__host__ __device__ constexpr int epi_warps(int epi, bool two) {
  (void)two;
  return epi == EPI_BIAS_RES_32 ? 4 : 8;
}
__host__ __device__ constexpr int num_threads(int epi, bool two) { return (EPI_WARP0 + epi_warps(epi, two)) * 32; }

constexpr int RES_BOX = 32 * 128;  // 32 rows x 32 fp32 = 4 KB TMA box, SWIZZLE_128B
constexpr int STAT_COLS = 128;     // LayerNorm producer: one (mean, sum of squared deviations) partial per row and 128 columns
constexpr int kMaxStatBlocks = 8;  // width <= 1024 (api.cu allocates the statistics buffer with the same bound)

template <int BN, int EPI, bool TWO>
struct SmemLayout {
  static constexpr bool kRes = is_res(EPI);
  static constexpr bool kWide = (EPI == EPI_RES_WIDE);
  // the residual epilogue trades operand stages for a 64 KB TMA ring: those GEMMs are bound by the fp32 residual
  // read-modify-write, not by the MMA pipe.  In a CTA pair each CTA stages only half of the W tile.
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int kResWarps = kWide ? 8 : 4;  // residual epilogue warps (rings)
// CTA-pair residual GEMM with a long main loop (c_proj, K = 4 D): 5 operand stages + 3-box rings.  Measured against
// 4 + 4 on the same boxes: c_proj 0.228 -> 0.215 and 0.231 -> 0.213 ms in the step (its epilogue is hidden behind the
// 48-k-block main loop, which wants the deeper operand pipeline), step +3.2 % / +0.1 %.
#ifndef AIHAB_RES_RING
#define AIHAB_RES_RING 3
#endif
#ifndef AIHAB_RES_STAGES
#define AIHAB_RES_STAGES 5
#endif
#ifndef AIHAB_WIDE_RING
#define AIHAB_WIDE_RING 3
#endif
#ifndef AIHAB_WIDE_STAGES
#define AIHAB_WIDE_STAGES 3
#endif
  static constexpr int kRS = kWide ? AIHAB_WIDE_RING : (TWO ? AIHAB_RES_RING : 4);  // ring slots per residual epilogue warp (kRS - 1 boxes in flight)
  static constexpr int kStages = kWide ? AIHAB_WIDE_STAGES :
      TWO ? (kRes ? AIHAB_RES_STAGES : 5) : ((BN == 256) ? 3 : (kRes ? (AIHAB_RES_STAGES < 4 ? AIHAB_RES_STAGES : 4) : 5));
  static constexpr int kStageBytes = kABytes + kBBytes;
  // residual rings (kResWarps x kRS boxes) + 2 KB per warp for the coalesced gamma*x store | 8 warps x 4 KB staging tile
  static constexpr int kStagingBytes = kRes ? kResWarps * (kRS * RES_BOX + 2048) : 8 * 4096;
  static constexpr int kBiasBytes = 4 * BN * 4;  // [2][BN] bias + [2][BN] auxiliary per-column vector (s_n / gamma)
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kStages * kABytes;
  static constexpr int kOffStaging = kStages * kStageBytes;
  static constexpr int kOffBias = kOffStaging + kStagingBytes;
  static constexpr int kOffBars = kOffBias + kBiasBytes;
  static constexpr int kNumBars = 2 * kStages + 4 + 8 * 4;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
  static_assert(kDynamic <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)  (clip/model.py:160-162)
#ifdef AIHAB_GELU_EXACT
  // ex2 + rcp (<= 2 ulp each): two MUFU operations per element
  return __fdividef(x, 1.0f + __expf(-1.702f * x));
#else
  // sigmoid(z) = 0.5 + 0.5 tanh(z / 2) with tanh.approx.f32 (max relative error 2^-11, i.e. at the precision of the
  // 16-bit value this feeds): ONE MUFU operation per element - the c_fc epilogue is MUFU-bound (32768 elements / tile).
  // CPU emulation of the whole tower with a 2^-11 tanh error moves max |dlogit| from 3.0e-3 to <= 3.8e-3.
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#endif
}

// TWO = CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, UMMA M = 256; each CTA loads its own
// 128 A rows and HALF of the W tile (the pair's MMA reads both halves), owns the accumulator rows of its A rows and
// runs its own epilogue.  Halves the W shared-memory fill and L2 -> SM traffic per FLOP.
template <int BN, int EPI_, bool TWO>
__global__ void __launch_bounds__(num_threads(EPI_, TWO), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
            const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using L = SmemLayout<BN, EPI_, TWO>;
  constexpr int EPI = (EPI_ == EPI_RES_WIDE) ? static_cast<int>(EPI_BIAS_RES_32)   // same arithmetic, wider warp layout
                      : (EPI_ == EPI_TOPK5) ? static_cast<int>(EPI_TOPK_32) : EPI_;
  constexpr int kStages = L::kStages;
  constexpr bool kLn = (EPI == EPI_LN_BIAS_16 || EPI == EPI_LN_BIAS_GELU_16);
  constexpr bool kGelu = (EPI == EPI_BIAS_GELU_16 || EPI == EPI_LN_BIAS_GELU_16);
  constexpr bool kSplit3 = (EPI == EPI_SPLIT3_16);
  constexpr bool kOut16 = (EPI == EPI_BIAS_16 || EPI == EPI_BIAS_GELU_16 || kLn || kSplit3);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024 B aligned, still a __shared__ pointer
  uint8_t* sA = smem + L::kOffA;
  uint8_t* sB = smem + L::kOffB;
  uint8_t* sStaging = smem + L::kOffStaging;
  float* sBias = reinterpret_cast<float*>(smem + L::kOffBias);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBars);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_full_bar = tmem_empty_bar + 2;  // [residual epilogue warps][RS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  // work units: CTAs (one 128-row tile each) or CTA pairs (two 128-row tiles m = 2 * mp + rank sharing the W tile)
  const int cta_rank = TWO ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int unit = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_units = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_tiles = (TWO ? (m_blocks + 1) / 2 : m_blocks) * n_blocks;
  auto tile_m = [&](int tile) {  // M block of this CTA for a tile (may be == m_blocks: all rows out of range)
    int mb = tile / n_blocks;
    if (p.reverse_m) mb = (TWO ? (m_blocks + 1) / 2 : m_blocks) - 1 - mb;
    return TWO ? 2 * mb + cta_rank : mb;
  };

  ptx::griddep_launch();  // the next kernel of the stream may be scheduled as soon as SM resources free up
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], epi_warps(EPI_, TWO) * (TWO ? 2 : 1));  // one arrive per epilogue warp (of the pair)
    }
    for (int i = 0; i < L::kResWarps * L::kRS; ++i) ptx::mbar_init(&res_full_bar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (TWO) {
      ptx::tmem_alloc_pair(tmem_slot, 2 * BN);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, 2 * BN);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (TWO) {
    __syncthreads();
    ptx::cluster_sync();  // the peer's barriers are initialised before anything signals them
  } else {
    __syncthreads();
  }
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::griddep_wait();  // everything above overlaps the tail of the previous kernel; global memory only from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol_w = ptx::policy_evict_last();  // weights are re-read by every M block
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_a = ptx::policy_evict_normal();
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        const int m_blk = tile_m(tile);
        const int n_blk = tile % n_blocks;
        int a_row = m_blk * BM;
        if constexpr (TWO) {
          if (p.ring_mode == 2) {  // consumer of a pipelined pair: the producer kernel has finished this pair-row
            ptx::wait_counter(p.ctr_done + tile / n_blocks, p.need_done);
            ptx::fence_proxy_async_all();  // its TMA stores (async proxy, other SMs) before our TMA loads
            a_row %= p.ring_rows;
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (TWO) {
            // both CTAs' loads complete on the leader's barrier: it expects the bytes of the whole pair
            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            ptx::tma_load_2d_pair(sA + stage * L::kABytes, &tmap_a, &full_bar[stage], kb * BK, a_row, pol_a);
            ptx::tma_load_2d_pair(sB + stage * L::kBBytes, &tmap_w, &full_bar[stage], kb * BK,
                                  n_blk * BN + cta_rank * (BN / 2), pol_w);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          ptx::tma_load_2d(sA + stage * L::kABytes, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          ptx::tma_load_2d_hint(sB + stage * L::kBBytes, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN, pol_w);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop (converged; one elected lane issues each tcgen05 instruction, ptx.cuh "_w"
    // variants) so that the descriptors live in uniform registers and the four MMAs of a k-block issue back to back.
    if (cta_rank == 0) {  // in a pair only the leader issues
      const uint32_t idesc = ptx::make_idesc_f16(p.ab_format, TWO ? 2 * BM : BM, BN);
      const uint32_t sA_u32 = ptx::smem_u32(sA), sB_u32 = ptx::smem_u32(sB);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_kmajor_sw128_desc(sA_u32 + stage * L::kABytes);
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sB_u32 + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 B (UMMA_K x 2 B) inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if constexpr (TWO)
              ptx::umma_f16_pair_w(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else
              ptx::umma_f16_w(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (TWO) {
            ptx::umma_commit_pair_w(&empty_bar[stage]);  // frees this smem slot in both CTAs
            if (kb == num_kb - 1) ptx::umma_commit_pair_w(&tmem_full_bar[as]);
          } else {
            ptx::umma_commit_w(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            if (kb == num_kb - 1) ptx::umma_commit_w(&tmem_full_bar[as]);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------ epilogue
    constexpr int kEpiThreads = epi_warps(EPI_, TWO) * 32;
    const int ew = warp & 3;                   // TMEM lanes [32*ew, 32*ew+32) (hardware: warp % 4)
    const int ehalf = (warp - EPI_WARP0) >> 2;  // 0, or 0/1 when two warps share a quadrant
    uint8_t* stg = sStaging + (warp - EPI_WARP0) * 4096;  // 16-bit / plain fp32 epilogues (not the residual rings)
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..kEpiThreads-1
    const bool bf16 = p.ab_format != 0;
    // EPI_BIAS_RES_32: every warp streams its 32-row slice of the fp32 residual through a private ring of 4 KB
    // TMA boxes (load -> add in place -> TMA store), prefetching RS-1 boxes ahead across tile boundaries.
    constexpr int RS = L::kRS;
    constexpr int CPT = BN / 32 / (L::kResWarps / 4);  // residual boxes per tile per warp: a run of CPT * 32 columns
    const int er = warp - EPI_WARP0;                   // ring of this warp (= ew + 4 * ehalf)
    const int c_off = ehalf * CPT;                     // first 32-column chunk of the tile owned by this warp
    uint64_t* my_full = res_full_bar + er * RS;
    uint8_t* my_ring = sStaging + er * (RS * RES_BOX);
    auto res_prefetch = [&](int q) {  // lane 0 only: issue the TMA load of this warp's q-th box, if it exists
      const int itq = q / CPT, cq = c_off + (q - itq * CPT);
      const long tq = static_cast<long>(unit) + static_cast<long>(itq) * num_units;
      if (tq >= num_tiles) return;
      const int mb = tile_m(static_cast<int>(tq));
      const int nb = static_cast<int>(tq % n_blocks);
      const int slot = q % RS;
      ptx::mbar_expect_tx(&my_full[slot], RES_BOX);
      ptx::tma_load_2d(my_ring + slot * RES_BOX, &tmap_c, &my_full[slot], nb * BN + cq * 32, mb * BM + ew * 32);
    };
    int q = 0;  // running box counter of this warp
    if constexpr (EPI == EPI_BIAS_RES_32) {
      if (lane == 0) {
        for (int i = 0; i < RS - 1; ++i) res_prefetch(i);
      }
    }
    // per-column vectors of a tile (bias; s_n or gamma) are fetched ONE TILE AHEAD into registers and parked in smem
    // at the start of the tile, so their global-load latency never sits in front of an epilogue
    constexpr int NV = (BN + kEpiThreads - 1) / kEpiThreads;
    const bool ln_prod = (EPI == EPI_BIAS_RES_32) && p.ln_gamma != nullptr;
    const float* aux = kLn ? p.ln_s : p.ln_gamma;
    float nxt_b[NV], nxt_x[NV];
    auto fetch_cols = [&](int tile) {
      const int nbase = (tile % n_blocks) * BN;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int i = et + v * kEpiThreads;
        const int n = nbase + i;
        const bool ok = tile < num_tiles && i < BN && n < p.N;
        nxt_b[v] = (ok && p.bias != nullptr) ? __ldg(p.bias + n) : 0.0f;
        nxt_x[v] = (ok && (kLn || ln_prod)) ? __ldg(aux + n) : 0.0f;
      }
    };
    if constexpr (EPI != EPI_PATCH_32) fetch_cols(unit);
    int it = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++it) {
      const int m_blk = tile_m(tile);
      const int n_blk = tile % n_blocks;
      const int m0 = m_blk * BM + ew * 32;
      const int n0 = n_blk * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      float* sb = sBias + as * BN;
      float* sx = sBias + (2 + as) * BN;  // s_n (LayerNorm consumer) or gamma (LayerNorm producer)
      if constexpr (EPI != EPI_PATCH_32) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int i = et + v * kEpiThreads;
          if (i < BN) {
            sb[i] = nxt_b[v];
            if (kLn || ln_prod) sx[i] = nxt_x[v];
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        fetch_cols(tile + num_units);  // in flight during this tile's epilogue
      }
      // LayerNorm consumer: statistics of this thread's row from the producer's per-tile partial sums
      float ln_r = 1.0f, ln_nrm = 0.0f;
      if constexpr (kLn) {
        const int grow = m0 + lane;
        if (grow < p.M) {
          const float* st = p.ln_stats + static_cast<size_t>(grow) * p.ln_nsb * 2;
          // per 128-column block the producer left (block mean, sum of squared deviations from it), both computed
          // around a pivot inside the block; merged here with the parallel-variance formula (Chan et al.): no
          // E[x^2] - mu^2 cancellation, whatever common offset or outlier channels the residual stream carries
          float2 sv[kMaxStatBlocks];
          float msum = 0.f;
#pragma unroll
          for (int b = 0; b < kMaxStatBlocks; ++b) {
            if (b < p.ln_nsb) {
              sv[b] = __ldg(reinterpret_cast<const float2*>(st) + b);
              msum += sv[b].x;
            }
          }
          const float mu = msum / static_cast<float>(p.ln_nsb);
          float m2 = 0.f;
#pragma unroll
          for (int b = 0; b < kMaxStatBlocks; ++b) {
            if (b < p.ln_nsb) {
              const float d = sv[b].x - mu;
              m2 += fmaf(static_cast<float>(STAT_COLS) * d, d, sv[b].y);
            }
          }
          const float var = fmaxf(m2 / static_cast<float>(p.K), 0.0f);
          ln_r = rsqrtf(var + 1e-5f);
          ln_nrm = -ln_r * mu;
        }
      }

      // EPI_SCALE_32 / EPI_TOPK_32 behind an EPI_SPLIT3_16 producer: F.normalize of the A rows applied to the accumulator
      // row instead (scale / max(||row||, 1e-12)); the chunk sums are added in ascending order
      [[maybe_unused]] float sc = p.scale;
      if constexpr (EPI == EPI_SCALE_32 || EPI == EPI_TOPK_32) {
        if (p.row_ss != nullptr) {
          const int grow = m0 + lane;
          float ssum = 0.f;
          if (grow < p.M)
            for (int b = 0; b < p.row_ss_n; ++b) ssum += __ldg(p.row_ss + static_cast<size_t>(grow) * p.row_ss_n + b);
          sc = p.scale / fmaxf(sqrtf(ssum), 1e-12f);
        }
      }
      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tc_fence_after();
      if constexpr (TWO && EPI == EPI_BIAS_RES_32) {
        // consumer of a pipelined pair: every MMA of the tile has retired, so this CTA's A rows of the ring are consumed
        if (p.ring_mode == 2 && warp == EPI_WARP0 && lane == 0) ptx::red_release_gpu_add(p.ctr_consumed + tile / n_blocks, 1u);
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      if (kOut16 && p.debug == 77) {  // DEBUG: drain without an epilogue (mainloop ceiling measurement)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (TWO) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
          else ptx::mbar_arrive(&tmem_empty_bar[as]);
        }
        continue;
      }

      int st_row = m0;
      if constexpr (kOut16 && TWO) {
        if (p.ring_mode == 1) {  // producer of a pipelined pair: the ring slot of this pair-row must have been consumed
          const int pr = tile / n_blocks, ring_pairs = p.ring_rows / (2 * BM);
          if (pr >= ring_pairs && lane == 0) ptx::wait_counter(p.ctr_consumed + pr - ring_pairs, p.need_consumed);
          __syncwarp();
          st_row = m0 % p.ring_rows;
        }
      }
      if constexpr (kOut16) {
        // 64-column chunks, dealt alternately to the two warps of a lane quadrant.  A chunk leaves as ONE TMA store of
        // a 32-row x 128-byte box: the staging tile is written in the SWIZZLE_128B pattern (16-byte units XOR row & 7,
        // conflict-free), the store writes complete 128-byte lines and clips the M / N tails itself.
#pragma unroll 1
        for (int c = ehalf; c < BN / 64; c += 2) {
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_32x32(taddr + c * 64, ra);
          ptx::tmem_ld_32x32(taddr + c * 64 + 32, rb);
          ptx::tmem_ld_wait();
          if (c + 2 >= BN / 64) {  // last chunk of this warp: the accumulator stage can go back to the MMA issuer now
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (TWO) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
              else ptx::mbar_arrive(&tmem_empty_bar[as]);
            }
          }
          const float* cb = sb + c * 64;
          const float* cx = sx + c * 64;
          if constexpr (kSplit3) {
            // hi / lo split of the chunk, its sum of squares, three stores (hi | hi | lo) through the one staging tile
            uint32_t pk[32];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v0 = __uint_as_float(j < 16 ? ra[2 * j] : rb[2 * j - 32]) + cb[2 * j];
              const float v1 = __uint_as_float(j < 16 ? ra[2 * j + 1] : rb[2 * j - 31]) + cb[2 * j + 1];
              ss = fmaf(v0, v0, ss);
              ss = fmaf(v1, v1, ss);
              pk[j] = ptx::pack2<false>(v0, v1);
              // keep the residuals in the accumulator registers for the second pass
              const __half2 h2 = *reinterpret_cast<const __half2*>(&pk[j]);
              const float l0 = v0 - __low2float(h2), l1 = v1 - __high2float(h2);
              if (j < 16) {
                ra[2 * j] = __float_as_uint(l0);
                ra[2 * j + 1] = __float_as_uint(l1);
              } else {
                rb[2 * j - 32] = __float_as_uint(l0);
                rb[2 * j - 31] = __float_as_uint(l1);
              }
            }
            const int grow = m0 + lane;
            if (grow < p.M && n0 + c * 64 < p.N) p.stats_out[static_cast<size_t>(grow) * ((p.N + 63) / 64) + (n0 / 64 + c)] = ss;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
              if (pass == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  pk[j] = ptx::pack2<false>(__uint_as_float(j < 16 ? ra[2 * j] : rb[2 * j - 32]),
                                            __uint_as_float(j < 16 ? ra[2 * j + 1] : rb[2 * j - 31]));
              }
              if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store of this warp has finished reading the tile
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                    make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              }
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                if (pass == 0) {
                  ptx::tma_store_2d(&tmap_c, stg, n0 + c * 64, st_row);
                  ptx::tma_store_2d(&tmap_c, stg, p.N + n0 + c * 64, st_row);
                } else {
                  ptx::tma_store_2d(&tmap_c, stg, 2 * p.N + n0 + c * 64, st_row);
                }
                ptx::bulk_commit();
              }
            }
            continue;
          }
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v0 = __uint_as_float(j < 16 ? ra[2 * j] : rb[2 * j - 32]);
            const float v1 = __uint_as_float(j < 16 ? ra[2 * j + 1] : rb[2 * j - 31]);
            float a0, a1;
            if constexpr (kLn) {
              a0 = fmaf(ln_r, v0, fmaf(ln_nrm, cx[2 * j], cb[2 * j]));
              a1 = fmaf(ln_r, v1, fmaf(ln_nrm, cx[2 * j + 1], cb[2 * j + 1]));
            } else {
              a0 = v0 + cb[2 * j];
              a1 = v1 + cb[2 * j + 1];
            }
            if constexpr (kGelu) {
              a0 = quick_gelu(a0);
              a1 = quick_gelu(a1);
            }
            pk[j] = bf16 ? ptx::pack2<true>(a0, a1) : ptx::pack2<false>(a0, a1);
          }
          if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store of this warp has finished reading the tile
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
          ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_c, stg, n0 + c * 64, st_row);
            ptx::bulk_commit();
          }
        }
        if constexpr (TWO) {
          if (p.ring_mode == 1 && lane == 0) {  // this warp's share of the tile is in global memory: count it
            ptx::bulk_wait_all();
            ptx::fence_proxy_async_all();
            ptx::red_release_gpu_add(p.ctr_done + tile / n_blocks, 1u);
          }
        }
      } else if constexpr (EPI == EPI_BIAS_RES_32) {
        float ln_s1 = 0.f, ln_s2 = 0.f, ln_piv = 0.f;  // LayerNorm producer: this row's pivoted sum / sum of squares over a 128-column block
        const int prow = m0 + lane;
#pragma unroll 1
        for (int cl = 0; cl < CPT; ++cl, ++q) {
          const int c = c_off + cl;
          const int slot = q % RS;
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr + c * 32, r);
          ptx::mbar_wait(&my_full[slot], (q / RS) & 1);
          ptx::tmem_ld_wait();
          uint8_t* box = my_ring + slot * RES_BOX;
          // all loads of the box first, then the arithmetic, then all stores: the shared-memory latency is paid once
          float4 xv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) xv[u] = *reinterpret_cast<const float4*>(box + lane * 128 + ((u ^ (lane & 7)) << 4));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 32 + 4 * u);
            float4 v = xv[u];
            v.x += __uint_as_float(r[4 * u]) + b4.x;
            v.y += __uint_as_float(r[4 * u + 1]) + b4.y;
            v.z += __uint_as_float(r[4 * u + 2]) + b4.z;
            v.w += __uint_as_float(r[4 * u + 3]) + b4.w;
            xv[u] = v;
            if (ln_prod) {  // reuse r[] for the packed gamma * x_new row segment
              // statistics around a pivot (the block's first element): sums of small deviations, no cancellation
              if (u == 0 && (c & (STAT_COLS / 32 - 1)) == 0) ln_piv = v.x;
              const float dx = v.x - ln_piv, dy = v.y - ln_piv, dz = v.z - ln_piv, dw = v.w - ln_piv;
              ln_s1 += (dx + dy) + (dz + dw);
              ln_s2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, ln_s2))));
              const float4 g4 = *reinterpret_cast<const float4*>(sx + c * 32 + 4 * u);
              r[2 * u] = bf16 ? ptx::pack2_sat<true>(g4.x * v.x, g4.y * v.y) : ptx::pack2_sat<false>(g4.x * v.x, g4.y * v.y);
              r[2 * u + 1] = bf16 ? ptx::pack2_sat<true>(g4.z * v.z, g4.w * v.w) : ptx::pack2_sat<false>(g4.z * v.z, g4.w * v.w);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(box + lane * 128 + ((u ^ (lane & 7)) << 4)) = xv[u];
          if (ln_prod) {  // gamma * x_new: transpose through smem so each store covers 8 complete 64 B row segments
            uint8_t* ast = sStaging + L::kResWarps * (RS * RES_BOX) + er * 2048;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              *reinterpret_cast<uint4*>(ast + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
            }
            __syncwarp();
            const int u = lane & 3;
            const int gcol = n0 + c * 32 + u * 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + (lane >> 2);
              const uint4 v = *reinterpret_cast<const uint4*>(ast + row * 64 + ((u ^ ((row >> 1) & 3)) << 4));
              if (m0 + row < p.M && gcol < p.N)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.a16_out) +
                                          static_cast<size_t>(m0 + row) * p.N + gcol) = v;
            }
          }
          ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_c, box, n0 + c * 32, m0);
            ptx::bulk_commit();
            ptx::bulk_wait_read<1>();   // the store issued one box ago has finished reading its slot ...
            res_prefetch(q + RS - 1);   // ... which is exactly the slot box q+RS-1 lands in
          }
          __syncwarp();
          // one partial per row and STAT_COLS columns, whatever the tile width or the number of epilogue warps: the
          // consumer adds the N / STAT_COLS partials of a row in order, so the statistics (and with them every
          // per-image result) do not depend on the launch configuration the batch size selects
          if (ln_prod && (c & (STAT_COLS / 32 - 1)) == STAT_COLS / 32 - 1) {
            const int sblk = (n0 + c * 32) / STAT_COLS;
            if (prow < p.M && sblk * STAT_COLS < p.N)
              *reinterpret_cast<float2*>(p.stats_out + (static_cast<size_t>(prow) * ((p.N + STAT_COLS - 1) / STAT_COLS) + sblk) * 2) =
                  make_float2(fmaf(ln_s1, 1.0f / STAT_COLS, ln_piv), fmaf(-ln_s1 * (1.0f / STAT_COLS), ln_s1, ln_s2));
            ln_s1 = ln_s2 = 0.f;
          }
        }
      } else {
        // EPI_PATCH_32: the 8 rows this lane stores are the same for every chunk of the tile - patch row -> token row
        // (one class-token row per image is skipped) and its positional-embedding row are resolved once per tile
        int tok_row[8], pos_row[8];
        if constexpr (EPI == EPI_PATCH_32) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int grow = m0 + i * 4 + (lane >> 3);
            const int img = grow / p.g2;
            pos_row[i] = 1 + grow - img * p.g2;
            tok_row[i] = img * (p.g2 + 1) + pos_row[i];
          }
        }
        if constexpr (EPI == EPI_TOPK_32) {
          // running top-S of this thread's accumulator row over the warp's chunks of the tile; columns arrive in
          // ascending order, so a strict '>' keeps the lower column ahead among equal values (torch.topk order).
          // BRANCH-FREE insertion (compare / select chain, ~5 instructions per slot and element): lanes hold different
          // rows, so a data-dependent insertion branch is taken by some lane for almost every element and the warp
          // pays the slow path every time (measured: 148 -> see DESIGN 4.4 per 32 k-row pass).
          constexpr int S = (EPI_ == EPI_TOPK5) ? 5 : TOPK_SLOTS;
          float tv[S];
          int ti[S];
#pragma unroll
          for (int i = 0; i < S; ++i) {
            tv[i] = -INFINITY;
            ti[i] = 0x7fffffff;
          }
#pragma unroll 1
          for (int c = ehalf; c < BN / 32; c += 2) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(taddr + c * 32, r);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const float v = col < p.N ? fmaf(__uint_as_float(r[j]), sc, sb[c * 32 + j]) : -INFINITY;
#pragma unroll
              for (int i = S - 1; i >= 1; --i) {
                const bool ci = v > tv[i], cm = v > tv[i - 1];
                ti[i] = cm ? ti[i - 1] : (ci ? col : ti[i]);
                tv[i] = cm ? tv[i - 1] : (ci ? v : tv[i]);
              }
              const bool c0 = v > tv[0];
              ti[0] = c0 ? col : ti[0];
              tv[0] = c0 ? v : tv[0];
            }
          }
          const int grow = m0 + lane;
          if (grow < p.M) {
            const size_t base = (static_cast<size_t>(grow) * (2 * n_blocks) + 2 * (n0 / BN) + ehalf) * TOPK_SLOTS;
            float ov[TOPK_SLOTS];
            int oi[TOPK_SLOTS];
#pragma unroll
            for (int i = 0; i < TOPK_SLOTS; ++i) {
              ov[i] = i < S ? tv[i < S ? i : 0] : -INFINITY;
              oi[i] = i < S ? ti[i < S ? i : 0] : 0x7fffffff;
            }
#pragma unroll
            for (int i = 0; i < TOPK_SLOTS; i += 4) {
              *reinterpret_cast<float4*>(p.cand_val + base + i) = make_float4(ov[i], ov[i + 1], ov[i + 2], ov[i + 3]);
              *reinterpret_cast<int4*>(p.cand_idx + base + i) = make_int4(oi[i], oi[i + 1], oi[i + 2], oi[i + 3]);
            }
          }
        } else {
#pragma unroll 1
        for (int c = ehalf; c < BN / 32; c += 2) {  // 32-column chunks dealt alternately to the quadrant's two warps
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 v;
            v.x = __uint_as_float(r[4 * u]);
            v.y = __uint_as_float(r[4 * u + 1]);
            v.z = __uint_as_float(r[4 * u + 2]);
            v.w = __uint_as_float(r[4 * u + 3]);
            if constexpr (EPI == EPI_BIAS_RES_32) {
              v.x += sb[c * 32 + 4 * u];
              v.y += sb[c * 32 + 4 * u + 1];
              v.z += sb[c * 32 + 4 * u + 2];
              v.w += sb[c * 32 + 4 * u + 3];
            } else if constexpr (EPI == EPI_SCALE_32) {
              v.x = fmaf(v.x, sc, sb[c * 32 + 4 * u]);
              v.y = fmaf(v.y, sc, sb[c * 32 + 4 * u + 1]);
              v.z = fmaf(v.z, sc, sb[c * 32 + 4 * u + 2]);
              v.w = fmaf(v.w, sc, sb[c * 32 + 4 * u + 3]);
            }
            *reinterpret_cast<float4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) = v;
          }
          __syncwarp();
          const int u = lane & 7;
          const int gcol = n0 + c * 32 + u * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3);
            float4 v = *reinterpret_cast<const float4*>(stg + row * 128 + ((u ^ (row & 7)) << 4));
            const int grow = m0 + row;
            if (grow < p.M && gcol < p.N) {
              if constexpr (EPI == EPI_BIAS_RES_32) {
                float4* dst = reinterpret_cast<float4*>(p.out32 + static_cast<size_t>(grow) * p.ldo + gcol);
                const float4 x = *dst;
                v.x += x.x;
                v.y += x.y;
                v.z += x.z;
                v.w += x.w;
                *dst = v;
              } else if constexpr (EPI == EPI_PATCH_32) {
                const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(pos_row[i]) * p.N + gcol));
                v.x += pe.x;
                v.y += pe.y;
                v.z += pe.z;
                v.w += pe.w;
                *reinterpret_cast<float4*>(p.out32 + static_cast<size_t>(tok_row[i]) * p.ldo + gcol) = v;
              } else {
                *reinterpret_cast<float4*>(p.out32 + static_cast<size_t>(grow) * p.ldo + gcol) = v;
              }
            }
          }
          __syncwarp();
        }
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer (the 16-bit epilogues did so after their last load)
      if constexpr (!kOut16) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (TWO) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);  // the leader's MMA issuer waits for both CTAs
          else ptx::mbar_arrive(&tmem_empty_bar[as]);
        }
      }
    }
    if constexpr (EPI == EPI_BIAS_RES_32 || kOut16) {
      if (lane == 0) ptx::bulk_wait_all();  // smem must outlive the last TMA stores
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (TWO) ptx::cluster_sync();  // neither CTA may exit (or free TMEM) while the pair's MMAs / signals are in flight
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (TWO) ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
    else ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused MLP: c_fc (LayerNorm fold + bias + QuickGELU, 16-bit hidden) and c_proj (+ fp32 residual, LayerNorm producer) as
// ONE persistent kernel of CTA pairs on the whole machine.  The host hands it a tile list that interleaves the two
// GEMMs' tiles so that a c_proj tile comes a few rounds after the c_fc tiles it reads: the [M, 4D] hidden activations
// live in a ring of ring_pairs 256-row pair-rows that stays in L2 and never make the HBM round trip.  Per-pair-row
// progress counters in global memory (release / acquire) order producers and consumers across CTAs.
// Roles as in gemm_kernel; all 8 epilogue warps work on c_fc tiles, warps 4..7 run the residual rings of c_proj tiles.
#ifndef AIHAB_FUSED_STAGES
#define AIHAB_FUSED_STAGES 4
#endif
#ifndef AIHAB_FUSED_RING
#define AIHAB_FUSED_RING 3
#endif
struct FusedSmem {
  static constexpr int kStages = AIHAB_FUSED_STAGES;
  static constexpr int kRS = AIHAB_FUSED_RING;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BM * BK * 2;  // half of the 256-wide W tile per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kStages * kABytes;
  static constexpr int kOffStaging = kStages * kStageBytes;  // 8 warps x 4 KB (c_fc); first 2 KB of warps 4..7: gamma * x transposition (c_proj)
  static constexpr int kOffRing = kOffStaging + 8 * 4096;
  static constexpr int kOffBias = kOffRing + 4 * kRS * RES_BOX;
  static constexpr int kOffBars = kOffBias + 4 * 256 * 4;
  static constexpr int kNumBars = 2 * kStages + 4 + 4 * kRS;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kDynamic <= 227 * 1024, "shared memory budget");
};

constexpr uint32_t MLP_TILE_NONE = 0xffffffffu;

__global__ void __launch_bounds__(384, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_wfc,
                 const __grid_constant__ CUtensorMap tmap_hld, const __grid_constant__ CUtensorMap tmap_wproj,
                 const __grid_constant__ CUtensorMap tmap_hst, const __grid_constant__ CUtensorMap tmap_x,
                 const MlpFusedParams p) {
  using L = FusedSmem;
  constexpr int kStages = L::kStages;
  constexpr int BN = 256;
  constexpr int RS = L::kRS;
  constexpr int CPT = BN / 32;  // residual boxes per c_proj tile and warp

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + L::kOffA;
  uint8_t* sB = smem + L::kOffB;
  uint8_t* sStaging = smem + L::kOffStaging;
  uint8_t* sRing = smem + L::kOffRing;
  float* sBias = reinterpret_cast<float*>(smem + L::kOffBias);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBars);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_full_bar = tmem_empty_bar + 2;  // [4 warps][RS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(ptx::cluster_ctarank());
  const int unit = static_cast<int>(blockIdx.x >> 1);
  const int num_units = static_cast<int>(gridDim.x >> 1);
  const int num_tiles = p.num_tiles;
  const int kb_fc = p.D / BK, kb_proj = 4 * p.D / BK;

  ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_y);
    ptx::prefetch_tmap(&tmap_wfc);
    ptx::prefetch_tmap(&tmap_hld);
    ptx::prefetch_tmap(&tmap_wproj);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 8 * 2);  // one arrive per epilogue warp of the pair, whatever the tile type
    }
    for (int i = 0; i < 4 * RS; ++i) ptx::mbar_init(&res_full_bar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(tmem_slot, 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol_w = ptx::policy_evict_last();
      const uint64_t pol_a = ptx::policy_evict_normal();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = unit; i < num_tiles; i += num_units) {
        const uint32_t d = __ldg(p.tiles + i);
        if (d == MLP_TILE_NONE) continue;
        const bool proj = (d >> 31) != 0;
        const int pr = static_cast<int>((d >> 8) & 0x7fffffu), nb = static_cast<int>(d & 0xffu);
        const CUtensorMap* ma = proj ? &tmap_hld : &tmap_y;
        const CUtensorMap* mw = proj ? &tmap_wproj : &tmap_wfc;
        int a_row = pr * (2 * BM) + cta_rank * BM;
        const int num_kb = proj ? kb_proj : kb_fc;
        if (proj) {  // every c_fc epilogue warp of this pair-row has finished its TMA stores
          ptx::wait_counter(p.ctr_done + pr, p.need_done);
          ptx::fence_proxy_async_all();
          a_row = (pr % p.ring_pairs) * (2 * BM) + cta_rank * BM;
        }
        const int w_row = nb * BN + cta_rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          ptx::tma_load_2d_pair(sA + stage * L::kABytes, ma, &full_bar[stage], kb * BK, a_row, pol_a);
          ptx::tma_load_2d_pair(sB + stage * L::kBBytes, mw, &full_bar[stage], kb * BK, w_row, pol_w);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, leader CTA only)
    if (cta_rank == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(p.ab_format, 2 * BM, BN);
      const uint32_t sA_u32 = ptx::smem_u32(sA), sB_u32 = ptx::smem_u32(sB);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int i = unit; i < num_tiles; i += num_units) {
        const uint32_t d = __ldg(p.tiles + i);
        if (d == MLP_TILE_NONE) continue;
        const int num_kb = (d >> 31) ? kb_proj : kb_fc;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ++it;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_kmajor_sw128_desc(sA_u32 + stage * L::kABytes);
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sB_u32 + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_f16_pair_w(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit_pair_w(&empty_bar[stage]);
          if (kb == num_kb - 1) ptx::umma_commit_pair_w(&tmem_full_bar[as]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------ epilogue
    constexpr int kEpiThreads = 256;
    const int ew = warp & 3;
    const int ehalf = (warp - EPI_WARP0) >> 2;
    uint8_t* stg = sStaging + (warp - EPI_WARP0) * 4096;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    const bool bf16 = p.ab_format != 0;
    const bool ln_prod = p.ln_gamma != nullptr;
    const int N_fc = 4 * p.D, N_proj = p.D;
    // residual rings of warps 4..7 (ehalf == 0): boxes of this unit's c_proj tiles, prefetched RS-1 ahead across tiles
    uint64_t* my_full = res_full_bar + ew * RS;
    uint8_t* my_ring = sRing + ew * (RS * RES_BOX);
    int pf_i = num_tiles, pf_c = 0;  // lane 0 of a ring warp: tile index / chunk of the next box to prefetch
    int pf_q = 0;
    auto pf_advance = [&]() {  // next c_proj tile of this unit after pf_i
      do {
        pf_i += num_units;
      } while (pf_i < num_tiles && ((__ldg(p.tiles + pf_i) >> 31) == 0 || __ldg(p.tiles + pf_i) == MLP_TILE_NONE));
    };
    auto res_prefetch = [&]() {
      if (pf_i >= num_tiles) return;
      const uint32_t d = __ldg(p.tiles + pf_i);
      const int pr = static_cast<int>((d >> 8) & 0x7fffffu), nb = static_cast<int>(d & 0xffu);
      const int slot = pf_q % RS;
      ptx::mbar_expect_tx(&my_full[slot], RES_BOX);
      ptx::tma_load_2d(my_ring + slot * RES_BOX, &tmap_x, &my_full[slot], nb * BN + pf_c * 32,
                       pr * (2 * BM) + cta_rank * BM + ew * 32);
      ++pf_q;
      if (++pf_c == CPT) {
        pf_c = 0;
        pf_advance();
      }
    };
    int q = 0;
    if (ehalf == 0 && lane == 0) {
      pf_i = unit - num_units;
      pf_advance();
      for (int i = 0; i < RS - 1; ++i) res_prefetch();
    }
    // per-column vectors of the NEXT tile are fetched one tile ahead (see gemm_kernel)
    float nxt_b, nxt_x;
    auto fetch_cols = [&](int i) {
      nxt_b = nxt_x = 0.0f;
      while (i < num_tiles && __ldg(p.tiles + i) == MLP_TILE_NONE) i += num_units;
      if (i >= num_tiles) return;
      const uint32_t d = __ldg(p.tiles + i);
      const int n = static_cast<int>(d & 0xffu) * BN + et;
      if (d >> 31) {
        nxt_b = __ldg(p.proj_bias + n);
        if (ln_prod) nxt_x = __ldg(p.ln_gamma + n);
      } else {
        nxt_b = __ldg(p.fc_bias + n);
        nxt_x = __ldg(p.fc_s + n);
      }
    };
    fetch_cols(unit);
    // Completion signal of a c_fc tile (this warp's TMA stores are in global memory -> ctr_done[pair-row] += 1).  Waiting
    // for the stores right after issuing them would sit in front of the next epilogue, so the signal is deferred: to
    // after the tmem_full wait of the next tile if that is a c_fc tile (its start waits for nothing this signal could
    // hold up), and to the top of the next tile otherwise (a c_proj tile may depend, through other pairs, on this
    // very signal; the list keeps >= 2 rounds between a pair-row's c_fc and c_proj tiles except in the final rounds).
    int pend_pr = -1;
    auto flush_done = [&]() {
      if (pend_pr >= 0 && lane == 0) {
        ptx::bulk_wait_all();
        ptx::fence_proxy_async_all();
        ptx::red_release_gpu_add(p.ctr_done + pend_pr, 1u);
      }
      pend_pr = -1;
    };
    int it = 0;
    for (int i = unit; i < num_tiles; i += num_units) {
      const uint32_t d = __ldg(p.tiles + i);
      if (d == MLP_TILE_NONE) continue;
      const bool proj = (d >> 31) != 0;
      if (proj) flush_done();
      const int pr = static_cast<int>((d >> 8) & 0x7fffffu), nb = static_cast<int>(d & 0xffu);
      const int m0 = pr * (2 * BM) + cta_rank * BM + ew * 32;
      const int n0 = nb * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ++it;

      float* sb = sBias + as * BN;
      float* sx = sBias + (2 + as) * BN;
      sb[et] = nxt_b;
      sx[et] = nxt_x;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      fetch_cols(i + num_units);

      float ln_r = 1.0f, ln_nrm = 0.0f;
      if (!proj) {  // LayerNorm consumer statistics of this thread's row (see gemm_kernel)
        const float* st = p.ln_stats + static_cast<size_t>(m0 + lane) * p.ln_nsb * 2;
        float2 sv[kMaxStatBlocks];
        float msum = 0.f;
#pragma unroll
        for (int b = 0; b < kMaxStatBlocks; ++b) {
          if (b < p.ln_nsb) {
            sv[b] = __ldg(reinterpret_cast<const float2*>(st) + b);
            msum += sv[b].x;
          }
        }
        const float mu = msum / static_cast<float>(p.ln_nsb);
        float m2 = 0.f;
#pragma unroll
        for (int b = 0; b < kMaxStatBlocks; ++b) {
          if (b < p.ln_nsb) {
            const float dd = sv[b].x - mu;
            m2 += fmaf(static_cast<float>(STAT_COLS) * dd, dd, sv[b].y);
          }
        }
        const float var = fmaxf(m2 / static_cast<float>(p.D), 0.0f);
        ln_r = rsqrtf(var + 1e-5f);
        ln_nrm = -ln_r * mu;
      }

      if (!proj && pr >= p.ring_pairs) {  // the ring slot's previous pair-row must have been consumed
        // (a BLOCKING wait never sits in front of a deferred completion signal: the consumer may be waiting for it)
        const unsigned* cc = p.ctr_cons + pr - p.ring_pairs;
        const bool block = __shfl_sync(0xffffffffu, lane == 0 && ptx::ld_acquire_gpu(cc) < p.need_cons, 0);
        if (block) {
          flush_done();
          if (lane == 0) ptx::wait_counter(cc, p.need_cons);
        }
        if (lane == 0) ptx::fence_proxy_async_all();
      }
      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tc_fence_after();
      if (!proj) flush_done();  // the previous c_fc tile's stores completed while this tile's MMAs ran: free to count now
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

      if (!proj) {
        // ---- c_fc tile: 64-column chunks dealt alternately to the two warps of a lane quadrant, TMA store into the ring
        __syncwarp();
        const int st_row = (pr % p.ring_pairs) * (2 * BM) + cta_rank * BM + ew * 32;
#pragma unroll 1
        for (int c = ehalf; c < BN / 64; c += 2) {
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_32x32(taddr + c * 64, ra);
          ptx::tmem_ld_32x32(taddr + c * 64 + 32, rb);
          ptx::tmem_ld_wait();
          if (c + 2 >= BN / 64) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
          }
          const float* cb = sb + c * 64;
          const float* cx = sx + c * 64;
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v0 = __uint_as_float(j < 16 ? ra[2 * j] : rb[2 * j - 32]);
            const float v1 = __uint_as_float(j < 16 ? ra[2 * j + 1] : rb[2 * j - 31]);
            const float a0 = quick_gelu(fmaf(ln_r, v0, fmaf(ln_nrm, cx[2 * j], cb[2 * j])));
            const float a1 = quick_gelu(fmaf(ln_r, v1, fmaf(ln_nrm, cx[2 * j + 1], cb[2 * j + 1])));
            pk[j] = bf16 ? ptx::pack2<true>(a0, a1) : ptx::pack2<false>(a0, a1);
          }
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_hst, stg, n0 + c * 64, st_row);
            ptx::bulk_commit();
          }
        }
        pend_pr = pr;  // counted in ctr_done once the stores have completed (flush_done)
      } else {
        // ---- c_proj tile: every MMA has retired, so this CTA's A rows of the ring are consumed
        if (warp == EPI_WARP0 && lane == 0) ptx::red_release_gpu_add(p.ctr_cons + pr, 1u);
        if (ehalf == 0) {
          if (ln_prod) {  // the gamma * x transposition reuses this warp's c_fc staging tile: its last TMA store has read it
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
          }
          float ln_s1 = 0.f, ln_s2 = 0.f, ln_piv = 0.f;
          const int prow = m0 + lane;
#pragma unroll 1
          for (int c = 0; c < CPT; ++c, ++q) {
            const int slot = q % RS;
            uint32_t r[32];
            ptx::tmem_ld_32x32(taddr + c * 32, r);
            ptx::mbar_wait(&my_full[slot], (q / RS) & 1);
            ptx::tmem_ld_wait();
            uint8_t* box = my_ring + slot * RES_BOX;
            float4 xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = *reinterpret_cast<const float4*>(box + lane * 128 + ((u ^ (lane & 7)) << 4));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 32 + 4 * u);
              float4 v = xv[u];
              v.x += __uint_as_float(r[4 * u]) + b4.x;
              v.y += __uint_as_float(r[4 * u + 1]) + b4.y;
              v.z += __uint_as_float(r[4 * u + 2]) + b4.z;
              v.w += __uint_as_float(r[4 * u + 3]) + b4.w;
              xv[u] = v;
              if (ln_prod) {
                if (u == 0 && (c & (STAT_COLS / 32 - 1)) == 0) ln_piv = v.x;
                const float dx = v.x - ln_piv, dy = v.y - ln_piv, dz = v.z - ln_piv, dw = v.w - ln_piv;
                ln_s1 += (dx + dy) + (dz + dw);
                ln_s2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, ln_s2))));
                const float4 g4 = *reinterpret_cast<const float4*>(sx + c * 32 + 4 * u);
                r[2 * u] = bf16 ? ptx::pack2_sat<true>(g4.x * v.x, g4.y * v.y) : ptx::pack2_sat<false>(g4.x * v.x, g4.y * v.y);
                r[2 * u + 1] = bf16 ? ptx::pack2_sat<true>(g4.z * v.z, g4.w * v.w) : ptx::pack2_sat<false>(g4.z * v.z, g4.w * v.w);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(box + lane * 128 + ((u ^ (lane & 7)) << 4)) = xv[u];
            if (ln_prod) {
              uint8_t* ast = stg;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                *reinterpret_cast<uint4*>(ast + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
              }
              __syncwarp();
              const int u = lane & 3;
              const int gcol = n0 + c * 32 + u * 8;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int row = k * 8 + (lane >> 2);
                const uint4 v = *reinterpret_cast<const uint4*>(ast + row * 64 + ((u ^ ((row >> 1) & 3)) << 4));
                *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.a16_out) + static_cast<size_t>(m0 + row) * N_proj + gcol) = v;
              }
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tmap_x, box, n0 + c * 32, m0);
              ptx::bulk_commit();
              ptx::bulk_wait_read<1>();
              res_prefetch();
            }
            __syncwarp();
            if (ln_prod && (c & (STAT_COLS / 32 - 1)) == STAT_COLS / 32 - 1) {
              const int sblk = (n0 + c * 32) / STAT_COLS;
              *reinterpret_cast<float2*>(p.stats_out + (static_cast<size_t>(prow) * (N_proj / STAT_COLS) + sblk) * 2) =
                  make_float2(fmaf(ln_s1, 1.0f / STAT_COLS, ln_piv), fmaf(-ln_s1 * (1.0f / STAT_COLS), ln_s1, ln_s2));
              ln_s1 = ln_s2 = 0.f;
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
      }
      (void)N_fc;
    }
    flush_done();
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

template <int BN, int EPI>
cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SmemLayout<BN, EPI, false>::kDynamic);
  if (e != cudaSuccess || BN != 256) return e;
  return cudaFuncSetAttribute(gemm_kernel<256, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SmemLayout<256, EPI, true>::kDynamic);
}

template <int BN, int EPI>
cudaError_t launch_one(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const GemmParams& p,
                       int grid, bool pair, cudaStream_t stream) {
  if constexpr (BN == 256) {
    if (pair)  // CTA pairs: cluster of 2, grid = 2 x #pairs
      return launch_kernel(gemm_kernel<256, EPI, true>, grid, num_threads(EPI, true), SmemLayout<256, EPI, true>::kDynamic,
                           stream, 2, true, ta, tw, tc, p);
  }
  return launch_kernel(gemm_kernel<BN, EPI, false>, grid, num_threads(EPI, false), SmemLayout<BN, EPI, false>::kDynamic, stream,
                       1, true, ta, tw, tc, p);
}

// the wide residual epilogue pays when the main loop is short (see EPI_RES_WIDE); AIHAB_RES_WIDE=0 disables it
bool res_wide(const GemmParams& p) {
  static const bool enabled = [] {
    const char* e = getenv("AIHAB_RES_WIDE");
    return e == nullptr || e[0] != '0';
  }();
  return enabled && p.K <= 1024 && p.ring_mode == 0;
}

template <int BN>
cudaError_t launch_bn(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const GemmParams& p,
                      int grid, bool pair, cudaStream_t stream) {
  switch (p.epilogue) {
    case EPI_BIAS_16: return launch_one<BN, EPI_BIAS_16>(ta, tw, tc, p, grid, pair, stream);
    case EPI_BIAS_GELU_16: return launch_one<BN, EPI_BIAS_GELU_16>(ta, tw, tc, p, grid, pair, stream);
    case EPI_BIAS_RES_32:
      if constexpr (BN == 256) {
        if (pair && res_wide(p))
          return launch_kernel(gemm_kernel<256, EPI_RES_WIDE, true>, grid, num_threads(EPI_RES_WIDE, true),
                               SmemLayout<256, EPI_RES_WIDE, true>::kDynamic, stream, 2, true, ta, tw, tc, p);
      }
      return launch_one<BN, EPI_BIAS_RES_32>(ta, tw, tc, p, grid, pair, stream);
    case EPI_PATCH_32: return launch_one<BN, EPI_PATCH_32>(ta, tw, tc, p, grid, pair, stream);
    case EPI_SCALE_32: return launch_one<BN, EPI_SCALE_32>(ta, tw, tc, p, grid, pair, stream);
    case EPI_SPLIT3_16: return launch_one<BN, EPI_SPLIT3_16>(ta, tw, tc, p, grid, pair, stream);
    case EPI_LN_BIAS_16: return launch_one<BN, EPI_LN_BIAS_16>(ta, tw, tc, p, grid, pair, stream);
    case EPI_LN_BIAS_GELU_16: return launch_one<BN, EPI_LN_BIAS_GELU_16>(ta, tw, tc, p, grid, pair, stream);
    case EPI_TOPK_32:
      if (p.topk_k >= 1 && p.topk_k <= 5) return launch_one<BN, EPI_TOPK5>(ta, tw, tc, p, grid, pair, stream);
      return launch_one<BN, EPI_TOPK_32>(ta, tw, tc, p, grid, pair, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t gemm_init() {
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorNotSupported;
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cudaError_t e;
#define AIHAB_SET(BN, EPI) \
  if ((e = set_attr<BN, EPI>()) != cudaSuccess) return e;
  AIHAB_SET(256, EPI_BIAS_16) AIHAB_SET(256, EPI_BIAS_GELU_16) AIHAB_SET(256, EPI_BIAS_RES_32)
  AIHAB_SET(256, EPI_PATCH_32) AIHAB_SET(256, EPI_SCALE_32) AIHAB_SET(256, EPI_LN_BIAS_16)
  AIHAB_SET(256, EPI_LN_BIAS_GELU_16) AIHAB_SET(128, EPI_LN_BIAS_16) AIHAB_SET(128, EPI_LN_BIAS_GELU_16)
  AIHAB_SET(128, EPI_BIAS_16) AIHAB_SET(128, EPI_BIAS_GELU_16) AIHAB_SET(128, EPI_BIAS_RES_32)
  AIHAB_SET(128, EPI_PATCH_32) AIHAB_SET(128, EPI_SCALE_32) AIHAB_SET(256, EPI_TOPK_32) AIHAB_SET(128, EPI_TOPK_32)
  AIHAB_SET(256, EPI_TOPK5) AIHAB_SET(128, EPI_TOPK5) AIHAB_SET(256, EPI_SPLIT3_16) AIHAB_SET(128, EPI_SPLIT3_16)
#undef AIHAB_SET
  if ((e = cudaFuncSetAttribute(gemm_kernel<256, EPI_RES_WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SmemLayout<256, EPI_RES_WIDE, true>::kDynamic)) != cudaSuccess)
    return e;
  if ((e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::kDynamic)) !=
      cudaSuccess)
    return e;
  return cudaSuccess;
}

cudaError_t make_tmap_2d_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t row_pitch_bytes, uint32_t box_rows, int ab_format) {
  if (g_encode == nullptr) {
    cudaError_t e = gemm_init();
    if (e != cudaSuccess) return e;
  }
  if ((row_pitch_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || box_rows > 256)
    return cudaErrorInvalidValue;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, ab_format ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_tmap_3d_16bit_seq(CUtensorMap* map, const void* base, uint64_t n_seq, uint64_t seq_rows, uint64_t cols,
                                   uint32_t box_rows, int ab_format) {
  if (g_encode == nullptr) {
    cudaError_t e = gemm_init();
    if (e != cudaSuccess) return e;
  }
  if (((cols * 2) & 15) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || box_rows > 256 || n_seq == 0)
    return cudaErrorInvalidValue;
  cuuint64_t gdim[3] = {cols, seq_rows, n_seq};
  cuuint64_t gstride[2] = {cols * 2, seq_rows * cols * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, ab_format ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int gemm_block_n(int M, int N, int num_sms) {
  if (N <= 128) return 128;
  const long m_blocks = (M + BM - 1) / BM;
  const long t256 = m_blocks * ((N + 255) / 256);
  const long t128 = m_blocks * ((N + 127) / 128);
  const long cost256 = ((t256 + num_sms - 1) / num_sms) * 256;
  const long cost128 = ((t128 + num_sms - 1) / num_sms) * 128;
  return cost128 < cost256 ? 128 : 256;
}

cudaError_t make_tmap_2d_f32_box32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                                   uint64_t row_pitch_bytes) {
  if (g_encode == nullptr) {
    cudaError_t e = gemm_init();
    if (e != cudaSuccess) return e;
  }
  if ((row_pitch_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return cudaErrorInvalidValue;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

void build_mlp_tiles(int pair_rows, int n_fc, int n_proj, int units, int ring_pairs, std::vector<uint32_t>* out, int lag,
                     int extra) {
  lag = std::max(lag, 2);  // the deferred completion signal of a c_fc tile needs two rounds of slack (mlp_fused_kernel)
  const long F = static_cast<long>(pair_rows) * n_fc, J = static_cast<long>(pair_rows) * n_proj;
  // one balanced round of c_proj tiles closes the kernel (if the ring can hold that many pair-rows)
  const long reserve = std::min<long>(std::min<long>(J, units), static_cast<long>(std::max(0, ring_pairs - 8)) * n_proj);
  const int base = std::max(1, units * n_proj / (n_fc + n_proj));
  std::vector<int> last_round(pair_rows, -1);  // round in which the pair-row's last c_fc tile was dealt
  out->clear();
  long f = 0, j = 0;
  int start = 0;
  for (int r = 0; f < F || j < J; ++r) {
    const int eff_lag = f < F ? lag : 0;
    long avail = 0;
    for (long jj = j; jj < J; ++jj) {
      const int lr = last_round[jj / n_proj];
      if (lr < 0 || lr > r - eff_lag) break;
      ++avail;
    }
    // a c_fc tile of pair-row pr overwrites the ring slot of pr - ring_pairs: deal it only after every c_proj tile of
    // that pair-row was dealt in an EARLIER round (deadlock freedom: waits only ever point to earlier rounds)
    const long fc_pr_limit = ring_pairs + j / n_proj;
    long quota;
    if (f < F) {
      const long cap = std::max<long>(0, (J - reserve) - j);
      quota = std::min<long>(std::min(avail, cap), base + (avail > base ? extra : 0));
    } else {
      quota = std::min<long>(avail, units);
    }
    const size_t row0 = out->size();
    out->resize(row0 + units, MLP_TILE_NONE);
    std::vector<char> chosen(units, 0);
    for (long k = 0; k < quota; ++k) chosen[(start + k) % units] = 1;
    start = static_cast<int>((start + quota) % units);
    long taken = 0;
    for (int u = 0; u < units; ++u) {
      const bool fc_ok = f < F && f / n_fc < fc_pr_limit;
      if (j < J && taken < avail && (chosen[u] || !fc_ok)) {
        (*out)[row0 + u] = 0x80000000u | (static_cast<uint32_t>(j / n_proj) << 8) | static_cast<uint32_t>(j % n_proj);
        ++j;
        ++taken;
      } else if (fc_ok) {
        const int pr = static_cast<int>(f / n_fc), n = static_cast<int>(f % n_fc);
        (*out)[row0 + u] = (static_cast<uint32_t>(pr) << 8) | static_cast<uint32_t>(n);
        ++f;
        if (n == n_fc - 1) last_round[pr] = r;
      }
    }
  }
}

cudaError_t launch_mlp_fused(const CUtensorMap& tmap_y, const CUtensorMap& tmap_wfc, const CUtensorMap& tmap_hld,
                             const CUtensorMap& tmap_wproj, const CUtensorMap& tmap_x, void* ring_base,
                             const MlpFusedParams& p, int num_sms, cudaStream_t stream) {
  if (p.M <= 0 || (p.M % (2 * BM)) || p.D <= 0 || (p.D % 256) || p.D / STAT_COLS > kMaxStatBlocks || p.tiles == nullptr ||
      p.num_tiles <= 0 || p.ring_pairs <= 0 || p.ctr_done == nullptr || p.ctr_cons == nullptr || p.fc_bias == nullptr ||
      p.fc_s == nullptr || p.ln_stats == nullptr || p.ln_nsb <= 0 || p.ln_nsb > kMaxStatBlocks || p.proj_bias == nullptr ||
      4 * p.D / 256 > 255 || num_sms < 2 || (reinterpret_cast<uintptr_t>(ring_base) & 15))
    return cudaErrorInvalidValue;
  if (p.ln_gamma != nullptr && (p.a16_out == nullptr || p.stats_out == nullptr)) return cudaErrorInvalidValue;
  if (g_encode == nullptr) {
    cudaError_t e = gemm_init();
    if (e != cudaSuccess) return e;
  }
  CUtensorMap hst;  // 64-column x 32-row store boxes over the ring
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(4 * p.D), static_cast<cuuint64_t>(p.ring_pairs) * 2 * BM};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(4 * p.D) * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  if (g_encode(&hst, p.ab_format ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ring_base, gdim,
               gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  const int units = num_sms / 2;
  if (p.num_tiles % units) return cudaErrorInvalidValue;  // the list was built for this many pairs
  return launch_kernel(mlp_fused_kernel, 2 * units, 384, FusedSmem::kDynamic, stream, 2, true, tmap_y, tmap_wfc, tmap_hld,
                       tmap_wproj, hst, tmap_x, p);
}

bool gemm_use_pair(int M, int N, int num_sms) {
  // CTA pairs finish a 256-wide tile ~5 % faster (half the W traffic per SM) but schedule in units of two M blocks:
  // take them unless wave quantisation costs more than that.
  if (gemm_block_n(M, N, num_sms) != 256 || num_sms < 2) return false;
  const long m_blocks = (M + BM - 1) / BM, n_blocks = (N + 255) / 256;
  const long t1 = m_blocks * n_blocks, t2 = ((m_blocks + 1) / 2) * n_blocks;
  const long u1 = num_sms, u2 = num_sms / 2;
  const double eff1 = static_cast<double>(t1) / (((t1 + u1 - 1) / u1) * u1);
  const double eff2 = static_cast<double>(m_blocks * n_blocks) / (((t2 + u2 - 1) / u2) * u2 * 2);
  return eff2 * 1.05 >= eff1;
}

cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w, const CUtensorMap* tmap_c,
                        const GemmParams& p, int block_n, int num_sms, cudaStream_t stream, bool pair) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return cudaErrorInvalidValue;
  if (p.ring_mode != 0) {  // pipelined pair: CTA pairs, whole pair-rows, forward traversal, counters present
    if (!pair || p.reverse_m || (p.M % (2 * BM)) || p.ring_rows <= 0 || (p.ring_rows % (2 * BM)) || p.ctr_done == nullptr ||
        p.ctr_consumed == nullptr || (p.ring_mode != 1 && p.ring_mode != 2))
      return cudaErrorInvalidValue;
  }
  const bool ln = (p.epilogue == EPI_LN_BIAS_16 || p.epilogue == EPI_LN_BIAS_GELU_16);
  const bool split3 = p.epilogue == EPI_SPLIT3_16;
  const bool out16 = (p.epilogue == EPI_BIAS_16 || p.epilogue == EPI_BIAS_GELU_16 || ln || split3);
  if (split3 && (p.stats_out == nullptr || p.ldo != 3 * p.N || (p.N & 63) || p.ring_mode != 0)) return cudaErrorInvalidValue;
  if (p.row_ss != nullptr && p.row_ss_n <= 0) return cudaErrorInvalidValue;
  if (ln && (p.ln_stats == nullptr || p.ln_s == nullptr || p.ln_nsb <= 0 || p.bias == nullptr)) return cudaErrorInvalidValue;
  if (p.epilogue == EPI_BIAS_RES_32 && p.ln_gamma != nullptr && (p.a16_out == nullptr || p.stats_out == nullptr))
    return cudaErrorInvalidValue;
  if (out16 && ((p.N & 7) || (p.ldo & 7) || p.out16 == nullptr)) return cudaErrorInvalidValue;
  const bool topk = p.epilogue == EPI_TOPK_32;
  if (topk && (p.cand_val == nullptr || p.cand_idx == nullptr)) return cudaErrorInvalidValue;
  if (!out16 && !topk && ((p.N & 3) || (p.ldo & 3) || p.out32 == nullptr)) return cudaErrorInvalidValue;
  if (p.epilogue == EPI_PATCH_32 && (p.pos == nullptr || p.g2 <= 0)) return cudaErrorInvalidValue;
  if (p.epilogue == EPI_BIAS_RES_32 && (tmap_c == nullptr || (p.N & 31))) return cudaErrorInvalidValue;
  CUtensorMap out_map;
  if (out16 && tmap_c == nullptr) {  // the 16-bit epilogues store through TMA: 64-column x 32-row boxes over out16
    if (reinterpret_cast<uintptr_t>(p.out16) & 15) return cudaErrorInvalidValue;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(split3 ? 3 * p.N : p.N), static_cast<cuuint64_t>(p.ring_mode == 1 ? p.ring_rows : p.M)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(p.ldo) * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    if (g_encode(&out_map, p.ab_format ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.out16,
                 gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    tmap_c = &out_map;
  }
  const CUtensorMap& tc = tmap_c ? *tmap_c : tmap_a;  // unused by the fp32 epilogues other than EPI_BIAS_RES_32
  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_blocks = (p.N + block_n - 1) / block_n;
  if (pair) {  // tmap_w must have a 128-row box (half of the 256-wide W tile per CTA)
    if (block_n != 256) return cudaErrorInvalidValue;
    const long pair_tiles = static_cast<long>((m_blocks + 1) / 2) * n_blocks;
    const long pairs = std::min<long>(pair_tiles, num_sms / 2);
    return launch_bn<256>(tmap_a, tmap_w, tc, p, static_cast<int>(2 * pairs), true, stream);
  }
  const long tiles = static_cast<long>(m_blocks) * n_blocks;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  if (block_n == 256) return launch_bn<256>(tmap_a, tmap_w, tc, p, grid, false, stream);
  if (block_n == 128) return launch_bn<128>(tmap_a, tmap_w, tc, p, grid, false, stream);
  return cudaErrorInvalidValue;
}

}  // namespace aihab

// Debug export (not part of include/aihab_clip.h): the fused-MLP tile list for the CPU tests of its invariants.
// Returns the number of entries (rounds x units); writes min(cap, entries) of them.
extern "C" __attribute__((visibility("default"))) long aihab_debug_mlp_tiles(int pair_rows, int n_fc, int n_proj, int units,
                                                                             int ring_pairs, uint32_t* out, long cap) {
  if (pair_rows <= 0 || n_fc <= 0 || n_proj <= 0 || units <= 0 || ring_pairs <= 0 || n_fc > 255 || n_proj > 255) return -1;
  std::vector<uint32_t> v;
  aihab::build_mlp_tiles(pair_rows, n_fc, n_proj, units, ring_pairs, &v);
  for (long i = 0; i < cap && i < static_cast<long>(v.size()); ++i) out[i] = v[i];
  return static_cast<long>(v.size());
}
