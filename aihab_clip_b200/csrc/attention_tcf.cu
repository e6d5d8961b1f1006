// Persistent flash-style tcgen05 / TMEM attention for long sequences (L > 224: ViT-L/14 L = 257, ViT-L/14@336px
// L = 577), head dim 64.  One CTA per SM walks (image, head, 128-query tile) items; the keys of an item are visited in
// nb blocks of <= 160 keys (multiples of 32, balanced) with an online softmax:
//   warp 0 (lane 0)  TMA producer : Q tile (double-buffered per item); K / V blocks through a 3-stage ring of
//                                   32-row boxes
//   warp 1 (lane 0)  MMA issuer   : S(g+1) = Q K_j^T goes into the other TMEM S buffer before it waits for P(g);
//                                   O (+)= P(g) V_j with A = P read from TMEM and V as an MN-major smem operand
//   warps 2..9       softmax      : two threads per query row (column halves of the block).  S is read from TMEM once,
//                                   block max exchanged through smem; the running reference max is only raised when
//                                   the block max exceeds it by more than 2^8 (P stays <= 256, exact in fp16/bf16
//                                   range), and only then O is rescaled in TMEM (tcgen05.ld / st) - rare after the
//                                   first block.  P is written over S as packed 16-bit pairs (tcgen05.st).
// TMEM: S0/P0 [0,160) S1/P1 [160,320) O [320,384).  smem: 2 x Q 16 KB + 3 x (K 20 KB + V 20 KB) = 152 KB.
// A tail of <= 16 query rows beyond the last full tile (L = 257: one row) is left to the mma.sync kernel
// (launch_attention_rows) instead of a 128-row tile of padding.
// Reference: clip/model.py:179-181 (nn.MultiheadAttention core: softmax(q k^T / sqrt(64)) v, no mask).
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace aihab {

namespace {

constexpr int F_THREADS = 320;
constexpr int KB_MAX = 160;  // keys per block: 2 x 160 S columns + 64 O columns of TMEM, 5 chunks of 16 per thread
constexpr int NST = 3;
constexpr int SQ_BYTES = 128 * 128;
constexpr int SK_BYTES = KB_MAX * 128;
constexpr int SKV_BYTES = 2 * SK_BYTES;
constexpr int OFF_Q = 0;
constexpr int OFF_KV = 2 * SQ_BYTES;
constexpr int OFF_BAR = OFF_KV + NST * SKV_BYTES;
constexpr int OFF_RED = OFF_BAR + 256;  // s_max[2][256], s_sum[2][256]
constexpr int F_SMEM = OFF_RED + 4096 + 1024;
constexpr int TM_S = KB_MAX;
constexpr int TM_O = 2 * KB_MAX;
constexpr int NCH = KB_MAX / 32;  // 16-column chunks per thread, at most
constexpr int MAX_BLOCKS = 8;
static_assert(OFF_KV % 1024 == 0 && SK_BYTES % 1024 == 0, "SWIZZLE_128B tiles need 1024 B alignment");
static_assert(F_SMEM <= 227 * 1024, "smem budget");

struct FlashParams {
  int L, H, nq, nb, total, reverse;
  int ub, ur;  // 32-key units per block: blocks j < ur have ub + 1 units, the others ub
  // first key of block j (j = nb: keys padded to 32)
  __host__ __device__ int kb_off(int j) const { return 32 * (j * ub + (j < ur ? j : ur)); }
};

template <bool BF16>
__global__ void __launch_bounds__(F_THREADS, 1)
attention_tcf_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     uint16_t* __restrict__ out, const FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024 B aligned, still a __shared__ pointer
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_q = bars + 0;        // [2] Q tile of an item landed
  uint64_t* bar_qfree = bars + 2;    // [2] every S MMA of the item has retired
  uint64_t* bar_k = bars + 4;        // [NST] K block landed
  uint64_t* bar_v = bars + 7;        // [NST] V block landed
  uint64_t* bar_kvfree = bars + 10;  // [NST] PV MMA of the step retired: stage reusable
  uint64_t* bar_sfull = bars + 13;   // [2] S buffer written
  uint64_t* bar_pvdone = bars + 15;  // [2] PV MMA of the step using S/P buffer b retired: the buffer is free
  uint64_t* bar_p = bars + 17;       // P written to TMEM (and O rescaled / read)
  uint64_t* bar_o = bars + 18;       // PV of a step retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* s_max = reinterpret_cast<float*>(smem + OFF_RED);  // [2 (step parity)][256]
  float* s_sum = s_max + 512;                               // [2 (item parity)][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int L = p.L, H = p.H, nb = p.nb;
  const int D = H * 64;
  const int n_items = static_cast<int>(blockIdx.x) < p.total ? (p.total - static_cast<int>(blockIdx.x) + G - 1) / G : 0;
  const int total_steps = n_items * nb;

  auto decode = [&](int it, int& img, int& h, int& qt) {
    const int idx = static_cast<int>(blockIdx.x) + it * G;
    int u = idx / p.nq;
    qt = idx - u * p.nq;
    if (p.reverse) u = p.total / p.nq - 1 - u;  // walk (image, head) units from the end: the producer's freshest rows first
    img = u / H;
    h = u - img * H;
  };

  ptx::griddep_launch();
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_kv);
      for (int i = 0; i < 17; ++i) ptx::mbar_init(&bars[i], 1);
      ptx::mbar_init(bar_p, 256);
      ptx::mbar_init(bar_o, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::griddep_wait();  // qkv comes from the previous kernel of the stream

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0, ring = 0;  // K/V stage of the step and how many times the ring has wrapped
      for (int it = 0; it < n_items; ++it) {
        int img, h, qt;
        decode(it, img, h, qt);
        const int qs = it & 1;
        const int row0 = img * L;
        if (it >= 2) ptx::mbar_wait(&bar_qfree[qs], ((it >> 1) - 1) & 1);
        ptx::mbar_expect_tx(&bar_q[qs], SQ_BYTES);
        ptx::tma_load_2d(smem + OFF_Q + qs * SQ_BYTES, &tmap_q, &bar_q[qs], h * 64, row0 + qt * 128);
        for (int j = 0; j < nb; ++j) {
          if (ring > 0) ptx::mbar_wait(&bar_kvfree[st], (ring - 1) & 1);
          const int key0 = row0 + p.kb_off(j);
          const int nbox = (p.kb_off(j + 1) - p.kb_off(j)) >> 5;
          uint8_t* kd = smem + OFF_KV + st * SKV_BYTES;
          ptx::mbar_expect_tx(&bar_k[st], nbox * 4096);
          for (int r = 0; r < nbox; ++r)
            ptx::tma_load_2d(kd + r * 4096, &tmap_kv, &bar_k[st], D + h * 64, key0 + r * 32);
          ptx::mbar_expect_tx(&bar_v[st], nbox * 4096);
          for (int r = 0; r < nbox; ++r)
            ptx::tma_load_2d(kd + SK_BYTES + r * 4096, &tmap_kv, &bar_v[st], 2 * D + h * 64, key0 + r * 32);
          if (++st == NST) {
            st = 0;
            ++ring;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (total_steps > 0) {  // whole warp, converged; one elected lane issues each tcgen05 instruction (ptx.cuh "_w")
      const uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, 64, /*b_mn_major=*/1);
      // S(g) = Q(it) K_j^T into S buffer g & 1
      auto issue_s = [&](int g, int it, int j, int st, int ring) {
        const int b = g & 1;
        ptx::mbar_wait(&bar_k[st], ring & 1);
        ptx::mbar_wait(&bar_q[it & 1], (it >> 1) & 1);
        if (g >= 2) ptx::mbar_wait(&bar_pvdone[b], ((g >> 1) - 1) & 1);  // P(g-2) lives there until PV(g-2) retires
        ptx::tc_fence_after();
        const int kb = p.kb_off(j + 1) - p.kb_off(j);
        const uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, kb);
        const uint64_t qd = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + OFF_Q + (it & 1) * SQ_BYTES));
        const uint64_t kd = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + OFF_KV + st * SKV_BYTES));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_w(tmem + b * TM_S, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
        ptx::umma_commit_w(&bar_sfull[b]);
        if (j == nb - 1) ptx::umma_commit_w(&bar_qfree[it & 1]);
      };
      int it = 0, j = 0, st = 0, ring = 0;      // step g
      int itn = 0, jn = 0, stn = 0, ringn = 0;  // step g + 1
      auto advance = [&](int& a_it, int& a_j, int& a_st, int& a_ring) {
        if (++a_j == nb) {
          a_j = 0;
          ++a_it;
        }
        if (++a_st == NST) {
          a_st = 0;
          ++a_ring;
        }
      };
      issue_s(0, 0, 0, 0, 0);
      advance(itn, jn, stn, ringn);
      for (int g = 0; g < total_steps; ++g) {
        if (g + 1 < total_steps) issue_s(g + 1, itn, jn, stn, ringn);
        const int b = g & 1;
        ptx::mbar_wait(bar_p, g & 1);
        ptx::mbar_wait(&bar_v[st], ring & 1);
        ptx::tc_fence_after();
        const int ksteps = (p.kb_off(j + 1) - p.kb_off(j)) >> 4;
        const uint32_t v_base = ptx::smem_u32(smem + OFF_KV + st * SKV_BYTES + SK_BYTES);
        for (int t = 0; t < ksteps; ++t) {
          const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + t * 2048);
          ptx::umma_f16_ts_w(tmem + TM_O, tmem + b * TM_S + t * 8, vd, idesc_o, (j | t) != 0);  // A = P from TMEM
        }
        ptx::umma_commit_w(bar_o);
        ptx::umma_commit_w(&bar_kvfree[st]);
        ptx::umma_commit_w(&bar_pvdone[b]);
        advance(it, j, st, ring);
        advance(itn, jn, stn, ringn);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + O rescale + epilogue (warps 2..9)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;

    // O(item) -> global: this thread owns 32 of the 64 output columns of its row
    auto epilogue = [&](int img, int h, int qt, int par) {
      if (qt * 128 + quad * 32 >= L) return;
      uint32_t o[32];
      ptx::tmem_ld_32x32(tmem + lane_off + TM_O + half * 32, o);
      ptx::tmem_ld_wait();
      const int grow = qt * 128 + row;
      if (grow < L) {
        const float inv_l = 1.0f / (s_sum[par * 256 + row] + s_sum[par * 256 + 128 + row]);
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
    };

    float m_ref = 0.f, l_part = 0.f;  // reference max of the row (raw S units), this thread's partial row sum
    int g = 0;
    int p_img = 0, p_h = 0, p_qt = 0;
    for (int it = 0; it < n_items; ++it) {
      int img, h, qt;
      decode(it, img, h, qt);
      const bool has_rows = qt * 128 + quad * 32 < L;
      for (int j = 0; j < nb; ++j, ++g) {
        const int b = g & 1;
        const int key0 = p.kb_off(j);
        const int nch = (p.kb_off(j + 1) - key0) >> 5;  // 16-column chunks per thread (1..5)
        const int c0 = half * nch;
        const uint32_t t_row = tmem + lane_off + b * TM_S;
        float* s_max_b = s_max + b * 256;

        ptx::mbar_wait(&bar_sfull[b], (g >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t r[NCH][16];
        if (has_rows) {
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            if (i < nch) ptx::tmem_ld_32x16(t_row + (c0 + i) * 16, r[i]);
          ptx::tmem_ld_wait();
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            if (i < nch) {
              const int col = key0 + (c0 + i) * 16;
              if (col + 16 <= L) {
#pragma unroll
                for (int jj = 0; jj < 16; jj += 2) {
                  m0 = fmaxf(m0, __uint_as_float(r[i][jj]));
                  m1 = fmaxf(m1, __uint_as_float(r[i][jj + 1]));
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                  if (col + jj < L) m0 = fmaxf(m0, __uint_as_float(r[i][jj]));
              }
            }
          }
          s_max_b[half * 128 + row] = fmaxf(m0, m1);
        }
        // after this barrier every S column of the block has been read into registers: P may overwrite S in place
        ptx::tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // the two warps sharing this lane quadrant
        ptx::tc_fence_after();
        float alpha = 1.0f;
        bool rescale = false;
        if (has_rows) {
          const float bm = fmaxf(s_max_b[row], s_max_b[128 + row]);
          if (j == 0) {
            m_ref = bm;
            l_part = 0.f;
          } else if ((bm - m_ref) * sl2 > 8.0f) {  // keep the old reference while P <= 2^8: no O rescale
            alpha = ptx::ex2_approx((m_ref - bm) * sl2);
            m_ref = bm;
            l_part *= alpha;
            rescale = true;
          }
          const float ms = m_ref * sl2;
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            if (i < nch) {
              const int col = key0 + (c0 + i) * 16;
              const bool full = col + 16 <= L;
              uint32_t pk[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                float p0 = ptx::ex2_approx(fmaf(__uint_as_float(r[i][2 * jj]), sl2, -ms));
                float p1 = ptx::ex2_approx(fmaf(__uint_as_float(r[i][2 * jj + 1]), sl2, -ms));
                if (!full) {
                  if (col + 2 * jj >= L) p0 = 0.f;
                  if (col + 2 * jj + 1 >= L) p1 = 0.f;
                }
                l0 += p0;
                l1 += p1;
                pk[jj] = ptx::pack2<BF16>(p0, p1);
              }
              ptx::tmem_st_32x8(t_row + (c0 + i) * 8, pk);  // 16 keys -> 8 packed columns of P, over the S buffer
            }
          }
          l_part += l0 + l1;
          if (j == nb - 1) s_sum[(it & 1) * 256 + half * 128 + row] = l_part;
          ptx::tmem_st_wait();
        }
        if (g > 0) {  // PV(g-1) was issued a whole softmax ago
          ptx::mbar_wait(bar_o, (g - 1) & 1);
          ptx::tc_fence_after();
          if (j == 0) {
            epilogue(p_img, p_h, p_qt, (it - 1) & 1);
          } else if (__any_sync(0xffffffffu, rescale)) {  // O *= alpha for the rows whose reference max moved
            uint32_t o[32];
            const uint32_t t_o = tmem + lane_off + TM_O + half * 32;
            ptx::tmem_ld_32x32(t_o, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(__uint_as_float(o[8 * u + e]) * alpha);
              ptx::tmem_st_32x8(t_o + 8 * u, w);
            }
            ptx::tmem_st_wait();
          }
        }
        ptx::tc_fence_before();   // P(g) stored, O rescaled / read ...
        ptx::mbar_arrive(bar_p);  // ... before PV(g) may read P and accumulate into / overwrite O
      }
      p_img = img;
      p_h = h;
      p_qt = qt;
    }
    if (n_items > 0) {
      ptx::mbar_wait(bar_o, (total_steps - 1) & 1);
      ptx::tc_fence_after();
      epilogue(p_img, p_h, p_qt, (n_items - 1) & 1);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

template <bool BF16>
cudaError_t launch_f(const CUtensorMap& tq, const CUtensorMap& tkv, uint16_t* out, const FlashParams& p, int grid,
                     cudaStream_t stream) {
  static bool attr_set[64] = {};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcf_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  return launch_kernel(attention_tcf_kernel<BF16>, grid, F_THREADS, F_SMEM, stream, 1, true, tq, tkv, out, p);
}

}  // namespace

bool attention_tcf_supported(int L) { return L > 128 && (L + 31) / 32 * 32 <= MAX_BLOCKS * KB_MAX; }

int attention_tcf_tail_rows(int L) {
  const int tail = L % 128;
  return (tail > 0 && tail <= 16) ? tail : 0;
}

cudaError_t launch_attention_tcf(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv32, const void* qkv, void* out,
                                 int n_img, int L, int H, int is_bf16, int num_sms, cudaStream_t stream, int reverse,
                                 bool run_tail) {
  if (n_img <= 0) return cudaSuccess;
  if (!attention_tcf_supported(L)) return cudaErrorInvalidValue;
  FlashParams p{};
  p.L = L;
  p.H = H;
  const int tail = attention_tcf_tail_rows(L);
  p.nq = tail ? L / 128 : (L + 127) / 128;
  const int units = (L + 31) / 32;            // 32-key units
  p.nb = (units * 32 + KB_MAX - 1) / KB_MAX;  // blocks of <= KB_MAX keys, balanced to within one unit
  p.ub = units / p.nb;
  p.ur = units % p.nb;
  p.total = n_img * H * p.nq;
  p.reverse = reverse;
  const int grid = p.total < num_sms ? p.total : num_sms;
  uint16_t* o = static_cast<uint16_t*>(out);
  cudaError_t e = is_bf16 ? launch_f<true>(tmap_q, tmap_kv32, o, p, grid, stream)
                          : launch_f<false>(tmap_q, tmap_kv32, o, p, grid, stream);
  if (e != cudaSuccess || tail == 0 || !run_tail) return e;
  return launch_attention_rows(qkv, out, n_img, L, H, is_bf16, p.nq * 128, stream);
}

}  // namespace aihab
