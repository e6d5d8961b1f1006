// Persistent, software-pipelined tcgen05 / TMEM attention for 64 < L <= 224 (ViT-B/16: L = 197, Lk = 208).
// One CTA per SM walks a static list of work items (image, head, 128-query tile):
//   warp 0 (lane 0)  TMA producer : Q [128x64], K [Lk x 64], V [Lk x 64] of item i+1 land in the other smem stage
//                                   while item i is in its softmax
//   warp 1 (converged) MMA issuer : S(i+1) = Q K^T is issued into the other TMEM S buffer BEFORE it waits for P(i), so
//                                   the softmax warps never wait for a QK^T; then O = P(i) V (V read as an MN-major
//                                   operand straight from its [key][d] tile)
//   warps 2..9       softmax      : two threads per query row; S is read from TMEM once (registers), row max exchanged
//                                   through smem, exp2 / row sum, and P is written back INTO TENSOR MEMORY as packed
//                                   16-bit pairs over the S columns (tcgen05.st) - the PV MMA takes A from TMEM, so P
//                                   never touches shared memory; the O(i-1) epilogue (TMEM -> 1/sum -> global) runs
//                                   before the hand-off, hiding the PV MMA of the previous item
// Short sequences (16 <= L <= 64, ViT-B/32: L = 50) share a tile TWO AT A TIME in 64-row SLOTS: image a of a pair owns
// query rows / keys [0, L), image b [64, 64 + L) (two 64-row TMA boxes per operand, the rows in between are padding), one
// S = Q K^T covers both and a block mask (query row r sees keys [64 (r / 64), 64 (r / 64) + L)) keeps them apart; P is
// zero outside the row's own slot, so O = P V over both slots is already the per-image result.  Because an image's keys
// always sit at the same offsets inside the 16-key MMA steps, whichever slot it lands in and whoever its neighbour is,
// its result is bit-identical for every batch composition (the contiguous packing of round 1 was not, and stayed opt-in).
// TMEM: S0/P0 [0,224) S1/P1 [224,448) O [448,512).  smem: NS x (Q 16 KB + K Lk x 128 B + V Lk x 128 B), NS = 2..4.
// Reference: clip/model.py:179-181 (nn.MultiheadAttention core: softmax(q k^T / sqrt(64)) v, no mask).
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace aihab {

namespace {

constexpr int P_THREADS = 320;
constexpr int KV_MAX = 224;
// shared memory: NS input stages of (Q 16 KB | K Lk x 128 B | V Lk x 128 B), sized by the launch for its Lk - three
// stages at Lk = 208, four for the packed short sequences (their per-item work is small, so the loads have to be
// requested further ahead) - then the barriers and the row max / sum exchange
constexpr int NS_MAX = 4;
constexpr int ST_Q = 0;
constexpr int ST_K = 16384;
__host__ __device__ constexpr int st_v(int Lk) { return ST_K + Lk * 128; }
__host__ __device__ constexpr int st_bytes(int Lk) { return ST_K + 2 * Lk * 128; }  // Lk = 224: 73728
constexpr int RED_BYTES = 4096;  // s_max[2][256], s_sum[2][256]
constexpr int BAR_BYTES = 256;
__host__ __device__ constexpr int smem_bytes(int Lk, int ns) { return ns * st_bytes(Lk) + BAR_BYTES + RED_BYTES + 1024; }
constexpr int SMEM_CAP = 227 * 1024;
constexpr int TM_S = 224;                      // TMEM columns per S buffer
constexpr int TM_O = 448;
static_assert(ST_K % 1024 == 0 && st_v(16) % 1024 == 0 && st_bytes(16) % 1024 == 0, "SWIZZLE_128B tiles need 1024 B alignment");
static_assert(smem_bytes(KV_MAX, 2) <= SMEM_CAP, "smem budget");

// Softmax of one query row half for one item, fully unrolled over NC 16-column chunks: the S values are read from
// TMEM ONCE (NC x tcgen05.ld.x16 in flight together) and stay in registers across the row-max exchange.
// Columns >= L are masked; only the last two chunks of a thread's range can contain such columns.
// MASK 1 (causal, text tower, clip/model.py:323-329): L is the row's own limit min(L, query index + 1), and any chunk
// can hold masked columns.  MASK 2 (block-diagonal, packed short sequences): the row sees columns [lo, L).
enum : int { MASK_NONE = 0, MASK_CAUSAL = 1, MASK_BLOCK = 2 };
template <bool BF16, int NC, int MASK>
__device__ __forceinline__ void softmax_item(uint32_t t_row, int c_begin, int lo, int L, float sl2, float* s_max_b,
                                             float* s_sum_b, int half, int row, bool has_rows) {
  constexpr bool CAUSAL = MASK != MASK_NONE;  // per-row limits: every chunk takes the masked path
  uint32_t r[NC][16];
  if (has_rows) {
#pragma unroll
    for (int i = 0; i < NC; ++i) ptx::tmem_ld_32x16(t_row + (c_begin + i) * 16, r[i]);
    ptx::tmem_ld_wait();
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (!CAUSAL && i < NC - 2) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(r[i][j]));
          m1 = fmaxf(m1, __uint_as_float(r[i][j + 1]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = (c_begin + i) * 16 + j;
          if (col < L && (MASK != MASK_BLOCK || col >= lo)) m0 = fmaxf(m0, __uint_as_float(r[i][j]));
        }
      }
    }
    s_max_b[half * 128 + row] = fmaxf(m0, m1);
  }
  // after this barrier every S column of the tile has been read into registers: P may overwrite S in place
  ptx::tc_fence_before();
  // only the two warps that share this lane quadrant (the two column halves of the same 32 rows) have to meet
  asm volatile("bar.sync %0, 64;" ::"r"(1 + (row >> 5)) : "memory");
  ptx::tc_fence_after();
  if (has_rows) {
    const float ms = fmaxf(s_max_b[row], s_max_b[128 + row]) * sl2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = c_begin + i;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p0 = ptx::ex2_approx(fmaf(__uint_as_float(r[i][2 * j]), sl2, -ms));
        float p1 = ptx::ex2_approx(fmaf(__uint_as_float(r[i][2 * j + 1]), sl2, -ms));
        if (CAUSAL || i >= NC - 2) {
          if (c * 16 + 2 * j >= L || (MASK == MASK_BLOCK && c * 16 + 2 * j < lo)) p0 = 0.f;
          if (c * 16 + 2 * j + 1 >= L || (MASK == MASK_BLOCK && c * 16 + 2 * j + 1 < lo)) p1 = 0.f;
        }
        l0 += p0;
        l1 += p1;
        pk[j] = ptx::pack2<BF16>(p0, p1);
      }
      ptx::tmem_st_32x8(t_row + c * 8, pk);  // 16 keys -> 8 packed columns of P, over the S buffer
    }
    s_sum_b[half * 128 + row] = l0 + l1;
    ptx::tmem_st_wait();
  }
}

// L = rows per sequence (one image, or g packed images of Lblk rows each with MASK_BLOCK); rows_total = valid rows
// of the qkv buffer (the last packed sequence may hold fewer than g images).
template <bool BF16, int NC, int MASK>
__global__ void __launch_bounds__(P_THREADS, 1)
attention_tcp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     uint16_t* __restrict__ out, int L, int H, int Lk, int nq, int total, int reverse, int Lblk,
                     int rows_total, int NS) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024 B aligned, still a __shared__ pointer
  const int ST_V = st_v(Lk), ST_BYTES = st_bytes(Lk);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * ST_BYTES);
  uint64_t* bar_qk = bars + 0;                   // [NS] Q + K of a stage landed
  uint64_t* bar_v = bars + NS_MAX;               // [NS] V of a stage landed
  uint64_t* bar_stfree = bars + 2 * NS_MAX;      // [NS] stage inputs consumed (commit after PV)
  uint64_t* bar_sfull = bars + 3 * NS_MAX;       // [2] S buffer written
  uint64_t* bar_pvdone = bars + 3 * NS_MAX + 2;  // [2] PV MMA of the item using S/P buffer b retired: the buffer is free
  uint64_t* bar_p = bars + 3 * NS_MAX + 4;       // P written to TMEM (and O of the previous item read)
  uint64_t* bar_o = bars + 3 * NS_MAX + 5;       // O written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NS_MAX + 6);
  float* s_max = reinterpret_cast<float*>(smem + NS * ST_BYTES + BAR_BYTES);  // [2][256]
  float* s_sum = s_max + 512;                               // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int D = H * 64;
  const int n_items = blockIdx.x < total ? (total - static_cast<int>(blockIdx.x) + G - 1) / G : 0;

  // item it of this CTA -> (image, head, query tile).  For two query tiles per head the tile parity alternates
  // with `it` so that every CTA sees the same mix of full (128-row) and partial tiles (G is even).
  auto decode = [&](int it, int& img, int& h, int& qt) {
    const int idx = static_cast<int>(blockIdx.x) + it * G;
    int u = idx;
    qt = 0;
    if (nq == 2) {
      u = idx >> 1;
      qt = (idx & 1) ^ (it & 1);
    }
    if (reverse) u = total / nq - 1 - u;  // walk (image, head) units from the end: the producer's freshest rows first
    img = u / H;
    h = u - img * H;
  };

  ptx::griddep_launch();
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_kv);
      for (int i = 0; i < NS; ++i) {
        ptx::mbar_init(&bar_qk[i], 1);
        ptx::mbar_init(&bar_v[i], 1);
        ptx::mbar_init(&bar_stfree[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&bar_sfull[i], 1);
        ptx::mbar_init(&bar_pvdone[i], 1);
      }
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
      for (int it = 0; it < n_items; ++it) {
        int img, h, qt;
        decode(it, img, h, qt);
        const int s = it % NS, k = it / NS;  // input stage and its use count
        uint8_t* st = smem + s * ST_BYTES;
        if (it >= NS) ptx::mbar_wait(&bar_stfree[s], (k - 1) & 1);
        ptx::mbar_expect_tx(&bar_qk[s], 128 * 128 + Lk * 128);
        if (MASK == MASK_BLOCK) {  // two 64-row slots: images 2 * img and 2 * img + 1 (tmap_kv has a 64-row box)
          const int ra = 2 * img * Lblk, rb = ra + Lblk;
          ptx::tma_load_2d(st + ST_Q, &tmap_kv, &bar_qk[s], h * 64, ra);
          ptx::tma_load_2d(st + ST_Q + 8192, &tmap_kv, &bar_qk[s], h * 64, rb);
          ptx::tma_load_2d(st + ST_K, &tmap_kv, &bar_qk[s], D + h * 64, ra);
          ptx::tma_load_2d(st + ST_K + 8192, &tmap_kv, &bar_qk[s], D + h * 64, rb);
          ptx::mbar_expect_tx(&bar_v[s], Lk * 128);
          ptx::tma_load_2d(st + ST_V, &tmap_kv, &bar_v[s], 2 * D + h * 64, ra);
          ptx::tma_load_2d(st + ST_V + 8192, &tmap_kv, &bar_v[s], 2 * D + h * 64, rb);
          continue;
        }
        const int row0 = img * L;
        ptx::tma_load_2d(st + ST_Q, &tmap_q, &bar_qk[s], h * 64, row0 + qt * 128);
        ptx::tma_load_2d(st + ST_K, &tmap_kv, &bar_qk[s], D + h * 64, row0);
        ptx::mbar_expect_tx(&bar_v[s], Lk * 128);
        ptx::tma_load_2d(st + ST_V, &tmap_kv, &bar_v[s], 2 * D + h * 64, row0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {  // the whole warp runs the loop (converged); one elected lane issues each tcgen05 instruction (ptx.cuh "_w")
      const uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, Lk);
      const uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1 : 0, 128, 64, /*b_mn_major=*/1);
      const int ksteps = Lk >> 4;
      auto issue_s = [&](int it) {
        const int s = it % NS, b = it & 1;  // input stage; S/P buffer in TMEM
        ptx::mbar_wait(&bar_qk[s], (it / NS) & 1);
        if (it >= 2) ptx::mbar_wait(&bar_pvdone[b], ((it >> 1) - 1) & 1);  // P(it-2) lives in this buffer until PV(it-2) retires
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + s * ST_BYTES);
        const uint64_t qd = ptx::make_kmajor_sw128_desc(st + ST_Q);
        const uint64_t kd = ptx::make_kmajor_sw128_desc(st + ST_K);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_w(tmem + b * TM_S, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
        ptx::umma_commit_w(&bar_sfull[b]);
      };
      if (n_items > 0) issue_s(0);
      for (int it = 0; it < n_items; ++it) {
        if (it + 1 < n_items) issue_s(it + 1);
        const int s = it % NS, b = it & 1;
        ptx::mbar_wait(bar_p, it & 1);
        ptx::mbar_wait(&bar_v[s], (it / NS) & 1);
        ptx::tc_fence_after();
        const uint32_t v_base = ptx::smem_u32(smem + s * ST_BYTES + ST_V);
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t vd = ptx::make_mnmajor_sw128_desc(v_base + j * 2048);
          ptx::umma_f16_ts_w(tmem + TM_O, tmem + b * TM_S + j * 8, vd, idesc_o, j != 0);  // A = P from TMEM
        }
        ptx::umma_commit_w(bar_o);
        ptx::umma_commit_w(&bar_stfree[s]);
        ptx::umma_commit_w(&bar_pvdone[b]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 2..9)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;
    // NC = ceil(Lk / 32) 16-column chunks per thread (3..7); a chunk past Lk is fully masked
    const int c_begin = half * NC;

    // O(prev) -> global: this thread owns 32 of the 64 output columns of its row
    auto epilogue = [&](int img, int h, int qt, int b) {
      if (qt * 128 + quad * 32 >= L) return;
      uint32_t o[32];
      ptx::tmem_ld_32x32(tmem + lane_off + TM_O + half * 32, o);
      ptx::tmem_ld_wait();
      const int grow = qt * 128 + row;
      // MASK_BLOCK: row -> (slot, token); image 2 * img + slot exists if it is below rows_total (= number of images)
      const int image = MASK == MASK_BLOCK ? 2 * img + (row >> 6) : img;
      const int tok = MASK == MASK_BLOCK ? (row & 63) : grow;
      const int seq_len = MASK == MASK_BLOCK ? Lblk : L;
      if (tok < seq_len && (MASK != MASK_BLOCK || image < rows_total)) {
        const float inv_l = 1.0f / (s_sum[b * 256 + row] + s_sum[b * 256 + 128 + row]);
        uint16_t* dst = out + (static_cast<size_t>(image) * seq_len + tok) * D + h * 64 + half * 32;
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

    int p_img = 0, p_h = 0, p_qt = 0;
    for (int it = 0; it < n_items; ++it) {
      int img, h, qt;
      decode(it, img, h, qt);
      const int b = it & 1, k = it >> 1;
      const bool has_rows = qt * 128 + quad * 32 < L;
      const uint32_t t_row = tmem + lane_off + b * TM_S;

      ptx::mbar_wait(&bar_sfull[b], k & 1);
      ptx::tc_fence_after();
      // valid key columns [lo, lim) of this thread's row
      int lo = 0, lim = L;
      if (MASK == MASK_CAUSAL) lim = min(L, qt * 128 + row + 1);
      if (MASK == MASK_BLOCK) {  // the row's own 64-key slot
        lo = (row >> 6) << 6;
        lim = lo + Lblk;
      }
      softmax_item<BF16, NC, MASK>(t_row, c_begin, lo, lim, sl2, s_max + b * 256, s_sum + b * 256, half, row, has_rows);
      if (it > 0) {  // O(it-1): its PV was issued a whole softmax ago
        ptx::mbar_wait(bar_o, (it - 1) & 1);
        ptx::tc_fence_after();
        epilogue(p_img, p_h, p_qt, b ^ 1);
      }
      ptx::tc_fence_before();                    // P(it) stored (wait::st) and O(it-1) read (wait::ld) ...
      ptx::mbar_arrive(bar_p);                   // ... before PV(it) may read P and overwrite O
      p_img = img;
      p_h = h;
      p_qt = qt;
    }
    if (n_items > 0) {
      ptx::mbar_wait(bar_o, (n_items - 1) & 1);
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

template <bool BF16, int NC, int MASK>
cudaError_t launch_nc(const CUtensorMap& tq, const CUtensorMap& tkv, uint16_t* out, int L, int H, int Lk, int nq,
                      int total, int grid, int reverse, int Lblk, int rows_total, cudaStream_t stream) {
  static bool attr_set[64] = {};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcp_kernel<BF16, NC, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  // as many input stages as fit (2..NS_MAX); AIHAB_ATTN_STAGES caps them for A/B runs
  int ns = NS_MAX;
  if (const char* e = getenv("AIHAB_ATTN_STAGES")) ns = atoi(e) < 2 ? 2 : (atoi(e) > NS_MAX ? NS_MAX : atoi(e));
  while (ns > 2 && smem_bytes(Lk, ns) > SMEM_CAP) --ns;
  return launch_kernel(attention_tcp_kernel<BF16, NC, MASK>, grid, P_THREADS, smem_bytes(Lk, ns), stream, 1, true, tq, tkv, out, L, H,
                       Lk, nq, total, reverse, Lblk, rows_total, ns);
}

template <bool BF16, int MASK>
cudaError_t launch_dt(int nc, const CUtensorMap& tq, const CUtensorMap& tkv, uint16_t* out, int L, int H, int Lk,
                      int nq, int total, int grid, int reverse, int Lblk, int rows_total, cudaStream_t stream) {
  switch (nc) {
    case 3: return launch_nc<BF16, 3, MASK>(tq, tkv, out, L, H, Lk, nq, total, grid, reverse, Lblk, rows_total, stream);
    case 4: return launch_nc<BF16, 4, MASK>(tq, tkv, out, L, H, Lk, nq, total, grid, reverse, Lblk, rows_total, stream);
    case 5: return launch_nc<BF16, 5, MASK>(tq, tkv, out, L, H, Lk, nq, total, grid, reverse, Lblk, rows_total, stream);
    case 6: return launch_nc<BF16, 6, MASK>(tq, tkv, out, L, H, Lk, nq, total, grid, reverse, Lblk, rows_total, stream);
    case 7: return launch_nc<BF16, 7, MASK>(tq, tkv, out, L, H, Lk, nq, total, grid, reverse, Lblk, rows_total, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

// images per tile: 2 (64-row slots) for 16 <= L <= 64, else 1.  AIHAB_ATTN_PACK=0 sends L <= 64 back to the mma.sync
// kernel (A/B runs).
int attention_tcp_pack(int L) {
  static const bool enabled = [] {  // read once: tensor maps built at create and launches must agree
    const char* e = getenv("AIHAB_ATTN_PACK");
    return !(e != nullptr && e[0] == '0');
  }();
  return (L > 64 || !enabled) ? 1 : 2;
}
bool attention_tcp_supported(int L) {
  const int g = attention_tcp_pack(L);
  return L >= 16 && (L > 64 || g > 1) && (g > 1 || (L + 15) / 16 * 16 <= KV_MAX);
}
int attention_tcp_key_rows(int L) { return attention_tcp_pack(L) > 1 ? 64 : (L + 15) / 16 * 16; }

cudaError_t launch_attention_tcp(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, void* out, int n_img, int L,
                                 int H, int is_bf16, int num_sms, cudaStream_t stream, int reverse, int causal,
                                 const CUtensorMap* tmap_out3) {
  if (n_img <= 0) return cudaSuccess;
  if (!attention_tcp_supported(L)) return cudaErrorInvalidValue;
  const int g = attention_tcp_pack(L);
  if (g > 1 && causal) return cudaErrorInvalidValue;  // the text tower has L = 77
  if (g == 1 && !causal && tmap_out3 != nullptr && attention_tcd_supported(L))  // two query tiles per unit: dual-stream kernel
    return launch_attention_tcd(tmap_q, tmap_kv, *tmap_out3, n_img, L, H, is_bf16, num_sms, stream, reverse);
  const int Ls = g > 1 ? 128 : L;             // rows per tile sequence (two 64-row slots when packed)
  const int n_seq = (n_img + g - 1) / g;
  const int Lk = (Ls + 15) / 16 * 16;
  const int nq = (Ls + 127) / 128;
  const int total = n_seq * H * nq;
  int grid = total < num_sms ? total : num_sms;
  if (nq == 2) grid &= ~1;  // even: a CTA's items alternate between the full and the partial query tile
  const int nc = ((Lk >> 4) + 1) >> 1;
  const int rows_total = g > 1 ? n_img : n_img * L;  // packed: the number of images (the last pair may be half empty)
  uint16_t* o = static_cast<uint16_t*>(out);
  if (g > 1)
    return is_bf16 ? launch_dt<true, MASK_BLOCK>(nc, tmap_q, tmap_kv, o, Ls, H, Lk, nq, total, grid, reverse, L, rows_total, stream)
                   : launch_dt<false, MASK_BLOCK>(nc, tmap_q, tmap_kv, o, Ls, H, Lk, nq, total, grid, reverse, L, rows_total, stream);
  if (causal)
    return is_bf16 ? launch_dt<true, MASK_CAUSAL>(nc, tmap_q, tmap_kv, o, L, H, Lk, nq, total, grid, reverse, L, rows_total, stream)
                   : launch_dt<false, MASK_CAUSAL>(nc, tmap_q, tmap_kv, o, L, H, Lk, nq, total, grid, reverse, L, rows_total, stream);
  return is_bf16 ? launch_dt<true, MASK_NONE>(nc, tmap_q, tmap_kv, o, L, H, Lk, nq, total, grid, reverse, L, rows_total, stream)
                 : launch_dt<false, MASK_NONE>(nc, tmap_q, tmap_kv, o, L, H, Lk, nq, total, grid, reverse, L, rows_total, stream);
}

}  // namespace aihab
