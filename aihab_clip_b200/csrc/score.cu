// Zero-shot scoring epilogue in exact fp32 (CUDA-core FMA): visual projection, L2 normalisation, x100 cosine
// logits and top-k.   Reference: methods/ProLIP.py:40 (x @ vit_proj), methods/utils.py:183-186
// (F.normalize -> 100. * f @ text_weights -> argmax), methods/utils.py:16-21 / aihab_utils/evaluation.py:261-273
// (topk, sorted, lowest index first among exact ties).
#include "kernels.cuh"
#include "ptx.cuh"

#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>
#include <stdlib.h>

namespace aihab {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// C = (alpha * A) @ B.  256 threads, 4x4 micro-tile per thread, k accumulated in ascending order.
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ C, int M, int N, int K, float alpha) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i - r * TK;
      const int gm = m0 + r, gk = k0 + c;
      As[c][r] = (gm < M && gk < K) ? alpha * A[static_cast<size_t>(gm) * K + gk] : 0.f;
    }
    for (int i = threadIdx.x; i < TK * TN; i += 256) {
      const int r = i / TN, c = i - r * TN;
      const int gk = k0 + r, gn = n0 + c;
      Bs[r][c] = (gk < K && gn < N) ? B[static_cast<size_t>(gk) * N + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[static_cast<size_t>(gm) * N + gn] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(128) l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int rows,
                                                     int cols, float eps) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = x + static_cast<size_t>(row) * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(src[c], src[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float denom = fmaxf(sqrtf(s), eps);
  float* dst = y + static_cast<size_t>(row) * cols;
  for (int c = lane; c < cols; c += 32) dst[c] = src[c] / denom;
}

// (value desc, index asc) ordering; `better(a, b)` = a precedes b.
__device__ __forceinline__ bool precedes(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(128) topk_kernel(const float* __restrict__ logits, int rows, int cols, int k,
                                                   int64_t* __restrict__ idx, float* __restrict__ val) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = logits + static_cast<size_t>(row) * cols;
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
      const float v = src[c];
      // candidates strictly after the previously selected element in the ordering
      if (precedes(last_v, last_i, v, c) && precedes(v, c, bv, bi)) {
        bv = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (precedes(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      idx[static_cast<size_t>(row) * k + j] = bi;
      if (val != nullptr) val[static_cast<size_t>(row) * k + j] = bv;
    }
    last_v = bv;
    last_i = bi;
  }
}

// Same selection with the row read ONCE (128-bit loads) and held in registers: 4 * NV values per lane (cols <= 128 * NV,
// cols % 4 == 0).  The k rounds compare registers only; the row's bytes cross HBM / L2 exactly once.
template <int NV>
__global__ void __launch_bounds__(128) topk_regs_kernel(const float* __restrict__ logits, int rows, int cols, int k,
                                                        int64_t* __restrict__ idx, float* __restrict__ val) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(logits + static_cast<size_t>(row) * cols);
  const int nvec = cols >> 2;
  float v[NV * 4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int q = i * 32 + lane;
    float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (q < nvec) t = __ldg(src + q);
    v[4 * i] = t.x;
    v[4 * i + 1] = t.y;
    v[4 * i + 2] = t.z;
    v[4 * i + 3] = t.w;
  }
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) {
      const int c = ((i >> 2) * 32 + lane) * 4 + (i & 3);
      if (c < cols && precedes(last_v, last_i, v[i], c) && precedes(v[i], c, bv, bi)) {
        bv = v[i];
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (precedes(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      idx[static_cast<size_t>(row) * k + j] = bi;
      if (val != nullptr) val[static_cast<size_t>(row) * k + j] = bv;
    }
    last_v = bv;
    last_i = bi;
  }
}

// Final selection over the candidates the EPI_TOPK_32 GEMM epilogue left per row: [rows, slots, 8] values and column
// indices (unused entries are -inf / INT_MAX).  One thread per row, k rounds in (value desc, index asc) order.
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                         int rows, int slots, int k, int64_t* __restrict__ idx,
                                                         float* __restrict__ val) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float4* cv = reinterpret_cast<const float4*>(cand_val + static_cast<size_t>(row) * slots * 8);
  const int4* ci = reinterpret_cast<const int4*>(cand_idx + static_cast<size_t>(row) * slots * 8);
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int q = 0; q < slots * 2; ++q) {
      const float4 v = __ldg(cv + q);
      const int4 c = __ldg(ci + q);
      if (precedes(last_v, last_i, v.x, c.x) && precedes(v.x, c.x, bv, bi)) { bv = v.x; bi = c.x; }
      if (precedes(last_v, last_i, v.y, c.y) && precedes(v.y, c.y, bv, bi)) { bv = v.y; bi = c.y; }
      if (precedes(last_v, last_i, v.z, c.z) && precedes(v.z, c.z, bv, bi)) { bv = v.z; bi = c.z; }
      if (precedes(last_v, last_i, v.w, c.w) && precedes(v.w, c.w, bv, bi)) { bv = v.w; bi = c.w; }
    }
    idx[static_cast<size_t>(row) * k + j] = bi;
    if (val != nullptr) val[static_cast<size_t>(row) * k + j] = bv;
    last_v = bv;
    last_i = bi;
  }
}

// The same selection for slots <= 8 with EIGHT lanes per row: lane l holds slot group l (8 candidates, already sorted by
// the epilogue in (value desc, index asc) order), every round is a 3-step butterfly over the eight list heads and the
// winning lane pops its head.  Loads are 128 contiguous bytes per row and array (the one-thread-per-row kernel reads
// each row k times with a 32 B-per-lane stride): 29 -> 8 us per 32 k rows.
__global__ void __launch_bounds__(256) topk_merge8_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                          int rows, int slots, int k, int64_t* __restrict__ idx,
                                                          float* __restrict__ val) {
  const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = t >> 3;
  const int sub = static_cast<int>(t & 7);
  const bool active = row < rows;
  float v[8];
  int c[8];
  if (active && sub < slots) {
    const float4* cv = reinterpret_cast<const float4*>(cand_val + (static_cast<size_t>(row) * slots + sub) * 8);
    const int4* ci = reinterpret_cast<const int4*>(cand_idx + (static_cast<size_t>(row) * slots + sub) * 8);
    const float4 v0 = __ldg(cv), v1 = __ldg(cv + 1);
    const int4 c0 = __ldg(ci), c1 = __ldg(ci + 1);
    v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = -INFINITY;
      c[i] = 0x7fffffff;
    }
  }
  for (int j = 0; j < k; ++j) {
    float bv = v[0];
    int bi = c[0], bl = sub;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off, 8);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off, 8);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, off, 8);
      // strict total order (value desc, index asc, lane asc): all eight lanes agree on the winner
      if (precedes(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) {
        bv = ov;
        bi = oi;
        bl = ol;
      }
    }
    const bool pop = sub == bl;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      v[i] = pop ? v[i + 1] : v[i];
      c[i] = pop ? c[i + 1] : c[i];
    }
    v[7] = pop ? -INFINITY : v[7];
    c[7] = pop ? 0x7fffffff : c[7];
    if (active && sub == 0) {
      idx[static_cast<size_t>(row) * k + j] = bi;
      if (val != nullptr) val[static_cast<size_t>(row) * k + j] = bv;
    }
  }
}

// Small-batch path (one extraction batch): proj -> L2 normalise -> scaled logits -> top-k in ONE launch.
// RPB rows per CTA share every visual.proj / text-weight element they load.  Thread layout for the projection:
// 128 column quads (float4 loads of one contiguous proj row per k) x 2 halves of K, 8 loads in flight per thread;
// per-row reduction order is fixed, so results do not depend on batch composition.
constexpr int RPB = 4;

__global__ void __launch_bounds__(256) score_fused_kernel(const float* __restrict__ feats, int n, int D,
                                                          const float* __restrict__ proj, int E,
                                                          const float* __restrict__ text_w, int C, float scale, int k,
                                                          float* __restrict__ emb_out, float* __restrict__ logits_out,
                                                          int64_t* __restrict__ topk_idx, float* __restrict__ topk_val) {
  extern __shared__ __align__(16) float sm[];
  float* sf = sm;                 // [RPB][D]
  float* se = sf + RPB * D;       // [2][RPB][E] partial sums, then [RPB][E] embeddings in the first half
  float* sl = se + 2 * RPB * E;   // [RPB][C]
  __shared__ float s_den[RPB];
  const int row0 = blockIdx.x * RPB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Warm L2 first: in the extraction step visual.proj and the text matrix were evicted by the ~26 GB the tower moved since
  // the last call, and all CTAs walk them in lockstep through dependent 8-load batches - a cold miss would be paid ~50
  // times in a row by every CTA.  One 128-byte line per thread across the grid instead: one HBM round trip in total.
  {
    const size_t gtid = static_cast<size_t>(blockIdx.x) * 256 + tid, gstride = static_cast<size_t>(gridDim.x) * 256;
    if (proj != nullptr) {
      const size_t lines = (static_cast<size_t>(D) * E * sizeof(float) + 127) / 128;
      for (size_t i = gtid; i < lines; i += gstride)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(proj) + i * 128));
    }
    if (text_w != nullptr) {
      const size_t lines = (static_cast<size_t>(E) * C * sizeof(float) + 127) / 128;
      for (size_t i = gtid; i < lines; i += gstride)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(text_w) + i * 128));
    }
  }
  for (int i = tid; i < RPB * D; i += 256) {
    const int r = i / D;
    sf[i] = (row0 + r < n) ? feats[static_cast<size_t>(row0 + r) * D + (i - r * D)] : 0.f;
  }
  __syncthreads();
  if (proj != nullptr) {
    const int half = tid >> 7, q = tid & 127;
    const int k0 = half * (D / 2), k1 = half ? D : D / 2;
    for (int e0 = q * 4; e0 < E; e0 += 512) {
      float4 acc[RPB];
#pragma unroll
      for (int r = 0; r < RPB; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      int kk = k0;
      // PU weight rows in flight per thread: the CTA is one dependent chain of D / 2 / PU round trips to L2
      constexpr int PU = 16;
      for (; kk + PU <= k1; kk += PU) {
        float4 w[PU];
#pragma unroll
        for (int j = 0; j < PU; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(proj + static_cast<size_t>(kk + j) * E + e0));
#pragma unroll
        for (int j = 0; j < PU; ++j)
#pragma unroll
          for (int r = 0; r < RPB; ++r) {
            const float f = sf[r * D + kk + j];
            acc[r].x = fmaf(f, w[j].x, acc[r].x);
            acc[r].y = fmaf(f, w[j].y, acc[r].y);
            acc[r].z = fmaf(f, w[j].z, acc[r].z);
            acc[r].w = fmaf(f, w[j].w, acc[r].w);
          }
      }
      for (; kk < k1; ++kk) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(proj + static_cast<size_t>(kk) * E + e0));
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
          const float f = sf[r * D + kk];
          acc[r].x = fmaf(f, w.x, acc[r].x);
          acc[r].y = fmaf(f, w.y, acc[r].y);
          acc[r].z = fmaf(f, w.z, acc[r].z);
          acc[r].w = fmaf(f, w.w, acc[r].w);
        }
      }
#pragma unroll
      for (int r = 0; r < RPB; ++r) *reinterpret_cast<float4*>(se + (half * RPB + r) * E + e0) = acc[r];
    }
    __syncthreads();
    for (int i = tid; i < RPB * E; i += 256) se[i] += se[RPB * E + i];
  } else {
    for (int i = tid; i < RPB * E; i += 256) se[i] = sf[(i / E) * D + (i % E)];
  }
  __syncthreads();
  if (warp < RPB) {  // F.normalize: x / max(||x||, 1e-12)
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s = fmaf(se[warp * E + c], se[warp * E + c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_den[warp] = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  for (int i = tid; i < RPB * E; i += 256) {
    const int r = i / E;
    const float v = se[i] / s_den[r];
    se[i] = v;
    if (emb_out != nullptr && row0 + r < n) emb_out[static_cast<size_t>(row0 + r) * E + (i - r * E)] = v;
  }
  __syncthreads();
  if (text_w == nullptr) return;
  if (C <= 64) {
    // few classes: one warp per class, lanes split E, shuffle-tree reduction
    for (int c = warp; c < C; c += 8) {
      float acc[RPB] = {};
      for (int e = lane; e < E; e += 32) {
        const float w = __ldg(text_w + static_cast<size_t>(e) * C + c);
#pragma unroll
        for (int r = 0; r < RPB; ++r) acc[r] = fmaf(scale * se[r * E + e], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RPB; ++r) sl[r * C + c] = acc[r];
      }
    }
  } else {
    // many classes: one thread per class, coalesced reads of the [E, C] text matrix
    for (int c = tid; c < C; c += 256) {
      float acc[RPB] = {};
      for (int e = 0; e < E; ++e) {
        const float w = __ldg(text_w + static_cast<size_t>(e) * C + c);
#pragma unroll
        for (int r = 0; r < RPB; ++r) acc[r] = fmaf(scale * se[r * E + e], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < RPB; ++r) sl[r * C + c] = acc[r];
    }
  }
  __syncthreads();
  if (logits_out != nullptr) {
    for (int i = tid; i < RPB * C; i += 256) {
      const int r = i / C;
      if (row0 + r < n) logits_out[static_cast<size_t>(row0 + r) * C + (i - r * C)] = sl[i];
    }
  }
  if (k > 0 && warp < RPB && row0 + warp < n) {
    const float* src = sl + warp * C;
    float last_v = INFINITY;
    int last_i = -1;
    for (int j = 0; j < k; ++j) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        const float v = src[c];
        if (precedes(last_v, last_i, v, c) && precedes(v, c, bv, bi)) {
          bv = v;
          bi = c;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (precedes(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        topk_idx[static_cast<size_t>(row0 + warp) * k + j] = bi;
        if (topk_val != nullptr) topk_val[static_cast<size_t>(row0 + warp) * k + j] = bv;
      }
      last_v = bv;
      last_i = bi;
    }
  }
}

// ---- mid-size batches (one extraction batch of 65..8192 rows, e.g. the 256 rows of the headline step) ---------------
// score_fused_kernel gives every CTA whole rows, so each of its n / 4 CTAs pulls ALL of visual.proj and the text matrix
// (3.5 MB at 768 -> 512 -> 1000) through one SM's L2 port: 82-100 us for 256 rows whatever the row count.  These two
// kernels slice the COLUMNS as well (16 rows x 32 embedding columns, 16 rows x 32 classes per CTA: 256 / 512 CTAs
// at n = 256, 1000 classes), so an SM loads a 1/16 slice of the weights.  Per output the k loop is one ascending fmaf chain - the
// order of sgemm_kernel - and the normalisation is l2norm_kernel's: results are bit-identical to the chunked path
// (aihab_score for n > 8192) and do not depend on the batch composition.
constexpr int MID_ROWS = 16;
constexpr int MID_KC = 64;   // weight rows per shared-memory chunk (64 x 32 fp32 = 8 KB)
constexpr int MID_NST = 4;   // chunk ring: MID_NST - 1 chunks in flight (the weights come from HBM once per step)

// One thread's 4 columns of  sum over k (ascending, ONE fmaf chain per output) of a_row[k] * B[k][c0 + 4q .. + 3]  for the
// CTA's 32-column slice of the row-major [K, ldb] matrix B.  The slice streams through a ring of MID_KC-row chunks in
// shared memory (16-byte cp.async pieces, MID_NST - 1 chunks in flight), so no global-load latency sits inside the
// dependent fmaf chain.  slice_issue / slice_dot are called by all 128 threads; one commit group per chunk index
// (empty past the end) keeps the wait_group arithmetic uniform.
struct Slice {
  const float* B;
  int K, ldb, ncols, c0;
  float* sw;  // [MID_NST][MID_KC][32]
};
__device__ __forceinline__ void slice_issue(const Slice& sl, int ch, int tid) {
  if (ch * MID_KC < sl.K) {
    float* dst = sl.sw + (ch % MID_NST) * MID_KC * 32;
    for (int i = tid; i < MID_KC * 8; i += 128) {
      const int kk = i >> 3, qq = i & 7;
      const int k = ch * MID_KC + kk, c = sl.c0 + 4 * qq;
      const bool ok = k < sl.K && c < sl.ncols;
      ptx::cp_async16(dst + kk * 32 + 4 * qq, ok ? sl.B + static_cast<size_t>(k) * sl.ldb + c : sl.B, ok);
    }
  }
  ptx::cp_async_commit();
}
// chunks 0 .. MID_NST-2 must have been issued (in order, nothing issued after them)
__device__ __forceinline__ float4 slice_dot(const Slice& sl, const float* a_row, int tid) {
  const int q = tid & 7;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nch = (sl.K + MID_KC - 1) / MID_KC;
  for (int ch = 0; ch < nch; ++ch) {
    slice_issue(sl, ch + MID_NST - 1, tid);  // into the slot consumed at iteration ch - 1 (all threads passed its barrier)
    ptx::cp_async_wait<MID_NST - 1>();
    __syncthreads();
    const float* w = sl.sw + (ch % MID_NST) * MID_KC * 32 + 4 * q;
    const float* a = a_row + ch * MID_KC;
    const int kmax = min(MID_KC, sl.K - ch * MID_KC);
    int kk = 0;
    for (; kk + 8 <= kmax; kk += 8) {
      // all shared-memory loads of 8 steps first (2 + 8 LDS.128), then the 32 fmaf: with two warps per scheduler
      // nothing else hides the load-to-use latency inside the chain
      const float4 f0 = *reinterpret_cast<const float4*>(a + kk), f1 = *reinterpret_cast<const float4*>(a + kk + 4);
      const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
      float4 wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = *reinterpret_cast<const float4*>(w + (kk + j) * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc.x = fmaf(fv[j], wv[j].x, acc.x);
        acc.y = fmaf(fv[j], wv[j].y, acc.y);
        acc.z = fmaf(fv[j], wv[j].z, acc.z);
        acc.w = fmaf(fv[j], wv[j].w, acc.w);
      }
    }
    for (; kk < kmax; ++kk) {
      const float f = a[kk];
      const float4 w4 = *reinterpret_cast<const float4*>(w + kk * 32);
      acc.x = fmaf(f, w4.x, acc.x);
      acc.y = fmaf(f, w4.y, acc.y);
      acc.z = fmaf(f, w4.z, acc.z);
      acc.w = fmaf(f, w4.w, acc.w);
    }
    __syncthreads();
  }
  return acc;
}

// rows [row0, row0 + MID_ROWS) of a row-major [n, cols] fp32 matrix into shared memory (zero rows past n); cols % 4 == 0.
// One commit group; the caller waits.
__device__ __forceinline__ void rows_issue(float* dst, const float* __restrict__ src, int n, int cols, int row0, int tid) {
  const int per_row = cols >> 2;
  for (int i = tid; i < MID_ROWS * per_row; i += 128) {
    const int r = i / per_row, c = (i - r * per_row) * 4;
    const bool ok = row0 + r < n;
    ptx::cp_async16(dst + r * cols + c, ok ? src + static_cast<size_t>(row0 + r) * cols + c : src, ok);
  }
  ptx::cp_async_commit();
}

__global__ void __launch_bounds__(128) score_proj_kernel(const float* __restrict__ feats, int n, int D,
                                                         const float* __restrict__ proj, int E, float* __restrict__ emb_raw) {
  extern __shared__ __align__(16) float sm[];  // [MID_ROWS][D] features | [MID_NST][MID_KC][32] weight chunks
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int row0 = blockIdx.y * MID_ROWS, c0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const Slice sl{proj, D, E, E, c0, sm + MID_ROWS * D};
  rows_issue(sm, feats, n, D, row0, tid);
  for (int ch = 0; ch < MID_NST - 1; ++ch) slice_issue(sl, ch, tid);
  const int r = tid >> 3, col = c0 + 4 * (tid & 7);
  const float4 acc = slice_dot(sl, sm + r * D, tid);  // its first wait_group also covers the (older) row group
  if (row0 + r < n) *reinterpret_cast<float4*>(emb_raw + static_cast<size_t>(row0 + r) * E + col) = acc;
}

__global__ void __launch_bounds__(128) score_logits_kernel(const float* __restrict__ emb_raw, int n, int E,
                                                           const float* __restrict__ text_w, int C, float scale,
                                                           float* __restrict__ emb_out, float* __restrict__ logits) {
  extern __shared__ __align__(16) float sm[];  // [MID_ROWS][E] embeddings | [MID_NST][MID_KC][32] weight chunks
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int row0 = blockIdx.y * MID_ROWS, c0 = blockIdx.x * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Slice sl{text_w, E, C, C, c0, sm + MID_ROWS * E};
  rows_issue(sm, emb_raw, n, E, row0, tid);
  for (int ch = 0; ch < MID_NST - 1; ++ch) slice_issue(sl, ch, tid);  // text-weight chunks fly during the normalisation
  ptx::cp_async_wait<MID_NST - 1>();
  __syncthreads();
  // F.normalize (x / max(||x||, 1e-12)) with the summation order of l2norm_kernel; every column-slice CTA of a row
  // group runs the same instructions on the same data, so they all hold the same normalised rows.  The rows stay in
  // shared memory multiplied by `scale`: (100. * f) @ w puts the scale onto the features first (methods/utils.py:185)
  for (int rr = 0; rr < MID_ROWS / 4; ++rr) {
    float* row = sm + (warp * (MID_ROWS / 4) + rr) * E;
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s = fmaf(row[c], row[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float denom = fmaxf(sqrtf(s), 1e-12f);
    const int grow = row0 + warp * (MID_ROWS / 4) + rr;
    for (int c = lane; c < E; c += 32) {
      const float v = row[c] / denom;
      row[c] = scale * v;
      if (emb_out != nullptr && blockIdx.x == 0 && grow < n) emb_out[static_cast<size_t>(grow) * E + c] = v;
    }
  }
  __syncthreads();
  const int r = tid >> 3, col = c0 + 4 * (tid & 7);
  const float4 acc = slice_dot(sl, sm + r * E, tid);
  if (row0 + r < n && col < C) *reinterpret_cast<float4*>(logits + static_cast<size_t>(row0 + r) * C + col) = acc;
}

// ---- tensor-core scoring helpers (aihab_score16) -------------------------------------------------------------
// 16-bit transpose: src [R, Cc] -> dst [Cc, R]   (visual.proj [D, E] -> [E, D], the K-major UMMA operand B)
__global__ void transpose16_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int R, int Cc) {
  __shared__ uint16_t t[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Cc) ? src[static_cast<size_t>(r) * Cc + c] : 0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) dst[static_cast<size_t>(c) * R + r] = t[threadIdx.x][i];
  }
}

__device__ __forceinline__ void split_hi_lo(float v, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}

// text weights fp32 [E, C] -> [C, 3E] fp16 rows (w_hi | w_lo | w_hi): the B operand matching A' = (e_hi | e_hi | e_lo)
__global__ void split_textw_kernel(const float* __restrict__ w, int E, int C, uint16_t* __restrict__ out) {
  const long total = static_cast<long>(E) * C;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<long>(e) * C);
    uint16_t hi, lo;
    split_hi_lo(w[i], hi, lo);
    uint16_t* row = out + static_cast<size_t>(c) * 3 * E;
    row[e] = hi;
    row[E + e] = lo;
    row[2 * E + e] = hi;
  }
}

// rows of emb fp32 [rows, E]: L2-normalise (F.normalize eps; skipped when normalize == 0: the rows are used as they
// are) -> optional fp32 copy + A' [rows, 3E] fp16 (hi | hi | lo)
__global__ void __launch_bounds__(128) l2norm_split_kernel(const float* __restrict__ emb, float* __restrict__ emb_out,
                                                           uint16_t* __restrict__ a3, int rows, int E, int normalize) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = emb + static_cast<size_t>(row) * E;
  float denom = 1.0f;
  if (normalize) {
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s = fmaf(src[c], src[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    denom = fmaxf(sqrtf(s), 1e-12f);
  }
  uint16_t* dst = a3 + static_cast<size_t>(row) * 3 * E;
  for (int c = lane; c < E; c += 32) {
    const float v = normalize ? src[c] / denom : src[c];
    if (emb_out != nullptr) emb_out[static_cast<size_t>(row) * E + c] = v;
    uint16_t hi, lo;
    split_hi_lo(v, hi, lo);
    dst[c] = hi;
    dst[E + c] = hi;
    dst[2 * E + c] = lo;
  }
}

// Same with the row in registers (128-bit loads, E % 4 == 0, E <= 128 * NV) and 8-byte stores of the three fp16 parts.
template <int NV>
__global__ void __launch_bounds__(128) l2norm_split_vec_kernel(const float* __restrict__ emb, float* __restrict__ emb_out,
                                                               uint16_t* __restrict__ a3, int rows, int E, int normalize) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(emb + static_cast<size_t>(row) * E);
  const int nvec = E >> 2;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int q = i * 32 + lane;
    v[i] = q < nvec ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float denom = 1.0f;
  if (normalize) {
    // same summation order as l2norm_split_kernel: lane-strided scalar order c = lane, lane + 32, ... is NOT what a
    // float4 layout gives, so the sum is accumulated per lane over ITS elements in column order and reduced by the same
    // butterfly; the result differs from the scalar kernel only in the last bits of the norm (both are valid
    // F.normalize evaluations; tests compare with tolerances)
#pragma unroll
    for (int i = 0; i < NV; ++i) s = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s))));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    denom = fmaxf(sqrtf(s), 1e-12f);
  }
  uint16_t* dst = a3 + static_cast<size_t>(row) * 3 * E;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int q = i * 32 + lane;
    if (q < nvec) {
      float4 t = v[i];
      if (normalize) {
        t.x /= denom;
        t.y /= denom;
        t.z /= denom;
        t.w /= denom;
      }
      if (emb_out != nullptr) reinterpret_cast<float4*>(emb_out + static_cast<size_t>(row) * E)[q] = t;
      uint16_t h[4], l[4];
      split_hi_lo(t.x, h[0], l[0]);
      split_hi_lo(t.y, h[1], l[1]);
      split_hi_lo(t.z, h[2], l[2]);
      split_hi_lo(t.w, h[3], l[3]);
      const uint2 hv = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
      const uint2 lv = make_uint2(l[0] | (static_cast<uint32_t>(l[1]) << 16), l[2] | (static_cast<uint32_t>(l[3]) << 16));
      reinterpret_cast<uint2*>(dst)[q] = hv;
      reinterpret_cast<uint2*>(dst + E)[q] = hv;
      reinterpret_cast<uint2*>(dst + 2 * E)[q] = lv;
    }
  }
}

// ---- L3 -> L2 aggregation + metrics epilogue (aihab_utils/evaluation.py:92-142, 186-221, 261-273), one launch.
// One warp per row: the L3 logits row is staged in smem; lane g accumulates L2 group g over the L3 ids IN ID ORDER
// (the reference's `for l3_id, l2_id in enumerate(...)` loop, so sum / mean are bit-identical to it);
// then top-k over the L2 logits, and top-3 + softmax probabilities over the L3 logits.
constexpr int L2M_MAX_C3 = 1024;
constexpr int L2M_MAX_C2 = 256;

__device__ __forceinline__ float logaddexp_ref(float a, float b) {
  // torch.logaddexp: equal infinities return themselves, otherwise max + log1p(exp(-|a - b|))
  if (isinf(a) && a == b) return a;
  return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}

// warp top-k of src[0..cols) in (value desc, index asc) order; lane 0 writes idx / val / prob (prob = exp(v - m) * inv)
__device__ __forceinline__ void warp_topk(const float* src, int cols, int k, int lane, int64_t* idx, float* val,
                                          float* prob, float m, float inv) {
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
      const float v = src[c];
      if (precedes(last_v, last_i, v, c) && precedes(v, c, bv, bi)) {
        bv = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (precedes(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      if (idx != nullptr) idx[j] = bi;
      if (val != nullptr) val[j] = bv;
      if (prob != nullptr) prob[j] = expf(bv - m) * inv;
    }
    last_v = bv;
    last_i = bi;
  }
}

__global__ void __launch_bounds__(128) l2_metrics_kernel(const float* __restrict__ logits_l3, int n, int C3,
                                                         const int* __restrict__ l3_to_l2, int C2, int reduce, int k,
                                                         float* __restrict__ logits_l2_out,
                                                         int64_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                                                         int64_t* __restrict__ top3_idx, float* __restrict__ top3_prob) {
  extern __shared__ float l2m_smem[];
  int* s_map = reinterpret_cast<int*>(l2m_smem);  // [C3]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_row = l2m_smem + C3 + warp * (C3 + C2);  // [C3] L3 logits, then [C2] L2 logits
  float* s_l2 = s_row + C3;
  for (int i = threadIdx.x; i < C3; i += blockDim.x) s_map[i] = l3_to_l2[i];
  const int row = blockIdx.x * 4 + warp;
  if (row < n) {
    const float* src = logits_l3 + static_cast<size_t>(row) * C3;
    for (int c = lane; c < C3; c += 32) s_row[c] = src[c];
  }
  __syncthreads();
  if (row >= n) return;
  for (int g = lane; g < C2; g += 32) {
    float acc = reduce == 2 ? -INFINITY : 0.0f;
    float cnt = 0.0f;
    for (int c = 0; c < C3; ++c) {
      if (s_map[c] == g) {
        acc = reduce == 2 ? logaddexp_ref(acc, s_row[c]) : acc + s_row[c];
        cnt += 1.0f;
      }
    }
    if (reduce == 1) acc = acc / fmaxf(cnt, 1.0f);
    s_l2[g] = acc;
    if (logits_l2_out != nullptr) logits_l2_out[static_cast<size_t>(row) * C2 + g] = acc;
  }
  __syncwarp();
  if (k > 0)
    warp_topk(s_l2, C2, k, lane, topk_idx + static_cast<size_t>(row) * k,
              topk_val != nullptr ? topk_val + static_cast<size_t>(row) * k : nullptr, nullptr, 0.f, 0.f);
  if (top3_idx != nullptr || top3_prob != nullptr) {
    float m = -INFINITY;
    for (int c = lane; c < C3; c += 32) m = fmaxf(m, s_row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < C3; c += 32) sum += expf(s_row[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int k3 = C3 < 3 ? C3 : 3;
    warp_topk(s_row, C3, k3, lane, top3_idx != nullptr ? top3_idx + static_cast<size_t>(row) * 3 : nullptr, nullptr,
              top3_prob != nullptr ? top3_prob + static_cast<size_t>(row) * 3 : nullptr, m, 1.0f / sum);
  }
}

// Thread-per-row variant for small class counts (the shipped 20 -> 11 map): 128 rows per CTA are staged through smem
// with coalesced global reads / writes (odd row strides: conflict-free both ways), every thread walks its own row.
// Same accumulation order as above (sum / mean bit-identical to the reference loop).
__device__ __forceinline__ void thread_topk(const float* src, int cols, int k, int64_t* idx, float* val, float* prob,
                                            float m, float inv) {
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = 0; c < cols; ++c) {
      const float v = src[c];
      if (precedes(last_v, last_i, v, c) && precedes(v, c, bv, bi)) {
        bv = v;
        bi = c;
      }
    }
    if (idx != nullptr) idx[j] = bi;
    if (val != nullptr) val[j] = bv;
    if (prob != nullptr) prob[j] = expf(bv - m) * inv;
    last_v = bv;
    last_i = bi;
  }
}

__global__ void __launch_bounds__(128) l2_metrics_small_kernel(const float* __restrict__ logits_l3, int n, int C3,
                                                               const int* __restrict__ l3_to_l2, int C2, int reduce,
                                                               int k, float* __restrict__ logits_l2_out,
                                                               int64_t* __restrict__ topk_idx,
                                                               float* __restrict__ topk_val,
                                                               int64_t* __restrict__ top3_idx,
                                                               float* __restrict__ top3_prob) {
  extern __shared__ float l2m_smem[];
  const int st3 = C3 | 1, st2 = C2 | 1;
  float* s_x = l2m_smem;               // [128][st3]
  float* s_l2 = s_x + 128 * st3;       // [128][st2]
  int* s_map = reinterpret_cast<int*>(s_l2 + 128 * st2);  // [C3]
  float* s_cnt = reinterpret_cast<float*>(s_map + C3);    // [C2]
  const int t = threadIdx.x;
  const long row0 = static_cast<long>(blockIdx.x) * 128;
  const int rows = static_cast<int>(min(static_cast<long>(128), n - row0));
  for (int i = t; i < C3; i += 128) s_map[i] = l3_to_l2[i];
  for (int i = t; i < C2; i += 128) s_cnt[i] = 0.f;
  const float* src = logits_l3 + row0 * C3;
  for (int e = t; e < rows * C3; e += 128) {
    const int r = e / C3;
    s_x[r * st3 + (e - r * C3)] = src[e];
  }
  __syncthreads();
  if (t == 0)
    for (int c = 0; c < C3; ++c) s_cnt[s_map[c]] += 1.0f;
  __syncthreads();
  if (t < rows) {
    const float* x = s_x + t * st3;
    float* l2 = s_l2 + t * st2;
    for (int g = 0; g < C2; ++g) l2[g] = reduce == 2 ? -INFINITY : 0.0f;
    for (int c = 0; c < C3; ++c) {
      const int g = s_map[c];
      l2[g] = reduce == 2 ? logaddexp_ref(l2[g], x[c]) : l2[g] + x[c];
    }
    if (reduce == 1)
      for (int g = 0; g < C2; ++g) l2[g] = l2[g] / fmaxf(s_cnt[g], 1.0f);
    const long row = row0 + t;
    if (k > 0) thread_topk(l2, C2, k, topk_idx + row * k, topk_val != nullptr ? topk_val + row * k : nullptr, nullptr, 0.f, 0.f);
    if (top3_idx != nullptr || top3_prob != nullptr) {
      float m = -INFINITY;
      for (int c = 0; c < C3; ++c) m = fmaxf(m, x[c]);
      float sum = 0.f;
      for (int c = 0; c < C3; ++c) sum += expf(x[c] - m);
      thread_topk(x, C3, C3 < 3 ? C3 : 3, top3_idx != nullptr ? top3_idx + row * 3 : nullptr, nullptr,
                  top3_prob != nullptr ? top3_prob + row * 3 : nullptr, m, 1.0f / sum);
    }
  }
  if (logits_l2_out != nullptr) {
    __syncthreads();
    float* dst = logits_l2_out + row0 * C2;
    for (int e = t; e < rows * C2; e += 128) {
      const int r = e / C2;
      dst[e] = s_l2[r * st2 + (e - r * C2)];
    }
  }
}

// ---- multi-prototype scoring over cached embeddings (tools/outlier_cleaning.py:553-668): per row of the similarity
// matrix sim [n, P] = emb @ prototypes^T: the best similarity among the prototypes of the row's own class (and that
// prototype's index inside its class block, first maximum), the best among all other classes (NaN if there is none)
// and the margin between the two.  owner[p] = class of prototype p; class blocks are contiguous.  One warp per row.
__global__ void __launch_bounds__(128) prototype_reduce_kernel(const float* __restrict__ sim, const int64_t* __restrict__ labels,
                                                               const int64_t* __restrict__ owner, int n, int P,
                                                               int ld, float* __restrict__ sim_own, int64_t* __restrict__ proto_id,
                                                               float* __restrict__ sim_other, float* __restrict__ margin) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* src = sim + static_cast<size_t>(row) * ld;
  const int64_t lab = labels[row];
  float bv = -INFINITY, ov = -INFINITY;
  int bi = 0x7fffffff, first_own = 0x7fffffff;
  for (int p = lane; p < P; p += 32) {
    const float v = src[p];
    if (owner[p] == lab) {
      first_own = min(first_own, p);
      if (precedes(v, p, bv, bi)) {
        bv = v;
        bi = p;
      }
    } else {
      ov = fmaxf(ov, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float tv = __shfl_xor_sync(0xffffffffu, bv, o);
    const int ti = __shfl_xor_sync(0xffffffffu, bi, o);
    if (precedes(tv, ti, bv, bi)) {
      bv = tv;
      bi = ti;
    }
    ov = fmaxf(ov, __shfl_xor_sync(0xffffffffu, ov, o));
    first_own = min(first_own, __shfl_xor_sync(0xffffffffu, first_own, o));
  }
  if (lane == 0) {
    const float other = isinf(ov) ? __int_as_float(0x7fc00000) : ov;  // masked_fill(isinf, nan)
    sim_own[row] = bv;
    if (proto_id != nullptr) proto_id[row] = bi == 0x7fffffff ? -1 : bi - first_own;
    if (sim_other != nullptr) sim_other[row] = other;
    if (margin != nullptr) margin[row] = bv - other;
  }
}

}  // namespace

cudaError_t launch_transpose16(const void* src, void* dst, int R, int Cc, cudaStream_t stream) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose16_kernel<<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), R, Cc);
  return cudaGetLastError();
}

cudaError_t launch_split_textw(const float* w, int E, int C, void* out, cudaStream_t stream) {
  const long total = static_cast<long>(E) * C;
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 8));
  split_textw_kernel<<<grid, 256, 0, stream>>>(w, E, C, static_cast<uint16_t*>(out));
  return cudaGetLastError();
}

cudaError_t launch_l2norm_split(const float* emb, float* emb_out, void* a3, int rows, int E, cudaStream_t stream,
                                int normalize) {
  if (rows <= 0) return cudaSuccess;
  const bool vec = (E & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0 && (reinterpret_cast<uintptr_t>(a3) & 7) == 0 &&
                   (emb_out == nullptr || (reinterpret_cast<uintptr_t>(emb_out) & 15) == 0);
  if (vec && E <= 512)
    l2norm_split_vec_kernel<4><<<(rows + 3) / 4, 128, 0, stream>>>(emb, emb_out, static_cast<uint16_t*>(a3), rows, E, normalize);
  else if (vec && E <= 1024)
    l2norm_split_vec_kernel<8><<<(rows + 3) / 4, 128, 0, stream>>>(emb, emb_out, static_cast<uint16_t*>(a3), rows, E, normalize);
  else
    l2norm_split_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(emb, emb_out, static_cast<uint16_t*>(a3), rows, E, normalize);
  return cudaGetLastError();
}

bool score_mid_supported(int n, int D, int E, int C) {
  static const bool enabled = [] {
    const char* e = getenv("AIHAB_SCORE_MID");
    return e == nullptr || e[0] != '0';
  }();
  // measured at n = 256, 768 -> 512: 1000 classes 68 us against 150 us for score_fused_kernel; 20 classes 54 against 44 us
  // (both kernels are then bound by the per-CTA fmaf / shared-memory chain of the projection), so few-class heads stay
  // on the one-launch kernel
  return enabled && n > 64 && n <= 8192 && D > 0 && (D % 4) == 0 && (E % 32) == 0 && C > 64 && (C % 4) == 0 &&
         static_cast<size_t>(MID_ROWS) * std::max(D, E) * sizeof(float) <= 96 * 1024;
}

cudaError_t launch_score_mid(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C, float scale,
                             float* emb_raw, float* emb_out, float* logits, cudaStream_t stream) {
  const size_t chunk = static_cast<size_t>(MID_NST) * MID_KC * 32 * sizeof(float);
  const size_t smem_a = static_cast<size_t>(MID_ROWS) * D * sizeof(float) + chunk, smem_b = static_cast<size_t>(MID_ROWS) * E * sizeof(float) + chunk;
  cudaError_t e;
  if (smem_a > 48 * 1024 &&
      (e = cudaFuncSetAttribute(score_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_a))) != cudaSuccess)
    return e;
  if (smem_b > 48 * 1024 &&
      (e = cudaFuncSetAttribute(score_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_b))) != cudaSuccess)
    return e;
  const int row_groups = (n + MID_ROWS - 1) / MID_ROWS;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cfg.stream = stream;
  cfg.blockDim = dim3(128);
  cfg.gridDim = dim3(E / 32, row_groups);
  cfg.dynamicSmemBytes = smem_a;
  if (proj == nullptr) emb_raw = const_cast<float*>(feats);  // features are already projected (E == D): normalise + logits only
  else if ((e = cudaLaunchKernelEx(&cfg, score_proj_kernel, feats, n, D, proj, E, emb_raw)) != cudaSuccess) return e;
  cfg.gridDim = dim3((C + 31) / 32, row_groups);
  cfg.dynamicSmemBytes = smem_b;
  return cudaLaunchKernelEx(&cfg, score_logits_kernel, static_cast<const float*>(emb_raw), n, E, text_w, C, scale, emb_out, logits);
}

bool score_fused_supported(int n, int D, int E, int C) {
  const size_t smem = static_cast<size_t>(RPB) * (D + 2 * E + (C > 0 ? C : 0)) * sizeof(float);
  static const int max_rows = [] {
    const char* e = getenv("AIHAB_SCORE_FUSED_MAX_ROWS");
    return e ? atoi(e) : 8192;
  }();
  return n <= max_rows && smem <= 160 * 1024 && (E % 4) == 0 && (D % 2) == 0;
}

cudaError_t launch_score_fused(const float* feats, int n, int D, const float* proj, int E, const float* text_w, int C,
                               float scale, int k, float* emb_out, float* logits_out, int64_t* topk_idx,
                               float* topk_val, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(RPB) * (D + 2 * E + (text_w ? C : 0)) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(score_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  score_fused_kernel<<<(n + RPB - 1) / RPB, 256, smem, stream>>>(feats, n, D, proj, E, text_w, C, scale, k, emb_out,
                                                                 logits_out, topk_idx, topk_val);
  return cudaGetLastError();
}

cudaError_t launch_sgemm(const float* A, const float* B, float* C, int M, int N, int K, float alpha,
                         cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  sgemm_kernel<<<grid, 256, 0, stream>>>(A, B, C, M, N, K, alpha);
  return cudaGetLastError();
}

cudaError_t launch_l2norm(const float* x, float* y, int rows, int cols, float eps, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  l2norm_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(x, y, rows, cols, eps);
  return cudaGetLastError();
}

cudaError_t launch_topk(const float* logits, int rows, int cols, int k, int64_t* idx, float* val,
                        cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (k <= 0 || k > cols) return cudaErrorInvalidValue;
  const bool vec = (cols & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (vec && cols <= 256)
    topk_regs_kernel<2><<<(rows + 3) / 4, 128, 0, stream>>>(logits, rows, cols, k, idx, val);
  else if (vec && cols <= 512)
    topk_regs_kernel<4><<<(rows + 3) / 4, 128, 0, stream>>>(logits, rows, cols, k, idx, val);
  else if (vec && cols <= 1024)
    topk_regs_kernel<8><<<(rows + 3) / 4, 128, 0, stream>>>(logits, rows, cols, k, idx, val);
  else
    topk_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(logits, rows, cols, k, idx, val);
  return cudaGetLastError();
}

cudaError_t launch_topk_merge(const float* cand_val, const int* cand_idx, int rows, int slots, int k, int64_t* idx,
                              float* val, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (k <= 0 || k > 8 || slots <= 0) return cudaErrorInvalidValue;
  if (slots <= 8) {
    const long threads = static_cast<long>(rows) * 8;
    topk_merge8_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(cand_val, cand_idx, rows, slots, k, idx,
                                                                                        val);
    return cudaGetLastError();
  }
  topk_merge_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(cand_val, cand_idx, rows, slots, k, idx, val);
  return cudaGetLastError();
}

cudaError_t launch_l2_metrics(const float* logits_l3, int n, int C3, const int* l3_to_l2, int C2, int reduce, int k,
                              float* logits_l2_out, int64_t* topk_idx, float* topk_val, int64_t* top3_idx,
                              float* top3_prob, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (C3 <= 0 || C3 > L2M_MAX_C3 || C2 <= 0 || C2 > L2M_MAX_C2 || reduce < 0 || reduce > 2 || k < 0 || k > C2 ||
      (k > 0 && topk_idx == nullptr))
    return cudaErrorInvalidValue;
  const size_t small = (static_cast<size_t>(128) * ((C3 | 1) + (C2 | 1)) + C3 + C2) * sizeof(float);
  if (small <= 48 * 1024) {  // thread per row (the shipped 20 -> 11 map: 16 KB)
    l2_metrics_small_kernel<<<(n + 127) / 128, 128, small, stream>>>(logits_l3, n, C3, l3_to_l2, C2, reduce, k,
                                                                     logits_l2_out, topk_idx, topk_val, top3_idx, top3_prob);
    return cudaGetLastError();
  }
  const size_t smem = (static_cast<size_t>(C3) + 4 * (C3 + C2)) * sizeof(float);
  l2_metrics_kernel<<<(n + 3) / 4, 128, smem, stream>>>(logits_l3, n, C3, l3_to_l2, C2, reduce, k, logits_l2_out, topk_idx,
                                                        topk_val, top3_idx, top3_prob);
  return cudaGetLastError();
}

cudaError_t launch_prototype_reduce(const float* sim, const int64_t* labels, const int64_t* owner, int n, int P, int ld,
                                    float* sim_own, int64_t* proto_id, float* sim_other, float* margin,
                                    cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  prototype_reduce_kernel<<<(n + 3) / 4, 128, 0, stream>>>(sim, labels, owner, n, P, ld, sim_own, proto_id, sim_other, margin);
  return cudaGetLastError();
}

}  // namespace aihab
