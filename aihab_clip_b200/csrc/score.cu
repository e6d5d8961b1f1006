// Zero-shot scoring epilogue in exact fp32 (CUDA-core FMA): visual projection, L2 normalisation, x100 cosine
// logits and top-k.   Reference: methods/ProLIP.py:40 (x @ vit_proj), methods/utils.py:183-186
// (F.normalize -> 100. * f @ text_weights -> argmax), methods/utils.py:16-21 / aihab_utils/evaluation.py:261-273
// (topk, sorted, lowest index first among exact ties).
#include "kernels.cuh"

#include <math.h>

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

}  // namespace

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
  topk_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(logits, rows, cols, k, idx, val);
  return cudaGetLastError();
}

}  // namespace aihab
