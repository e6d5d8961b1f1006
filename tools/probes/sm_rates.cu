// Single-warp issue / pipe rates on one SM sub-partition (round-2 probe for the attention softmax):
//   MUFU.EX2, FFMA, the degree-3 exp2 polynomial, tcgen05.ld.x32 latency and pipelined rate, tcgen05.st.x16.
// W warps per scheduler run the same loop (W = 1, 2); cycles per warp-instruction as seen by one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aihab_clip_b200/csrc tools/probes/sm_rates.cu -o tools/probes/sm_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = __fadd_rd(x, 12582912.0f);
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.077119089663028717041015625f, f, 0.227564394474029541015625f);
  p = fmaf(p, f, 0.695146143436431884765625f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

__global__ void __launch_bounds__(256, 1) probe(int mode, int iters, float seed, long long* cyc, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t taddr = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed * (i + 1) - threadIdx.x * 1e-3f;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {  // MUFU.EX2, 16 independent chains
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ptx::ex2_approx(v[i]);
    }
  } else if (mode == 1) {  // FFMA
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], 0.999f, 0.001f);
    }
  } else if (mode == 2) {  // polynomial exp2 (10 instructions)
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = exp2_poly(v[i] * 0.5f - 1.0f);
    }
  } else if (mode == 3) {  // the softmax inner loop: FFMA + MUFU + FADD per element, F2FP per pair
    float l0 = 0.f, l1 = 0.f;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = ptx::ex2_approx(fmaf(v[i], 0.18f, -seed));
        const float p1 = ptx::ex2_approx(fmaf(v[i + 1], 0.18f, -seed));
        l0 += p0;
        l1 += p1;
        acc ^= ptx::pack2<false>(p0, p1);
      }
    }
    v[0] = l0 + l1 + __uint_as_float(acc & 0xff);
  } else if (mode == 4) {  // tcgen05.ld.x32 + wait: latency
    for (int it = 0; it < iters; ++it) {
      ptx::tmem_ld_32x32(taddr, r);
      ptx::tmem_ld_wait();
    }
  } else if (mode == 5) {  // two tcgen05.ld.x32 in flight per wait: pipelined rate
    uint32_t r2[32];
    for (int it = 0; it < iters; it += 2) {
      ptx::tmem_ld_32x32(taddr, r);
      ptx::tmem_ld_32x32(taddr + 32, r2);
      ptx::tmem_ld_wait();
      r[0] ^= r2[0];
    }
  } else if (mode == 6) {  // tcgen05.st.x16 + wait::st
    for (int it = 0; it < iters; ++it) {
      ptx::tmem_st_32x16(taddr, reinterpret_cast<uint32_t(&)[16]>(r[0]));
      ptx::tmem_st_wait();
    }
  } else if (mode == 7) {  // tcgen05.st.x16 back to back, one wait at the end
    for (int it = 0; it < iters; ++it) ptx::tmem_st_32x16(taddr + (it & 3) * 16, reinterpret_cast<uint32_t(&)[16]>(r[0]));
    ptx::tmem_st_wait();
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (s == 12345.678f) sink[0] = s + r[0];
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_slot, 512);
  }
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  const char* names[] = {"MUFU.EX2 (16 chains)", "FFMA (16 chains)", "exp2 polynomial (per exp2)", "softmax element (FFMA+MUFU+FADD+F2FP/2)",
                         "tcgen05.ld.x32 + wait (latency)", "tcgen05.ld.x32 two in flight (per load)", "tcgen05.st.x16 + wait::st",
                         "tcgen05.st.x16 back to back"};
  const int per_iter[] = {16, 16, 16, 16, 1, 1, 1, 1};
  for (int threads : {128, 256}) {
    printf("== %d warp(s) per scheduler\n", threads / 128);
    for (int mode = 0; mode < 8; ++mode) {
      const int iters = mode < 4 ? 2048 : 4096;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<148, threads>>>(mode, iters, 0.37f, d, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) {
          printf("mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError()));
          return 1;
        }
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("  %-45s %8.2f cycles per warp-instruction (per warp)\n", names[mode], double(h) / iters / per_iter[mode]);
    }
  }
  return 0;
}
