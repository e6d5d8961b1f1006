// The exp2 pass of the attention softmax in isolation: per 32-column chunk tcgen05.ld.x32 (one ahead) -> 32 x (FFMA,
// MUFU.EX2) -> row sum + pack -> tcgen05.st.x16, one warp per scheduler (W = 1) or two (W = 2), nothing else on the SM.
// Tells whether the ~540 cycles per chunk seen inside attention_tcd_kernel are inherent to this instruction stream or
// come from interference (polling warps, MMA traffic on TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aihab_clip_b200/csrc tools/probes/softmax_chunk.cu -o tools/probes/softmax_chunk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(int iters, float ms, long long* cyc, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t t_row = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  const float sl2 = 0.18033688f;
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  uint32_t ra[32], rb[32];
  auto exp32 = [&](uint32_t (&r)[32], int c) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(ptx::ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -ms)));
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      l0 += __uint_as_float(r[2 * j]);
      l1 += __uint_as_float(r[2 * j + 1]);
      l2 += __uint_as_float(r[2 * j + 2]);
      l3 += __uint_as_float(r[2 * j + 3]);
      pk[j] = ptx::pack2<false>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      pk[j + 1] = ptx::pack2<false>(__uint_as_float(r[2 * j + 2]), __uint_as_float(r[2 * j + 3]));
    }
    if (MODE & 1) ptx::tmem_st_32x16(t_row + c * 16, pk);
    else l0 += __uint_as_float(pk[0] ^ pk[7] ^ pk[15]);
  };
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE & 2) {
      ptx::tmem_ld_32x32(t_row, ra);
#pragma unroll 1
      for (int c = 0; c < 6; c += 2) {
        ptx::tmem_ld_wait_regs(ra);
        ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
        exp32(ra, c);
        ptx::tmem_ld_wait_regs(rb);
        if (c + 2 < 6) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, ra);
        exp32(rb, c + 1);
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < 6; c += 2) {
        exp32(ra, c);
        exp32(rb, c + 1);
      }
    }
    if (MODE & 1) ptx::tmem_st_wait();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (l0 + l1 + l2 + l3 == 12345.678f) sink[0] = l0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_slot, 512);
  }
}

template <int MODE>
void run(const char* name, long long* d, float* sink) {
  for (int threads : {128, 256}) {
    long long h = 0;
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) {
      probe<MODE><<<148, threads>>>(iters, 3.0f, d, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError()));
        return;
      }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %d warp(s)/scheduler: %7.1f cycles per 32-column chunk per warp\n", name, threads / 128, double(h) / iters / 6);
  }
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  run<0>("exp2 + sum + pack (registers only)", d, sink);
  run<1>("exp2 + sum + pack + tcgen05.st", d, sink);
  run<2>("tcgen05.ld (one ahead) + exp2 + sum + pack", d, sink);
  run<3>("tcgen05.ld + exp2 + sum + pack + tcgen05.st", d, sink);
  return 0;
}
