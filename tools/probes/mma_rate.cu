// Probe for round 2 (DESIGN.md 4.1 / 4.2 open questions): how many cycles does a tcgen05.mma of a given shape take when
// issued back to back, alone on the SM and next to shared-memory traffic?
//   * SS (A and B from shared memory) M = 128, N = 256 / 208 / 128 / 64, K = 16   -> the GEMM and S = Q K^T shapes
//   * TS (A from tensor memory)       M = 128, N = 64, K = 16                      -> the O = P V shape
//   * each of them again while the other warps of the CTA stream st.shared / ld.shared over a 64 KB window
//     (stand-in for the TMA fills and the epilogue staging that share the banks with the operand reads)
// One CTA per SM; one thread issues `reps` MMAs, commits to an mbarrier and waits; clock64 around issue -> completion.
// Operand contents do not matter (whatever is in shared / tensor memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aihab_clip_b200/csrc tools/probes/mma_rate.cu -o tools/probes/mma_rate
//   gpurun -- ./tools/probes/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"


// issue by a CONVERGED warp: every lane computes the (uniform) descriptors, one elected lane issues.  The compiler can
// then keep the operands in uniform registers instead of the R2UR.BROADCAST + ELECT waterfall it emits around a
// tcgen05.mma sitting in a divergent `if (lane == 0)` region.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int THREADS = 192;  // warp 0: issuer, warp 1: TMEM allocator, warps 2..5: optional smem traffic
constexpr int SMEM = 200 * 1024;

__global__ void __launch_bounds__(THREADS, 1)
probe(int n, int ts, int reps, int traffic, long long* cycles_out, int mode = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    stop = 0;
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && mode == 4) {  // converged-warp issue, k walk, one accumulator (compare with the base case)
    const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + 16384));
    const uint32_t idesc = ptx::make_idesc_f16(0, 128, n, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int k = r & 3;
      umma_f16_elect(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, r != 0);
    }
    if (lane == 0) {
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      cycles_out[blockIdx.x] = clock64() - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp == 0) {
    if (lane == 0) {
      // A tile: 128 rows x 64 (K-major, SW128) at smem + 0; B tile: up to 256 rows x 64 at smem + 16 KB
      const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
      const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + 16384));
      const uint32_t idesc = ptx::make_idesc_f16(0, 128, n, ts ? 1 : 0);
      const long long t0 = clock64();
      if (mode == 1) {         // two accumulators, alternating: is it the dependent accumulate into ONE D tile?
        for (int r = 0; r < reps; ++r) {
          const int k = r & 3;
          ptx::umma_f16(tmem + ((r >> 2) & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, r > 7);
        }
      } else if (mode == 2) {  // issue loop unrolled by 4, constant predicate: is it the issuing thread?
        for (int r = 0; r < reps; r += 4) {
          ptx::umma_f16(tmem, adesc + 0, bdesc + 0, idesc, 1);
          ptx::umma_f16(tmem, adesc + 2, bdesc + 2, idesc, 1);
          ptx::umma_f16(tmem, adesc + 4, bdesc + 4, idesc, 1);
          ptx::umma_f16(tmem, adesc + 6, bdesc + 6, idesc, 1);
        }
      } else if (mode == 3) {  // same operands every time (no k walk): is it the operand fetch?
        for (int r = 0; r < reps; ++r) ptx::umma_f16(tmem, adesc, bdesc, idesc, 1);
      } else {
      for (int r = 0; r < reps; ++r) {
        const int k = r & 3;  // walk the four K = 16 slices of the 64-wide tile like the real kernels
        if (ts) ptx::umma_f16_ts(ts == 2 ? tmem : tmem + 256, tmem + (ts == 2 ? 256 : 0) + 8 * k, bdesc + 2 * k, idesc, r != 0);  // A = TMEM cols
        else ptx::umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, r != 0);
      }
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      const long long t1 = clock64();
      cycles_out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp >= 2 && traffic) {
    // 4 warps x 128-bit accesses over a 64 KB window above the operand tiles
    uint4* w = reinterpret_cast<uint4*>(smem + 96 * 1024);
    uint4 v = make_uint4(lane, warp, 0, 0);
    int i = (warp - 2) * 32 + lane;
    while (!stop) {
#pragma unroll 8
      for (int u = 0; u < 8; ++u) {
        if (traffic & 1) w[i & 4095] = v;
        if (traffic & 2) v.x ^= w[(i + 2048) & 4095].x;
        i += 128;
      }
    }
    if (v.x == 0xdeadbeef) cycles_out[0] = 0;  // keep the loads alive
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  long long* d = nullptr;
  cudaMalloc(&d, sms * sizeof(long long));
  long long* h = new long long[sms];
  const int reps = 4096;
  struct Case { const char* name; int n, ts; };
  const Case cases[] = {{"SS M128 N256", 256, 0}, {"SS M128 N208", 208, 0}, {"SS M128 N128", 128, 0},
                        {"SS M128 N64 ", 64, 0},  {"TS M128 N64 ", 64, 1}};
  printf("SMs %d, %d MMAs (K = 16) per measurement, cycles per MMA (mean over SMs)\n", sms, reps);
  for (const Case& c : cases) {
    for (int traffic = 0; traffic <= 3; ++traffic) {
      for (int rep = 0; rep < 2; ++rep) {  // second run is the measurement
        probe<<<sms, THREADS, SMEM>>>(c.n, c.ts, reps, traffic, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("%s: %s\n", c.name, cudaGetErrorString(e));
          return 1;
        }
      }
      cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < sms; ++i) s += h[i];
      const char* tn[] = {"alone", "+ st.shared", "+ ld.shared", "+ st/ld.shared"};
      printf("%s  %-15s %7.1f cycles/MMA   (full rate would be %d)\n", c.name, tn[traffic], s / sms / reps, c.n / 2);
    }
  }
  struct Extra { const char* name; int n, ts, mode; };
  const Extra extra[] = {{"SS M128 N256 two accumulators alternating", 256, 0, 1}, {"SS M128 N256 issue unrolled x4", 256, 0, 2},
                         {"SS M128 N256 same operand slice", 256, 0, 3},          {"TS M128 N256 (A from TMEM)", 256, 2, 0},
                         {"SS M128 N256 converged warp + elect, k walk", 256, 0, 4}, {"SS M128 N128 converged warp + elect, k walk", 128, 0, 4},
                         {"SS M128 N128 issue unrolled x4", 128, 0, 2},           {"SS M128 N64 issue unrolled x4", 64, 0, 2}};
  for (const Extra& c : extra) {
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<sms, THREADS, SMEM>>>(c.n, c.ts, reps, 0, d, c.mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", c.name, cudaGetErrorString(e));
        return 1;
      }
    }
    cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < sms; ++i) s += h[i];
    printf("%-45s %7.1f cycles/MMA   (nominal floor %d)\n", c.name, s / sms / reps, c.n / 2);
  }
  return 0;
}
