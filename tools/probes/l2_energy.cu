// tools/probes/l2_energy.cu - joules per GB delivered from L2 (or HBM) to shared memory by bulk copies, plain and with
// cluster multicast (one L2 read delivered to the shared memory of BOTH CTAs of a cluster of 2).  Decides whether a
// 4-CTA-cluster GEMM that multicasts the A tile to two CTA pairs is worth building at the power cap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/probes/l2_energy.so tools/probes/l2_energy.cu
//   python tools/probes/l2_energy.py
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
constexpr int CHUNK = 16384;
constexpr int NS = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20000;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void arrive_remote(uint64_t* b, uint32_t cta) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// MC = 0: every CTA streams its own slice.  MC = 1: clusters of 2, the two CTAs take turns issuing ONE bulk copy that lands
// in both CTAs' shared memory.
template <int MC>
__global__ void __launch_bounds__(32, 1) stream_kernel(const uint8_t* buf, size_t slice_bytes, int passes) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * CHUNK);
  uint64_t* empty = full + NS;
  const uint32_t rank = MC ? cta_rank() : 0;
  const int unit = MC ? blockIdx.x >> 1 : blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MC ? 2 : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    const uint8_t* src = buf + static_cast<size_t>(unit) * slice_bytes;
    const long per_pass = static_cast<long>(slice_bytes / CHUNK);
    const long n = per_pass * passes;
    auto issue = [&](long j) {
      if (j >= n) return;
      if (MC && (j & 1) != rank) return;
      const int slot = static_cast<int>(j % NS);
      if (j >= NS) mbar_wait(&empty[slot], ((j / NS) - 1) & 1);
      const uint8_t* g = src + (j % per_pass) * CHUNK;
      if (MC) {
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                smem_u32(smem + slot * CHUNK)),
            "l"(g), "r"(CHUNK), "r"(smem_u32(&full[slot])), "h"(static_cast<uint16_t>(3))
            : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + slot * CHUNK)),
                     "l"(g), "r"(CHUNK), "r"(smem_u32(&full[slot]))
                     : "memory");
      }
    };
    for (long j = 0; j < NS - 1; ++j) issue(j);
    for (long j = 0; j < n; ++j) {
      const int slot = static_cast<int>(j % NS);
      mbar_expect(&full[slot], CHUNK);
      issue(j + NS - 1);
      mbar_wait(&full[slot], (j / NS) & 1);
      if (MC) arrive_remote(&empty[slot], static_cast<uint32_t>(j & 1));  // to the CTA that issues chunk j + NS
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
    }
  }
  __syncthreads();
  if (MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
}  // namespace

// returns bytes delivered to shared memory per launch (all CTAs), or -1 on error
extern "C" long long l2_stream(const void* buf, unsigned long long slice_bytes, int passes, int multicast, int ctas,
                               void* stream) {
  const size_t smem = NS * CHUNK + 2 * NS * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (multicast) {
    cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, stream_kernel<1>, static_cast<const uint8_t*>(buf), static_cast<size_t>(slice_bytes), passes) !=
        cudaSuccess)
      return -1;
  } else {
    cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    stream_kernel<0><<<ctas, 32, smem, s>>>(static_cast<const uint8_t*>(buf), static_cast<size_t>(slice_bytes), passes);
    if (cudaGetLastError() != cudaSuccess) return -1;
  }
  return static_cast<long long>(ctas) * (slice_bytes / CHUNK) * CHUNK * passes;
}
