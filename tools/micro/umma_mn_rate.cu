// Micro-benchmark: execution time of tcgen05.mma (kind::f16, M=128, K=16, SS operands, no swizzle) with MN-MAJOR operands
// (the weight-gradient kernels' layout: 8 M/N elements contiguous in 16 bytes, K positions 16 bytes apart, LBO = 128 B to the
// next K group, SBO = stride between 8-element M/N groups) as a function of the two SBOs and N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/umma_mn_rate tools/micro/umma_mn_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(64) k(int N, int sbo_a, int sbo_b, int mn_a, int mn_b, int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t bar_a = smem_u32(&bar);
  if (warp == 0) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)mn_a << 15) | ((uint32_t)mn_b << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    // MN-major: LBO = 128 (next K group), SBO = group stride.  K-major: LBO = next K group (sbo arg), SBO = 128 (next 8 rows)
    auto desc = [](uint32_t addr, int mn, int s) {
      const uint32_t lbo = mn ? 128u : (uint32_t)s, sbo = mn ? (uint32_t)s : 128u;
      return ((uint64_t)(((sbo >> 4) & 0x3FFF) | (1u << 14)) << 32) | (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16);
    };
    const uint64_t ad = desc(smem_u32(smem), mn_a, sbo_a), bd = desc(smem_u32(smem) + 100 * 1024, mn_b, sbo_b);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma(tm + (uint32_t)((u & 1) * 128), ad + (uint64_t)((u & 3) * 16), bd, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a));
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_a));
    const long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

void run(int N, int sbo_a, int sbo_b, int mn_a, int mn_b) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 128, smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<148, 64, smem>>>(N, sbo_a, sbo_b, mn_a, mn_b, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("A %s SBO/LBO %5d   B %s SBO/LBO %5d   N=%3d: %.1f cyc/MMA  (%s)\n", mn_a ? "MN" : "K ", sbo_a, mn_b ? "MN" : "K ", sbo_b, N,
         (double)h[0] / (iters * 8), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int sbos[] = {512, 528, 544, 576, 640, 1040, 4096, 4112, 4128};
  printf("# K-major reference\n");
  for (int N : {32, 112}) run(N, 2048, N * 16, 0, 0);
  printf("# MN-major A (SBO varies), K-major B\n");
  for (int s : sbos) run(32, s, 32 * 16, 1, 0);
  printf("# K-major A, MN-major B (SBO varies), N = 112\n");
  for (int s : sbos) run(112, 2048, s, 0, 1);
  printf("# both MN-major, same SBO\n");
  for (int N : {32, 80, 112})
    for (int s : sbos) run(N, s, s, 1, 1);
  return 0;
}
