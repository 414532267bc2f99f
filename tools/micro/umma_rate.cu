// Micro-benchmark: issue rate / execution time of tcgen05.mma (kind::f16, M=128, K=16, SS operands, no swizzle)
// as a function of N, accumulator rotation, and CTAs per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}

template <int ROT, int STEP>
__global__ void __launch_bounds__(64) k(int N, int iters, long long* out) {
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
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 40 * 1024) >> 4;
    const uint64_t ad = hi | a0 | (uint64_t)(2048 >> 4) << 16, bd = hi | b0 | (uint64_t)((N * 16) >> 4) << 16;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (leader) umma(tm + (uint32_t)((u % ROT) * 64), ad + (uint64_t)(u * STEP), bd, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a));
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_a));
    const long long t2 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

template <int ROT, int STEP>
void run(const char* name, int N, int ctas_per_sm) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 256, smem = ctas_per_sm == 1 ? 120 * 1024 : 60 * 1024;
  cudaFuncSetAttribute(k<ROT, STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<ROT, STEP><<<148 * ctas_per_sm, 64, smem>>>(N, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-22s N=%3d ctas/SM=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", name, N, ctas_per_sm,
         (double)h[0] / (iters * 8), (double)h[1] / (iters * 8), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int Ns[] = {16, 32, 48, 64, 128, 256};
  for (int c = 1; c <= 2; ++c)
    for (int N : Ns) {
      if (N > 64) { run<1, 0>("same acc, same A", N, c); continue; }
      run<1, 0>("same acc, same A", N, c);
      run<4, 0>("4 accs, same A", N, c);
      run<4, 8>("4 accs, A += 128 B", N, c);
    }
  return 0;
}
