// Adam on the flat parameter buffer: ONE launch per step.
//
// hcunet_b200.FlatParameters makes every nn.Parameter a view of one fp32 buffer whose gradient is the engine's flat gradient buffer
// (after the data-parallel all-reduce: the same values on every rank).  torch's path over that buffer is a fill + a multi-tensor
// non-finite check + the fused Adam kernel + step-counter kernels -- ~0.1 ms per step for 727 009 parameters (2.9 MB), i.e. pure
// launch latency.  Here one kernel does the skip-step check of fp16-storage training and the update:
//   phase 1: every CTA scans its slice of the gradient for inf / NaN and raises a device flag;
//   grid barrier (the grid is at most one small CTA per SM: all resident);
//   phase 2: flag clear -> m = lerp(m, g, 1 - b1), v = b2 v + (1 - b2) g^2, p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//            (torch.optim.Adam's arithmetic, fp32, bias corrections in double); flag set -> nothing is touched, t is not advanced.
// Replaces: torch.optim.Adam(...).step() of the reference's training scripts (tests/r_unet_test.py:24) + GradScaler's skip-step.
#include <algorithm>

#include "common.cuh"

namespace hcu {
namespace opt {

struct AdamParams {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  float lr, beta1, beta2, eps, wd;
  int* step;
  float* found_inf;     // scratch[0]
  unsigned* barrier;    // scratch[1]
};

__device__ __forceinline__ bool finite4(const float4& a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(a.w); }

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamParams& a, float step_size, float bc2_sqrt) {
  if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  m = m + (1.f - a.beta1) * (g - m);
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(256) adam_flat_kernel(const AdamParams a) {
  const long long n4 = a.n >> 2;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  // ---- phase 1 ----
  int bad = 0;
  for (long long i = tid; i < n4; i += stride) bad |= !finite4(__ldg(g4 + i));
  for (long long i = (n4 << 2) + tid; i < a.n; i += stride) bad |= !isfinite(__ldg(a.g + i));
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    if (bad) atomicExch(reinterpret_cast<unsigned*>(a.found_inf), __float_as_uint(1.0f));
    if (blockIdx.x == 0) atomicAdd(a.step, 1);   // every CTA reads t after the barrier; taken back below when the step is skipped
    __threadfence();
    atomicAdd(a.barrier, 1u);
    // bounded: a grid that is not fully resident (it always is: at most one 256-thread CTA per SM) traps instead of hanging the GPU
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned*>(a.barrier) < gridDim.x) {
      __nanosleep(32);
      if (clock64() - t0 > 4000000000ll) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  // ---- phase 2 ----
  if (*reinterpret_cast<volatile float*>(a.found_inf) != 0.f) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicSub(a.step, 1);
    return;
  }
  const int t = *reinterpret_cast<volatile int*>(a.step);
  const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
  const float step_size = (float)((double)a.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  float4* p4 = reinterpret_cast<float4*>(a.p);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  for (long long i = tid; i < n4; i += stride) {
    const float4 g = __ldg(g4 + i);
    float4 p = p4[i], m = m4[i], v = v4[i];
    adam1(p.x, g.x, m.x, v.x, a, step_size, bc2_sqrt);
    adam1(p.y, g.y, m.y, v.y, a, step_size, bc2_sqrt);
    adam1(p.z, g.z, m.z, v.z, a, step_size, bc2_sqrt);
    adam1(p.w, g.w, m.w, v.w, a, step_size, bc2_sqrt);
    p4[i] = p; m4[i] = m; v4[i] = v;
  }
  for (long long i = (n4 << 2) + tid; i < a.n; i += stride) adam1(a.p[i], __ldg(a.g + i), a.m[i], a.v[i], a, step_size, bc2_sqrt);
}

}  // namespace opt
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int32_t* step, float* scratch, void* stream) {
  HCU_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && scratch && n > 0, "adam_flat: bad arguments");
  HCU_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_flat: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st);   // [non-finite flag, grid barrier]
  if (e != cudaSuccess) { set_error("adam_flat: memset: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
  opt::AdamParams a;
  a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.step = step; a.found_inf = scratch; a.barrier = reinterpret_cast<unsigned*>(scratch + 1);
  // at most one CTA per SM: the grid barrier needs every CTA resident
  const long long want = (n / 4 + 255) / 256;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, num_sms()));
  opt::adam_flat_kernel<<<grid, 256, 0, st>>>(a);
  HCU_CHECK_LAUNCH("adam_flat");
  return 0;
}
