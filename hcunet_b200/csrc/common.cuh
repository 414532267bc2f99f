// Shared helpers for the hcunet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hcunet_b200.h"

namespace hcu {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();

#define HCU_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      hcu::set_error(__VA_ARGS__);        \
      return HCU_ERR_INVALID;             \
    }                                     \
  } while (0)

// to be used right after a kernel launch
#define HCU_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaPeekAtLastError();                                     \
    if (e__ != cudaSuccess) {                                                    \
      cudaGetLastError();                                                        \
      hcu::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return HCU_ERR_CUDA;                                                       \
    }                                                                            \
    hcu::count_launch();                                                         \
  } while (0)

// ---- dtype helpers --------------------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float> { static constexpr int id = HCU_F32; };
template <> struct DT<__nv_bfloat16> { static constexpr int id = HCU_BF16; };
template <> struct DT<__half> { static constexpr int id = HCU_F16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 4 consecutive elements <-> float4 (pointer must be 4-element aligned)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 load4(const __half* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&r.x);
  __half2 b = *reinterpret_cast<__half2*>(&r.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ void store4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y);
  __half2 b = __floats2half2_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// element e of a packed [g][jx][jy][jz][a][b] weight tensor -> index into the reference-layout parameter
// (-1: a structural zero of a block-diagonal map)
__device__ __forceinline__ long long wm_index(const HcuWeightMap& m, long long e) {
  const int nph = m.phase_on ? m.ph[0] * m.ph[1] * m.ph[2] : 1;
  const int nbf = (m.phase_on == 2 ? m.nb * nph : m.nb) * (m.bdiag > 1 ? m.bdiag : 1),
            naf = (m.phase_on == 1 ? m.na * nph : m.na) * (m.bdiag > 1 ? m.bdiag : 1);
  int b = (int)(e % nbf); e /= nbf;
  int a = (int)(e % naf); e /= naf;
  const int jz = (int)(e % m.j[2]); e /= m.j[2];
  const int jy = (int)(e % m.j[1]); e /= m.j[1];
  const int jx = (int)(e % m.j[0]);
  int g = (int)(e / m.j[0]);
  long long idx = m.base;
  if (m.bdiag > 1) {  // dense view of a grouped weight: off-diagonal blocks do not exist in the reference tensor
    g = a / m.na;
    if (b / m.nb != g) return -1;
    a -= g * m.na; b -= g * m.nb;
  }
  if (m.phase_on) {
    int phi;
    if (m.phase_on == 1) { phi = a / m.na; a -= phi * m.na; } else { phi = b / m.nb; b -= phi * m.nb; }
    const int pz = phi % m.ph[2]; phi /= m.ph[2];
    const int py = phi % m.ph[1], px = phi / m.ph[1];
    idx += px * m.pst[0] + py * m.pst[1] + pz * m.pst[2];
  }
  return idx + g * m.sg + a * m.sa + b * m.sb + (long long)(m.t0[0] + jx * m.tstep[0]) * m.st[0] +
         (long long)(m.t0[1] + jy * m.tstep[1]) * m.st[1] + (long long)(m.t0[2] + jz * m.tstep[2]) * m.st[2];
}

// 32-bit index arithmetic flavour (the batched kernels: element counts < 2^31; 64-bit '/' '%' cost ~10x more)
// element e of a packed [g][jx][jy][jz][a][b] weight tensor -> index into the reference-layout parameter
__device__ __forceinline__ long long wm_index32(const HcuWeightMap& m, uint32_t e) {
  const int nph = m.phase_on ? m.ph[0] * m.ph[1] * m.ph[2] : 1;
  const int nbf = (m.phase_on == 2 ? m.nb * nph : m.nb) * (m.bdiag > 1 ? m.bdiag : 1),
            naf = (m.phase_on == 1 ? m.na * nph : m.na) * (m.bdiag > 1 ? m.bdiag : 1);
  int b = (int)(e % nbf); e /= nbf;
  int a = (int)(e % naf); e /= naf;
  const int jz = (int)(e % m.j[2]); e /= m.j[2];
  const int jy = (int)(e % m.j[1]); e /= m.j[1];
  const int jx = (int)(e % m.j[0]);
  int g = (int)(e / m.j[0]);
  long long idx = m.base;
  if (m.bdiag > 1) {  // dense view of a grouped weight: off-diagonal blocks do not exist in the reference tensor
    g = a / m.na;
    if (b / m.nb != g) return -1;
    a -= g * m.na; b -= g * m.nb;
  }
  if (m.phase_on) {
    int phi;
    if (m.phase_on == 1) { phi = a / m.na; a -= phi * m.na; } else { phi = b / m.nb; b -= phi * m.nb; }
    const int pz = phi % m.ph[2]; phi /= m.ph[2];
    const int py = phi % m.ph[1], px = phi / m.ph[1];
    idx += px * m.pst[0] + py * m.pst[1] + pz * m.pst[2];
  }
  return idx + g * m.sg + a * m.sa + b * m.sb + (long long)(m.t0[0] + jx * m.tstep[0]) * m.st[0] +
         (long long)(m.t0[1] + jy * m.tstep[1]) * m.st[1] + (long long)(m.t0[2] + jz * m.tstep[2]) * m.st[2];
}

// Division of a 31-bit unsigned by a runtime constant in 3 instructions (64-bit '/' and '%' cost ~100 each and made the
// index decomposition of the memory-bound kernels the bottleneck): q = (umulhi(n, m) + n) >> s, valid for n < 2^31.
struct FastDiv {
  uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.s = s;
  f.m = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.m) + n) >> f.s; }
__device__ __forceinline__ void fdivmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
  q = fdiv(n, f);
  r = n - q * f.d;
}

// BatchNorm(+ReLU) backward, per-channel finalize from the binned fp64 sums (sum g, sum g*xhat): dgamma, dbeta, the conv-bias
// gradient and the coefficients of dy = c1*g + c2*y + c3.  Shared by bn_bwd_stats' fused tail (elementwise.cu) and by the
// data-gradient convolution that computes the statistics in its epilogue (conv_tc_kernel.cuh).
__device__ __forceinline__ void bn_bwd_finalize_one(const double* sums, int c, int i, double count, const float* gamma,
                                                    const float* mean, const float* invstd, int training, float grad_scale,
                                                    float* dgamma, float* dbeta, float* dbias, float* coef) {
  double sg = 0.0, sgx = 0.0;
  for (int b = 0; b < HCU_STAT_BINS; ++b) {
    sg += __ldcg(&sums[(size_t)b * 2 * c + i]);
    sgx += __ldcg(&sums[(size_t)b * 2 * c + c + i]);
  }
  if (dgamma != nullptr) dgamma[i] = (float)(sgx * grad_scale);
  if (dbeta != nullptr) dbeta[i] = (float)(sg * grad_scale);
  const double s = (double)gamma[i] * (double)invstd[i];
  double c1, c2, c3;
  if (training) {
    const double mg = sg / count, mgx = sgx / count;
    c1 = s;
    c2 = -s * (double)invstd[i] * mgx;
    c3 = s * (double)invstd[i] * mgx * (double)mean[i] - s * mg;
    // conv bias feeds a batch-stat BN: its gradient sum(dy) is analytically zero
    if (dbias != nullptr) dbias[i] = (float)((c1 * sg + c2 * count * (double)mean[i] + c3 * count) * grad_scale);
  } else {
    c1 = s; c2 = 0.0; c3 = 0.0;
    if (dbias != nullptr) dbias[i] = (float)(s * sg * grad_scale);
  }
  coef[i] = (float)c1;
  coef[c + i] = (float)c2;
  coef[2 * c + i] = (float)c3;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A kernel launched with launch_pdl() may START while its predecessor in the stream is still running: everything it does
// before pdl_wait() (barrier init, TMEM allocation, index tables built from the kernel parameters -- no global memory) overlaps
// the predecessor's tail; pdl_wait() returns once the predecessor grid has completed and its writes are visible.  Every such
// kernel calls pdl_launch_dependents() only AFTER its own pdl_wait(): by induction everything older than the immediate
// predecessor is complete when a kernel starts.  Works inside CUDA-graph capture (programmatic dependency edges).
// Without the launch attribute pdl_wait() is a no-op.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// HCU_PDL is a bit mask: 1 = the tensor-core conv kernels (default), 2 = the BatchNorm-backward / max-pool kernels.
// Measured on the bench step (README 3D model, CUDA graph): convs only 3.19 -> 3.13 ms; with the memory-bound kernels as
// well 3.26 ms -- their early-resident CTAs spin in pdl_wait() on slots the side stream's weight-gradient CTAs would use.
inline int pdl_mask() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("HCU_PDL"); on = e ? atoi(e) : 1; }
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// dispatch helpers on HcuDType
#define HCU_DISPATCH_DTYPE(dt, T, ...)                         \
  switch (dt) {                                                \
    case HCU_F32: { using T = float; __VA_ARGS__; } break;     \
    case HCU_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    case HCU_F16: { using T = __half; __VA_ARGS__; } break;    \
    default: hcu::set_error("bad dtype %d", (int)(dt)); return HCU_ERR_INVALID; \
  }
#define HCU_DISPATCH_ACT(dt, T, ...)                           \
  switch (dt) {                                                \
    case HCU_F32: { using T = float; __VA_ARGS__; } break;     \
    case HCU_F16: { using T = __half; __VA_ARGS__; } break;    \
    default: hcu::set_error("activation dtype must be f32 or f16, got %d", (int)(dt)); return HCU_ERR_INVALID; \
  }

}  // namespace hcu
