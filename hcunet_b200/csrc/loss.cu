// hcat/loss.py on device: pixel-weighted BCE-with-logits (fwd + bwd) and the dice / L1 / MSE
// reductions.  One memory-bound pass each way: 128-thread-per-warp-shuffle block reductions in
// fp32, fp64 atomics across CTAs, mask / pwl read in their storage dtype with the origin crop
// folded into the index arithmetic (loss.py:51-56) so no cropped copy is ever made.
#include "common.cuh"

namespace hcu {

struct LossIdx {
  int b, c, x, y, z, mx, my, mz;
  long long total;
};

__device__ __forceinline__ long long mask_index(const LossIdx& d, long long e, int& zi) {
  zi = (int)(e % d.z);
  long long r = e / d.z;
  const int yi = (int)(r % d.y); r /= d.y;
  const int xi = (int)(r % d.x);
  const long long bc = r / d.x;
  return ((bc * d.mx + xi) * d.my + yi) * d.mz + zi;
}

// (pwl + 1) evaluated in pwl's own dtype, like `loss * (pwl + 1)` does under type promotion
__device__ __forceinline__ float weight_of(float w) { return w + 1.f; }
__device__ __forceinline__ float weight_of(__half w) { return __half2float(__float2half_rn(__half2float(w) + 1.f)); }
__device__ __forceinline__ float weight_of(__nv_bfloat16 w) {
  return __bfloat162float(__float2bfloat16_rn(__bfloat162float(w) + 1.f));
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
// ATen: loss = (1 - t) * x - log_sigmoid(x),  log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float bce_logits(float x, float t) {
  return (1.f - t) * x - (fminf(x, 0.f) - log1pf(expf(-fabsf(x))));
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

template <typename TM, typename TW, bool HASW>
__global__ void wbce_fwd_kernel(LossIdx d, int mode, const float* __restrict__ pred, const TM* __restrict__ mask,
                                const TW* __restrict__ pwl, double* __restrict__ out_sum,
                                double* __restrict__ zsums) {
  extern __shared__ float zs[];  // [z] when zsums
  __shared__ float red[32];
  if (zsums != nullptr) {
    for (int i = threadIdx.x; i < d.z; i += blockDim.x) zs[i] = 0.f;
    __syncthreads();
  }
  float acc = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < d.total;
       e += (long long)gridDim.x * blockDim.x) {
    int zi;
    const long long mi = mask_index(d, e, zi);
    float x = pred[e];
    if (mode == 1) x = sigmoidf_acc(x);
    const float t = to_f(mask[mi]);
    const float w = HASW ? weight_of(pwl[mi]) : 2.f;
    const float l = bce_logits(x, t) * w;
    acc += l;
    if (zsums != nullptr) atomicAdd(&zs[zi], l);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out_sum, (double)tot);
  if (zsums != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < d.z; i += blockDim.x) atomicAdd(&zsums[i], (double)zs[i]);
  }
}

template <typename TM, typename TW, bool HASW>
__global__ void wbce_bwd_kernel(LossIdx d, int mode, const float* __restrict__ pred, const TM* __restrict__ mask,
                                const TW* __restrict__ pwl, const float* __restrict__ gout, float mult,
                                const float* __restrict__ zscale, float* __restrict__ dpred) {
  const float gs = gout[0] * mult;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < d.total;
       e += (long long)gridDim.x * blockDim.x) {
    int zi;
    const long long mi = mask_index(d, e, zi);
    const float x = pred[e];
    const float t = to_f(mask[mi]);
    const float w = HASW ? weight_of(pwl[mi]) : 2.f;
    float g;
    if (mode == 1) {
      const float s = sigmoidf_acc(x);
      g = (sigmoidf_acc(s) - t) * s * (1.f - s);
    } else {
      g = sigmoidf_acc(x) - t;
    }
    g *= w * gs;
    if (zscale != nullptr) g *= zscale[zi];
    dpred[e] = g;
  }
}

template <typename TM>
__global__ void pair_reduce_kernel(LossIdx d, int kind, const float* __restrict__ pred, const TM* __restrict__ mask,
                                   double* __restrict__ sums) {
  __shared__ float red[32];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < d.total;
       e += (long long)gridDim.x * blockDim.x) {
    int zi;
    const long long mi = mask_index(d, e, zi);
    const float x = pred[e];
    const float t = to_f(mask[mi]);
    if (kind == 0) {
      const float s = sigmoidf_acc(x);
      a0 = fmaf(s, t, a0); a1 += s; a2 += t;
    } else if (kind == 1) {
      a0 += fabsf(x - t);
    } else {
      a0 = fmaf(x - t, x - t, a0);
    }
  }
  float r0 = block_sum(a0, red);
  float r1 = 0.f, r2 = 0.f;
  if (kind == 0) { r1 = block_sum(a1, red); r2 = block_sum(a2, red); }
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], (double)r0);
    if (kind == 0) { atomicAdd(&sums[1], (double)r1); atomicAdd(&sums[2], (double)r2); }
  }
}

template <typename TM>
__global__ void pair_bwd_kernel(LossIdx d, int kind, const float* __restrict__ pred, const TM* __restrict__ mask,
                                const float* __restrict__ coef, float* __restrict__ dpred) {
  const float c0 = coef[0], c1 = coef[1];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < d.total;
       e += (long long)gridDim.x * blockDim.x) {
    int zi;
    const long long mi = mask_index(d, e, zi);
    const float x = pred[e];
    const float t = to_f(mask[mi]);
    float g;
    if (kind == 0) {
      const float s = sigmoidf_acc(x);
      g = s * (1.f - s) * fmaf(c0, t, c1);
    } else if (kind == 1) {
      const float df = x - t;
      g = df > 0.f ? c0 : (df < 0.f ? -c0 : 0.f);
    } else {
      g = c0 * (x - t);
    }
    dpred[e] = g;
  }
}

static int fill_idx(const HcuLossDesc* d, LossIdx& k) {
  HCU_CHECK_ARG(d != nullptr, "loss: null descriptor");
  HCU_CHECK_ARG(d->b > 0 && d->c > 0 && d->x > 0 && d->y > 0 && d->z > 0, "loss: empty pred");
  HCU_CHECK_ARG(d->mx >= d->x && d->my >= d->y && d->mz >= d->z, "loss: mask smaller than pred");
  k.b = d->b; k.c = d->c; k.x = d->x; k.y = d->y; k.z = d->z;
  k.mx = d->mx; k.my = d->my; k.mz = d->mz;
  k.total = (long long)d->b * d->c * d->x * d->y * d->z;
  return 0;
}

static inline int loss_grid(long long total) {
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace hcu

using namespace hcu;

extern "C" int hcu_wbce_fwd(const HcuLossDesc* d, const float* pred, const void* mask, const void* pwl,
                            double* out_sum, double* zsums, void* stream) {
  LossIdx k;
  int rc = fill_idx(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(pred && mask && out_sum, "wbce_fwd: null pointer");
  HCU_CHECK_ARG(d->mode == 0 || d->mode == 1, "wbce_fwd: mode must be 0 (pixel) or 1 (sigmoid)");
  HCU_CHECK_ARG(zsums == nullptr || d->z <= 8192, "wbce_fwd: z too large for per-z sums");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = loss_grid(k.total);
  const size_t sm = zsums ? sizeof(float) * d->z : 0;
  if (pwl == nullptr) {
    HCU_DISPATCH_DTYPE(d->dtype_mask, TM,
        wbce_fwd_kernel<TM, float, false><<<grid, 256, sm, st>>>(k, d->mode, pred, (const TM*)mask, nullptr, out_sum, zsums));
  } else {
    HCU_DISPATCH_DTYPE(d->dtype_mask, TM, HCU_DISPATCH_DTYPE(d->dtype_pwl, TW,
        wbce_fwd_kernel<TM, TW, true><<<grid, 256, sm, st>>>(k, d->mode, pred, (const TM*)mask, (const TW*)pwl, out_sum, zsums)));
  }
  HCU_CHECK_LAUNCH("wbce_fwd");
  return 0;
}

extern "C" int hcu_wbce_bwd(const HcuLossDesc* d, const float* pred, const void* mask, const void* pwl,
                            const float* gout, float mult, const float* zscale, float* dpred, void* stream) {
  LossIdx k;
  int rc = fill_idx(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(pred && mask && gout && dpred, "wbce_bwd: null pointer");
  HCU_CHECK_ARG(d->mode == 0 || d->mode == 1, "wbce_bwd: mode must be 0 (pixel) or 1 (sigmoid)");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = loss_grid(k.total);
  if (pwl == nullptr) {
    HCU_DISPATCH_DTYPE(d->dtype_mask, TM,
        wbce_bwd_kernel<TM, float, false><<<grid, 256, 0, st>>>(k, d->mode, pred, (const TM*)mask, nullptr, gout, mult, zscale, dpred));
  } else {
    HCU_DISPATCH_DTYPE(d->dtype_mask, TM, HCU_DISPATCH_DTYPE(d->dtype_pwl, TW,
        wbce_bwd_kernel<TM, TW, true><<<grid, 256, 0, st>>>(k, d->mode, pred, (const TM*)mask, (const TW*)pwl, gout, mult, zscale, dpred)));
  }
  HCU_CHECK_LAUNCH("wbce_bwd");
  return 0;
}

extern "C" int hcu_pair_reduce(const HcuLossDesc* d, int32_t kind, const float* pred, const void* mask, double* sums,
                               void* stream) {
  LossIdx k;
  int rc = fill_idx(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(pred && mask && sums && kind >= 0 && kind <= 2, "pair_reduce: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_DTYPE(d->dtype_mask, TM,
      pair_reduce_kernel<TM><<<loss_grid(k.total), 256, 0, st>>>(k, kind, pred, (const TM*)mask, sums));
  HCU_CHECK_LAUNCH("pair_reduce");
  return 0;
}

extern "C" int hcu_pair_bwd(const HcuLossDesc* d, int32_t kind, const float* pred, const void* mask,
                            const float* coef, float* dpred, void* stream) {
  LossIdx k;
  int rc = fill_idx(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(pred && mask && coef && dpred && kind >= 0 && kind <= 2, "pair_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_DTYPE(d->dtype_mask, TM,
      pair_bwd_kernel<TM><<<loss_grid(k.total), 256, 0, st>>>(k, kind, pred, (const TM*)mask, coef, dpred));
  HCU_CHECK_LAUNCH("pair_bwd");
  return 0;
}
