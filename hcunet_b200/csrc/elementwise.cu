// Memory-bound kernels of the hot path: boundary layout transforms, weight gathers, BatchNorm
// finalize / apply / backward, fused BN+ReLU+MaxPool, max-pool backward, column sums.
// All are grid-stride kernels sized to a multiple of the SM count, 128-bit accesses where the
// channel pitch allows it, fp32 math with fp64 cross-CTA accumulation.
#include <stdarg.h>

#include <atomic>

#include <cstring>

#include <algorithm>

#include "common.cuh"

namespace hcu {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// launch geometry whose total thread count is a multiple of c, so each thread keeps ONE channel
static inline void channel_fixed_geometry(long long total, int c, int& threads, int& grid) {
  threads = 256;
  if (c <= 256 && 256 % c != 0) threads = (256 / c) * c;
  if (threads < 32) threads = 256;
  grid = grid_for(total, threads * 8, 4);
  if (c > 256 && c % 256 == 0) {
    const int q = c / 256;
    grid = (grid + q - 1) / q * q;
  }
}

// ---- layout ---------------------------------------------------------------------------------
// [N][C][S] -> [N][S][cpitch]; 32x32 smem transpose tiles over (c, s) so both sides coalesce.
template <typename TS, typename TD>
__global__ void nc_to_cl_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n, int c, long long s,
                                int cpitch, const float* __restrict__ dscale) {
  __shared__ float tile[32][33];
  const float mul = dscale != nullptr ? dscale[0] : 1.f;
  const long long s_tiles = (s + 31) / 32;
  const int c_tiles = (cpitch + 31) / 32;
  const long long total = n * s_tiles * c_tiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % c_tiles);
    long long r = t / c_tiles;
    const long long stile = r % s_tiles;
    const long long b = r / s_tiles;
    const long long s0 = stile * 32;
    const int c0 = ct * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int cc = c0 + j;
      const long long ss = s0 + threadIdx.x;
      float v = 0.f;
      if (cc < c && ss < s) v = to_f(src[(b * c + cc) * s + ss]) * mul;
      tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const long long ss = s0 + j;
      const int cc = c0 + threadIdx.x;
      if (ss < s && cc < cpitch) dst[(b * s + ss) * cpitch + cc] = from_f<TD>(tile[threadIdx.x][j]);
    }
    __syncthreads();
  }
}

// single channel: both layouts are the same array -- a cast (and the optional device scale), nothing to transpose
template <typename TS, typename TD>
__global__ void cast_scale_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long total, int cpitch,
                                  const float* __restrict__ dscale) {
  const float mul = dscale != nullptr ? dscale[0] : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    dst[i * cpitch] = from_f<TD>(to_f(src[i]) * mul);
    for (int k = 1; k < cpitch; ++k) dst[i * cpitch + k] = from_f<TD>(0.f);  // zero padding channels
  }
}

template <typename TS, typename TD>
__global__ void cl_to_nc_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n, int c, long long s,
                                int cpitch, const float* __restrict__ dscale) {
  __shared__ float tile[32][33];
  const float mul = dscale != nullptr ? dscale[0] : 1.f;
  const long long s_tiles = (s + 31) / 32;
  const int c_tiles = (c + 31) / 32;
  const long long total = n * s_tiles * c_tiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % c_tiles);
    long long r = t / c_tiles;
    const long long stile = r % s_tiles;
    const long long b = r / s_tiles;
    const long long s0 = stile * 32;
    const int c0 = ct * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const long long ss = s0 + j;
      const int cc = c0 + threadIdx.x;
      float v = 0.f;
      if (ss < s && cc < c) v = to_f(src[(b * s + ss) * cpitch + cc]) * mul;
      tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int cc = c0 + j;
      const long long ss = s0 + threadIdx.x;
      if (cc < c && ss < s) dst[(b * c + cc) * s + ss] = from_f<TD>(tile[threadIdx.x][j]);
    }
    __syncthreads();
  }
}

// small channel counts (the network input): one thread per spatial site gathers its channels (each channel read is
// coalesced along s) and writes 16-byte groups of 8 fp16 channels
template <typename TS>
__global__ void nc_to_cl_h8_kernel(const TS* __restrict__ src, __half* __restrict__ dst, long long n, int c, long long s,
                                   int cpitch, const float* __restrict__ dscale) {
  const float mul = dscale != nullptr ? dscale[0] : 1.f;
  const long long total = n * s;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / s, ss = e - b * s;
    const TS* sp = src + b * c * s + ss;
    for (int c0 = 0; c0 < cpitch; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < c) ? to_f(sp[(long long)(c0 + j) * s]) * mul : 0.f;
      __half2 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>(dst + e * cpitch + c0) = *reinterpret_cast<uint4*>(h);
    }
  }
}

// ---- weights --------------------------------------------------------------------------------
__global__ void weight_gather_kernel(HcuWeightMap m, const float* __restrict__ ref, float* __restrict__ packed,
                                     long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long idx = wm_index(m, e);
    float v = 0.f;
    if (idx >= 0) {
      v = ref[idx];
      if (m.fold) v += ref[idx + m.fold_stride];
    }
    packed[e] = v;
  }
}

__global__ void weight_scatter_kernel(HcuWeightMap m, const float* __restrict__ partial, int nsplit,
                                      long long split_stride, float scale, const float* __restrict__ dscale,
                                      int accumulate, float* __restrict__ ref, long long total) {
  if (dscale != nullptr) scale *= dscale[0];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < nsplit; ++i) s += partial[(long long)i * split_stride + e];
    s *= scale;
    const long long idx = wm_index(m, e);
    if (idx < 0) continue;
    if (accumulate) {
      ref[idx] += s;
      if (m.fold) ref[idx + m.fold_stride] += s;
    } else {
      ref[idx] = s;
      if (m.fold) ref[idx + m.fold_stride] = s;
    }
  }
}


// ---- batched weight-gradient scatter: one launch for every parameter of a step ---------------------------------
struct ScatterJob {
  HcuWeightMap m;
  long long part_off, ref_off, total;
  int nsplit, block0, nblocks, pad;
};
static_assert(sizeof(ScatterJob) <= HCU_BATCH_JOB_BYTES, "ScatterJob does not fit its table slot");
constexpr int kScatterPerBlock = 1024;

__global__ void __launch_bounds__(256) weight_scatter_batch_kernel(const unsigned char* __restrict__ jobs, int n,
                                                                   const float* __restrict__ partial, float scale,
                                                                   const float* __restrict__ dscale, float* __restrict__ grads) {
  __shared__ ScatterJob J;
  __shared__ int jidx;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const ScatterJob* pj = reinterpret_cast<const ScatterJob*>(jobs + (size_t)mid * HCU_BATCH_JOB_BYTES);
      if (pj->block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    jidx = lo;
  }
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + (size_t)jidx * HCU_BATCH_JOB_BYTES);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = threadIdx.x; i < (int)(sizeof(ScatterJob) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  if (dscale != nullptr) scale *= dscale[0];
  const float* part = partial + J.part_off;
  float* ref = grads + J.ref_off;
  const long long base = (long long)(blockIdx.x - J.block0) * kScatterPerBlock;
  for (int k = threadIdx.x; k < kScatterPerBlock; k += 256) {
    const long long e = base + k;
    if (e >= J.total) break;
    float s = 0.f;
    for (int i = 0; i < J.nsplit; ++i) s += part[(long long)i * J.total + e];
    s *= scale;
    const long long idx = wm_index32(J.m, (uint32_t)e);
    if (idx < 0) continue;
    ref[idx] = s;
    if (J.m.fold) ref[idx + J.m.fold_stride] = s;
  }
}

// ---- BatchNorm --------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int c, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean,
                                   float* invstd, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = 0; b < HCU_STAT_BINS; ++b) { s1 += stats[(size_t)b * 2 * c + i]; s2 += stats[(size_t)b * 2 * c + c + i]; }
  const double mu = s1 / count;
  double var = s2 / count - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  const float muf = (float)mu;
  mean[i] = muf;
  invstd[i] = is;
  const float sc = gamma[i] * is;
  scale[i] = sc;
  shift[i] = beta[i] - muf * sc;
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * muf;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_affine_kernel(int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      const float* __restrict__ conv_bias, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float is = 1.f / sqrtf(rv[i] + eps);
  const float sc = gamma[i] * is;
  scale[i] = sc;
  const float b = conv_bias != nullptr ? conv_bias[i] : 0.f;
  shift[i] = beta[i] + (b - rm[i]) * sc;
}

template <typename TY, typename TA, bool VEC>
__global__ void bn_relu_apply_kernel(const TY* __restrict__ y, TA* __restrict__ a, long long total, int c,
                                     const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
  if (VEC) {
    const long long nv = total >> 2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nv;
         e += (long long)gridDim.x * blockDim.x) {
      const int ch = (int)((e << 2) % c);
      float4 v = load4(y + (e << 2));
      const float4 sc = *reinterpret_cast<const float4*>(scale + ch);
      const float4 sh = *reinterpret_cast<const float4*>(shift + ch);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
      v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
      if (relu) { v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w; }
      store4(a + (e << 2), v);
    }
  } else {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
      const int ch = (int)(e % c);
      float v = fmaf(to_f(y[e]), scale[ch], shift[ch]);
      if (relu) v = v < 0.f ? 0.f : v;
      a[e] = from_f<TA>(v);
    }
  }
}

// one thread per (pooled voxel, channel); channels fastest => coalesced on both sides
template <typename TY, typename TP>
__global__ void bn_relu_maxpool_kernel(const TY* __restrict__ y, TP* __restrict__ pooled,
                                       uint8_t* __restrict__ argmax, int n, int ix, int iy, int iz, int c, int px,
                                       int py, int pz, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int relu) {
  const int ox = ix / px, oy = iy / py, oz = iz / pz;
  const long long total = (long long)n * ox * oy * oz * c;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    long long r = e / c;
    const int z = (int)(r % oz); r /= oz;
    const int yy = (int)(r % oy); r /= oy;
    const int x = (int)(r % ox);
    const int b = (int)(r / ox);
    const float sc = scale != nullptr ? scale[ch] : 1.f;
    const float sh = scale != nullptr ? shift[ch] : 0.f;
    float best = -INFINITY;
    int bi = 0;
    for (int wx = 0; wx < px; ++wx)
      for (int wy = 0; wy < py; ++wy)
        for (int wz = 0; wz < pz; ++wz) {
          const long long src = ((((long long)b * ix + (x * px + wx)) * iy + (yy * py + wy)) * iz + (z * pz + wz)) * c + ch;
          float v = to_f(y[src]);
          if (scale != nullptr) v = fmaf(v, sc, sh);
          if (relu) v = v < 0.f ? 0.f : v;  // NaN stays NaN, like ATen's relu
          // ATen max_pool3d_with_indices: take if (val > max) || isnan(val); ties keep the first
          if (v > best || v != v) {
            best = v;
            bi = (wx * py + wy) * pz + wz;
          }
        }
    pooled[e] = from_f<TP>(best);
    argmax[e] = (uint8_t)bi;
  }
}

// fp16, channels in groups of 8: one thread per (pooled voxel, 8 channels), 16-byte accesses
__global__ void bn_relu_maxpool_h8_kernel(const __half* __restrict__ y, __half* __restrict__ pooled,
                                          uint8_t* __restrict__ argmax, int n, int ix, int iy, int iz, int c, int px,
                                          int py, int pz, const float* __restrict__ scale,
                                          const float* __restrict__ shift, int relu, FastDiv dc8, FastDiv doz, FastDiv doy,
                                          FastDiv dox, uint32_t total) {
  pdl_wait();  // PDL: the launch overlapped the previous kernel's tail; nothing above touched global memory
  pdl_launch_dependents();
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    uint32_t r, ucg, uz, uy, ux, ub;
    fdivmod(e, dc8, r, ucg);
    fdivmod(r, doz, r, uz);
    fdivmod(r, doy, r, uy);
    fdivmod(r, dox, ub, ux);
    const int cg = (int)ucg, z = (int)uz, yy = (int)uy, x = (int)ux, b = (int)ub;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
    if (scale != nullptr) {
      const float4 a0 = *reinterpret_cast<const float4*>(scale + cg * 8), a1 = *reinterpret_cast<const float4*>(scale + cg * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(shift + cg * 8), b1 = *reinterpret_cast<const float4*>(shift + cg * 8 + 4);
      sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
      sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
    }
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    auto take = [&](const uint4& raw, uint32_t w) {
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        if (scale != nullptr) { f.x = fmaf(f.x, sc[2 * j], sh[2 * j]); f.y = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]); }
        if (relu) { f.x = f.x < 0.f ? 0.f : f.x; f.y = f.y < 0.f ? 0.f : f.y; }
        if (f.x > best[2 * j] || f.x != f.x) { best[2 * j] = f.x; bi[2 * j] = w; }
        if (f.y > best[2 * j + 1] || f.y != f.y) { best[2 * j + 1] = f.y; bi[2 * j + 1] = w; }
      }
    };
    const int nwin = px * py * pz;
    if (nwin <= 8) {
      // every load of the window issued before the first comparison: with one load per loop trip a full SM had ~32 KB in flight
      // and the pass ran at 2.8 TB/s (Little's law).  Same visiting order as below (x, then y, then z): same ties, same NaN rule.
      uint4 raw[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        if (w < nwin) {
          const int wz = w % pz, wq = w / pz, wy = wq % py, wx = wq / py;
          const long long src = ((((long long)b * ix + (x * px + wx)) * iy + (yy * py + wy)) * iz + (z * pz + wz)) * c + cg * 8;
          raw[w] = *reinterpret_cast<const uint4*>(y + src);
        }
      }
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (w < nwin) take(raw[w], (uint32_t)w);
    } else {
      for (int wx = 0; wx < px; ++wx)
        for (int wy = 0; wy < py; ++wy)
          for (int wz = 0; wz < pz; ++wz) {
            const long long src = ((((long long)b * ix + (x * px + wx)) * iy + (yy * py + wy)) * iz + (z * pz + wz)) * c + cg * 8;
            take(*reinterpret_cast<const uint4*>(y + src), (uint32_t)((wx * py + wy) * pz + wz));
          }
    }
    __half2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2half2_rn(best[2 * j], best[2 * j + 1]);
    *reinterpret_cast<uint4*>(pooled + (size_t)e * 8) = *reinterpret_cast<uint4*>(o);
    uint2 a;
    a.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    a.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(argmax + (size_t)e * 8) = a;
  }
}

// covers every voxel inside a pooling window; the caller zero-fills dfull first when the input has a remainder
__global__ void maxpool_bwd_h8_kernel(const __half* __restrict__ dpooled, const uint8_t* __restrict__ argmax,
                                      __half* __restrict__ dfull, int n, int ix, int iy, int iz, int c, int px, int py,
                                      int pz) {
  const int ox = ix / px, oy = iy / py, oz = iz / pz, c8 = c >> 3;
  const long long total = (long long)n * ox * oy * oz * c8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(e % c8);
    long long r = e / c8;
    const int z = (int)(r % oz); r /= oz;
    const int yy = (int)(r % oy); r /= oy;
    const int x = (int)(r % ox);
    const int b = (int)(r / ox);
    const uint4 g = *reinterpret_cast<const uint4*>(dpooled + e * 8);
    const uint2 a = *reinterpret_cast<const uint2*>(argmax + e * 8);
    const unsigned short* gh = reinterpret_cast<const unsigned short*>(&g);
    for (int wx = 0; wx < px; ++wx)
      for (int wy = 0; wy < py; ++wy)
        for (int wz = 0; wz < pz; ++wz) {
          const uint32_t w = (uint32_t)((wx * py + wy) * pz + wz);
          unsigned short o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
            o[j] = aj == w ? gh[j] : (unsigned short)0;
          }
          const long long dst = ((((long long)b * ix + (x * px + wx)) * iy + (yy * py + wy)) * iz + (z * pz + wz)) * c + cg * 8;
          *reinterpret_cast<uint4*>(dfull + dst) = *reinterpret_cast<uint4*>(o);
        }
  }
}

template <typename TDP, typename TDF>
__global__ void maxpool_bwd_kernel(const TDP* __restrict__ dpooled, const uint8_t* __restrict__ argmax,
                                   TDF* __restrict__ dfull, int n, int ix, int iy, int iz, int c, int px, int py,
                                   int pz) {
  const int ox = ix / px, oy = iy / py, oz = iz / pz;
  const long long total = (long long)n * ix * iy * iz * c;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    long long r = e / c;
    const int z = (int)(r % iz); r /= iz;
    const int yy = (int)(r % iy); r /= iy;
    const int x = (int)(r % ix);
    const int b = (int)(r / ix);
    const int qx = x / px, qy = yy / py, qz = z / pz;
    float v = 0.f;
    if (qx < ox && qy < oy && qz < oz) {
      const long long pe = ((((long long)b * ox + qx) * oy + qy) * oz + qz) * c + ch;
      const int w = ((x - qx * px) * py + (yy - qy * py)) * pz + (z - qz * pz);
      if ((int)argmax[pe] == w) v = to_f(dpooled[pe]);
    }
    dfull[e] = from_f<TDF>(v);
  }
}

// ---- fp16, 8 channels per thread (16-byte accesses); optional fused max-pool backward -------------------------
// When `argmax` is given, `da` is the gradient of the POOLED tensor: the gradient of voxel (x,y,z) is
// dpooled[window] if this voxel was the window's argmax, else 0 (nothing full-resolution is materialised).
struct PoolGeom {
  int n, ix, iy, iz, px, py, pz, ox, oy, oz;
  FastDiv diz, diy, dix, dpx, dpy, dpz;
};

__device__ __forceinline__ void load_g8(const __half* __restrict__ da, const uint8_t* __restrict__ argmax, const PoolGeom& pg,
                                        uint32_t pix, int cg, int c, float* g) {
  uint4 raw;
  if (argmax == nullptr) {
    raw = *reinterpret_cast<const uint4*>(da + (size_t)pix * c + cg * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); g[2 * j] = f.x; g[2 * j + 1] = f.y; }
    return;
  }
  uint32_t r, z, y, x, b;
  fdivmod(pix, pg.diz, r, z);
  fdivmod(r, pg.diy, r, y);
  fdivmod(r, pg.dix, b, x);
  const uint32_t qx = fdiv(x, pg.dpx), qy = fdiv(y, pg.dpy), qz = fdiv(z, pg.dpz);
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  if ((int)qx >= pg.ox || (int)qy >= pg.oy || (int)qz >= pg.oz) return;
  const size_t pe = ((((size_t)b * pg.ox + qx) * pg.oy + qy) * pg.oz + qz) * c + cg * 8;
  const uint32_t w = ((x - qx * pg.px) * pg.py + (y - qy * pg.py)) * pg.pz + (z - qz * pg.pz);
  raw = *reinterpret_cast<const uint4*>(da + pe);
  const uint2 a = *reinterpret_cast<const uint2*>(argmax + pe);
  const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
    if (aj == w) g[j] = __half2float(h[j]);
  }
}


__global__ void __launch_bounds__(256, 3) bn_bwd_stats_h8_kernel(const __half* __restrict__ da, const __half* __restrict__ y,
                                                             long long npix, int c, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int relu,
                                                             const uint8_t* __restrict__ argmax, PoolGeom pg,
                                                             double* __restrict__ sums, HcuBnBwdFin fin) {
  // [8 warps][2][c]: one slot per warp (single writer), added in a fixed order -> run-to-run reproducible sums
  extern __shared__ float sh[];
  for (int i = threadIdx.x; i < 16 * c; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  pdl_wait();  // PDL: the launch overlapped the previous kernel's tail; nothing above touched global memory
  pdl_launch_dependents();
  const int c8 = c >> 3, lc8 = __ffs(c8) - 1;  // c8 is a power of two (c8 | 256)
  const uint32_t total = (uint32_t)(npix * c8), stride = gridDim.x * blockDim.x;  // stride % c8 == 0
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(e & (uint32_t)(c8 - 1));
  // the loop accumulates sum(g) and sum(g * y) only; sum(g * xhat) = invstd * (sum(g*y) - mean * sum(g)) is formed per
  // block in fp64 when the block's partials are flushed (fewer registers -> more blocks per SM, half the arithmetic)
  float sc[8], sf[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[cg * 8 + j]; sf[j] = shift[cg * 8 + j];
    s1[j] = 0.f; s2[j] = 0.f;
  }
  auto accum = [&](const uint4& yr, const float* g) {
    const __half2* yh = reinterpret_cast<const __half2*>(&yr);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 yv = __half22float2(yh[j]);
      float g0 = g[2 * j], g1 = g[2 * j + 1];
      if (relu && fmaf(yv.x, sc[2 * j], sf[2 * j]) <= 0.f) g0 = 0.f;
      if (relu && fmaf(yv.y, sc[2 * j + 1], sf[2 * j + 1]) <= 0.f) g1 = 0.f;
      s1[2 * j] += g0; s1[2 * j + 1] += g1;
      s2[2 * j] = fmaf(g0, yv.x, s2[2 * j]);
      s2[2 * j + 1] = fmaf(g1, yv.y, s2[2 * j + 1]);
    }
  };
  // two elements per trip: four independent 16-byte loads in flight per thread
  for (; e + stride < total; e += 2 * stride) {
    const uint32_t pix0 = e >> lc8, pix1 = (e + stride) >> lc8;
    const uint4 y0 = *reinterpret_cast<const uint4*>(y + (size_t)pix0 * c + cg * 8);
    const uint4 y1 = *reinterpret_cast<const uint4*>(y + (size_t)pix1 * c + cg * 8);
    float g0[8], g1[8];
    load_g8(da, argmax, pg, pix0, cg, c, g0);
    load_g8(da, argmax, pg, pix1, cg, c, g1);
    accum(y0, g0);
    accum(y1, g1);
  }
  if (e < total) {
    const uint32_t pix = e >> lc8;
    const uint4 yr = *reinterpret_cast<const uint4*>(y + (size_t)pix * c + cg * 8);
    float g[8];
    load_g8(da, argmax, pg, pix, cg, c, g);
    accum(yr, g);
  }
  // lanes l and l' hold the same channel group when l = l' (mod c8): butterfly over the other lane bits first
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    for (int off = 16; off >= c8; off >>= 1) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
    }
  }
  if ((threadIdx.x & 31) < c8 || c8 > 16) {  // after the butterfly: one lane per channel group in this warp
    float* sw = sh + (threadIdx.x >> 5) * 2 * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sw[cg * 8 + j] = s1[j];
      sw[c + cg * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  double* sb = sums + (size_t)(blockIdx.x % HCU_STAT_BINS) * 2 * c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { q1 += sh[w * 2 * c + i]; q2 += sh[w * 2 * c + c + i]; }
    const double sg = (double)q1, sgy = (double)q2;
    atomicAdd(&sb[i], sg);
    atomicAdd(&sb[c + i], (double)invstd[i] * (sgy - (double)mean[i] * sg));
  }
  if (fin.counter != nullptr) {  // fused hcu_bn_bwd_finalize: the block that takes the last ticket sees every partial sum
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
      __threadfence();
      float gs = fin.grad_scale;
      if (fin.dscale != nullptr) gs *= fin.dscale[0];
      for (int i = threadIdx.x; i < c; i += blockDim.x)
        bn_bwd_finalize_one(sums, c, i, fin.count, fin.gamma, mean, invstd, fin.training, gs, fin.dgamma, fin.dbeta,
                            fin.dbias, fin.coef);
    }
  }
}

// Statistics pass of a POOLED layer at POOLED resolution.  The gradient of a full-resolution voxel is the pooled gradient if the
// voxel was its window's argmax, else 0 -- so both sums only need, per pooled voxel and channel, the pooled gradient and the
// ONE y value at the argmax position: a quarter of the threads of the full-resolution pass (kernel (2,2,1)), no per-voxel
// window search (bn_bwd_stats_h8_kernel with `argmax` spent ~60 instructions per full-resolution 8-channel group on index
// arithmetic: d0.conv2 of the bench step 86 us alone, 153 us beside a weight gradient).  The y values are gathered with 2-byte
// loads (each channel has its own argmax); a window's voxels share 32-byte sectors, so the DRAM traffic stays one pass over y.
__global__ void __launch_bounds__(256, 3) bn_bwd_stats_pool_h8_kernel(const __half* __restrict__ dpool, const uint8_t* __restrict__ argmax,
                                                                      const __half* __restrict__ y, long long npool, int c,
                                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                      int relu, PoolGeom pg, FastDiv doz, FastDiv doy, FastDiv dox,
                                                                      double* __restrict__ sums, HcuBnBwdFin fin) {
  extern __shared__ float sh[];                       // [8 warps][2][c] partial sums, then the window offset table
  int* woff = reinterpret_cast<int*>(sh + 16 * c);    // [px * py * pz]: element offset of window position w from the window origin
  for (int i = threadIdx.x; i < 16 * c; i += blockDim.x) sh[i] = 0.f;
  const int nwin = pg.px * pg.py * pg.pz;
  for (int w = threadIdx.x; w < nwin; w += blockDim.x) {
    const int wz = w % pg.pz, wq = w / pg.pz;
    const int wy = wq % pg.py, wx = wq / pg.py;
    woff[w] = ((wx * pg.iy + wy) * pg.iz + wz) * c;
  }
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  const int c8 = c >> 3, lc8 = __ffs(c8) - 1;
  const uint32_t total = (uint32_t)(npool * c8), stride = gridDim.x * blockDim.x;  // stride % c8 == 0
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(e & (uint32_t)(c8 - 1));
  float sc[8], sf[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[cg * 8 + j]; sf[j] = shift[cg * 8 + j];
    s1[j] = 0.f; s2[j] = 0.f;
  }
  for (; e < total; e += stride) {
    const uint32_t q = e >> lc8;                      // pooled voxel
    uint32_t r, qz, qy, qx, b;
    fdivmod(q, doz, r, qz);
    fdivmod(r, doy, r, qy);
    fdivmod(r, dox, b, qx);
    const size_t pe = (size_t)q * c + cg * 8;
    const uint4 graw = *reinterpret_cast<const uint4*>(dpool + pe);
    const uint2 a = *reinterpret_cast<const uint2*>(argmax + pe);
    const __half* yb = y + ((((size_t)b * pg.ix + qx * pg.px) * pg.iy + qy * pg.py) * pg.iz + qz * pg.pz) * c + cg * 8;
    const __half* gh = reinterpret_cast<const __half*>(&graw);
    float yv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
      yv[j] = __half2float(yb[woff[aj] + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = __half2float(gh[j]);
      if (relu && fmaf(yv[j], sc[j], sf[j]) <= 0.f) g = 0.f;
      s1[j] += g;
      s2[j] = fmaf(g, yv[j], s2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    for (int off = 16; off >= c8; off >>= 1) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
    }
  }
  if ((threadIdx.x & 31) < c8 || c8 > 16) {
    float* sw = sh + (threadIdx.x >> 5) * 2 * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sw[cg * 8 + j] = s1[j];
      sw[c + cg * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  double* sb = sums + (size_t)(blockIdx.x % HCU_STAT_BINS) * 2 * c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { q1 += sh[w * 2 * c + i]; q2 += sh[w * 2 * c + c + i]; }
    const double sg = (double)q1, sgy = (double)q2;
    atomicAdd(&sb[i], sg);
    atomicAdd(&sb[c + i], (double)invstd[i] * (sgy - (double)mean[i] * sg));
  }
  if (fin.counter != nullptr) {
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
      __threadfence();
      float gs = fin.grad_scale;
      if (fin.dscale != nullptr) gs *= fin.dscale[0];
      for (int i = threadIdx.x; i < c; i += blockDim.x)
        bn_bwd_finalize_one(sums, c, i, fin.count, fin.gamma, mean, invstd, fin.training, gs, fin.dgamma, fin.dbeta,
                            fin.dbias, fin.coef);
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_h8_kernel(const __half* __restrict__ da, const __half* __restrict__ y,
                                                             __half* __restrict__ dy, long long npix, int c,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             int relu, const float* __restrict__ coef,
                                                             const uint8_t* __restrict__ argmax, PoolGeom pg) {
  pdl_wait();  // PDL: the launch overlapped the previous kernel's tail; nothing above touched global memory
  pdl_launch_dependents();
  const int c8 = c >> 3, lc8 = __ffs(c8) - 1;
  const uint32_t total = (uint32_t)(npix * c8), stride = gridDim.x * blockDim.x;
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(e & (uint32_t)(c8 - 1));
  float sc[8], sf[8], c1[8], c2[8], c3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[cg * 8 + j]; sf[j] = shift[cg * 8 + j];
    c1[j] = coef[cg * 8 + j]; c2[j] = coef[c + cg * 8 + j]; c3[j] = coef[2 * c + cg * 8 + j];
  }
  auto apply = [&](const uint4& yr, const float* g, uint32_t pix) {
    const __half2* yh = reinterpret_cast<const __half2*>(&yr);
    __half2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 yv = __half22float2(yh[j]);
      float g0 = g[2 * j], g1 = g[2 * j + 1];
      if (relu && fmaf(yv.x, sc[2 * j], sf[2 * j]) <= 0.f) g0 = 0.f;
      if (relu && fmaf(yv.y, sc[2 * j + 1], sf[2 * j + 1]) <= 0.f) g1 = 0.f;
      o[j] = __floats2half2_rn(fmaf(c1[2 * j], g0, fmaf(c2[2 * j], yv.x, c3[2 * j])),
                               fmaf(c1[2 * j + 1], g1, fmaf(c2[2 * j + 1], yv.y, c3[2 * j + 1])));
    }
    *reinterpret_cast<uint4*>(dy + (size_t)pix * c + cg * 8) = *reinterpret_cast<uint4*>(o);
  };
  // (two pixels per trip with all four loads issued first was measured: 81 -> 101 us on d0.conv2 -- the extra registers cost
  // more resident warps than the deeper per-thread queue gained)
  for (; e < total; e += stride) {
    const uint32_t pix = e >> lc8;
    const uint4 yr = *reinterpret_cast<const uint4*>(y + (size_t)pix * c + cg * 8);
    float g[8];
    load_g8(da, argmax, pg, pix, cg, c, g);
    apply(yr, g, pix);
  }
}

// pass 1 of BN(+ReLU) backward: per-channel sums of g and g*xhat.
// Each thread walks pixels for a fixed channel group so per-thread partials stay in registers.
template <typename TD, typename TY>
__global__ void bn_bwd_stats_kernel(const TD* __restrict__ da, const TY* __restrict__ y, long long npix, int c,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                    double* __restrict__ sums) {
  // thread -> channel (tid % c) when c <= blockDim; generic: loop channels
  extern __shared__ float sh[];  // [2][c]
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const long long total = npix * c;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // make the per-thread channel constant across iterations when stride % c == 0
  const bool fixed = (stride % c) == 0;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (fixed) {
    const int ch = (int)(e % c);
    const float sc = scale[ch], sf = shift[ch], mu = mean[ch], is = invstd[ch];
    float s1 = 0.f, s2 = 0.f;
    for (; e < total; e += stride) {
      const float yv = to_f(y[e]);
      float g = to_f(da[e]);
      if (relu && fmaf(yv, sc, sf) <= 0.f) g = 0.f;
      s1 += g;
      s2 = fmaf(g, (yv - mu) * is, s2);
    }
    atomicAdd(&sh[ch], s1);
    atomicAdd(&sh[c + ch], s2);
  } else {
    for (; e < total; e += stride) {
      const int ch = (int)(e % c);
      const float yv = to_f(y[e]);
      float g = to_f(da[e]);
      if (relu && fmaf(yv, scale[ch], shift[ch]) <= 0.f) g = 0.f;
      atomicAdd(&sh[ch], g);
      atomicAdd(&sh[c + ch], g * (yv - mean[ch]) * invstd[ch]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x)
    atomicAdd(&sums[(size_t)(blockIdx.x % HCU_STAT_BINS) * 2 * c + i], (double)sh[i]);
}


__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int c, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int training, float grad_scale,
                                       const float* __restrict__ dscale, float* dgamma, float* dbeta, float* dbias,
                                       float* coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  if (dscale != nullptr) grad_scale *= dscale[0];
  bn_bwd_finalize_one(sums, c, i, count, gamma, mean, invstd, training, grad_scale, dgamma, dbeta, dbias, coef);
}

template <typename TD, typename TY, typename TO, bool VEC>
__global__ void bn_bwd_apply_kernel(const TD* __restrict__ da, const TY* __restrict__ y, TO* __restrict__ dy,
                                    long long total, int c, const float* __restrict__ scale,
                                    const float* __restrict__ shift, int relu, const float* __restrict__ coef) {
  if (VEC) {
    const long long nv = total >> 2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nv;
         e += (long long)gridDim.x * blockDim.x) {
      const int ch = (int)((e << 2) % c);
      const float4 yv = load4(y + (e << 2));
      float4 g = load4(da + (e << 2));
      const float4 sc = *reinterpret_cast<const float4*>(scale + ch);
      const float4 sf = *reinterpret_cast<const float4*>(shift + ch);
      const float4 c1 = *reinterpret_cast<const float4*>(coef + ch);
      const float4 c2 = *reinterpret_cast<const float4*>(coef + c + ch);
      const float4 c3 = *reinterpret_cast<const float4*>(coef + 2 * c + ch);
      if (relu) {
        if (fmaf(yv.x, sc.x, sf.x) <= 0.f) g.x = 0.f;
        if (fmaf(yv.y, sc.y, sf.y) <= 0.f) g.y = 0.f;
        if (fmaf(yv.z, sc.z, sf.z) <= 0.f) g.z = 0.f;
        if (fmaf(yv.w, sc.w, sf.w) <= 0.f) g.w = 0.f;
      }
      float4 o;
      o.x = fmaf(c1.x, g.x, fmaf(c2.x, yv.x, c3.x));
      o.y = fmaf(c1.y, g.y, fmaf(c2.y, yv.y, c3.y));
      o.z = fmaf(c1.z, g.z, fmaf(c2.z, yv.z, c3.z));
      o.w = fmaf(c1.w, g.w, fmaf(c2.w, yv.w, c3.w));
      store4(dy + (e << 2), o);
    }
  } else {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
      const int ch = (int)(e % c);
      const float yv = to_f(y[e]);
      float g = to_f(da[e]);
      if (relu && fmaf(yv, scale[ch], shift[ch]) <= 0.f) g = 0.f;
      dy[e] = from_f<TO>(fmaf(coef[ch], g, fmaf(coef[c + ch], yv, coef[2 * c + ch])));
    }
  }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long npix, int cpitch, int c_off, int c,
                              double* __restrict__ scratch, int nbins) {
  extern __shared__ float sh[];  // [c]
  for (int i = threadIdx.x; i < c; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const long long total = npix * c;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (stride % c == 0) {
    const int ch = (int)(e % c);
    float s = 0.f;
    for (; e < total; e += stride) s += to_f(x[(e / c) * cpitch + c_off + ch]);
    atomicAdd(&sh[ch], s);
  } else {
    for (; e < total; e += stride) atomicAdd(&sh[(int)(e % c)], to_f(x[(e / c) * cpitch + c_off + (int)(e % c)]));
  }
  __syncthreads();
  double* sb = scratch + (size_t)(blockIdx.x % nbins) * c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&sb[i], (double)sh[i]);
}

// fp16, 8-channel aligned: one 16-byte load per thread per step, a thread keeps ONE group of 8 channels (the total thread
// count is a multiple of c / 8), no division inside the loop; the CTA's partial sums go through shared memory
__global__ void __launch_bounds__(256) colsum_h8_kernel(const __half* __restrict__ x, long long npix, int cpitch, int c_off,
                                                        int c, double* __restrict__ scratch, int nbins) {
  extern __shared__ float sh[];  // [c]
  for (int i = threadIdx.x; i < c; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int g = c >> 3;                                   // 8-channel groups per pixel
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;  // multiple of g
  const int grp = (int)(tid % g);
  const long long pstep = nthr / g;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  const __half* src = x + c_off + grp * 8;
  for (long long pix = tid / g; pix < npix; pix += pstep) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + pix * cpitch));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      s[2 * k] += f.x;
      s[2 * k + 1] += f.y;
    }
  }
  // lanes holding the same group: 32 % g == 0 -> lanes l, l + g, ...; reduce across them with shuffles when g divides 32
  if (g <= 32 && (32 % g) == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      for (int o = 16; o >= g; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    if ((threadIdx.x & 31) < g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sh[grp * 8 + j], s[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[grp * 8 + j], s[j]);
  }
  __syncthreads();
  double* sb = scratch + (size_t)(blockIdx.x % nbins) * c;  // binned: same-address fp64 atomics of hundreds of CTAs serialise
  for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&sb[i], (double)sh[i]);
}

__global__ void colsum_finish_kernel(const double* __restrict__ scratch, int c, int nbins, float scale,
                                     const float* __restrict__ dscale, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (dscale != nullptr) scale *= dscale[0];
  if (i < c) {
    double sum = 0.0;
    for (int b = 0; b < nbins; ++b) sum += scratch[(size_t)b * c + i];
    out[i] = (float)(sum * scale);
  }
}

// loss scaling for the fp16 backward: scales[0] = S = 2^k with max|g| * S ~ target, scales[1] = 1/S
__global__ void absmax_kernel(const float* __restrict__ g, long long n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(g[e]);
    if (v > m && v <= 3.0e38f) m = v;  // ignore inf / NaN: they propagate on their own
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));  // non-negative floats order as uints
}
__global__ void grad_scale_finish_kernel(const unsigned int* __restrict__ amax_bits, float target, float* scales) {
  const float amax = __uint_as_float(amax_bits[0]);
  float s = 1.f;
  if (amax > 0.f) {
    int e;
    frexpf(target / amax, &e);          // target/amax = f * 2^e, f in [0.5, 1)
    e = max(-60, min(60, e - 1));
    s = ldexpf(1.f, e);
  }
  scales[0] = s;
  scales[1] = 1.f / s;
}

}  // namespace hcu

using namespace hcu;

extern "C" int hcu_grad_scale(const float* g, int64_t n, float target, unsigned int* scratch, float* scales,
                              void* stream) {
  HCU_CHECK_ARG(g && scratch && scales && n > 0 && target > 0.f, "grad_scale: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) { set_error("grad_scale: memset: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
  absmax_kernel<<<grid_for(n, 256 * 4), 256, 0, st>>>(g, n, scratch);
  HCU_CHECK_LAUNCH("absmax");
  grad_scale_finish_kernel<<<1, 1, 0, st>>>(scratch, target, scales);
  HCU_CHECK_LAUNCH("grad_scale_finish");
  return 0;
}

extern "C" int hcu_abi_version(void) { return HCU_ABI_VERSION; }
extern "C" const char* hcu_last_error(void) { return g_err; }
extern "C" long long hcu_launch_count(void) { return g_launches.load(); }

// Strided tile gather straight out of pinned (UVA-mapped) host memory: every thread moves 16-byte units, 4 in flight, so
// the PCIe reads of a whole grid are outstanding at once.  (cudaMemcpy2DAsync issues one DMA per row: measured 3.6 GB/s
// for 179 KB rows at a 512 KB pitch; this kernel is bounded by the link.)
__global__ void __launch_bounds__(256) h2d_tile_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long planes,
                                                       long long plane_pitch16, long long rows, long long row_pitch16,
                                                       long long row16) {
  const long long total = planes * rows * row16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long u0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; u0 < total; u0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long u = u0 + k * stride;
      if (u < total) {
        const long long r = u / row16, c = u - r * row16;
        const long long p = r / rows, rr = r - p * rows;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w)
                     : "l"(src + p * plane_pitch16 + rr * row_pitch16 + c));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long u = u0 + k * stride;
      if (u < total) dst[u] = v[k];
    }
  }
}

extern "C" int hcu_h2d_tile(const void* src, int64_t planes, int64_t src_plane_pitch, int64_t rows, int64_t src_row_pitch,
                            int64_t row_bytes, void* dst, void* stream) {
  HCU_CHECK_ARG(src && dst && planes > 0 && rows > 0 && row_bytes > 0 && src_row_pitch >= row_bytes, "h2d_tile: bad arguments");
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)src_plane_pitch |
                     (uintptr_t)src_row_pitch | (uintptr_t)row_bytes) & 15) == 0;
  static int use_kernel = -1;
  if (use_kernel < 0) { const char* e = getenv("HCU_H2D_KERNEL"); use_kernel = e ? atoi(e) : 1; }
  if (vec && use_kernel) {
    const long long total = planes * rows * (row_bytes / 16);
    const int grid = (int)std::min<long long>((total + 1023) / 1024, (long long)num_sms() * 8);
    h2d_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, planes, src_plane_pitch / 16, rows,
                                                            src_row_pitch / 16, row_bytes / 16);
    HCU_CHECK_LAUNCH("h2d_tile");
    return 0;
  }
  for (int64_t p = 0; p < planes; ++p) {
    cudaError_t e = cudaMemcpy2DAsync((char*)dst + p * rows * row_bytes, (size_t)row_bytes, (const char*)src + p * src_plane_pitch,
                                      (size_t)src_row_pitch, (size_t)row_bytes, (size_t)rows, cudaMemcpyHostToDevice,
                                      (cudaStream_t)stream);
    if (e != cudaSuccess) {
      set_error("h2d_tile: %s", cudaGetErrorString(e));
      return HCU_ERR_CUDA;
    }
  }
  return 0;
}

extern "C" int hcu_zero(void* ptr, size_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("hcu_zero: %s", cudaGetErrorString(e));
    return HCU_ERR_CUDA;
  }
  return 0;
}

extern "C" int hcu_nc_to_cl(const void* src, int32_t dtype_src, void* dst, int32_t dtype_dst, int64_t n, int32_t c,
                            int64_t s, int32_t cpitch, const float* dscale, void* stream) {
  HCU_CHECK_ARG(src && dst && n > 0 && c > 0 && s > 0 && cpitch >= c, "nc_to_cl: bad arguments");
  if (dtype_dst == HCU_F16 && cpitch % 8 == 0 && cpitch <= 32 && (((uintptr_t)dst) & 15) == 0) {
    const int g = grid_for(n * s, 256, 16);
    HCU_DISPATCH_DTYPE(dtype_src, TS,
        nc_to_cl_h8_kernel<TS><<<g, 256, 0, (cudaStream_t)stream>>>((const TS*)src, (__half*)dst, n, c, s, cpitch, dscale));
    HCU_CHECK_LAUNCH("nc_to_cl_h8");
    return 0;
  }
  long long tiles = n * ((s + 31) / 32) * ((cpitch + 31) / 32);
  int grid = (int)(tiles < (long long)num_sms() * 16 ? tiles : (long long)num_sms() * 16);
  dim3 block(32, 8);
  if (c == 1 && cpitch <= 8) {
    HCU_DISPATCH_DTYPE(dtype_src, TS, HCU_DISPATCH_ACT(dtype_dst, TD,
        cast_scale_kernel<TS, TD><<<grid_for(n * s, 256), 256, 0, (cudaStream_t)stream>>>((const TS*)src, (TD*)dst, n * s, cpitch,
                                                                                      dscale)));
    HCU_CHECK_LAUNCH("nc_to_cl(cast)");
    return 0;
  }
  HCU_DISPATCH_DTYPE(dtype_src, TS, HCU_DISPATCH_ACT(dtype_dst, TD,
      nc_to_cl_kernel<TS, TD><<<grid, block, 0, (cudaStream_t)stream>>>((const TS*)src, (TD*)dst, n, c, s, cpitch, dscale)));
  HCU_CHECK_LAUNCH("nc_to_cl");
  return 0;
}

extern "C" int hcu_cl_to_nc(const void* src, int32_t dtype_src, void* dst, int32_t dtype_dst, int64_t n, int32_t c,
                            int64_t s, int32_t cpitch, const float* dscale, void* stream) {
  HCU_CHECK_ARG(src && dst && n > 0 && c > 0 && s > 0 && cpitch >= c, "cl_to_nc: bad arguments");
  long long tiles = n * ((s + 31) / 32) * ((c + 31) / 32);
  int grid = (int)(tiles < (long long)num_sms() * 16 ? tiles : (long long)num_sms() * 16);
  dim3 block(32, 8);
  HCU_DISPATCH_ACT(dtype_src, TS, HCU_DISPATCH_DTYPE(dtype_dst, TD,
      cl_to_nc_kernel<TS, TD><<<grid, block, 0, (cudaStream_t)stream>>>((const TS*)src, (TD*)dst, n, c, s, cpitch, dscale)));
  HCU_CHECK_LAUNCH("cl_to_nc");
  return 0;
}

static long long wm_total(const HcuWeightMap* m) {
  const long long nph = m->phase_on ? (long long)m->ph[0] * m->ph[1] * m->ph[2] : 1;
  const long long bd = m->bdiag > 1 ? m->bdiag : 1;
  return (long long)m->groups * m->j[0] * m->j[1] * m->j[2] * m->na * m->nb * nph * bd * bd;
}

extern "C" int hcu_weight_gather(const HcuWeightMap* m, const float* ref, float* packed, void* stream) {
  HCU_CHECK_ARG(m && ref && packed, "weight_gather: null pointer");
  const long long total = wm_total(m);
  HCU_CHECK_ARG(total > 0, "weight_gather: empty map");
  weight_gather_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*m, ref, packed, total);
  HCU_CHECK_LAUNCH("weight_gather");
  return 0;
}

extern "C" int hcu_weight_scatter(const HcuWeightMap* m, const float* partial, int32_t nsplit, int64_t split_stride,
                                  float scale, const float* dscale, int32_t accumulate, float* ref, void* stream) {
  HCU_CHECK_ARG(m && ref && partial && nsplit > 0, "weight_scatter: bad arguments");
  const long long total = wm_total(m);
  HCU_CHECK_ARG(total > 0 && split_stride >= total, "weight_scatter: bad sizes");
  weight_scatter_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*m, partial, nsplit, split_stride,
                                                                                scale, dscale, accumulate, ref, total);
  HCU_CHECK_LAUNCH("weight_scatter");
  return 0;
}


extern "C" int hcu_weight_scatter_batch_build(const HcuWeightMap* maps, const int32_t* nsplit, const int64_t* part_off,
                                              const int64_t* ref_off, int32_t n, void* host_jobs, int32_t* blocks) {
  HCU_CHECK_ARG(maps && nsplit && part_off && ref_off && host_jobs && blocks && n > 0, "weight_scatter_batch_build: bad arguments");
  int b0 = 0;
  for (int i = 0; i < n; ++i) {
    ScatterJob j;
    memset(&j, 0, sizeof(j));
    j.m = maps[i]; j.part_off = part_off[i]; j.ref_off = ref_off[i]; j.total = wm_total(&maps[i]);
    HCU_CHECK_ARG(j.total > 0 && j.total < 0x7fffffffLL && nsplit[i] > 0, "weight_scatter_batch_build: job %d: bad sizes", i);
    j.nsplit = nsplit[i]; j.block0 = b0;
    j.nblocks = (int)((j.total + kScatterPerBlock - 1) / kScatterPerBlock);
    b0 += j.nblocks;
    unsigned char* slot = reinterpret_cast<unsigned char*>(host_jobs) + (size_t)i * HCU_BATCH_JOB_BYTES;
    memset(slot, 0, HCU_BATCH_JOB_BYTES);
    memcpy(slot, &j, sizeof(j));
  }
  *blocks = b0;
  return 0;
}

extern "C" int hcu_weight_scatter_batch(const void* dev_jobs, int32_t n, int32_t blocks, const float* partial, float scale,
                                        const float* dscale, float* grads, void* stream) {
  HCU_CHECK_ARG(dev_jobs && partial && grads && n > 0 && blocks > 0, "weight_scatter_batch: bad arguments");
  weight_scatter_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)dev_jobs, n, partial, scale,
                                                                        dscale, grads);
  HCU_CHECK_LAUNCH("weight_scatter_batch");
  return 0;
}

extern "C" int hcu_bn_finalize(const double* stats, int32_t c, double count, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, float* mean,
                               float* invstd, float* scale, float* shift, void* stream) {
  HCU_CHECK_ARG(stats && gamma && beta && mean && invstd && scale && shift && c > 0 && count > 0,
                "bn_finalize: bad arguments");
  HCU_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats come together");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, c, count, gamma, beta, eps, momentum,
                                                                       running_mean, running_var, mean, invstd,
                                                                       scale, shift);
  HCU_CHECK_LAUNCH("bn_finalize");
  return 0;
}

extern "C" int hcu_bn_eval_affine(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, const float* conv_bias, float* scale,
                                  float* shift, void* stream) {
  HCU_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && c > 0, "bn_eval_affine: bad args");
  bn_eval_affine_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, gamma, beta, running_mean,
                                                                          running_var, eps, conv_bias, scale, shift);
  HCU_CHECK_LAUNCH("bn_eval_affine");
  return 0;
}

extern "C" int hcu_bn_relu_apply(const void* y, int32_t dtype_y, void* a, int32_t dtype_a, int64_t npix, int32_t c,
                                 const float* scale, const float* shift, int32_t relu, void* stream) {
  HCU_CHECK_ARG(y && a && scale && shift && npix > 0 && c > 0, "bn_relu_apply: bad arguments");
  const long long total = npix * c;
  const bool vec = (c % 4 == 0) && aligned16(y) && aligned16(a) && aligned16(scale) && aligned16(shift);
  const int grid = grid_for(vec ? total / 4 : total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_ACT(dtype_y, TY, HCU_DISPATCH_ACT(dtype_a, TA, {
    if (vec) bn_relu_apply_kernel<TY, TA, true><<<grid, 256, 0, st>>>((const TY*)y, (TA*)a, total, c, scale, shift, relu);
    else bn_relu_apply_kernel<TY, TA, false><<<grid, 256, 0, st>>>((const TY*)y, (TA*)a, total, c, scale, shift, relu);
  }));
  HCU_CHECK_LAUNCH("bn_relu_apply");
  return 0;
}

extern "C" int hcu_bn_relu_maxpool(const void* y, int32_t dtype_y, void* pooled, int32_t dtype_p, uint8_t* argmax,
                                   int32_t n, int32_t ix, int32_t iy, int32_t iz, int32_t c, int32_t px, int32_t py,
                                   int32_t pz, const float* scale, const float* shift, int32_t relu, void* stream) {
  HCU_CHECK_ARG(y && pooled && argmax && n > 0 && c > 0 && px > 0 && py > 0 && pz > 0, "maxpool: bad arguments");
  HCU_CHECK_ARG(ix / px > 0 && iy / py > 0 && iz / pz > 0, "maxpool: input smaller than the pooling window");
  HCU_CHECK_ARG(px * py * pz <= 255, "maxpool: window too large for uint8 argmax");
  HCU_CHECK_ARG((scale == nullptr) == (shift == nullptr), "maxpool: scale/shift come together");
  const long long total = (long long)n * (ix / px) * (iy / py) * (iz / pz) * c;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_y == HCU_F16 && dtype_p == HCU_F16 && c % 8 == 0 && aligned16(y) && aligned16(pooled) &&
      (((uintptr_t)argmax) & 7) == 0 && (scale == nullptr || (aligned16(scale) && aligned16(shift)))) {
    const long long work = (long long)n * (ix / px) * (iy / py) * (iz / pz) * (c / 8);
    HCU_CHECK_ARG(work < 0x7fffffffLL, "bn_relu_maxpool: more than 2^31 pooled vectors");
    launch_pdl(2, bn_relu_maxpool_h8_kernel, dim3(grid_for(work, 256, 16)), dim3(256), (size_t)0, st,
               (const __half*)y, (__half*)pooled, argmax, n, ix, iy, iz, c, px, py, pz, scale, shift, relu, make_fastdiv(c / 8),
               make_fastdiv(iz / pz), make_fastdiv(iy / py), make_fastdiv(ix / px), (uint32_t)work);
    HCU_CHECK_LAUNCH("bn_relu_maxpool_h8");
    return 0;
  }
  HCU_DISPATCH_ACT(dtype_y, TY, HCU_DISPATCH_ACT(dtype_p, TP,
      bn_relu_maxpool_kernel<TY, TP><<<grid_for(total, 256), 256, 0, st>>>((const TY*)y, (TP*)pooled, argmax, n, ix,
                                                                           iy, iz, c, px, py, pz, scale, shift,
                                                                           relu)));
  HCU_CHECK_LAUNCH("bn_relu_maxpool");
  return 0;
}

extern "C" int hcu_maxpool_bwd(const void* dpooled, int32_t dtype_dp, const uint8_t* argmax, void* dfull,
                               int32_t dtype_df, int32_t n, int32_t ix, int32_t iy, int32_t iz, int32_t c, int32_t px,
                               int32_t py, int32_t pz, void* stream) {
  HCU_CHECK_ARG(dpooled && argmax && dfull && n > 0 && c > 0 && px > 0 && py > 0 && pz > 0, "maxpool_bwd: bad args");
  const long long total = (long long)n * ix * iy * iz * c;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_dp == HCU_F16 && dtype_df == HCU_F16 && c % 8 == 0 && aligned16(dpooled) && aligned16(dfull) &&
      (((uintptr_t)argmax) & 7) == 0) {
    if (ix % px || iy % py || iz % pz) {
      cudaError_t e = cudaMemsetAsync(dfull, 0, (size_t)total * 2, st);
      if (e != cudaSuccess) { set_error("maxpool_bwd: memset: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
    }
    const long long work = (long long)n * (ix / px) * (iy / py) * (iz / pz) * (c / 8);
    maxpool_bwd_h8_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>((const __half*)dpooled, argmax, (__half*)dfull, n, ix, iy,
                                                                   iz, c, px, py, pz);
    HCU_CHECK_LAUNCH("maxpool_bwd_h8");
    return 0;
  }
  HCU_DISPATCH_ACT(dtype_dp, TDP, HCU_DISPATCH_ACT(dtype_df, TDF,
      maxpool_bwd_kernel<TDP, TDF><<<grid_for(total, 256), 256, 0, st>>>((const TDP*)dpooled, argmax, (TDF*)dfull, n,
                                                                         ix, iy, iz, c, px, py, pz)));
  HCU_CHECK_LAUNCH("maxpool_bwd");
  return 0;
}

static int fill_pool(const HcuPoolGeom* g, int64_t npix, PoolGeom& pg, const char* who) {
  HCU_CHECK_ARG(g != nullptr && g->n > 0 && g->px > 0 && g->py > 0 && g->pz > 0, "%s: bad pool geometry", who);
  HCU_CHECK_ARG((int64_t)g->n * g->ix * g->iy * g->iz == npix, "%s: pool geometry does not match npix", who);
  pg.n = g->n; pg.ix = g->ix; pg.iy = g->iy; pg.iz = g->iz; pg.px = g->px; pg.py = g->py; pg.pz = g->pz;
  pg.ox = g->ix / g->px; pg.oy = g->iy / g->py; pg.oz = g->iz / g->pz;
  pg.diz = make_fastdiv(g->iz); pg.diy = make_fastdiv(g->iy); pg.dix = make_fastdiv(g->ix);
  pg.dpx = make_fastdiv(g->px); pg.dpy = make_fastdiv(g->py); pg.dpz = make_fastdiv(g->pz);
  return 0;
}

extern "C" int hcu_bn_bwd_finalize(const double* sums, int32_t c, double count, const float* gamma, const float* mean,
                                   const float* invstd, int32_t training, float grad_scale, const float* dscale,
                                   float* dgamma, float* dbeta, float* dbias, float* coef, void* stream);

static int bn_bwd_stats_impl(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, int64_t npix,
                             int32_t c, const float* scale, const float* shift, const float* mean,
                             const float* invstd, int32_t relu, const uint8_t* argmax, const HcuPoolGeom* pool,
                             double* sums, const HcuBnBwdFin* fin, void* stream) {
  HCU_CHECK_ARG(da && y && scale && shift && mean && invstd && sums && npix > 0 && c > 0, "bn_bwd_stats: bad args");
  HCU_CHECK_ARG(c <= 4096, "bn_bwd_stats: too many channels");
  if (dtype_da == HCU_F16 && dtype_y == HCU_F16 && c % 8 == 0 && c <= 512 && 256 % (c / 8) == 0 && aligned16(da) && aligned16(y) &&
      npix * (c / 8) < 0x7fffffffLL && (argmax == nullptr || (((uintptr_t)argmax) & 7) == 0)) {
    PoolGeom pg = {};
    if (argmax != nullptr) { int rc = fill_pool(pool, npix, pg, "bn_bwd_stats"); if (rc) return rc; }
    HcuBnBwdFin f;
    memset(&f, 0, sizeof(f));
    if (fin != nullptr) f = *fin;
    static int pooled_on = -1;
    if (pooled_on < 0) { const char* e = getenv("HCU_BN_POOLED_STATS"); pooled_on = e ? atoi(e) : 1; }
    if (argmax != nullptr && pooled_on && pg.px * pg.py * pg.pz <= 256 && (((uintptr_t)argmax) & 7) == 0) {
      // pooled layer: the sums over the full-resolution gradient only need the pooled gradient and y at the argmax positions
      const long long npool = (long long)pg.n * pg.ox * pg.oy * pg.oz;
      const int grid = grid_for(npool * (c / 8), 256 * 4, 12);
      const size_t smem = 16 * c * sizeof(float) + (size_t)pg.px * pg.py * pg.pz * sizeof(int);
      launch_pdl(2, bn_bwd_stats_pool_h8_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, (const __half*)da, argmax,
                 (const __half*)y, npool, c, scale, shift, mean, invstd, relu, pg, make_fastdiv((uint32_t)pg.oz),
                 make_fastdiv((uint32_t)pg.oy), make_fastdiv((uint32_t)pg.ox), sums, f);
      HCU_CHECK_LAUNCH("bn_bwd_stats_pool_h8");
      return 0;
    }
    const int grid = grid_for(npix * (c / 8), 256 * 4, 12);
    launch_pdl(2, bn_bwd_stats_h8_kernel, dim3(grid), dim3(256), 16 * c * sizeof(float), (cudaStream_t)stream,
               (const __half*)da, (const __half*)y, npix, c, scale, shift, mean, invstd, relu, argmax, pg, sums, f);
    HCU_CHECK_LAUNCH("bn_bwd_stats_h8");
    return 0;
  }
  HCU_CHECK_ARG(argmax == nullptr, "bn_bwd_stats: the fused max-pool backward needs fp16 tensors with c %% 8 == 0");
  const long long total = npix * c;
  int threads, grid;
  channel_fixed_geometry(total, c, threads, grid);
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_ACT(dtype_da, TD, HCU_DISPATCH_ACT(dtype_y, TY,
      bn_bwd_stats_kernel<TD, TY><<<grid, threads, 2 * c * sizeof(float), st>>>((const TD*)da, (const TY*)y, npix, c,
                                                                                 scale, shift, mean, invstd, relu,
                                                                                 sums)));
  HCU_CHECK_LAUNCH("bn_bwd_stats");
  if (fin != nullptr)
    return hcu_bn_bwd_finalize(sums, c, fin->count, fin->gamma, mean, invstd, fin->training, fin->grad_scale, fin->dscale,
                               fin->dgamma, fin->dbeta, fin->dbias, fin->coef, stream);
  return 0;
}

extern "C" int hcu_bn_bwd_stats(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, int64_t npix,
                                int32_t c, const float* scale, const float* shift, const float* mean,
                                const float* invstd, int32_t relu, const uint8_t* argmax, const HcuPoolGeom* pool,
                                double* sums, void* stream) {
  return bn_bwd_stats_impl(da, dtype_da, y, dtype_y, npix, c, scale, shift, mean, invstd, relu, argmax, pool, sums, nullptr,
                           stream);
}

extern "C" int hcu_bn_bwd_stats_fin(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, int64_t npix,
                                    int32_t c, const float* scale, const float* shift, const float* mean,
                                    const float* invstd, int32_t relu, const uint8_t* argmax, const HcuPoolGeom* pool,
                                    double* sums, const HcuBnBwdFin* fin, void* stream) {
  HCU_CHECK_ARG(fin && fin->gamma && fin->coef && fin->counter && fin->count > 0, "bn_bwd_stats_fin: bad finalize arguments");
  return bn_bwd_stats_impl(da, dtype_da, y, dtype_y, npix, c, scale, shift, mean, invstd, relu, argmax, pool, sums, fin,
                           stream);
}

extern "C" int hcu_bn_bwd_finalize(const double* sums, int32_t c, double count, const float* gamma, const float* mean,
                                   const float* invstd, int32_t training, float grad_scale, const float* dscale,
                                   float* dgamma, float* dbeta, float* dbias, float* coef, void* stream) {
  HCU_CHECK_ARG(sums && gamma && mean && invstd && coef && c > 0 && count > 0, "bn_bwd_finalize: bad args");
  bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, c, count, gamma, mean, invstd,
                                                                           training, grad_scale, dscale, dgamma, dbeta,
                                                                           dbias, coef);
  HCU_CHECK_LAUNCH("bn_bwd_finalize");
  return 0;
}

extern "C" int hcu_bn_bwd_apply(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, void* dy,
                                int32_t dtype_dy, int64_t npix, int32_t c, const float* scale, const float* shift,
                                int32_t relu, const float* coef, const uint8_t* argmax, const HcuPoolGeom* pool,
                                void* stream) {
  HCU_CHECK_ARG(da && y && dy && scale && shift && coef && npix > 0 && c > 0, "bn_bwd_apply: bad args");
  if (dtype_da == HCU_F16 && dtype_y == HCU_F16 && dtype_dy == HCU_F16 && c % 8 == 0 && 256 % (c / 8) == 0 && aligned16(da) &&
      aligned16(y) && aligned16(dy) && npix * (c / 8) < 0x7fffffffLL && (argmax == nullptr || (((uintptr_t)argmax) & 7) == 0)) {
    PoolGeom pg = {};
    if (argmax != nullptr) { int rc = fill_pool(pool, npix, pg, "bn_bwd_apply"); if (rc) return rc; }
    const int grid = grid_for(npix * (c / 8), 256 * 2, 16);
    launch_pdl(2, bn_bwd_apply_h8_kernel, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)stream, (const __half*)da, (const __half*)y,
               (__half*)dy, npix, c, scale, shift, relu, coef, argmax, pg);
    HCU_CHECK_LAUNCH("bn_bwd_apply_h8");
    return 0;
  }
  HCU_CHECK_ARG(argmax == nullptr, "bn_bwd_apply: the fused max-pool backward needs fp16 tensors with c %% 8 == 0");
  const long long total = npix * c;
  const bool vec = (c % 4 == 0) && aligned16(da) && aligned16(y) && aligned16(dy) && aligned16(scale) &&
                   aligned16(shift) && aligned16(coef);
  const int grid = grid_for(vec ? total / 4 : total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_ACT(dtype_da, TD, HCU_DISPATCH_ACT(dtype_y, TY, HCU_DISPATCH_ACT(dtype_dy, TO, {
    if (vec) bn_bwd_apply_kernel<TD, TY, TO, true><<<grid, 256, 0, st>>>((const TD*)da, (const TY*)y, (TO*)dy, total, c, scale, shift, relu, coef);
    else bn_bwd_apply_kernel<TD, TY, TO, false><<<grid, 256, 0, st>>>((const TD*)da, (const TY*)y, (TO*)dy, total, c, scale, shift, relu, coef);
  })));
  HCU_CHECK_LAUNCH("bn_bwd_apply");
  return 0;
}

extern "C" int hcu_colsum(const void* x, int32_t dtype_x, int64_t npix, int32_t cpitch, int32_t c_off, int32_t c,
                          float scale, const float* dscale, double* scratch, float* out, void* stream) {
  HCU_CHECK_ARG(x && scratch && out && npix > 0 && c > 0 && c_off >= 0 && c_off + c <= cpitch && c <= 4096,
                "colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int nbins = std::max(1, std::min(32, 4096 / c));  // the caller's scratch holds 4096 doubles
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * c * nbins, st);
  if (e != cudaSuccess) { set_error("colsum: memset: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
  int threads, grid;
  const int g8 = c / 8;
  if (dtype_x == HCU_F16 && c % 8 == 0 && cpitch % 8 == 0 && c_off % 8 == 0 && aligned16(x) && g8 <= 256 &&
      (g8 & (g8 - 1)) == 0) {  // power-of-two group count: divides the 256 threads (and a warp, up to 32 groups)
    threads = 256;
    grid = grid_for(npix * g8, threads * 8, 2);
    colsum_h8_kernel<<<grid, threads, c * sizeof(float), st>>>((const __half*)x, npix, cpitch, c_off, c, scratch, nbins);
  } else {
    channel_fixed_geometry(npix * c, c, threads, grid);
    HCU_DISPATCH_ACT(dtype_x, T,
        colsum_kernel<T><<<grid, threads, c * sizeof(float), st>>>((const T*)x, npix, cpitch, c_off, c, scratch, nbins));
  }
  HCU_CHECK_LAUNCH("colsum");
  colsum_finish_kernel<<<(c + 127) / 128, 128, 0, st>>>(scratch, c, nbins, scale, dscale, out);
  HCU_CHECK_LAUNCH("colsum_finish");
  return 0;
}
