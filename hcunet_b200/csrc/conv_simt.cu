// Generic gather-convolution engine, SIMT (FFMA, fp32 accumulate) flavour.
//
// This is the strict-fp32 path (rel-L2 <= 1e-5 vs the reference) and the general fallback for every
// shape the tcgen05 implicit-GEMM kernel (conv_tc.cu) does not take (odd channel counts, groups
// with tiny channel slices, transposed-conv phases with ragged taps ...).
//
// One CTA computes a BM x BN tile of the implicit GEMM  out[m, n] = sum_k A[m, k] * W[k, n]
//   m = flattened (n, ox, oy, oz) output position, k = (tap, ci), n = co
// A is never materialised in global memory: K is walked in chunks of 8, each chunk is gathered
// from the channels-last input into shared memory as As[k][m] (register-staged prefetch of the next
// chunk overlaps the FMAs of the current one).  Each thread owns an 8 (pixels) x 8 (channels)
// accumulator tile.
#include "common.cuh"

namespace hcu {

struct ConvK {
  int N, IX, IY, IZ, OX, OY, OZ;
  int cin, cout, groups;
  int KX, KY, KZ, ntaps;
  int dil[3], pad[3], istep[3];
  long long in_sn, in_sx, in_sy, in_sz;  // element strides of the input tensor
  int in_c_off, in_c_gstep;
  long long out_sn, out_sx, out_sy, out_sz, out_base;  // ostep / ooff folded in
  int out_c_off;
  long long M;
  int K;
  int in_relu, out_relu, padded;
};

static int fill_convk(const HcuConvDesc* d, ConvK& k) {
  HCU_CHECK_ARG(d != nullptr, "conv: null descriptor");
  HCU_CHECK_ARG(d->batch > 0 && d->cin > 0 && d->cout > 0 && d->groups > 0, "conv: bad batch/cin/cout/groups");
  for (int i = 0; i < 3; ++i) {
    HCU_CHECK_ARG(d->in_size[i] > 0 && d->out_size[i] > 0 && d->out_tsize[i] > 0, "conv: non-positive size");
    HCU_CHECK_ARG(d->taps[i] > 0 && d->dil[i] > 0 && d->istep[i] > 0 && d->ostep[i] > 0 && d->pad[i] >= 0 &&
                      d->ooff[i] >= 0,
                  "conv: bad taps/dil/step/pad");
    HCU_CHECK_ARG((long long)(d->out_size[i] - 1) * d->ostep[i] + d->ooff[i] < d->out_tsize[i],
                  "conv: output grid exceeds output tensor in dim %d", i);
  }
  HCU_CHECK_ARG(d->in_c_off >= 0 && d->in_c_off + (long long)(d->groups - 1) * d->in_c_gstep + d->cin <= d->in_cpitch,
                "conv: input channel slice out of range");
  HCU_CHECK_ARG(d->out_c_off >= 0 && d->out_c_off + (long long)d->groups * d->cout <= d->out_cpitch,
                "conv: output channel slice out of range");
  k.N = d->batch;
  k.IX = d->in_size[0]; k.IY = d->in_size[1]; k.IZ = d->in_size[2];
  k.OX = d->out_size[0]; k.OY = d->out_size[1]; k.OZ = d->out_size[2];
  k.cin = d->cin; k.cout = d->cout; k.groups = d->groups;
  k.KX = d->taps[0]; k.KY = d->taps[1]; k.KZ = d->taps[2];
  k.ntaps = k.KX * k.KY * k.KZ;
  bool padded = false;
  for (int i = 0; i < 3; ++i) {
    k.dil[i] = d->dil[i]; k.pad[i] = d->pad[i]; k.istep[i] = d->istep[i];
    long long hi = (long long)(d->out_size[i] - 1) * d->istep[i] - d->pad[i] + (long long)(d->taps[i] - 1) * d->dil[i];
    if (d->pad[i] > 0 || hi >= d->in_size[i]) padded = true;
  }
  k.padded = padded ? 1 : 0;
  k.in_sz = d->in_cpitch;
  k.in_sy = k.in_sz * k.IZ;
  k.in_sx = k.in_sy * k.IY;
  k.in_sn = k.in_sx * k.IX;
  k.in_c_off = d->in_c_off; k.in_c_gstep = d->in_c_gstep;
  long long tz = d->out_cpitch, ty = tz * d->out_tsize[2], tx = ty * d->out_tsize[1], tn = tx * d->out_tsize[0];
  k.out_sn = tn;
  k.out_sx = tx * d->ostep[0]; k.out_sy = ty * d->ostep[1]; k.out_sz = tz * d->ostep[2];
  k.out_base = tx * d->ooff[0] + ty * d->ooff[1] + tz * d->ooff[2];
  k.out_c_off = d->out_c_off;
  k.M = (long long)k.N * k.OX * k.OY * k.OZ;
  k.K = k.ntaps * k.cin;
  k.in_relu = d->in_relu; k.out_relu = d->out_relu;
  return 0;
}

constexpr int kThreads = 128;
constexpr int KC = 8;

// ---------------------------------------------------------------------------------------------
// forward / dgrad / transposed-phase kernel
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO, int BN, int CV>
__global__ void __launch_bounds__(kThreads) conv_simt_kernel(ConvK p, const TI* __restrict__ in,
                                                             const float* __restrict__ W,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ in_scale,
                                                             const float* __restrict__ in_shift,
                                                             const float* __restrict__ out_scale,
                                                             const float* __restrict__ out_shift,
                                                             TO* __restrict__ out, double* __restrict__ stats,
                                                             int stats_pitch) {
  constexpr int NT = BN / 8;            // thread columns
  constexpr int MT = kThreads / NT;     // thread rows
  constexpr int BM = MT * 8;            // pixels per CTA
  constexpr int PPT = BM / kThreads;    // pixels per thread during the gather
  static_assert(BM % kThreads == 0, "tile");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);                 // [KC][BM]
  float* Bs = As + KC * BM;                                       // [KC][BN]
  long long* pbase = reinterpret_cast<long long*>(Bs + KC * BN);  // [BM] input base offset (or -1)
  long long* obase = pbase + BM;                                  // [BM] output offset
  int* pc = reinterpret_cast<int*>(obase + BM);                   // [BM][3] start coords (padded only)
  float* sred = reinterpret_cast<float*>(pc + 3 * BM);            // [2][BN]

  const int tid = threadIdx.x;
  const int ntn = (p.cout + BN - 1) / BN;
  const int g = blockIdx.y / ntn;
  const int n0 = (blockIdx.y % ntn) * BN;
  const long long m0 = (long long)blockIdx.x * BM;
  const float* Wg = W + (long long)g * p.K * p.cout;
  const int cbase = p.in_c_off + g * p.in_c_gstep;

  // ---- per-pixel bookkeeping ----------------------------------------------------------------
  for (int i = tid; i < BM; i += kThreads) {
    long long m = m0 + i;
    if (m < p.M) {
      int oz = (int)(m % p.OZ);
      long long r = m / p.OZ;
      int oy = (int)(r % p.OY);
      r /= p.OY;
      int ox = (int)(r % p.OX);
      int n = (int)(r / p.OX);
      int x0 = ox * p.istep[0] - p.pad[0], y0 = oy * p.istep[1] - p.pad[1], z0 = oz * p.istep[2] - p.pad[2];
      pbase[i] = n * p.in_sn + x0 * p.in_sx + y0 * p.in_sy + z0 * p.in_sz + cbase;
      obase[i] = p.out_base + n * p.out_sn + ox * p.out_sx + oy * p.out_sy + oz * p.out_sz + p.out_c_off +
                 (long long)g * p.cout;
      pc[3 * i + 0] = x0; pc[3 * i + 1] = y0; pc[3 * i + 2] = z0;
    } else {
      pbase[i] = -1;  // flagged through pc instead (pbase can legitimately be negative when padded)
      obase[i] = -1;
      pc[3 * i + 0] = -(1 << 29); pc[3 * i + 1] = 0; pc[3 * i + 2] = 0;
    }
  }
  if (tid < 2 * BN) sred[tid] = 0.f;
  __syncthreads();

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nchunks = (p.K + KC - 1) / KC;
  const int tn = tid % NT, tm = tid / NT;

  // register staging for the next chunk
  float areg[PPT * KC];
  float breg[(KC * BN + kThreads - 1) / kThreads];

  auto gather = [&](int ch) {
    const int k0 = ch * KC;
    if (CV == 4) {
      // two groups of 4 channels
      const int c4n = p.cin >> 2;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gk = (k0 >> 2) + h;
        const int tap = gk / c4n;
        const int c = (gk - tap * c4n) << 2;
        const bool kvalid = tap < p.ntaps;
        const int tz = tap % p.KZ, tq = tap / p.KZ;
        const int ty = tq % p.KY, tx = tq / p.KY;
        const int dx = tx * p.dil[0], dy = ty * p.dil[1], dz = tz * p.dil[2];
        const long long toff = dx * p.in_sx + dy * p.in_sy + dz * p.in_sz + c;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_scale != nullptr && kvalid) {
          sc = *reinterpret_cast<const float4*>(in_scale + cbase + c);
          sh = *reinterpret_cast<const float4*>(in_shift + cbase + c);
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          const int pi = tid + i * kThreads;
          const int x = pc[3 * pi + 0] + dx, y = pc[3 * pi + 1] + dy, z = pc[3 * pi + 2] + dz;
          bool ok = kvalid && x > -(1 << 28);
          if (p.padded) ok = ok && x >= 0 && x < p.IX && y >= 0 && y < p.IY && z >= 0 && z < p.IZ;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            v = load4(in + pbase[pi] + toff);
            if (in_scale != nullptr) {
              v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
              v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
              if (p.in_relu) {
                v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
                v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
              }
              // the transformed operand has the precision of the storage type (fp16 path: same arithmetic as the
              // tensor-core kernels, whose operand is the fp16-rounded BatchNorm + ReLU output)
              v.x = to_f(from_f<TI>(v.x)); v.y = to_f(from_f<TI>(v.y));
              v.z = to_f(from_f<TI>(v.z)); v.w = to_f(from_f<TI>(v.w));
            }
          }
          areg[(i * 2 + h) * 4 + 0] = v.x; areg[(i * 2 + h) * 4 + 1] = v.y;
          areg[(i * 2 + h) * 4 + 2] = v.z; areg[(i * 2 + h) * 4 + 3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int k = k0 + kk;
        const int tap = k / p.cin;
        const int c = k - tap * p.cin;
        const bool kvalid = k < p.K;
        const int tz = tap % p.KZ, tq = tap / p.KZ;
        const int ty = tq % p.KY, tx = tq / p.KY;
        const int dx = tx * p.dil[0], dy = ty * p.dil[1], dz = tz * p.dil[2];
        const long long toff = dx * p.in_sx + dy * p.in_sy + dz * p.in_sz + c;
        float sc = 1.f, sh = 0.f;
        if (in_scale != nullptr && kvalid) { sc = in_scale[cbase + c]; sh = in_shift[cbase + c]; }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          const int pi = tid + i * kThreads;
          const int x = pc[3 * pi + 0] + dx, y = pc[3 * pi + 1] + dy, z = pc[3 * pi + 2] + dz;
          bool ok = kvalid && x > -(1 << 28);
          if (p.padded) ok = ok && x >= 0 && x < p.IX && y >= 0 && y < p.IY && z >= 0 && z < p.IZ;
          float v = 0.f;
          if (ok) {
            v = to_f(in[pbase[pi] + toff]);
            if (in_scale != nullptr) {
              v = fmaf(v, sc, sh);
              if (p.in_relu) v = v < 0.f ? 0.f : v;
              v = to_f(from_f<TI>(v));
            }
          }
          areg[i * KC + kk] = v;
        }
      }
    }
    // weights: Bs[kk][n]
#pragma unroll
    for (int j = 0; j < (KC * BN + kThreads - 1) / kThreads; ++j) {
      const int e = tid + j * kThreads;
      const int kk = e / BN, n = e - kk * BN;
      float w = 0.f;
      if (e < KC * BN && k0 + kk < p.K && n0 + n < p.cout) w = Wg[(long long)(k0 + kk) * p.cout + n0 + n];
      breg[j] = w;
    }
  };

  auto commit = [&]() {
    if (CV == 4) {
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const int pi = tid + i * kThreads;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 4; ++e) As[(h * 4 + e) * BM + pi] = areg[(i * 2 + h) * 4 + e];
      }
    } else {
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const int pi = tid + i * kThreads;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) As[kk * BM + pi] = areg[i * KC + kk];
      }
    }
#pragma unroll
    for (int j = 0; j < (KC * BN + kThreads - 1) / kThreads; ++j) {
      const int e = tid + j * kThreads;
      if (e < KC * BN) Bs[e] = breg[j];
    }
  };

  gather(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    commit();
    __syncthreads();
    if (ch + 1 < nchunks) gather(ch + 1);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BM + tm * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BM + BM / 2 + tm * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * BN + tn * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * BN + tn * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue -------------------------------------------------------------------------------
  const int nb = n0 + tn * 8;  // first channel (within the group) of this thread
  float bv[8], osc[8], osh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = nb + j;
    const bool nv = n < p.cout;
    bv[j] = (bias != nullptr && nv) ? bias[g * p.cout + n] : 0.f;
    osc[j] = (out_scale != nullptr && nv) ? out_scale[g * p.cout + n] : 1.f;
    osh[j] = (out_scale != nullptr && nv) ? out_shift[g * p.cout + n] : 0.f;
  }
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }

  const bool vec_store = (p.cout % 4 == 0) && ((p.out_c_off % 4) == 0) && (p.out_sz % 4 == 0) &&
                         (p.out_sn % 4 == 0) && (p.out_base % 4 == 0) && (p.out_sx % 4 == 0) && (p.out_sy % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pi = (i < 4) ? (tm * 4 + i) : (BM / 2 + tm * 4 + (i - 4));
    const long long ob = obase[pi];
    if (ob < 0) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = acc[i][j] + bv[j];
      ssum[j] += t;
      ssq[j] = fmaf(t, t, ssq[j]);
      if (out_scale != nullptr) t = fmaf(t, osc[j], osh[j]);
      if (p.out_relu) t = t < 0.f ? 0.f : t;
      v[j] = t;
    }
    TO* op = out + ob + nb;
    if (vec_store && nb + 8 <= p.cout) {
      store4(op, make_float4(v[0], v[1], v[2], v[3]));
      store4(op + 4, make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (nb + j < p.cout) op[j] = from_f<TO>(v[j]);
    }
  }

  if (stats != nullptr) {
    // reduce over lanes that share tn (lane % NT), then over warps via shared atomics
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o >= NT; o >>= 1) {
        ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], o);
        ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
      }
    }
    const int lane = tid & 31;
    if (lane < NT) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sred[lane * 8 + j], ssum[j]);
        atomicAdd(&sred[BN + lane * 8 + j], ssq[j]);
      }
    }
    __syncthreads();
    if (tid < BN && n0 + tid < p.cout) {
      const int ch = p.out_c_off + g * p.cout + n0 + tid;
      atomicAdd(&stats[ch], (double)sred[tid]);
      atomicAdd(&stats[stats_pitch + ch], (double)sred[BN + tid]);
    }
  }
}

template <int BN>
constexpr size_t conv_smem_bytes() {
  constexpr int BM = (kThreads / (BN / 8)) * 8;
  return (size_t)(KC * BM + KC * BN) * 4 + (size_t)BM * 16 + (size_t)BM * 12 + (size_t)2 * BN * 4;
}

template <typename TI, typename TO, int BN, int CV>
static int launch_conv(const ConvK& k, const void* in, const float* W, const float* bias, const float* in_scale,
                       const float* in_shift, const float* out_scale, const float* out_shift, void* out,
                       double* stats, int stats_pitch, cudaStream_t st) {
  constexpr int BM = (kThreads / (BN / 8)) * 8;
  constexpr size_t smem = conv_smem_bytes<BN>();
  auto kern = conv_simt_kernel<TI, TO, BN, CV>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv_simt: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return HCU_ERR_CUDA;
    }
    attr_done = true;
  }
  long long mt = (k.M + BM - 1) / BM;
  int ntn = (k.cout + BN - 1) / BN;
  HCU_CHECK_ARG(mt <= 0x7fffffffLL && (long long)ntn * k.groups <= 65535, "conv_simt: grid too large");
  dim3 grid((unsigned)mt, (unsigned)(ntn * k.groups));
  kern<<<grid, kThreads, smem, st>>>(k, (const TI*)in, W, bias, in_scale, in_shift, out_scale, out_shift, (TO*)out,
                                     stats, stats_pitch);
  HCU_CHECK_LAUNCH("conv_simt");
  return 0;
}

template <typename TI, typename TO>
static int dispatch_conv(const ConvK& k, bool vec, const void* in, const float* W, const float* bias,
                         const float* in_scale, const float* in_shift, const float* out_scale,
                         const float* out_shift, void* out, double* stats, int stats_pitch, cudaStream_t st) {
#define HCU_CONV_CASE(BN)                                                                                          \
  return vec ? launch_conv<TI, TO, BN, 4>(k, in, W, bias, in_scale, in_shift, out_scale, out_shift, out, stats,   \
                                          stats_pitch, st)                                                        \
             : launch_conv<TI, TO, BN, 1>(k, in, W, bias, in_scale, in_shift, out_scale, out_shift, out, stats,   \
                                          stats_pitch, st)
  if (k.cout <= 8) { HCU_CONV_CASE(8); }
  if (k.cout <= 16) { HCU_CONV_CASE(16); }
  if (k.cout <= 32) { HCU_CONV_CASE(32); }
  HCU_CONV_CASE(64);
#undef HCU_CONV_CASE
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel: one warp per (group, tap, 8 a-channels, 8 b-channels) role, lanes walk m
// ---------------------------------------------------------------------------------------------
template <typename TA, typename TB, bool VA, bool VB>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(ConvK p, const TA* __restrict__ a,
                                                         const float* __restrict__ a_scale,
                                                         const float* __restrict__ a_shift,
                                                         const TB* __restrict__ b, int b_cpitch, int b_c_off,
                                                         float* __restrict__ partial, long long chunk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nca = (p.cin + 7) >> 3, ncb = (p.cout + 7) >> 3;
  const long long roles_per_group = (long long)p.ntaps * nca * ncb;
  const long long role = (long long)blockIdx.y * 8 + warp;
  if (role >= roles_per_group * p.groups) return;
  const int g = (int)(role / roles_per_group);
  long long r = role - (long long)g * roles_per_group;
  const int cbt = (int)(r % ncb); r /= ncb;
  const int cat = (int)(r % nca);
  const int tap = (int)(r / nca);
  const int ca0 = cat * 8, cb0 = cbt * 8;
  const int tz = tap % p.KZ, tq = tap / p.KZ;
  const int ty = tq % p.KY, tx = tq / p.KY;
  const int dx = tx * p.dil[0], dy = ty * p.dil[1], dz = tz * p.dil[2];
  const int acb = p.in_c_off + g * p.in_c_gstep + ca0;  // absolute a channel
  const int bcb = b_c_off + g * p.cout + cb0;           // absolute b channel
  const int na = min(8, p.cin - ca0), nb = min(8, p.cout - cb0);

  float asc[8], ash[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    asc[i] = (a_scale != nullptr && i < na) ? a_scale[acb + i] : 1.f;
    ash[i] = (a_scale != nullptr && i < na) ? a_shift[acb + i] : 0.f;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const long long mbeg = (long long)blockIdx.x * chunk;
  const long long mend = min(p.M, mbeg + chunk);
  long long m = mbeg + lane;
  // decode once, then advance incrementally by 32
  int oz = 0, oy = 0, ox = 0, n = 0;
  if (m < mend) {
    oz = (int)(m % p.OZ);
    long long q = m / p.OZ;
    oy = (int)(q % p.OY); q /= p.OY;
    ox = (int)(q % p.OX);
    n = (int)(q / p.OX);
  }
  for (; m < mend; m += 32) {
    const int x = ox * p.istep[0] - p.pad[0] + dx, y = oy * p.istep[1] - p.pad[1] + dy,
              z = oz * p.istep[2] - p.pad[2] + dz;
    bool ok = true;
    if (p.padded) ok = x >= 0 && x < p.IX && y >= 0 && y < p.IY && z >= 0 && z < p.IZ;
    if (ok) {
      const TA* ap = a + n * p.in_sn + x * p.in_sx + y * p.in_sy + z * p.in_sz + acb;
      const TB* bp = b + m * b_cpitch + bcb;
      float av[8], bv[8];
      if (VA) {
        float4 v0 = load4(ap);
        float4 v1 = (na > 4) ? load4(ap + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
        av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = (i < na) ? to_f(ap[i]) : 0.f;
      }
      if (a_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = fmaf(av[i], asc[i], ash[i]);
          if (p.in_relu) t = t < 0.f ? 0.f : t;
          av[i] = (i < na) ? to_f(from_f<TA>(t)) : 0.f;
        }
      }
      if (VB) {
        float4 v0 = load4(bp);
        float4 v1 = (nb > 4) ? load4(bp + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
        bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = (j < nb) ? to_f(bp[j]) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    // advance (n, ox, oy, oz) by 32 positions
    oz += 32;
    while (oz >= p.OZ) {
      oz -= p.OZ;
      if (++oy == p.OY) {
        oy = 0;
        if (++ox == p.OX) { ox = 0; ++n; }
      }
    }
  }
  // warp reduction of the 64 accumulators
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = warp_sum(acc[i][j]);
  if (lane == 0) {
    // partial[split][g][tap][ca][cb]
    float* dst = partial + ((long long)blockIdx.x * p.groups + g) * ((long long)p.K * p.cout) +
                 ((long long)tap * p.cin + ca0) * p.cout + cb0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i < na && j < nb) dst[(long long)i * p.cout + j] = acc[i][j];
  }
}

template <typename TA, typename TB>
static int dispatch_wgrad(const ConvK& k, bool va, bool vb, const void* a, const float* a_scale, const float* a_shift,
                          const void* b, int b_cpitch, int b_c_off, float* partial, int nsplit, cudaStream_t st) {
  const int nca = (k.cin + 7) >> 3, ncb = (k.cout + 7) >> 3;
  long long roles = (long long)k.ntaps * nca * ncb * k.groups;
  long long gy = (roles + 7) / 8;
  HCU_CHECK_ARG(gy <= 65535, "wgrad: too many roles (%lld)", roles);
  long long chunk = (k.M + nsplit - 1) / nsplit;
  chunk = (chunk + 31) / 32 * 32;
  dim3 grid((unsigned)nsplit, (unsigned)gy);
#define HCU_WG(VA_, VB_)                                                                                         \
  wgrad_simt_kernel<TA, TB, VA_, VB_><<<grid, 256, 0, st>>>(k, (const TA*)a, a_scale, a_shift, (const TB*)b,     \
                                                            b_cpitch, b_c_off, partial, chunk)
  if (va && vb) HCU_WG(true, true);
  else if (va) HCU_WG(true, false);
  else if (vb) HCU_WG(false, true);
  else HCU_WG(false, false);
#undef HCU_WG
  HCU_CHECK_LAUNCH("wgrad_simt");
  return 0;
}

}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_fwd(const HcuConvDesc* d, const void* in, const float* W, const float* bias,
                            const float* in_scale, const float* in_shift, const float* out_scale,
                            const float* out_shift, void* out, double* stats, void* stream) {
  ConvK k;
  int rc = fill_convk(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(in && W && out, "conv_fwd: null pointer");
  HCU_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "conv_fwd: in_scale/in_shift must come together");
  HCU_CHECK_ARG((out_scale == nullptr) == (out_shift == nullptr), "conv_fwd: out_scale/out_shift must come together");
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = d->dtype_in == HCU_F32 ? 4 : 2;
  const bool vec = (d->cin % 4 == 0) && (d->in_cpitch % 4 == 0) && (d->in_c_off % 4 == 0) &&
                   (d->in_c_gstep % 4 == 0) && (((uintptr_t)in) % (4 * esz) == 0) && (k.K % 4 == 0);
  if (d->dtype_in == HCU_F32 && d->dtype_out == HCU_F32)
    return dispatch_conv<float, float>(k, vec, in, W, bias, in_scale, in_shift, out_scale, out_shift, out, stats,
                                       d->out_cpitch, st);
  if (d->dtype_in == HCU_F16 && d->dtype_out == HCU_F16)
    return dispatch_conv<__half, __half>(k, vec, in, W, bias, in_scale, in_shift, out_scale, out_shift, out, stats,
                                         d->out_cpitch, st);
  if (d->dtype_in == HCU_F16 && d->dtype_out == HCU_F32)
    return dispatch_conv<__half, float>(k, vec, in, W, bias, in_scale, in_shift, out_scale, out_shift, out, stats,
                                        d->out_cpitch, st);
  set_error("conv_fwd: unsupported dtype combination %d -> %d", d->dtype_in, d->dtype_out);
  return HCU_ERR_UNSUPPORTED;
}

extern "C" int hcu_conv_wgrad_partial(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                      const void* b, float* partial, int32_t nsplit, void* stream) {
  ConvK k;
  int rc = fill_convk(d, k);
  if (rc) return rc;
  HCU_CHECK_ARG(a && b && partial && nsplit > 0, "wgrad: null pointer / bad nsplit");
  for (int i = 0; i < 3; ++i)
    HCU_CHECK_ARG(d->ostep[i] == 1 && d->ooff[i] == 0 && d->out_tsize[i] == d->out_size[i],
                  "wgrad: the b tensor must be dense over the output grid");
  HCU_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "wgrad: a_scale/a_shift must come together");
  cudaStream_t st = (cudaStream_t)stream;
  const int esa = d->dtype_in == HCU_F32 ? 4 : 2, esb = d->dtype_out == HCU_F32 ? 4 : 2;
  const bool va = (d->cin % 4 == 0) && (d->in_cpitch % 4 == 0) && (d->in_c_off % 4 == 0) &&
                  (d->in_c_gstep % 4 == 0) && (((uintptr_t)a) % (4 * esa) == 0);
  const bool vb = (d->cout % 4 == 0) && (d->out_cpitch % 4 == 0) && (d->out_c_off % 4 == 0) &&
                  (((uintptr_t)b) % (4 * esb) == 0);
  if (d->dtype_in == HCU_F32 && d->dtype_out == HCU_F32)
    return dispatch_wgrad<float, float>(k, va, vb, a, a_scale, a_shift, b, d->out_cpitch, d->out_c_off, partial,
                                        nsplit, st);
  if (d->dtype_in == HCU_F16 && d->dtype_out == HCU_F16)
    return dispatch_wgrad<__half, __half>(k, va, vb, a, a_scale, a_shift, b, d->out_cpitch, d->out_c_off, partial,
                                          nsplit, st);
  set_error("wgrad: unsupported dtype combination %d / %d", d->dtype_in, d->dtype_out);
  return HCU_ERR_UNSUPPORTED;
}
