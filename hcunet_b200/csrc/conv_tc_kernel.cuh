// conv_tc_kernel: the x-march implicit-GEMM convolution kernel (tcgen05.mma / TMEM) and its launch parameters.  The kernel is a
// template over compile-time variants (struct Var below); conv_tc.cu instantiates the generic one, conv_tc_mb{1..4}.cu the
// specialised ones (one translation unit per M-block count so that they compile in parallel).  See conv_tc.cu for the
// formulation ("flat shift"), the roles and the shared packed weight layout.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace hcu {

namespace tc {
using namespace ptx;

constexpr int kIssuers = 1;     // MMA-issuing warps (each issues the M-blocks mb = w, w + kIssuers, ...)
constexpr int kThreads = 256 + 32 * kIssuers + 32;  // warps 0-3 epilogue, 4-7 producers (bulk mode: epilogue too), 8 MMA, 9 bulk-copy producer
constexpr int kMaxPairs = 64;   // (ty, tz, channel-plane) K16 steps per tx
constexpr int kMaxRing = 12;
constexpr int kMaxTab = 160;    // KX * npairs K16 steps of one output plane (constant-bank table in the kernel parameters)
constexpr int kSmemLimit = 227 * 1024;

struct Params {
  const __half* in;
  const __half* wp;  // packed weights [nsplit][E][Nc][8] fp16
  void* out;
  const float* bias;
  const float* in_scale;
  const float* in_shift;
  const float* out_scale;
  const float* out_shift;
  double* stats;
  int stats_pitch;
  HcuBnFin fin;  // optional fused BatchNorm finalize (fin.counter == nullptr: none)
  // Fused BatchNorm(+ReLU)-backward statistics (data-gradient launches, flag bit 16): this launch's output is the gradient g
  // with respect to a = relu(bn(y)) of the PREVIOUS layer; its epilogue also reads that layer's raw output y (same shape as
  // the output) and accumulates sum(g') and sum(g' * y), g' = g where bn(y) > 0 else 0 -- what hcu_bn_bwd_stats would read g
  // and y again for -- then the per-channel finalize runs in the CTA that takes the last ticket.
  struct BnBwdFuse {
    const __half* y;
    const float* scale; const float* shift; const float* mean; const float* invstd;
    double* sums;    // binned [HCU_STAT_BINS][2][c]
    HcuBnBwdFin fin;
  } bnb;
  int N, IX, IY, IZ, Cp, P;
  int OX, OY, OZ;
  int KX, KY, KZ, dx, dy, dz, px, py, pz;
  int Yv, Zv;
  int M, MB, RUN, PS, SLOT, R;
  int Nc, nsplit, cout, E_tx, npairs, E;
  int n_runs, Lx, n_xseg;
  long long out_sn, out_sx, out_sy, out_sz, out_base;
  int out_c_off, out_f32, in_relu, out_relu;
  // input addressing (elements): coarse strides + per-phase offsets (iphase folds stride phases into the channel planes)
  long long in_ns, in_xs;
  int in_ys, in_zs, Pc, ips[3];
  long long in_ph[3];
  // output stride phases (ophase): cout = [nph][cpp]
  int ops[3], cpp;
  long long out_ph[3];
  // shared memory carve-up (bytes from the 128-aligned base)
  int off_w, off_a, off_bar, off_tab, off_stat, smem_bytes;
  int tmem_cols;
  int wide;      // x-fused MMAs: one MMA feeds up to KX output planes (N = w * Nc), 4 accumulator slots per M-block
  int D;         // producer look-ahead in planes (loads in flight beyond the published planes)
  int epi_fast;  // fp16 output, 8-channel aligned, no stride phases: vector epilogue
  // K16 step (tx, e): x = (byte offset of the first K8 slab inside a ring slot >> 4) | (offset of the second slab >> 4) << 16
  //                   y = byte offset of the B tile inside the packed weights >> 4
  uint2 tab[kMaxTab];
  int debug;  // HCU_TC_DEBUG bits (profiling experiments only): 1 no global loads, 2 no epilogue math/stores, 4 no MMAs
  // ---- K-streamed kernel (conv_ks_kernel, the channel-rich levels) ----
  int ks;        // this descriptor runs on conv_ks_kernel
  int PC, NCH;   // channel planes per A stage / B tile, chunks = P / PC
  int RA, RB;    // ring depths of the A stages and the B tiles
  int BT;        // bytes of one B tile (one tap of one chunk) = PC * Nc * 16
  int TB;        // B tiles (consecutive taps of a chunk) per ring slot / barrier round
  int n_last;    // flat positions of one x-plane that hold outputs (all images stacked: rows = N * Yv)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------

// Bounded wait: a protocol bug traps (kernel error) after ~2 s instead of hanging the GPU.  The suspend-time hint lets
// the hardware park the thread until the phase completes, so waiting warps do not burn issue slots.



// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, M=128, K=16

// K-major, no swizzle: rows of one 8x(16 B) core matrix are 16 B apart; SBO = next 8 rows, LBO = next 8 K


// 8 accumulator columns of this thread's TMEM lane, no wait: pair with tmem_wait_ld() + tmem_pin8()
// orders every later use of r[0..7] after the preceding (volatile) wait

// sum over the 32 lanes of 8 per-lane values; every lane of a group of 4 (lane >> 2) ends up with channel lane >> 2



// sum over the 32 lanes of 16 per-lane values; lane l ends up with channel (l >> 1) & 15 (both lanes of a pair)



// Stage timing (compile with -DHCU_TC_PROF; experiments only): cycles spent waiting at each barrier vs. total per role
#ifdef HCU_TC_PROF
#define PROF_DECL long long pw0 = 0, pw1 = 0, pw2 = 0, pt0 = clock64()
#define PROF_WAIT(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#define PROF_REPORT(role, n0, n1, iters)                                                                   \
  if (blockIdx.x == gridDim.x / 2 && lane == 0)                                                            \
    printf("conv_tc prof %-8s warp %d: total %lld cyc, %s %lld, %s %lld, fence %lld, iters %d -> %lld cyc/iter busy\n", role, warp, \
           clock64() - pt0, n0, pw0, n1, pw1, pw2, iters, (clock64() - pt0 - pw0 - pw1) / max(1, iters))
// CTA timeline: global-timer stamps (ns) at kernel entry, after the prologue and at exit, for a sample of CTAs
__device__ __forceinline__ unsigned long long prof_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PROF_CTA_DECL const unsigned long long pg0 = prof_gtime(); unsigned long long pg1 = 0
#define PROF_CTA_MID pg1 = prof_gtime()
#define PROF_CTA_END                                                                                                        \
  if (threadIdx.x == 0 && (blockIdx.x % 37 == 0 || blockIdx.x == gridDim.x - 1))                                             \
    printf("conv_tc cta %4d/%d: entry %llu ns, prologue %llu ns, roles %llu ns\n", blockIdx.x, gridDim.x, pg0 % 100000000ull,  \
           pg1 - pg0, prof_gtime() - pg1)
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) stmt
#define PROF_REPORT(role, n0, n1, iters)
#define PROF_CTA_DECL
#define PROF_CTA_MID
#define PROF_CTA_END
#endif

// ---------------------------------------------------------------------------------------------------
// compile-time chunk counts for the producer's unrolled loops: a predicated-off chunk still costs its issue slots (the
// 12-way unrolled loops executed ~1400 instructions per plane for 5 live chunks), so the loops are instantiated per count
template <int N>
struct IC { static constexpr int value = N; };
#define HCU_DISPATCH_NCHUNK(n, fn)            \
  do {                                        \
    if ((n) <= 4) fn(IC<4>{});                \
    else if ((n) <= 5) fn(IC<5>{});           \
    else if ((n) <= 6) fn(IC<6>{});           \
    else if ((n) <= 8) fn(IC<8>{});           \
    else if ((n) <= 10) fn(IC<10>{});         \
    else fn(IC<12>{});                        \
  } while (0)
constexpr int kMaxChunk = 12;  // producer fast path: <= 12 16-byte chunks per thread per x-plane, addresses precomputed

// BatchNorm affine + ReLU of 8 fp16 channels: fp32 multiply-add (one rounding, like the reference's fp32 op followed
// by the fp16 store), ReLU on the packed halves.  A pure-half formulation ((x - c) * s + d, 4 instead of 6 instructions per
// channel pair) was measured: +25 % end-to-end logit error on the ill-conditioned 5-level fixture, so it is not used.
struct BnH8 {
  float sc[8], sh[8];
};
__device__ __forceinline__ void bn_h8_setup(BnH8& b, const float* scale, const float* shift) {
#pragma unroll
  for (int j = 0; j < 8; ++j) { b.sc[j] = scale[j]; b.sh[j] = shift[j]; }
}
__device__ __forceinline__ uint4 bn_relu8(uint4 v, const BnH8& b, int relu) {
  __half2* h = reinterpret_cast<__half2*>(&v);
  const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = __half22float2(h[k]);
    f.x = fmaf(f.x, b.sc[2 * k], b.sh[2 * k]);
    f.y = fmaf(f.y, b.sc[2 * k + 1], b.sh[2 * k + 1]);
    h[k] = __floats2half2_rn(f.x, f.y);
    if (relu) h[k] = __hmax2_nan(h[k], zero);
  }
  return v;
}

// Per-channel statistics of a CTA (one slot per epilogue warp, added in a fixed order) -> the caller's binned fp64
// accumulators, then the optional fused hcu_bn_finalize: the CTA that takes the last ticket sees every CTA's partial sums.
// Called by the nthr epilogue threads (row = 0 .. nthr-1).
template <typename PT>
__device__ __forceinline__ void stats_tail(const PT& p, float* sstat, int row, int ns, int Nc, int nthr = 128, bool eight = false) {
  const int cout = p.cout;
  named_bar_sync(1, nthr);
  for (int c = row; c < Nc; c += nthr) {
    const int ch = ns * Nc + c;
    if (ch < cout) {
      double* sb = p.stats + (size_t)(blockIdx.x % HCU_STAT_BINS) * 2 * p.stats_pitch;
      float q1 = ((sstat[c] + sstat[2 * Nc + c]) + sstat[4 * Nc + c]) + sstat[6 * Nc + c];
      float q2 = ((sstat[Nc + c] + sstat[3 * Nc + c]) + sstat[5 * Nc + c]) + sstat[7 * Nc + c];
      if (eight) {  // bulk mode: eight epilogue warps, slots 4..7 added in the same fixed order
        q1 += ((sstat[8 * Nc + c] + sstat[10 * Nc + c]) + sstat[12 * Nc + c]) + sstat[14 * Nc + c];
        q2 += ((sstat[9 * Nc + c] + sstat[11 * Nc + c]) + sstat[13 * Nc + c]) + sstat[15 * Nc + c];
      }
      atomicAdd(&sb[p.out_c_off + ch], (double)q1);
      atomicAdd(&sb[p.stats_pitch + p.out_c_off + ch], (double)q2);
    }
  }
  if (p.fin.counter != nullptr) {
    __threadfence();
    named_bar_sync(1, nthr);
    if (row == 0) sstat[0] = (atomicAdd(p.fin.counter, 1u) == gridDim.x - 1) ? 1.f : 0.f;
    named_bar_sync(1, nthr);
    if (sstat[0] != 0.f) {
      __threadfence();
      const int pitch = p.stats_pitch;
      for (int ch = row; ch < cout; ch += nthr) {
        double s1 = 0.0, s2 = 0.0;
        for (int b = 0; b < HCU_STAT_BINS; ++b) {
          s1 += __ldcg(&p.stats[(size_t)b * 2 * pitch + p.out_c_off + ch]);
          s2 += __ldcg(&p.stats[(size_t)b * 2 * pitch + pitch + p.out_c_off + ch]);
        }
        const double mu = s1 / p.fin.count;
        double var = s2 / p.fin.count - mu * mu;
        if (var < 0.0) var = 0.0;
        const float is = (float)(1.0 / sqrt(var + (double)p.fin.eps));
        const float muf = (float)mu;
        p.fin.mean[ch] = muf;
        p.fin.invstd[ch] = is;
        const float sc = p.fin.gamma[ch] * is;
        p.fin.scale[ch] = sc;
        p.fin.shift[ch] = p.fin.beta[ch] - muf * sc;
        if (p.fin.running_mean != nullptr) {
          const double unbiased = p.fin.count > 1.0 ? var * p.fin.count / (p.fin.count - 1.0) : var;
          p.fin.running_mean[ch] = (1.f - p.fin.momentum) * p.fin.running_mean[ch] + p.fin.momentum * muf;
          p.fin.running_var[ch] = (1.f - p.fin.momentum) * p.fin.running_var[ch] + p.fin.momentum * (float)unbiased;
        }
      }
    }
  }
}


// Fused BatchNorm-backward statistics: per-warp partial sums (same slot layout as the forward statistics) -> the binned fp64
// accumulators (sum g', invstd * (sum g'y - mean * sum g')), then hcu_bn_bwd_finalize's arithmetic in the last CTA.
template <typename PT>
__device__ __forceinline__ void bnb_tail(const PT& p, float* sstat, int row, int Nc, int C, int nthr, bool eight) {
  named_bar_sync(1, nthr);
  for (int c = row; c < C; c += nthr) {
    float q1 = ((sstat[c] + sstat[2 * Nc + c]) + sstat[4 * Nc + c]) + sstat[6 * Nc + c];
    float q2 = ((sstat[Nc + c] + sstat[3 * Nc + c]) + sstat[5 * Nc + c]) + sstat[7 * Nc + c];
    if (eight) {
      q1 += ((sstat[8 * Nc + c] + sstat[10 * Nc + c]) + sstat[12 * Nc + c]) + sstat[14 * Nc + c];
      q2 += ((sstat[9 * Nc + c] + sstat[11 * Nc + c]) + sstat[13 * Nc + c]) + sstat[15 * Nc + c];
    }
    double* sb = p.bnb.sums + (size_t)(blockIdx.x % HCU_STAT_BINS) * 2 * C;
    const double sg = (double)q1, sgy = (double)q2;
    atomicAdd(&sb[c], sg);
    atomicAdd(&sb[C + c], (double)p.bnb.invstd[c] * (sgy - (double)p.bnb.mean[c] * sg));
  }
  if (p.bnb.fin.counter != nullptr) {
    __threadfence();
    named_bar_sync(1, nthr);
    if (row == 0) sstat[0] = (atomicAdd(p.bnb.fin.counter, 1u) == gridDim.x - 1) ? 1.f : 0.f;
    named_bar_sync(1, nthr);
    if (sstat[0] != 0.f) {
      __threadfence();
      float gs = p.bnb.fin.grad_scale;
      if (p.bnb.fin.dscale != nullptr) gs *= p.bnb.fin.dscale[0];
      for (int i = row; i < C; i += nthr)
        bn_bwd_finalize_one(p.bnb.sums, C, i, p.bnb.fin.count, p.bnb.fin.gamma, p.bnb.mean, p.bnb.invstd, p.bnb.fin.training, gs,
                            p.bnb.fin.dgamma, p.bnb.fin.dbeta, p.bnb.fin.dbias, p.bnb.fin.coef);
    }
  }
}

// Compile-time variants of conv_tc_kernel.  The source-level profile of the generic kernel on the 8-channel levels
// (profiles/r02_conv_tc_sass_regions.txt) showed 317 warp instructions per epilogue warp per plane of which ~85 did arithmetic:
// the rest re-tested launch-invariant flags (bias / statistics / affine / ReLU / debug bits / M-block count / wide mode)
// inside the unrolled loops.  A variant fixes them: SPEC = false is the generic kernel (everything read from Params);
// SPEC = true fixes the M-block count, wide mode, Nc = 16, the vector epilogue (EPI 1: 8 real output channels, 2: 16) and
// the epilogue's flag set (bit 0 bias, 1 statistics, 2 affine, 3 ReLU).  configure() / the launcher pick a variant when the
// descriptor matches one of the three flag sets the U-Net uses (training forward, data gradient, inference), else generic.
//
// BULK (only with SPEC): the input needs no transform (no BatchNorm + ReLU pending on it: the first conv of every block, every
// data gradient, inference) and has 8 channels per pixel, so a run of an x-plane in memory IS the UMMA operand layout
// ([pixel][8 x fp16], rows 16 bytes apart).  Warp 9 streams the planes with cp.async.bulk (one copy per contiguous row
// segment, completion on the slot's mbarrier; padding positions are zero-filled once, they never change), the four producer
// warps have nothing to do and become a second set of epilogue warps (a warp reads the TMEM lanes 32 * (warp % 4) ..: warps
// w and w + 4 share a lane quadrant and split the M-blocks).
template <bool SPEC_, int MB_, int EPI_, int FLAGS_, bool BULK_ = false>
struct Var {
  static constexpr bool kSpec = SPEC_, kBulk = BULK_;
  static constexpr int kMB = MB_, kEpi = EPI_, kFlags = FLAGS_;
};
using VarGeneric = Var<false, 0, 0, 0>;
constexpr int kFlagsTrain = 1 | 2, kFlagsPlain = 0, kFlagsEval = 4 | 8, kFlagsPlainBnb = 16;

template <class V>
__global__ void __launch_bounds__(kThreads, 2) conv_tc_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool S = V::kSpec, BULK = V::kBulk;
  PROF_CTA_DECL;
  const bool wide = S ? true : (p.wide != 0);
  const int debug = S ? 0 : p.debug;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // barrier map: full[R], empty[R], tfull[4], tempty[4], wbar (2 accumulator buffers normally, 4 slots in wide mode)
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * p.R, bar_tfull = bar_empty + 8 * p.R,
                 bar_tempty = bar_tfull + 32, bar_w = bar_tempty + 32;
  const int NB = wide ? 4 : 2;  // accumulator slots an output plane rotates through
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * (2 * p.R + 9));
  // per-channel sum / sum of squares, one slot PER EPILOGUE WARP: every slot has a single writer and the four are added
  // in a fixed order, so the statistics (hence the whole forward) do not depend on the order warps happen to run in
  float* sstat = reinterpret_cast<float*>(smem + p.off_stat);  // [8 warps][2][Nc] (slots 4..7: bulk mode)
  const int Nc = S ? 16 : p.Nc;
  float* sbias = sstat + (BULK ? 16 : 8) * Nc;                 // [3][Nc]: bias, out_scale, out_shift of this column chunk
  const uint32_t a_base = smem_u32(smem + p.off_a), w_base = smem_u32(smem + p.off_w);
  const int R = p.R, MB = S ? V::kMB : p.MB;

  // ---- work item ---------------------------------------------------------------------------------
  int item = blockIdx.x;
  const int ns = item % p.nsplit; item /= p.nsplit;
  const int run = item % p.n_runs; item /= p.n_runs;
  const int xs = item % p.n_xseg;
  const int n = item / p.n_xseg;
  const int x0 = xs * p.Lx;
  const int nout = min(p.Lx, p.OX - x0);
  const int nplanes = nout + (p.KX - 1) * p.dx;
  const int q0 = run * p.M;

  // ---- one-time setup --------------------------------------------------------------------------------
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < R; ++i) {
        mbar_init(bar_full + 8 * i, BULK ? 1 : 4);   // one arrive per producer warp (bulk: the expect_tx arrive)
        mbar_init(bar_empty + 8 * i, kIssuers);  // tcgen05.commit of every issuer
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(bar_tfull + 8 * i, kIssuers);   // tcgen05.commit of every issuer
        mbar_init(bar_tempty + 8 * i, BULK ? 8 : 4);  // one arrive per epilogue warp
      }
      mbar_init(bar_w, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  }
  for (int i = threadIdx.x; i < (BULK ? 16 : 8) * Nc; i += kThreads) sstat[i] = 0.f;
  if (BULK) {  // padding positions of the ring are never written by the bulk copies: zero everything once
    uint4* a4 = reinterpret_cast<uint4*>(smem + p.off_a);
    for (int i = threadIdx.x; i < (R * p.SLOT) >> 4; i += kThreads) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  // ---- everything above touched no global memory: it overlapped the previous kernel's tail (PDL) ----
  pdl_wait();
  pdl_launch_dependents();
  if (S) {
    // rotating accumulator window (see the MMA issuer): KX copies of the weights, copy r = the order of the KX accumulator
    // slots when the newest output plane sits in slot r: [r][K8 slab][slot s][Nc][8] = packed[slab][KX-1-tx][Nc][8] with
    // tx = (r - s) mod KX.  A few KB (the specialised variants are the 8/16-channel levels), gathered from L2 by all threads.
    const int KX = p.KX, per_rot = p.E_tx * KX * Nc;  // 16-byte units
    const uint4* src = reinterpret_cast<const uint4*>(p.wp);
    uint4* dst = reinterpret_cast<uint4*>(smem + p.off_w);
    for (int u = threadIdx.x; u < KX * per_rot; u += kThreads) {
      const int r = u / per_rot, rem = u - r * per_rot;
      const int slab = rem / (KX * Nc), rem2 = rem - slab * (KX * Nc);
      const int sl = rem2 / Nc, c = rem2 - sl * Nc;
      int tx = r - sl;
      tx += tx < 0 ? KX : 0;
      dst[u] = __ldg(src + (slab * KX + (KX - 1 - tx)) * Nc + c);
    }
    fence_proxy_async();
  } else if (warp == 8 && lane == 0) {
    const uint32_t wbytes = (uint32_t)p.E * Nc * 16u;
    fence_barrier_init();
    mbar_expect_tx(bar_w, wbytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wp) + (size_t)ns * wbytes;
    for (uint32_t o = 0; o < wbytes; o += 32768u) bulk_g2s(w_base + o, src + o, min(32768u, wbytes - o), bar_w);
  }
  for (int i = threadIdx.x; i < Nc; i += kThreads) {
    const int ch = (blockIdx.x % p.nsplit) * Nc + i;  // ns
    const bool in = ch < p.cout;
    sbias[i] = (in && p.bias != nullptr) ? p.bias[ch % p.cpp] : 0.f;
    sbias[Nc + i] = (in && p.out_scale != nullptr) ? p.out_scale[ch] : 1.f;
    sbias[2 * Nc + i] = (in && p.out_shift != nullptr) ? p.out_shift[ch] : 0.f;
    if (S ? (V::kFlags & 16) != 0 : p.bnb.y != nullptr) {   // fused BN-backward statistics: that layer's scale / shift (ReLU mask)
      sbias[Nc + i] = in ? p.bnb.scale[ch] : 1.f;
      sbias[2 * Nc + i] = in ? p.bnb.shift[ch] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PROF_CTA_MID;

  if (!BULK && warp >= 4 && warp < 8) {
    // =========================================== PRODUCERS ===========================================
    // cp.async (16 B, zero-fill for padding) straight into the UMMA layout, D = R - span planes in flight per thread
    // beyond the ones the MMA is using; a landed plane is transformed IN PLACE (previous layer's BatchNorm + ReLU:
    // every thread touches only the chunks it copied itself), made visible to the async proxy and published.
    PROF_DECL;
    const int ptid = threadIdx.x - 128;
    const int plane = ptid % p.P;
    const int pix0 = ptid / p.P, pstep = 128 / p.P;
    const int nchunk = (p.RUN - pix0 + pstep - 1) / pstep;  // pixels this thread copies per x-plane
    const int nmax = (p.RUN + pstep - 1) / pstep;           // CTA-uniform upper bound of nchunk
    const bool xf = p.in_scale != nullptr;
    const int relu = p.in_relu;
    BnH8 bn;
    if (xf) bn_h8_setup(bn, p.in_scale + plane * 8, p.in_shift + plane * 8);
    // (yv, zv) of this thread's first pixel
    const int qf = q0 + pix0;
    const int yv0 = qf / p.Zv, zv0 = qf - yv0 * p.Zv;
    const int ystep = pstep / p.Zv, zstep = pstep - ystep * p.Zv;
    const size_t xstride = (size_t)p.in_xs;
    // channel plane -> (stride phase, 8-channel group): the phase selects a sub-lattice of the full-resolution tensor
    long long plane_off = (long long)plane * 8;
    if (p.ips[0] * p.ips[1] * p.ips[2] > 1) {
      int phi = plane / p.Pc;
      const int cg = plane - phi * p.Pc;
      const int fz = phi % p.ips[2]; phi /= p.ips[2];
      const int fy = phi % p.ips[1], fx = phi / p.ips[1];
      plane_off = (long long)cg * 8 + fx * p.in_ph[0] + fy * p.in_ph[1] + fz * p.in_ph[2];
    }
    const __half* in_n = p.in + (size_t)n * p.in_ns + plane_off;
    const int sstep = pstep * 16;
    const uint32_t dst0 = a_base + (uint32_t)(plane * p.PS + pix0 * 16);
    const bool fast = nmax <= kMaxChunk;  // CTA-uniform
    int goff[kMaxChunk];  // element offset of each chunk inside an x-plane; -1 = zero fill, -2 = not this thread's
    if (fast) {
      int yv = yv0, zv = zv0;
#pragma unroll
      for (int c = 0; c < kMaxChunk; ++c) {
        const int ym = yv - p.py, zm = zv - p.pz;
        const bool ok = ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ;
        goff[c] = c < nchunk ? (ok ? ym * p.in_ys + zm * p.in_zs : -1) : -2;
        zv += zstep; yv += ystep;
        if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
      }
    }
    const int D = p.D;  // planes in flight (1 .. 3; 0 = synchronous); the host keeps R >= span + D + slack
    int slot_i = 0, slot_f = 0;
    uint32_t par = 1;
    auto finish = [&](int jf) {  // plane jf has landed: transform in place, publish
      if (xf) {
        const int xm = x0 + jf - p.px;
        if (xm >= 0 && xm < p.IX) {  // warp-uniform
          unsigned char* dp = smem + p.off_a + slot_f * p.SLOT + plane * p.PS + pix0 * 16;
          if (fast) {
            auto body = [&](auto NC) {
              constexpr int N = decltype(NC)::value;
#pragma unroll
              for (int g = 0; g < N; g += 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  v[u] = make_uint4(0u, 0u, 0u, 0u);
                  if (g + u < N && goff[g + u] >= 0) v[u] = *reinterpret_cast<const uint4*>(dp + (g + u) * sstep);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (g + u < N) v[u] = bn_relu8(v[u], bn, relu);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (g + u < N && goff[g + u] >= 0) *reinterpret_cast<uint4*>(dp + (g + u) * sstep) = v[u];
              }
            };
            HCU_DISPATCH_NCHUNK(nmax, body);
          } else {
            int yv = yv0, zv = zv0;
            for (int c = 0; c < nchunk; ++c) {
              const int ym = yv - p.py, zm = zv - p.pz;
              if (ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ) {
                uint4* q = reinterpret_cast<uint4*>(dp + (size_t)c * sstep);
                *q = bn_relu8(*q, bn, relu);
              }
              zv += zstep; yv += ystep;
              if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
            }
          }
        }
      }
      PROF_WAIT(pw2, fence_proxy_async());  // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * slot_f);
      if (++slot_f == R) slot_f = 0;
    };
    for (int j = 0; j < nplanes + D; ++j) {
      // groups committed so far = planes 0 .. j-1; publish plane j-D before (possibly) blocking on the ring
      if (D > 0 && j >= D) {
        PROF_WAIT(pw1, {
          if (D == 1) cp_async_wait<0>();
          else if (D == 2) cp_async_wait<1>();
          else cp_async_wait<2>();
        });
        finish(j - D);
      }
      if (j < nplanes) {
        PROF_WAIT(pw0, mbar_wait(bar_empty + 8 * slot_i, par));
        const int xm = x0 + j - p.px;  // memory x of this virtual plane
        const bool xok = xm >= 0 && xm < p.IX;
        const __half* in_x = in_n + (size_t)(xok ? xm : 0) * xstride;
        const uint32_t dst = dst0 + (uint32_t)(slot_i * p.SLOT);
        if (debug & 1) {
        } else if (fast) {
          auto body = [&](auto NC) {
            constexpr int N = decltype(NC)::value;
#pragma unroll
            for (int c = 0; c < N; ++c) {
              if (goff[c] != -2) {
                const bool ok = xok && goff[c] >= 0;
                cp_async16(dst + c * sstep, ok ? in_x + goff[c] : in_n, ok ? 16u : 0u);
              }
            }
          };
          HCU_DISPATCH_NCHUNK(nmax, body);
        } else {
          int yv = yv0, zv = zv0;
          for (int c = 0; c < nchunk; ++c) {
            const int ym = yv - p.py, zm = zv - p.pz;
            const bool ok = xok && ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ;
            cp_async16(dst + c * sstep, ok ? in_x + ((size_t)ym * p.in_ys + (size_t)zm * p.in_zs) : in_n, ok ? 16u : 0u);
            zv += zstep; yv += ystep;
            if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
          }
        }
        if (++slot_i == R) { slot_i = 0; par ^= 1; }
      }
      cp_async_commit();  // one group per iteration (possibly empty) keeps the group arithmetic uniform
      if (D == 0) {
        cp_async_wait<0>();
        finish(j);
      }
    }
    PROF_REPORT("producer", "wait_empty", "wait_cpasync", nplanes);
  } else if (warp == 8) {
    // =========================================== MMA ISSUER ==========================================
    // The whole warp runs this (warp-uniform values -> descriptors live in uniform registers, no per-thread
    // waterfall); one elected lane issues tcgen05.mma / tcgen05.commit.  Everything is table driven: one 16-byte
    // shared-memory read per MMA.
    const uint32_t idesc = (1u << 4) | ((uint32_t)(Nc >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, descriptor version 1 (bit 46)
    const bool leader = elect_one();
    const uint32_t a_desc0 = a_base >> 4;
    const uint32_t wdesc = (w_base >> 4) | ((((uint32_t)(Nc * 16)) >> 4) << 16);  // B: LBO = next K8 slab
    PROF_DECL;
    if (S) {
      // Rotating accumulator window.  Output plane i accumulates in slot i mod KX of its M-block (TMEM columns
      // (mb * KX + slot) * Nc); input plane j feeds the outputs j-KX+1 .. j = ALL KX slots, so every K16 step is ONE MMA per
      // M-block with N = KX * Nc against weight copy r = j mod KX (whose slot order matches): no window ever wraps and no
      // output needs a separate accumulate = 0 MMA -- the epilogue zeroes a slot (tcgen05.st) when it drains it, every MMA
      // accumulates.  The x-fused schedule below (4 slots, wrap splits + a separate first MMA per output) issued 15 MMAs per
      // plane on the (3,3,1) 8-channel layers where this one issues 8, and the tensor pipe -- bound by the shared-memory fetch
      // of the A operand, ~40 clk per MMA shared by the SM's two CTAs -- was the busiest role (profiles/r02_conv_tc_roles.txt).
      // The first KX-1 planes complete "virtual" outputs (i < 0) the epilogue drains without storing.
      const int KX = p.KX;
      const uint32_t wlbo = (((uint32_t)(KX * Nc * 16)) >> 4) << 16;    // next K8 slab
      const uint32_t idw = (1u << 4) | ((128u >> 4) << 24) | ((((uint32_t)(KX * Nc)) >> 3) << 17);
      const uint32_t per_rot = (uint32_t)(p.E_tx * KX * Nc);
      const uint32_t mbcols = (uint32_t)(KX * Nc);
      int ring = 0, snew = 0;
      uint32_t rpar = 0;
      // EVERY plane touches all KX slots.  Plane 0 needs all of them zeroed (the epilogue's initial arrivals: phase 0 of every
      // tempty barrier); plane j >= 1 needs the slot of its new output, j mod KX, drained: its previous occupant (output j - KX,
      // virtual while negative) completed with plane j - 1, so that is the slot's NEXT phase every time.
      for (int sl = 0; sl < KX; ++sl) mbar_wait(bar_tempty + 8 * sl, 0u);
      uint32_t te_phase = 0xffffffffu;
      for (int j = 0; j < nplanes; ++j) {
        PROF_WAIT(pw0, mbar_wait(bar_full + 8 * ring, rpar));
        if (j > 0) {
          PROF_WAIT(pw1, mbar_wait(bar_tempty + 8 * snew, (te_phase >> snew) & 1u));
          te_phase ^= 1u << snew;
        }
        tc_fence_after();
        const uint32_t abase = a_desc0 + (uint32_t)(ring * (p.SLOT >> 4));
        const uint32_t wb = (w_base >> 4) + (uint32_t)snew * per_rot;
        for (int e = 0; e < p.npairs; ++e) {
          const uint2 t = p.tab[e];
          const uint64_t ad = desc_hi | (uint64_t)(abase + t.x);
          const uint64_t bd = desc_hi | (uint64_t)((wb + t.y) | wlbo);
          if (elect_one()) {
            umma_f16(tmem_base, ad, bd, idw, 1u);
            if (MB > 1) umma_f16(tmem_base + mbcols, ad + 128u, bd, idw, 1u);
            if (MB > 2) umma_f16(tmem_base + 2u * mbcols, ad + 256u, bd, idw, 1u);
            if (MB > 3) umma_f16(tmem_base + 3u * mbcols, ad + 384u, bd, idw, 1u);
          }
          __syncwarp();
        }
        const int sdone = snew + 1 == KX ? 0 : snew + 1;  // slot of output j - KX + 1, complete with this plane
        if (elect_one()) {
          umma_commit(bar_empty + 8 * ring);
          umma_commit(bar_tfull + 8 * sdone);
        }
        __syncwarp();
        if (++ring == R) { ring = 0; rpar ^= 1; }
        snew = sdone;
      }
      PROF_REPORT("mma", "wait_full", "wait_tempty", nplanes);
    } else if (wide) {
      mbar_wait(bar_w, 0);
      // x-fused schedule: input plane j feeds the output planes i = j - tx (tx = 0 .. KX-1) in ONE MMA per K16 step whose
      // N spans their accumulator slots (4 slots per M-block, consecutive outputs in consecutive columns) against the
      // weights laid out [K8 slab][KX-1-tx][Nc]: the A operand is fetched once per input plane instead of once per
      // (input plane, tx).  A window that wraps around the slot ring is split in two, and the first K16 step of a new
      // output plane (tx = 0) is issued on its own with accumulate = 0.
      const int KX = p.KX;
      const uint32_t wrow = (uint32_t)Nc;                               // 16-byte units per (KX-1-tx) block
      const uint32_t wlbo = (((uint32_t)(KX * Nc * 16)) >> 4) << 16;    // next K8 slab
      const uint32_t idesc0 = (1u << 4) | ((128u >> 4) << 24);
      int wslot2 = 0;
      uint32_t wpar2 = 0;
      for (int j = 0; j < nplanes; ++j) {
        PROF_WAIT(pw0, mbar_wait(bar_full + 8 * wslot2, wpar2));
        if (j < nout) PROF_WAIT(pw1, mbar_wait(bar_tempty + 8 * (j & 3), ((j >> 2) & 1) ^ 1));
        tc_fence_after();
        const int i_lo = max(0, j - (KX - 1)), i_hi = min(j, nout - 1);
        const uint32_t abase = a_desc0 + (uint32_t)(wslot2 * (p.SLOT >> 4));
        // one MMA group (all M-blocks) for the outputs [ia, ib] (no slot wrap inside), K16 step e
        auto emit = [&](int ia, int ib, const uint2 t, uint32_t acc) {
          const uint32_t w = (uint32_t)(ib - ia + 1);
          const uint32_t idw = idesc0 | (((w * (uint32_t)Nc) >> 3) << 17);
          const uint32_t col = tmem_base + (uint32_t)((ia & 3) * Nc);
          const uint64_t ad = desc_hi | (uint64_t)(abase + t.x);
          const uint64_t bd = desc_hi | (uint64_t)(((w_base >> 4) + t.y + (uint32_t)(KX - 1 - (j - ia)) * wrow) | wlbo);
          if (elect_one()) {
            umma_f16(col, ad, bd, idw, acc);
            if (MB > 1) umma_f16(col + (uint32_t)(4 * Nc), ad + 128u, bd, idw, acc);
            if (MB > 2) umma_f16(col + (uint32_t)(8 * Nc), ad + 256u, bd, idw, acc);
            if (MB > 3) umma_f16(col + (uint32_t)(12 * Nc), ad + 384u, bd, idw, acc);
          }
          __syncwarp();
        };
        auto emit_range = [&](int ia, int ib, const uint2 t, uint32_t acc) {  // split where the slot ring wraps
          if (ia > ib) return;
          const int first = min(ib - ia + 1, 4 - (ia & 3));
          emit(ia, ia + first - 1, t, acc);
          if (ia + first <= ib) emit(ia + first, ib, t, acc);
        };
        if (!(debug & 4) && i_lo <= i_hi) {
          for (int e = 0; e < p.npairs; ++e) {
            const uint2 t = p.tab[e];
            if (e == 0 && i_hi == j) {  // output plane j starts here
              emit_range(i_lo, j - 1, t, 1u);
              emit(j, j, t, 0u);
            } else {
              emit_range(i_lo, i_hi, t, 1u);
            }
          }
        }
        if (elect_one()) {
          umma_commit(bar_empty + 8 * wslot2);                              // this input plane is consumed
          if (j >= KX - 1) umma_commit(bar_tfull + 8 * ((j - (KX - 1)) & 3));  // output j-(KX-1) is complete
        }
        __syncwarp();
        if (++wslot2 == R) { wslot2 = 0; wpar2 ^= 1; }
      }
    } else {
      mbar_wait(bar_w, 0);
      const int lastoff = (p.KX - 1) * p.dx;
      int next_wait = 0, wslot = 0;
      uint32_t wpar = 0;
      int i_mod = 0;
      for (int i = 0; i < nout; ++i) {
        for (; next_wait <= i + lastoff; ++next_wait) {
          PROF_WAIT(pw0, mbar_wait(bar_full + 8 * wslot, wpar));
          if (++wslot == R) { wslot = 0; wpar ^= 1; }
        }
        const int buf = i & 1;
        PROF_WAIT(pw1, mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tb = tmem_base + (uint32_t)(buf * MB * Nc);
        // every operand below is warp-uniform (kernel parameters + uniform counters): the descriptors are built in the
        // uniform datapath straight from the constant-bank table; the M-blocks are unrolled so that the four MMAs of a
        // K16 step issue back to back (they accumulate into different TMEM tiles)
        if (!(debug & 4)) {
          for (int tx = 0; tx < p.KX; ++tx) {
            int sl = i_mod + tx * p.dx;
            sl -= sl >= R ? R : 0;
            const uint32_t abase = a_desc0 + (uint32_t)(sl * (p.SLOT >> 4));
            const uint2* T = p.tab + tx * p.npairs;
  #pragma unroll 1
            for (int e = 0; e < p.npairs; ++e) {
              const uint2 t = T[e];
              const uint32_t acc = (uint32_t)(tx | e);
              const uint64_t ad = desc_hi | (uint64_t)(abase + t.x), bd = desc_hi | (uint64_t)(wdesc + t.y);
              if (elect_one()) {
                umma_f16(tb, ad, bd, idesc, acc);
                if (MB > 1) umma_f16(tb + (uint32_t)Nc, ad + 128u, bd, idesc, acc);
                if (MB > 2) umma_f16(tb + (uint32_t)(2 * Nc), ad + 256u, bd, idesc, acc);
                if (MB > 3) umma_f16(tb + (uint32_t)(3 * Nc), ad + 384u, bd, idesc, acc);
              }
              __syncwarp();
            }
          }
        }
        if (leader) {
          umma_commit(bar_empty + 8 * i_mod);  // plane i is not needed by later outputs
          umma_commit(bar_tfull + 8 * buf);
        }
        __syncwarp();
        i_mod = i_mod + 1 == R ? 0 : i_mod + 1;
      }
    }
    if (!S) PROF_REPORT("mma", "wait_full", "wait_tempty", nout);
  } else if (warp == 9) {
    // ====================================== BULK-COPY PRODUCER (bulk mode) ===========================
    if (BULK) {
      const int Zv = p.Zv, q1 = q0 + p.RUN;
      // contiguous: the virtual plane is the memory plane shifted by py rows (no z padding) -> ONE segment per x-plane
      const bool merged = p.pz == 0 && p.IZ == Zv;
      const int yfirst = q0 / Zv;
      const int nrows = merged ? 1 : (q1 - 1) / Zv - yfirst + 1;
      const __half* in_n = p.in + (size_t)n * p.in_ns;
      // segment r of the run: flat positions [qa, qb) and the element offset of qa inside the x-plane
      auto seg = [&](int r, int& qa, int& qb, int& goff) {
        if (merged) {
          qa = max(q0, p.py * Zv);
          qb = min(q1, (p.py + p.IY) * Zv);
          goff = (qa - p.py * Zv) * p.in_zs;
        } else {
          const int yv = yfirst + r, ym = yv - p.py;
          qa = max(q0, yv * Zv + p.pz);
          qb = (ym >= 0 && ym < p.IY) ? min(q1, yv * Zv + p.pz + p.IZ) : qa;
          goff = ym * p.in_ys + (qa - yv * Zv - p.pz) * p.in_zs;
        }
        if (qb < qa) qb = qa;
      };
      int slot = 0;
      uint32_t par = 1;
      for (int j = 0; j < nplanes; ++j) {
        mbar_wait(bar_empty + 8 * slot, par);
        const int xm = x0 + j - p.px;
        const uint32_t sdst = a_base + (uint32_t)(slot * p.SLOT);
        const uint32_t bar = bar_full + 8 * slot;
        if (xm >= 0 && xm < p.IX) {
          const __half* in_x = in_n + (size_t)xm * (size_t)p.in_xs;
          uint32_t bytes = 0;
          for (int r = lane; r < nrows; r += 32) {
            int qa, qb, goff;
            seg(r, qa, qb, goff);
            bytes += (uint32_t)(qb - qa) * 16u;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
          if (lane == 0) {
            if (bytes != 0u) mbar_expect_tx(bar, bytes); else mbar_arrive(bar);
          }
          __syncwarp();
          for (int r = lane; r < nrows; r += 32) {
            int qa, qb, goff;
            seg(r, qa, qb, goff);
            if (qb > qa) bulk_g2s(sdst + (uint32_t)(qa - q0) * 16u, in_x + goff, (uint32_t)(qb - qa) * 16u, bar);
          }
        } else {
          // x padding: the whole plane is zero (the slot may hold an earlier plane's data)
          uint4* z4 = reinterpret_cast<uint4*>(smem + p.off_a + slot * p.SLOT);
          for (int i = lane; i < p.RUN; i += 32) z4[i] = make_uint4(0u, 0u, 0u, 0u);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar);
        }
        if (++slot == R) { slot = 0; par ^= 1; }
      }
    }
  } else {
    // =========================================== EPILOGUE ============================================
    // warps 0-3; in bulk mode warps 4-7 as well: warp w and w + 4 read the same TMEM lane quadrant and split the M-blocks
    PROF_DECL;
    const int half = BULK ? (warp >> 2) : 0;
    constexpr int MBSTEP = BULK ? 2 : 1;
    const int row = (warp & 3) * 32 + lane;  // accumulator row within an M-block == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const bool do_stats = S ? (V::kFlags & 2) != 0 : p.stats != nullptr;
    const bool affine = S ? (V::kFlags & 4) != 0 : p.out_scale != nullptr;
    const bool has_bias = S ? (V::kFlags & 1) != 0 : p.bias != nullptr;
    const int out_relu = S ? ((V::kFlags & 8) != 0 ? 1 : 0) : p.out_relu, cout = p.cout;
    constexpr bool BNB = S && (V::kFlags & 16) != 0;   // fused BatchNorm-backward statistics (data-gradient launches)
    const bool bnb_rt = !S && p.bnb.y != nullptr;      // the same in the generic kernel's vector epilogue (32 / 64-channel levels)
    const int nch = S ? (V::kEpi == 1 ? 8 : 16) : min(Nc, cout - ns * Nc);  // real output channels of this CTA's column chunk
    // per M-block: element offset of this thread's pixel inside an output x-plane (-1: wrap-around / out of range)
    // (slot k of this warp = M-block half + k * MBSTEP)
    int poff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int mb = half + k * MBSTEP;
      const int q = q0 + mb * 128 + row;
      const int oy = q / p.Zv, oz = q - oy * p.Zv;
      poff[k] = (k * MBSTEP < MB && mb < MB && oy < p.OY && oz < p.OZ) ? (int)(oy * p.out_sy + oz * p.out_sz) : -1;
    }
    const long long obase0 = p.out_base + n * p.out_sn + p.out_c_off + ns * Nc;

    // Rotating window (specialised variants): zero this warp's share of every accumulator slot once, then hand the slots to
    // the MMA issuer (its first KX waits on tempty complete with these arrivals).
    const int KXr = p.KX;
    if (S) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int mb = half + k * MBSTEP;
        if (k * MBSTEP < MB && mb < MB)
          for (int sl = 0; sl < KXr; ++sl) tmem_zero16(tmem_base + lane_base + (uint32_t)((mb * KXr + sl) * Nc));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int sl = 0; sl < KXr; ++sl) mbar_arrive(bar_tempty + 8 * sl);
    }
    // iteration `it` drains: generic -- output it from buffer it mod NB; rotating window -- output it - (KX-1) (virtual, not
    // stored, while negative) from slot (it + 1) mod KX
    const int nit = S ? nplanes : nout;
    int rslot = 1 % KXr;
    uint32_t tf_phase = 0;

    if (S ? V::kEpi == 1 : (p.epi_fast && nch == 8)) {
      // ---- 8 output channels (the HBM-bound first / last levels): every M-block's 8 columns are fetched with ONE wait,
      // the accumulator buffer is released before the arithmetic, statistics stay in registers until the end
      float t1[8], t2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] = 0.f; t2[j] = 0.f; }
      float bs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bs[j] = sbias[j];
      // BNB: t1 = sum g', t2 = sum g' * y; scale / shift of the layer whose backward statistics these are
      float bsc[8], bsh[8];
      if (BNB) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { bsc[j] = p.bnb.scale[j]; bsh[j] = p.bnb.shift[j]; }
      }
      for (int it = 0; it < nit; ++it) {
        const int i = S ? it - (KXr - 1) : it;
        const int buf = S ? rslot : (it & (NB - 1));
        const uint32_t par = S ? ((tf_phase >> rslot) & 1u) : (uint32_t)((it / NB) & 1);
        if (S) {
          tf_phase ^= 1u << rslot;
          rslot = rslot + 1 == KXr ? 0 : rslot + 1;
        }
        // BNB: this plane's y values, requested before the wait for the accumulators (their latency hides behind it)
        uint4 yv[4];
        if (BNB && i >= 0) {
          const __half* yplane = p.bnb.y + obase0 + (long long)(x0 + i) * p.out_sx;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k * MBSTEP < MB && poff[k] >= 0) yv[k] = ldg_nc16(yplane + poff[k]);
        }
        PROF_WAIT(pw0, mbar_wait(bar_tfull + 8 * buf, par));
        tc_fence_after();
        if (debug & 2) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
          continue;
        }
        uint32_t r[4][8];
        if (!S || i >= 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int mb = half + k * MBSTEP;
            if (k * MBSTEP < MB && mb < MB)
              tmem_ld8_nowait(tmem_base + lane_base + (uint32_t)((S ? mb * KXr + buf : (wide ? mb * 4 + buf : buf * MB + mb)) * Nc), r[k]);
          }
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k * MBSTEP < MB && half + k * MBSTEP < MB) tmem_pin8(r[k]);
        }
        if (S) {  // the slot's next output starts from zero: every MMA accumulates
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int mb = half + k * MBSTEP;
            if (k * MBSTEP < MB && mb < MB) tmem_zero8(tmem_base + lane_base + (uint32_t)((mb * KXr + buf) * Nc));
          }
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        if (S && i < 0) continue;
        __half* oplane = reinterpret_cast<__half*>(p.out) + obase0 + (long long)(x0 + i) * p.out_sx;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {  // mb: slot index of this warp
          if (mb * MBSTEP < MB && poff[mb] >= 0) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[mb][j]);
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += bs[j];
            }
            if (do_stats) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                t1[j] += v[j];
                t2[j] = fmaf(v[j], v[j], t2[j]);
              }
            }
            if (affine) {
#pragma unroll
              for (int j = 0; j < 8; j += 4) {
                const float4 a = *reinterpret_cast<const float4*>(&sbias[Nc + j]);
                const float4 b = *reinterpret_cast<const float4*>(&sbias[2 * Nc + j]);
                v[j] = fmaf(v[j], a.x, b.x); v[j + 1] = fmaf(v[j + 1], a.y, b.y);
                v[j + 2] = fmaf(v[j + 2], a.z, b.z); v[j + 3] = fmaf(v[j + 3], a.w, b.w);
              }
            }
            if (out_relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            __half2 h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            if (!(debug & 16)) *reinterpret_cast<uint4*>(oplane + poff[mb]) = *reinterpret_cast<uint4*>(h);
            if (BNB) {  // the STORED (fp16) gradient is what the statistics pass would have read
              const __half2* yh = reinterpret_cast<const __half2*>(&yv[mb]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 g = __half22float2(h[j]);
                const float2 yy = __half22float2(yh[j]);
                if (fmaf(yy.x, bsc[2 * j], bsh[2 * j]) <= 0.f) g.x = 0.f;
                if (fmaf(yy.y, bsc[2 * j + 1], bsh[2 * j + 1]) <= 0.f) g.y = 0.f;
                t1[2 * j] += g.x; t1[2 * j + 1] += g.y;
                t2[2 * j] = fmaf(g.x, yy.x, t2[2 * j]);
                t2[2 * j + 1] = fmaf(g.y, yy.y, t2[2 * j + 1]);
              }
            }
          }
        }
      }
      if (do_stats || BNB) {
        const float r1 = reduce8(t1, lane);
        const float r2 = reduce8(t2, lane);
        if ((lane & 3) == 0) {
          sstat[warp * 2 * Nc + (lane >> 2)] = r1;
          sstat[warp * 2 * Nc + Nc + (lane >> 2)] = r2;
        }
      }
    } else if (S ? V::kEpi == 2 : (p.epi_fast != 0)) {
      // ---- multiples of 16 channels (an odd 8-channel tail is handled by the store predicate), fp16 vector stores
      const bool local_stats = do_stats && Nc == 16;  // one chunk: keep the sums in registers until the end
      float t1[16], t2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { t1[j] = 0.f; t2[j] = 0.f; }
      for (int it = 0; it < nit; ++it) {
        const int i = S ? it - (KXr - 1) : it;
        const int buf = S ? rslot : (it & (NB - 1));
        const uint32_t par = S ? ((tf_phase >> rslot) & 1u) : (uint32_t)((it / NB) & 1);
        if (S) {
          tf_phase ^= 1u << rslot;
          rslot = rslot + 1 == KXr ? 0 : rslot + 1;
        }
        uint4 yv[4][2];
        if (BNB && i >= 0) {
          const __half* yplane = p.bnb.y + obase0 + (long long)(x0 + i) * p.out_sx;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k * MBSTEP < MB && poff[k] >= 0) {
              yv[k][0] = ldg_nc16(yplane + poff[k]);
              yv[k][1] = ldg_nc16(yplane + poff[k] + 8);
            }
        }
        PROF_WAIT(pw0, mbar_wait(bar_tfull + 8 * buf, par));
        tc_fence_after();
        __half* oplane = reinterpret_cast<__half*>(p.out) + obase0 + (long long)(x0 + i) * p.out_sx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int mb = half + k * MBSTEP;  // warp-uniform
          if (k * MBSTEP >= MB || mb >= MB) break;
          const bool valid = poff[k] >= 0 && (!S || i >= 0);
          for (int cc = 0; cc < nch; cc += 16) {
            float v[16];
            const uint32_t taddr = tmem_base + lane_base + (uint32_t)((S ? mb * KXr + buf : (wide ? mb * 4 + buf : buf * MB + mb)) * Nc + cc);
            tmem_ld16(taddr, v);
            if (S) tmem_zero16(taddr);  // ordered behind the load by its wait; completion awaited before the slot is released
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(&sbias[cc + j]);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (local_stats) {
              if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  t1[j] += v[j];
                  t2[j] = fmaf(v[j], v[j], t2[j]);
                }
              }
            } else if (do_stats) {
              float s1[16], s2[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                s1[j] = valid ? v[j] : 0.f;
                s2[j] = s1[j] * s1[j];
              }
              const float r1 = reduce16(s1, lane);
              const float r2 = reduce16(s2, lane);
              if ((lane & 1) == 0) {
                sstat[warp * 2 * Nc + cc + (lane >> 1)] += r1;   // this lane is the slot's only writer
                sstat[warp * 2 * Nc + Nc + cc + (lane >> 1)] += r2;
              }
            }
            float b1[16], b2[16];   // generic kernel's fused BN-backward statistics: this row's masked gradient and its product with y
            if (bnb_rt) {
#pragma unroll
              for (int j = 0; j < 16; ++j) { b1[j] = 0.f; b2[j] = 0.f; }
            }
            if (valid) {
              if (affine) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float4 a = *reinterpret_cast<const float4*>(&sbias[Nc + cc + j]);
                  const float4 b = *reinterpret_cast<const float4*>(&sbias[2 * Nc + cc + j]);
                  v[j] = fmaf(v[j], a.x, b.x); v[j + 1] = fmaf(v[j + 1], a.y, b.y);
                  v[j + 2] = fmaf(v[j + 2], a.z, b.z); v[j + 3] = fmaf(v[j + 3], a.w, b.w);
                }
              }
              if (out_relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              __half2 h[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              __half* o = oplane + poff[k] + cc;
              *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(&h[0]);
              if (cc + 8 < nch) *reinterpret_cast<uint4*>(o + 8) = *reinterpret_cast<uint4*>(&h[4]);
              if (bnb_rt) {  // generic kernel: any multiple of 16 channels, one shuffle reduction per chunk into the warp's slot
                const __half* yp = p.bnb.y + obase0 + (long long)(x0 + i) * p.out_sx + poff[k] + cc;
                const uint4 y0 = ldg_nc16(yp), y1 = ldg_nc16(yp + 8);
                const __half2* ya = reinterpret_cast<const __half2*>(&y0);
                const __half2* yb = reinterpret_cast<const __half2*>(&y1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float2 g = __half22float2(h[j]);
                  const float2 yy = __half22float2(j < 4 ? ya[j] : yb[j - 4]);
                  if (fmaf(yy.x, sbias[Nc + cc + 2 * j], sbias[2 * Nc + cc + 2 * j]) <= 0.f) g.x = 0.f;
                  if (fmaf(yy.y, sbias[Nc + cc + 2 * j + 1], sbias[2 * Nc + cc + 2 * j + 1]) <= 0.f) g.y = 0.f;
                  b1[2 * j] = g.x; b1[2 * j + 1] = g.y;
                  b2[2 * j] = g.x * yy.x; b2[2 * j + 1] = g.y * yy.y;
                }
              }
              if (BNB) {  // 16 channels, one chunk (Nc == 16): sums in t1 / t2, scale / shift from shared memory
                const __half2* yh = reinterpret_cast<const __half2*>(&yv[k][0]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float2 g = __half22float2(h[j]);
                  const float2 yy = __half22float2(yh[j]);
                  if (fmaf(yy.x, sbias[Nc + 2 * j], sbias[2 * Nc + 2 * j]) <= 0.f) g.x = 0.f;
                  if (fmaf(yy.y, sbias[Nc + 2 * j + 1], sbias[2 * Nc + 2 * j + 1]) <= 0.f) g.y = 0.f;
                  t1[2 * j] += g.x; t1[2 * j + 1] += g.y;
                  t2[2 * j] = fmaf(g.x, yy.x, t2[2 * j]);
                  t2[2 * j + 1] = fmaf(g.y, yy.y, t2[2 * j + 1]);
                }
              }
            }
            if (bnb_rt) {   // every lane takes part in the reduction (rows without an output contribute zeros)
              const float r1 = reduce16(b1, lane);
              const float r2 = reduce16(b2, lane);
              if ((lane & 1) == 0) {
                sstat[warp * 2 * Nc + cc + (lane >> 1)] += r1;
                sstat[warp * 2 * Nc + Nc + cc + (lane >> 1)] += r2;
              }
            }
          }
        }
        if (S) tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      if (local_stats || BNB) {
        const float r1 = reduce16(t1, lane);
        const float r2 = reduce16(t2, lane);
        if ((lane & 1) == 0) {
          sstat[warp * 2 * Nc + (lane >> 1)] = r1;
          sstat[warp * 2 * Nc + Nc + (lane >> 1)] = r2;
        }
      }
    } else if (!S) {
      // ---- generic: fp32 output, channel counts that are not a multiple of 8, stride-phase scatter
      const int out_f32 = p.out_f32;
      const bool phased = p.ops[0] * p.ops[1] * p.ops[2] > 1;
      for (int i = 0; i < nout; ++i) {
        const int buf = i & (NB - 1);
        PROF_WAIT(pw0, mbar_wait(bar_tfull + 8 * buf, (i / NB) & 1));
        tc_fence_after();
        const long long obase = obase0 + (long long)(x0 + i) * p.out_sx;
#pragma unroll 1
        for (int mb = 0; mb < MB; ++mb) {
          const bool valid = poff[mb] >= 0;
#pragma unroll 1
          for (int cc = 0; cc < nch; cc += 16) {
            float v[16];
            tmem_ld16(tmem_base + lane_base + (uint32_t)((wide ? mb * 4 + buf : buf * MB + mb) * Nc + cc), v);
            const int ch0 = ns * Nc + cc;  // first output channel of this chunk
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += sbias[cc + j];
            if (do_stats) {
              float s1[16], s2[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                s1[j] = valid ? v[j] : 0.f;
                s2[j] = s1[j] * s1[j];
              }
              const float r1 = reduce16(s1, lane);
              const float r2 = reduce16(s2, lane);
              if ((lane & 1) == 0) {
                sstat[warp * 2 * Nc + cc + (lane >> 1)] += r1;   // this lane is the slot's only writer
                sstat[warp * 2 * Nc + Nc + cc + (lane >> 1)] += r2;
              }
            }
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], sbias[Nc + cc + j], sbias[2 * Nc + cc + j]);
              if (out_relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              const int nv = min(16, cout - ch0);
              if (phased) {
                // each 8-channel half of the chunk belongs to one stride phase: its own spatial offset
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int ch = ch0 + 8 * hh;
                  if (ch < cout) {
                    int phi = ch / p.cpp;
                    const int co = ch - phi * p.cpp;
                    const int fz = phi % p.ops[2]; phi /= p.ops[2];
                    const int fy = phi % p.ops[1], fx = phi / p.ops[1];
                    __half* o = reinterpret_cast<__half*>(p.out) + (obase - ns * Nc) + poff[mb] + fx * p.out_ph[0] +
                                fy * p.out_ph[1] + fz * p.out_ph[2] + co;
                    __half2 h[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[8 * hh + 2 * j], v[8 * hh + 2 * j + 1]);
                    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
                  }
                }
              } else if (out_f32) {
                float* o = reinterpret_cast<float*>(p.out) + obase + poff[mb] + cc;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < nv) o[j] = v[j];
              } else {
                __half* o = reinterpret_cast<__half*>(p.out) + obase + poff[mb] + cc;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < nv) o[j] = __float2half_rn(v[j]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
    }
    PROF_REPORT("epilogue", "wait_tfull", "-", nout);
    if (do_stats) stats_tail(p, sstat, BULK ? (int)threadIdx.x : row, ns, Nc, BULK ? 256 : 128, BULK);
    if (BNB || bnb_rt) bnb_tail(p, sstat, BULK ? (int)threadIdx.x : row, Nc, nch, BULK ? 256 : 128, BULK);
  }


  // ---- teardown --------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  PROF_CTA_END;
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


using ConvTcFn = void (*)(const Params);
// specialised instantiations, one translation unit per M-block count: epi 1 | 2, flag set 0 train | 1 plain | 2 eval |
// 3 plain + fused BatchNorm-backward statistics, bulk 0 | 1
ConvTcFn conv_tc_variant_mb1(int epi, int fi, int bulk);
ConvTcFn conv_tc_variant_mb2(int epi, int fi, int bulk);
ConvTcFn conv_tc_variant_mb3(int epi, int fi, int bulk);
ConvTcFn conv_tc_variant_mb4(int epi, int fi, int bulk);

#define HCU_TC_VARIANT_ROW(mb, epi, bulk)                                            \
  if (fi == 0) return conv_tc_kernel<Var<true, mb, epi, kFlagsTrain, bulk>>;         \
  if (fi == 1) return conv_tc_kernel<Var<true, mb, epi, kFlagsPlain, bulk>>;         \
  if (fi == 3) return conv_tc_kernel<Var<true, mb, epi, kFlagsPlainBnb, bulk>>;      \
  return conv_tc_kernel<Var<true, mb, epi, kFlagsEval, bulk>>;
#define HCU_TC_DEFINE_VARIANTS(mb)                                                   \
  ConvTcFn conv_tc_variant_mb##mb(int epi, int fi, int bulk) {                       \
    if (epi == 1 && !bulk) { HCU_TC_VARIANT_ROW(mb, 1, false) }                      \
    if (epi == 1) { HCU_TC_VARIANT_ROW(mb, 1, true) }                                \
    if (!bulk) { HCU_TC_VARIANT_ROW(mb, 2, false) }                                  \
    HCU_TC_VARIANT_ROW(mb, 2, true)                                                  \
  }

}  // namespace tc
}  // namespace hcu
