// Weight gradient of the gather-convolution on the 5th-generation tensor cores, for the CHANNEL-RICH levels.
//
//   dW[tap][ci][co] = sum_{n, q} act(a[n, q + shift(tap), ci]) * dy[n, q, co]
//
// Same "flat shift" staging as conv_tc.cu / wgrad_mma.cu: an x-plane is a flat array of (y, z) pixels stored in
// shared memory as [channel-plane of 8][pixel][8 x fp16].  Read with the PIXEL index as the GEMM's K dimension this is
// exactly the canonical no-swizzle MN-MAJOR UMMA operand: 8 channels (M or N) contiguous in 16 bytes, the 8 pixels of
// a K-group 16 bytes apart (LBO = 128 B to the next group), channel planes SBO = plane stride apart.  A filter tap
// only moves the START ADDRESS of the A operand by a whole number of pixels, so
//   D[tap] (M = Cin rows, N = Cout columns, fp32 in TMEM) += A(a-plane shifted by the tap)^T * B(dy-plane)
// is one tcgen05.mma (M = 128, K = 16 pixels) per tap and 16-pixel chunk, with nothing re-arranged.  Rows >= Cin of the
// M = 128 tile read whatever follows in shared memory and are never stored.
//
// A CTA owns (image n, x-segment, run of M flat positions, group of taps): the taps x Cout accumulators of a layer
// (up to 18 x 128 columns) exceed the 512 TMEM columns, so the taps are split into groups of <= 512 / Cout over CTAs --
// on these levels the pixels are few and the extra staging is cheap, the tap split is what fills the SMs.
// Roles (288 threads): warps 0-3 epilogue (TMEM lane = ci; vector red.add into the caller's zeroed accumulator),
// warps 4-7 producers (cp.async, previous layer's BatchNorm + ReLU applied in place, wrap-around / out-of-range dy
// positions zero-filled), warp 8 TMEM allocation + MMA issue.
//
// wgrad_mma.cu (mma.sync) stays the kernel of the 8/16-channel levels, where M = 128 rows would be 94 % padding.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace hcu {
namespace wg5 {
using namespace ptx;

constexpr int kThreads = 288;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxTaps = 64;

struct Params {
  const __half* a;
  const __half* dy;
  float* wacc;  // fp32 [taps][cin][cout], zeroed by the caller
  const float* a_scale;
  const float* a_shift;
  int N, IX, IY, IZ, Cp, P, cin;
  int OX, OY, OZ, Cop, Po, cout;
  int KX, KY, KZ, dx, dy_, dz, px, py, pz;
  int Yv, Zv;
  int M, RUN, PS, SLOT, DPS, DSLOT, R, RD, D;
  int Nc, TG, NG, taps;
  // dy addressing (elements): image / x / y / z strides of the COARSE grid, stride-phase decomposition of the channel planes
  long long d_is, d_xs, d_ph[3];
  int d_ys, d_zs, dps[3], Pc_o;
  int flat;        // 2D: all images stacked into ONE flat plane, tap groups = filter rows, the CTA marches over runs: all images stacked into ONE flat plane, the CTA marches over runs of M positions (OX = runs)
  int n_cb, n_ob;  // channel blocks over CTAs: 128 input channels (16 planes) x Nc output channels each; P / Po = planes per block
  int n_runs, Lx, n_xseg;
  int in_relu, vec4;
  int off_d, off_bar, smem_bytes, tmem_cols;
  // per tap: tx and the byte offset (>> 4) of its (ty, tz) shift inside a ring slot
  int tap_tx[kMaxTaps];
  int tap_off[kMaxTaps];
};

// ---- PTX wrappers (see conv_tc.cu for the commented originals) -------------------------------------------------

__device__ __forceinline__ uint4 bn_relu8(uint4 v, const float* sc, const float* sh, int relu) {
  __half2* h = reinterpret_cast<__half2*>(&v);
  const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = __half22float2(h[k]);
    f.x = fmaf(f.x, sc[2 * k], sh[2 * k]);
    f.y = fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]);
    h[k] = __floats2half2_rn(f.x, f.y);
    if (relu) h[k] = __hmax2_nan(h[k], zero);
  }
  return v;
}

// MN-major, no swizzle: LBO = next 8-row K group, SBO = next 8-element M/N group (channel plane)
__device__ __forceinline__ uint64_t desc_hi_mn(uint32_t sbo_bytes) {
  return (uint64_t)(((sbo_bytes >> 4) & 0x3FFF) | (1u << 14)) << 32;  // SBO | descriptor version 1 (bit 46)
}

// NPW = producer warps: 4 (288 threads, two CTAs per SM) or 8 (416 threads) for the configurations whose shared memory
// leaves one CTA per SM anyway -- there the staging (cp.async + index arithmetic + BatchNorm in place) by 4 warps took as
// long as the MMAs of a step (ncu: tensor pipe 13 % active, L2 25 %).
template <int NPW>
__global__ void __launch_bounds__(160 + 32 * NPW, NPW == 4 ? 2 : 1) wgrad_tc5_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R, RD = p.RD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // barrier map: full_a[R], empty_a[R], full_d[RD], empty_d[RD], done
  const uint32_t bar_fa = smem_u32(bars), bar_ea = bar_fa + 8 * R, bar_fd = bar_ea + 8 * R, bar_ed = bar_fd + 8 * RD,
                 bar_done = bar_ed + 8 * RD;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * (2 * R + 2 * RD + 1));
  const uint32_t a_base = smem_u32(smem), d_base = smem_u32(smem + p.off_d);

  // ---- work item ---------------------------------------------------------------------------------
  int item = blockIdx.x;
  const int grp = item % p.NG; item /= p.NG;
  const int cb = item % p.n_cb; item /= p.n_cb;
  const int ob = item % p.n_ob; item /= p.n_ob;
  const int run = item % p.n_runs; item /= p.n_runs;
  const int xs = item % p.n_xseg;
  const int n = item / p.n_xseg;
  const int x0 = xs * p.Lx;
  const int nout = min(p.Lx, p.OX - x0);
  const int q0 = run * p.M;
  const int t_lo = grp * p.TG, t_hi = min(p.taps, t_lo + p.TG);
  const int txlo = p.tap_tx[t_lo], txhi = p.tap_tx[t_hi - 1];
  const int span = (txhi - txlo) * p.dx + 1;   // a-planes one output plane needs
  const int nplanes = nout + span - 1;          // a-planes this CTA stages: virtual x = x0 + txlo*dx + j

  if (warp == 4 + NPW) {
    if (lane == 0) {
      for (int i = 0; i < R; ++i) { mbar_init(bar_fa + 8 * i, NPW); mbar_init(bar_ea + 8 * i, 1); }
      for (int i = 0; i < RD; ++i) { mbar_init(bar_fd + 8 * i, NPW); mbar_init(bar_ed + 8 * i, 1); }
      mbar_init(bar_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 4 + NPW) {
    // =========================================== PRODUCERS ===========================================
    const int ptid = threadIdx.x - 128;
    // a: [P planes][RUN pixels]
    const int plane = ptid % p.P, pix0 = ptid / p.P, pstep = (32 * NPW) / p.P;
    const int nchunk = (p.RUN - pix0 + pstep - 1) / pstep;
    const bool xf = p.a_scale != nullptr;
    const int relu = p.in_relu;
    float sc[8], sh[8];
    if (xf) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = p.a_scale[(cb * p.P + plane) * 8 + j]; sh[j] = p.a_shift[(cb * p.P + plane) * 8 + j]; }
    }
    const int ystep = pstep / p.Zv, zstep = pstep - ystep * p.Zv;
    const int flat = p.flat;
    const __half* a_n = p.a + (size_t)n * p.IX * p.IY * p.IZ * p.Cp + (cb * p.P + plane) * 8;
    const size_t a_xs = flat ? 0 : (size_t)p.IY * p.IZ * p.Cp;
    const size_t a_is = (size_t)p.IY * p.IZ * p.Cp;   // flat: image stride (IX == 1)
    // dy: [Po planes][M pixels]
    const int dplane = ptid % p.Po, dpix0 = ptid / p.Po, dstep = (32 * NPW) / p.Po;
    const int nchunk_d = (p.M - dpix0 + dstep - 1) / dstep;
    const int dystep = dstep / p.Zv, dzstep = dstep - dystep * p.Zv;
    // channel plane -> (stride phase, 8-channel group): a phase selects a sub-lattice of the full-resolution dy tensor
    long long dplane_off = (long long)(ob * p.Po + dplane) * 8;
    if (p.dps[0] * p.dps[1] * p.dps[2] > 1) {
      const int vp = ob * p.Po + dplane;
      int phi = vp / p.Pc_o;
      const int cg = vp - phi * p.Pc_o;
      const int fz = phi % p.dps[2]; phi /= p.dps[2];
      const int fy = phi % p.dps[1], fx = phi / p.dps[1];
      dplane_off = (long long)cg * 8 + fx * p.d_ph[0] + fy * p.d_ph[1] + fz * p.d_ph[2];
    }
    const __half* d_n = p.dy + (size_t)n * (size_t)p.d_is + dplane_off;
    const size_t d_xs = flat ? 0 : (size_t)p.d_xs;
    const size_t d_is = (size_t)p.d_is;
    // first position of this thread in step j: (image, virtual row, virtual column).  Normal mode: the run is fixed and the
    // step moves along x; flat mode: the step IS the run (x0 + j), positions run over the stacked images.
    const int a_shift = flat ? p.tap_off[t_lo] : 0;  // flat: the group's filter row moves the staged window, not the taps
    auto first = [&](int step, int px0, int& im, int& yv, int& zv) {
      const int qf = (flat ? (x0 + step) * p.M : q0) + px0;
      const int r = qf / p.Zv;
      zv = qf - r * p.Zv;
      if (flat) { im = r / p.Yv; yv = r - im * p.Yv; } else { im = 0; yv = r; }
    };

    const int D = p.D;
    // One pipeline over "steps": step j stages a-plane j (j < nplanes) and dy-plane j - (span - 1) (when >= 0).
    int sa_i = 0, sa_f = 0, sd_i = 0, sd_f = 0;
    uint32_t par_a = 1, par_d = 1;
    auto finish = [&](int jf) {
      // a-plane jf
      if (xf) {
        const int xm = x0 + txlo * p.dx + jf - p.px;
        if (flat || (xm >= 0 && xm < p.IX)) {
          unsigned char* dp = smem + sa_f * p.SLOT + plane * p.PS + pix0 * 16;
          int im, yv, zv;
          first(jf, pix0 + a_shift, im, yv, zv);
          for (int c = 0; c < nchunk; ++c) {
            const int ym = yv - p.py, zm = zv - p.pz;
            if (im < p.N && ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ) {
              uint4* q = reinterpret_cast<uint4*>(dp + (size_t)c * pstep * 16);
              *q = bn_relu8(*q, sc, sh, relu);
            }
            zv += zstep; yv += ystep;
            if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
            if (flat && yv >= p.Yv) { yv -= p.Yv; ++im; }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_fa + 8 * sa_f);
        if (jf >= span - 1) mbar_arrive(bar_fd + 8 * sd_f);
      }
      if (++sa_f == R) sa_f = 0;
      if (jf >= span - 1 && ++sd_f == RD) sd_f = 0;
    };
    for (int j = 0; j < nplanes + D; ++j) {
      if (D > 0 && j >= D) {
        if (D == 1) cp_async_wait<0>();
        else cp_async_wait<1>();
        finish(j - D);
      }
      if (j < nplanes) {
        mbar_wait(bar_ea + 8 * sa_i, par_a);
        {
          const int xm = x0 + txlo * p.dx + j - p.px;
          const bool xok = flat || (xm >= 0 && xm < p.IX);
          const __half* a_x = a_n + (size_t)(xok ? xm : 0) * a_xs;
          const uint32_t dst = a_base + (uint32_t)(sa_i * p.SLOT + plane * p.PS + pix0 * 16);
          int im, yv, zv;
          first(j, pix0 + a_shift, im, yv, zv);
          for (int c = 0; c < nchunk; ++c) {
            const int ym = yv - p.py, zm = zv - p.pz;
            const bool ok = xok && im < p.N && ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ;
            cp_async16(dst + c * pstep * 16, ok ? a_x + (size_t)im * a_is + ((size_t)ym * p.IZ + zm) * p.Cp : a_n, ok ? 16u : 0u);
            zv += zstep; yv += ystep;
            if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
            if (flat && yv >= p.Yv) { yv -= p.Yv; ++im; }
          }
        }
        if (++sa_i == R) { sa_i = 0; par_a ^= 1; }
        if (j >= span - 1) {
          mbar_wait(bar_ed + 8 * sd_i, par_d);
          const int i = j - (span - 1);  // output plane
          const __half* d_x = d_n + (size_t)(x0 + i) * d_xs;
          const uint32_t dst = d_base + (uint32_t)(sd_i * p.DSLOT + dplane * p.DPS + dpix0 * 16);
          int im, oy, oz;
          first(i, dpix0, im, oy, oz);
          for (int c = 0; c < nchunk_d; ++c) {
            const bool ok = im < p.N && oy < p.OY && oz < p.OZ;
            cp_async16(dst + c * dstep * 16, ok ? d_x + (size_t)im * d_is + ((size_t)oy * p.d_ys + (size_t)oz * p.d_zs) : d_n, ok ? 16u : 0u);
            oz += dzstep; oy += dystep;
            if (oz >= p.Zv) { oz -= p.Zv; ++oy; }
            if (flat && oy >= p.Yv) { oy -= p.Yv; ++im; }
          }
          if (++sd_i == RD) { sd_i = 0; par_d ^= 1; }
        }
      }
      cp_async_commit();
      if (D == 0) {
        cp_async_wait<0>();
        finish(j);
      }
    }
  } else if (warp == 4 + NPW) {
    // =========================================== MMA ISSUER ==========================================
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Nc >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_hi = desc_hi_mn((uint32_t)p.PS), b_hi = desc_hi_mn((uint32_t)p.DPS);
    const uint32_t lbo = (128u >> 4) << 16;
    const int nchunks = p.M / 16;
    const int a_shift_m = p.flat ? p.tap_off[t_lo] : 0;
    int wa = 0, wd = 0, next_a = 0;
    uint32_t pa = 0, pd = 0;
    int i_mod = 0;  // ring slot of a-plane i (the oldest plane output i needs)
    for (int i = 0; i < nout; ++i) {
      for (; next_a <= i + span - 1; ++next_a) {
        mbar_wait(bar_fa + 8 * wa, pa);
        if (++wa == R) { wa = 0; pa ^= 1; }
      }
      mbar_wait(bar_fd + 8 * wd, pd);
      tc_fence_after();
      const uint32_t bbase = ((d_base + (uint32_t)(wd * p.DSLOT)) >> 4) | lbo;
      for (int t = t_lo; t < t_hi; ++t) {
        int sl = i_mod + (p.tap_tx[t] - txlo) * p.dx;
        sl -= sl >= R ? R : 0;
        const uint32_t abase = (((a_base + (uint32_t)(sl * p.SLOT)) >> 4) + (uint32_t)(p.tap_off[t] - a_shift_m)) | lbo;
        const uint32_t tcol = tmem_base + (uint32_t)((t - t_lo) * p.Nc);
        const uint32_t first = (uint32_t)i;  // accumulate flag of chunk 0: overwrite only on the CTA's first plane
        if (elect_one()) {
          // nchunks is a multiple of 4 (M = 64 / 128 / 256): four MMAs per trip, descriptors advance by constants
          for (int c = 0; c < nchunks; c += 4) {
            const uint64_t ad = a_hi | (uint64_t)(abase + (uint32_t)(c * 16)), bd = b_hi | (uint64_t)(bbase + (uint32_t)(c * 16));
            umma_f16(tcol, ad, bd, idesc, first | (uint32_t)c);
            umma_f16(tcol, ad + 16u, bd + 16u, idesc, 1u);
            umma_f16(tcol, ad + 32u, bd + 32u, idesc, 1u);
            umma_f16(tcol, ad + 48u, bd + 48u, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(bar_ea + 8 * i_mod);  // a-plane i is not needed by later outputs
        umma_commit(bar_ed + 8 * wd);
        if (i == nout - 1) umma_commit(bar_done);
      }
      __syncwarp();
      if (++wd == RD) { wd = 0; pd ^= 1; }
      i_mod = i_mod + 1 == R ? 0 : i_mod + 1;
    }
  } else {
    // =========================================== EPILOGUE ============================================
    const int ci = cb * 128 + threadIdx.x;  // TMEM lane == accumulator row == input channel (of this CTA's block)
    const int co0 = ob * p.Nc;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int t = t_lo; t < t_hi; ++t) {
      for (int cc = 0; cc < p.Nc; cc += 16) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)((t - t_lo) * p.Nc + cc), v);
        if (ci < p.cin) {
          float* o = p.wacc + ((size_t)t * p.cin + ci) * p.cout + co0 + cc;
          if (p.vec4) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (co0 + cc + j < p.cout) red_add_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (co0 + cc + j < p.cout) atomicAdd(o + j, v[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 + NPW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static const char* configure(const HcuConvDesc* d, Params& p) {
  if (d->dtype_in != HCU_F16 || d->dtype_out != HCU_F16) return "fp16 only";
  if (d->groups != 1) return "groups != 1";
  if (d->iphase) return "iphase";
  for (int i = 0; i < 3; ++i) p.dps[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff);
  const int nph = p.dps[0] * p.dps[1] * p.dps[2];
  if (nph > 1 && (d->cout != d->out_cpitch || d->cout % nph || (d->cout / nph) % 8)) return "ophase needs 8-channel aligned phases";
  if (d->in_cpitch % 8 != 0 || d->in_c_off != 0 || d->cin > d->in_cpitch) return "input channel layout";
  if (d->out_cpitch % 8 != 0 || d->out_c_off != 0 || d->cout > d->out_cpitch) return "dy channel layout";
  const int Pt = d->in_cpitch / 8, Pot = d->out_cpitch / 8;  // channel planes of the whole tensors
  // channel blocks: a CTA takes <= 128 input channels (the M = 128 rows of the accumulator) and <= 128 output channels
  const int P = std::min(Pt, 16), Po = std::min(Pot, 16);
  if (P != 1 && P != 2 && P != 4 && P != 8 && P != 16) return "input channel pitch";
  if (Po != 1 && Po != 2 && Po != 4 && Po != 8 && Po != 16) return "dy channel pitch";
  if (Pt % P != 0 || Pot % Po != 0) return "channel pitch not a whole number of blocks";
  p.n_cb = Pt / P; p.n_ob = Pot / Po;
  if ((p.n_cb > 1 && d->cin != d->in_cpitch) || (p.n_ob > 1 && d->cout != d->out_cpitch)) return "channel blocks need dense channels";
  for (int i = 0; i < 3; ++i)
    if (d->istep[i] != 1 || d->ostep[i] != p.dps[i] || d->ooff[i] != 0 || d->out_tsize[i] != d->out_size[i] * p.dps[i]) return "strided";
  p.N = d->batch;
  p.Cp = d->in_cpitch; p.P = P; p.cin = d->cin;
  p.Cop = d->out_cpitch; p.Po = Po; p.cout = d->cout;
  // A 2D problem is re-read as one flat plane of rows x columns with all images of the batch stacked (rows = N * Yv): a
  // filter row becomes a flat shift of Zv positions, the CTA marches over consecutive RUNS of M positions (the kernel's
  // "x" axis is the run index) and every 16-pixel K chunk of an MMA is full, instead of one image row per step (a
  // 30-pixel row of the classic U-Net's bottom levels fills 12 % of a 256-position run, a 138-pixel row 54 %).  The tap
  // groups are whole filter rows: the group's row only moves the staged window, so a run carries KZ - 1 positions of halo.
  static int flat_on = -1;
  if (flat_on < 0) { const char* e = getenv("HCU_WG5_FLAT"); flat_on = e ? atoi(e) : 1; }
  const int nc_blk = (Pot / Po) > 1 ? Po * 8 : round_up(d->cout, 16);
  const bool flat2d = flat_on && d->in_size[2] == 1 && d->out_size[2] == 1 && d->taps[2] == 1 && d->pad[2] == 0 && d->dil[2] == 1 &&
                      nc_blk <= 256 && 512 / nc_blk >= d->taps[1] &&
                      d->in_cpitch >= 64;  // 32-channel levels: the row-by-row march stages every plane once (measured 1.13 vs 1.49 ms)
  p.flat = flat2d ? 1 : 0;
  {
    // dy (ophase: the `cout` channels are [nph][cout / nph], phase phi of coarse position o lives at o * s + phi of the
    // full-resolution tensor whose channel pitch is cout / nph)
    const long long cr = d->out_cpitch / nph;
    const long long tz = cr, ty = tz * d->out_tsize[2], tx = ty * d->out_tsize[1];
    const long long cs[3] = {tx * p.dps[0], ty * p.dps[1], tz * p.dps[2]};  // coarse strides per descriptor axis
    if (cs[0] >= 0x7fffffffLL) return "dy plane too large";
    p.d_is = tx * d->out_tsize[0];
    p.d_ph[0] = tx; p.d_ph[1] = ty; p.d_ph[2] = tz;
    p.Pc_o = (int)(cr / 8);
    if (flat2d) { p.d_xs = 0; p.d_ys = (int)cs[0]; p.d_zs = (int)cs[1]; }
    else { p.d_xs = cs[0]; p.d_ys = (int)cs[1]; p.d_zs = (int)cs[2]; }
  }
  if (flat2d) {
    p.IX = 1; p.IY = d->in_size[0]; p.IZ = d->in_size[1];
    p.OX = 1; p.OY = d->out_size[0]; p.OZ = d->out_size[1];
    p.KX = 1; p.KY = d->taps[0]; p.KZ = d->taps[1];
    p.dx = 1; p.dy_ = d->dil[0]; p.dz = d->dil[1];
    p.px = 0; p.py = d->pad[0]; p.pz = d->pad[1];
  } else {
    p.IX = d->in_size[0]; p.IY = d->in_size[1]; p.IZ = d->in_size[2];
    p.OX = d->out_size[0]; p.OY = d->out_size[1]; p.OZ = d->out_size[2];
    p.KX = d->taps[0]; p.KY = d->taps[1]; p.KZ = d->taps[2];
    p.dx = d->dil[0]; p.dy_ = d->dil[1]; p.dz = d->dil[2];
    p.px = d->pad[0]; p.py = d->pad[1]; p.pz = d->pad[2];
  }
  p.Yv = p.OY + (p.KY - 1) * p.dy_;
  p.Zv = p.OZ + (p.KZ - 1) * p.dz;
  p.taps = p.KX * p.KY * p.KZ;
  if (p.taps > kMaxTaps) return "too many taps";
  p.Nc = p.n_ob > 1 ? Po * 8 : round_up(d->cout, 16);
  if (p.Nc > 256) return "too many output channels";
  p.TG = std::min(p.taps, 512 / p.Nc);
  {
    static int tgmax = -1;
    if (tgmax < 0) { const char* e = getenv("HCU_WG5_TG"); tgmax = e ? atoi(e) : 0; }
    if (tgmax > 0) p.TG = std::min(p.TG, tgmax);
  }
  p.NG = (p.taps + p.TG - 1) / p.TG;
  p.TG = (p.taps + p.NG - 1) / p.NG;  // balance the groups
  int flat_rpg = 1;
  if (p.flat) {  // whole filter rows per group; several rows only while their halo stays small
    flat_rpg = std::max(1, std::min(p.KY, (512 / p.Nc) / p.KZ));
    if ((flat_rpg - 1) * p.dy_ * p.Zv + (p.KZ - 1) * p.dz > 64) flat_rpg = 1;
    p.TG = flat_rpg * p.KZ;
    p.NG = (p.KY + flat_rpg - 1) / flat_rpg;
  }
  {
    int cols = p.TG * p.Nc, t = 32;
    while (t < cols) t <<= 1;
    p.tmem_cols = t;
  }
  const int halo = p.flat ? (flat_rpg - 1) * p.dy_ * p.Zv + (p.KZ - 1) * p.dz : (p.KY - 1) * p.dy_ * p.Zv + (p.KZ - 1) * p.dz;
  // flat: positions of the stacked images that hold outputs (the tail rows of the last image are never needed)
  const long long plane_qq = p.flat ? ((long long)(p.N - 1) * p.Yv + p.OY - 1) * p.Zv + p.OZ : (long long)p.Yv * p.Zv;
  if (plane_qq >= 0x3fffffffLL) return "plane too large";
  const int plane_q = (int)plane_qq;
  // widest x extent a tap group can have
  int max_span = 1;
  for (int g = 0; g < p.NG; ++g) {
    const int lo = g * p.TG, hi = std::min(p.taps, lo + p.TG) - 1;
    max_span = std::max(max_span, (hi / (p.KY * p.KZ) - lo / (p.KY * p.KZ)) * p.dx + 1);
  }
  if (max_span > 8) return "x extent";
  const int m_cands[3] = {256, 128, 64};
  const int want[3] = {3, 2, 0};  // ring slack beyond the span: (look-ahead D, published slack) = (2,1), (1,1), (0,0)
  // normal mode: deepest ring first, then the longest run; flat mode: the longest run that still gets a pipelined ring
  for (int it = 0; it < 9; ++it) {
    {
      const int wi = p.flat ? it % 3 : it / 3, mi = p.flat ? it / 3 : it % 3;
      if (p.flat && wi == 2 && mi < 2) continue;  // a synchronous ring only for the shortest run
      const int M = m_cands[mi];
      if (M > 64 && M / 2 >= plane_q) continue;
      const int run = M + halo;
      int ps = run * 16, dps = M * 16;
      // spread the channel planes over the banks for the producers' 16-byte accesses (conv_tc.cu does the same):
      // plane stride = g (mod 2g), g = max(16, 128 / planes)
      if (P > 1) { const int g = P >= 8 ? 16 : 128 / P; ps = round_up(ps, 2 * g) + g; }
      if (Po > 1) { const int g = Po >= 8 ? 16 : 128 / Po; dps = round_up(dps, 2 * g) + g; }
      const int slot = ps * P, dslot = dps * Po;
      const int R = max_span + want[wi], RD = want[wi] >= 2 ? 3 : 2;
      const int D = want[wi] >= 3 ? 2 : (want[wi] >= 2 ? 1 : 0);
      // the M = 128 / N = Nc tiles read 16 (Nc / 8) planes from a slot's base: keep those reads inside the allocation
      const int a_end = (R - 1) * slot + 16 * ps;
      const int off_d = round_up(std::max(R * slot, a_end), 128);
      const int d_end = off_d + (RD - 1) * dslot + (p.Nc / 8) * dps;
      const int off_bar = round_up(std::max(off_d + RD * dslot, d_end), 128);
      const int total = off_bar + 8 * (2 * R + 2 * RD + 1) + 16 + 128;
      if (total > kSmemLimit) continue;
      if (ps / 16 > 0x3FFF || dps / 16 > 0x3FFF) continue;
      p.M = M; p.RUN = run; p.PS = ps; p.SLOT = slot; p.DPS = dps; p.DSLOT = dslot; p.R = R; p.RD = RD; p.D = D;
      p.off_d = off_d; p.off_bar = off_bar; p.smem_bytes = total;
      p.n_runs = (plane_q + M - 1) / M;
      if (p.flat) {  // the run index becomes the kernel's x axis (segmented over CTAs by the launch code)
        p.OX = p.n_runs;
        p.n_runs = 1;
      }
      for (int t = 0; t < p.taps; ++t) {
        const int tz = t % p.KZ, tq = t / p.KZ;
        const int ty = tq % p.KY, tx = tq / p.KY;
        p.tap_tx[t] = tx;
        p.tap_off[t] = ty * p.dy_ * p.Zv + tz * p.dz;  // pixels == 16-byte units
      }
      return nullptr;
    }
  }
  return "does not fit in shared memory";
}

}  // namespace wg5
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_wgrad_tc5_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  wg5::Params p;
  return wg5::configure(d, p) == nullptr ? 1 : 0;
}

extern "C" int hcu_conv_wgrad_tc5_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                      const void* dy, float* wacc, void* stream) {
  HCU_CHECK_ARG(d && a && dy && wacc, "wgrad_tc5: null pointer");
  HCU_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "wgrad_tc5: a_scale/a_shift must come together");
  wg5::Params p;
  const char* why = wg5::configure(d, p);
  if (why != nullptr) {
    set_error("wgrad_tc5: unsupported descriptor (%s)", why);
    return HCU_ERR_UNSUPPORTED;
  }
  p.a = (const __half*)a; p.dy = (const __half*)dy; p.wacc = wacc; p.a_scale = a_scale; p.a_shift = a_shift;
  p.in_relu = d->in_relu;
  p.vec4 = (d->cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(wacc) & 15) == 0);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wg5::wgrad_tc5_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg5::kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wg5::wgrad_tc5_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg5::kSmemLimit);
    if (e != cudaSuccess) { set_error("wgrad_tc5: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
    attr = true;
  }
  // x segmentation: about two waves of CTAs, segments no shorter than 4 planes
  const long long base_items = (long long)(p.flat ? 1 : p.N) * p.n_runs * p.NG * p.n_cb * p.n_ob;
  const int per_sm = std::max(1, std::min(233472 / (p.smem_bytes + 1024), 512 / p.tmem_cols));
  const long long target = 2LL * per_sm * num_sms();
  int nseg = (int)((target + base_items - 1) / base_items);
  nseg = std::max(1, std::min(nseg, (p.OX + 3) / 4));
  {
    static int fseg = -1;
    if (fseg < 0) { const char* e = getenv("HCU_WG5_NSEG"); fseg = e ? atoi(e) : 0; }
    if (fseg > 0) nseg = std::min(fseg, p.OX);
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("HCU_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (dbg & 8) fprintf(stderr, "wgrad_tc5: M %d R %d RD %d D %d Nc %d TG %d NG %d n_runs %d nseg %d smem %d tmem %d per_sm %d\n", p.M, p.R, p.RD, p.D, p.Nc, p.TG, p.NG, p.n_runs, nseg, p.smem_bytes, p.tmem_cols, per_sm);
  }
  p.Lx = (p.OX + nseg - 1) / nseg;
  p.n_xseg = (p.OX + p.Lx - 1) / p.Lx;
  const long long grid = base_items * p.n_xseg;
  HCU_CHECK_ARG(grid <= 0x7fffffffLL, "wgrad_tc5: grid too large");
  static int npw8 = -1;
  if (npw8 < 0) { const char* e = getenv("HCU_WG5_NPW8"); npw8 = e ? atoi(e) : 1; }
  if (npw8 && per_sm == 1)   // one CTA per SM anyway: twice the producer warps
    wg5::wgrad_tc5_kernel<8><<<(unsigned)grid, 160 + 32 * 8, p.smem_bytes, (cudaStream_t)stream>>>(p);
  else
    wg5::wgrad_tc5_kernel<4><<<(unsigned)grid, 160 + 32 * 4, p.smem_bytes, (cudaStream_t)stream>>>(p);
  HCU_CHECK_LAUNCH("wgrad_tc5");
  return 0;
}
