// Weight gradient of the gather-convolution for the 8/16-channel levels: warp-specialised mma.sync pipeline.
//
//   dW[tap][ci][co] = sum_{n, q} act(a[n, q + shift(tap), ci]) * dy[n, q, co]
//
// Same arithmetic and shared-memory layout as wgrad_mma.cu (flat-shift planes [channel-plane of 8][pixel][8 x fp16],
// ldmatrix.trans fragments, m16n8k16, fp32 accumulate), but the roles are split like conv_tc.cu:
//   warps 0-3  consumers: per 16-pixel block one ldmatrix of dy + (ldmatrix + MMA) per m-tile, nothing else;
//   warps 4-7  producers: cp.async of the next activation / dy planes (zero-fill for padding and wrap-around), the previous
//              layer's BatchNorm + ReLU applied in place, planes published through mbarriers, D planes in flight.
// wgrad_mma.cu interleaves both jobs in every thread: ~1200 instructions per warp and plane with two block-wide barriers,
// 164 registers (3 CTAs of 4 warps per SM) -- ~10 k cycles per plane.  Here a consumer warp runs ~200 instructions per
// plane and never waits for global memory.  These levels are bound by the shared-memory reads of the A fragments
// (each tap re-reads the plane: taps * Cin * 2 B per pixel), see DESIGN.md.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace hcu {
namespace wgs {
using namespace ptx;

constexpr int kThreads = 256;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxChunk = 6;

struct Params {
  const __half* a;
  const __half* dy;
  float* wacc;  // fp32 [taps][cin][cout], zeroed by the caller; accumulated with atomics
  const float* a_scale;
  const float* a_shift;
  int N, IX, IY, IZ, Cp, P, cin;
  int OX, OY, OZ, Cop, Po, cout;
  int KX, KY, KZ, dx, dy_, dz, px, py, pz;
  int Yv, Zv;
  int M, RUN, PS, SLOT, DPS, DSLOT, R, RD, D;
  int E, MTOT, NTOT;
  int n_runs, Lx, n_xseg, n_mchunk;
  int in_relu;
  int off_d, off_bar, smem_bytes;
};


// MTC m-tiles (16 rows = 2 (tap, channel-plane) slots) x NTC n-tiles (8 output channels) per consumer warp; the four
// consumer warps split the 16-pixel blocks of a plane and are reduced in shared memory at the end.
template <int MTC, int NTC>
__global__ void __launch_bounds__(kThreads, 2) wgrad_ws_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, RD = p.RD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // barrier map: full[R] (a-plane j and, from step span-1 on, dy-plane j-(span-1)), empty_a[R], empty_d[RD]
  const uint32_t bar_f = smem_u32(bars), bar_ea = bar_f + 8 * R, bar_ed = bar_ea + 8 * R;
  const uint32_t a_base = smem_u32(smem), d_base = smem_u32(smem + p.off_d);

  int item = blockIdx.x;
  const int run = item % p.n_runs; item /= p.n_runs;
  const int xs = item % p.n_xseg;
  const int n = item / p.n_xseg;
  const int mt0 = blockIdx.y * MTC;
  const int x0 = xs * p.Lx;
  const int nout = min(p.Lx, p.OX - x0);
  const int span = (p.KX - 1) * p.dx + 1;
  const int nplanes = nout + span - 1;
  const int q0 = run * p.M;

  if (tid == 0) {
    for (int i = 0; i < R; ++i) { mbar_init(bar_f + 8 * i, 4); mbar_init(bar_ea + 8 * i, 4); }
    for (int i = 0; i < RD; ++i) mbar_init(bar_ed + 8 * i, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float acc[MTC][NTC][4];
#pragma unroll
  for (int m = 0; m < MTC; ++m)
#pragma unroll
    for (int nn = 0; nn < NTC; ++nn)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[m][nn][k] = 0.f;

  if (warp >= 4) {
    // =========================================== PRODUCERS ===========================================
    const int ptid = tid - 128;
    const int plane = ptid % p.P, pix0 = ptid / p.P, pstep = 128 / p.P;
    const int nchunk = (p.RUN - pix0 + pstep - 1) / pstep;
    const int nmax = (p.RUN + pstep - 1) / pstep;
    const bool xf = p.a_scale != nullptr;
    float sc[8], sh[8];
    if (xf) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = p.a_scale[plane * 8 + j]; sh[j] = p.a_shift[plane * 8 + j]; }
    }
    const int qf = q0 + pix0;
    const int yv0 = qf / p.Zv, zv0 = qf - yv0 * p.Zv;
    const int ystep = pstep / p.Zv, zstep = pstep - ystep * p.Zv;
    const __half* a_n = p.a + (size_t)n * p.IX * p.IY * p.IZ * p.Cp + plane * 8;
    const size_t a_xs = (size_t)p.IY * p.IZ * p.Cp;
    int goff[kMaxChunk];  // element offset inside an x-plane; -1 zero fill; -2 not this thread's
    {
      int yv = yv0, zv = zv0;
#pragma unroll
      for (int c = 0; c < kMaxChunk; ++c) {
        const int ym = yv - p.py, zm = zv - p.pz;
        const bool ok = ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ;
        goff[c] = c < nchunk ? (ok ? (ym * p.IZ + zm) * p.Cp : -1) : -2;
        zv += zstep; yv += ystep;
        if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
      }
    }
    // dy: [Po planes][M pixels]; positions that wrap around a row / fall outside the output are zero
    const int dplane = ptid % p.Po, dpix0 = ptid / p.Po, dstep = 128 / p.Po;
    const int nchunk_d = (p.M - dpix0 + dstep - 1) / dstep;  // <= kMaxChunk (M <= 512, checked on the host)
    const __half* d_n = p.dy + (size_t)n * p.OX * p.OY * p.OZ * p.Cop + dplane * 8;
    const size_t d_xs = (size_t)p.OY * p.OZ * p.Cop;
    int doff[kMaxChunk];
    {
      const int dqf = q0 + dpix0;
      int oy = dqf / p.Zv, oz = dqf - oy * p.Zv;
      const int dystep = dstep / p.Zv, dzstep = dstep - dystep * p.Zv;
#pragma unroll
      for (int c = 0; c < kMaxChunk; ++c) {
        doff[c] = c < nchunk_d ? ((oy < p.OY && oz < p.OZ) ? (oy * p.OZ + oz) * p.Cop : -1) : -2;
        oz += dzstep; oy += dystep;
        if (oz >= p.Zv) { oz -= p.Zv; ++oy; }
      }
    }
    const int D = p.D;
    int sa_i = 0, sa_f = 0, sd_i = 0;
    uint32_t par_a = 1, par_d = 1;
    const __half2 zero = __float2half2_rn(0.f);
    auto finish = [&](int jf) {  // a-plane jf (and the dy plane of the same step) has landed
      if (xf) {
        const int xm = x0 + jf - p.px;
        if (xm >= 0 && xm < p.IX) {
          unsigned char* dp = smem + sa_f * p.SLOT + plane * p.PS + pix0 * 16;
#pragma unroll
          for (int c = 0; c < kMaxChunk; ++c) {
            if (c < nmax && goff[c] >= 0) {
              uint4* q = reinterpret_cast<uint4*>(dp + c * pstep * 16);
              uint4 v = *q;
              __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 f = __half22float2(h[k]);
                f.x = fmaf(f.x, sc[2 * k], sh[2 * k]);
                f.y = fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]);
                h[k] = __floats2half2_rn(f.x, f.y);
                if (p.in_relu) h[k] = __hmax2_nan(h[k], zero);
              }
              *q = v;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_f + 8 * sa_f);  // release: this warp's copies and stores of the step
      if (++sa_f == R) sa_f = 0;
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
          const int xm = x0 + j - p.px;
          const bool xok = xm >= 0 && xm < p.IX;
          const __half* a_x = a_n + (size_t)(xok ? xm : 0) * a_xs;
          const uint32_t dst = a_base + (uint32_t)(sa_i * p.SLOT + plane * p.PS + pix0 * 16);
#pragma unroll
          for (int c = 0; c < kMaxChunk; ++c) {
            if (c < nmax && goff[c] != -2) {
              const bool ok = xok && goff[c] >= 0;
              cp_async16(dst + c * pstep * 16, ok ? a_x + goff[c] : a_n, ok ? 16u : 0u);
            }
          }
        }
        if (++sa_i == R) { sa_i = 0; par_a ^= 1; }
        if (j >= span - 1) {
          mbar_wait(bar_ed + 8 * sd_i, par_d);
          const __half* d_x = d_n + (size_t)(x0 + j - (span - 1)) * d_xs;
          const uint32_t dst = d_base + (uint32_t)(sd_i * p.DSLOT + dplane * p.DPS + dpix0 * 16);
#pragma unroll
          for (int c = 0; c < kMaxChunk; ++c) {
            if (doff[c] != -2) cp_async16(dst + c * dstep * 16, doff[c] >= 0 ? d_x + doff[c] : d_n, doff[c] >= 0 ? 16u : 0u);
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
  } else {
    // =========================================== CONSUMERS ===========================================
    // A (x4.trans): matrix mi = lane >> 3: slot (mi & 1) of the m-tile's pair, pixel half (mi >> 1); row = lane & 7
    // slot e = tap * P + plane, tap = (tx*KY + ty)*KZ + tz
    int a_off[MTC];
    int a_tx[MTC];
#pragma unroll
    for (int m = 0; m < MTC; ++m) {
      const int mt = mt0 + m;
      a_off[m] = 0; a_tx[m] = 0;
      {
        int e = 2 * min(mt, p.MTOT - 1) + ((lane >> 3) & 1);
        if (e >= p.E) e = p.E - 1;  // odd tail: duplicate the last slot, its rows are discarded
        const int tap = e / p.P, pl = e - tap * p.P;
        const int tz = tap % p.KZ, tq = tap / p.KZ;
        const int ty = tq % p.KY, tx = tq / p.KY;
        a_tx[m] = tx;
        a_off[m] = pl * p.PS + ((ty * p.dy_ * p.Zv + tz * p.dz) + (lane >> 4) * 8 + (lane & 7)) * 16;
      }
    }
    const int b_off = (lane & 15) * 16;
    const int nblk = p.M / 16;
    int wf = 0, next_f = 0, wd = 0, i_mod = 0;
    uint32_t pf = 0;
    for (int i = 0; i < nout; ++i) {
      for (; next_f <= i + span - 1; ++next_f) {  // planes up to i + span - 1 (that step also carries dy-plane i)
        mbar_wait(bar_f + 8 * wf, pf);
        if (++wf == R) { wf = 0; pf ^= 1; }
      }
      const uint32_t dyb = d_base + (uint32_t)(wd * p.DSLOT);
      uint32_t slot_addr[MTC];
#pragma unroll
      for (int m = 0; m < MTC; ++m) {
        int sl = i_mod + a_tx[m] * p.dx;
        sl -= sl >= R ? R : 0;
        slot_addr[m] = a_base + (uint32_t)(sl * p.SLOT + a_off[m]);
      }
      for (int blk = warp; blk < nblk; blk += 4) {
        uint32_t bf[NTC][2];
#pragma unroll
        for (int nn = 0; nn < NTC; ++nn) {
          const int nt = min(nn, p.NTOT - 1);
          ldsm_x2_t(dyb + (uint32_t)(nt * p.DPS + blk * 256 + b_off), bf[nn][0], bf[nn][1]);
        }
#pragma unroll
        for (int m = 0; m < MTC; ++m) {
          if (mt0 + m < p.MTOT) {  // warp-uniform
            uint32_t af[4];
            ldsm_x4_t(slot_addr[m] + (uint32_t)(blk * 256), af[0], af[1], af[2], af[3]);
#pragma unroll
            for (int nn = 0; nn < NTC; ++nn) mma16816(acc[m][nn], af, bf[nn]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_ea + 8 * i_mod);  // a-plane i is not needed by later outputs
        mbar_arrive(bar_ed + 8 * wd);
      }
      if (++wd == RD) wd = 0;
      i_mod = i_mod + 1 == R ? 0 : i_mod + 1;
    }
  }

  // ---- reduce the four consumer warps (pixel slices) in shared memory; one atomic per element into the global accumulator
  __syncthreads();  // every plane consumed, every copy landed: the staging buffers are free
  constexpr int RW = NTC * 8, RH = MTC * 16;
  float* red = reinterpret_cast<float*>(smem);
  for (int e = tid; e < RH * RW; e += kThreads) red[e] = 0.f;
  __syncthreads();
  if (warp < 4) {
    const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int m = 0; m < MTC; ++m)
#pragma unroll
      for (int nn = 0; nn < NTC; ++nn) {
        float* r0 = red + (m * 16 + g) * RW + nn * 8 + t2;
        atomicAdd(r0, acc[m][nn][0]);
        atomicAdd(r0 + 1, acc[m][nn][1]);
        atomicAdd(r0 + 8 * RW, acc[m][nn][2]);
        atomicAdd(r0 + 8 * RW + 1, acc[m][nn][3]);
      }
  }
  __syncthreads();
  for (int e = tid; e < RH * RW; e += kThreads) {
    const int row = e / RW, col = e - row * RW;
    const int mt = mt0 + row / 16;
    const int slot = 2 * mt + ((row & 15) >> 3);
    if (mt < p.MTOT && slot < p.E && col < p.cout && col / 8 < p.NTOT) {
      const int tap = slot / p.P, pl = slot - tap * p.P;
      const int ci = pl * 8 + (row & 7);
      if (ci < p.cin) atomicAdd(&p.wacc[((size_t)tap * p.cin + ci) * p.cout + col], red[e]);
    }
  }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static const char* configure(const HcuConvDesc* d, Params& p, int& mtc, int& ntc) {
  if (d->dtype_in != HCU_F16 || d->dtype_out != HCU_F16) return "fp16 only";
  if (d->groups != 1) return "groups != 1";
  if (d->ophase || d->iphase) return "stride phases";
  if (d->in_cpitch % 8 != 0 || d->in_c_off != 0 || d->cin > d->in_cpitch) return "input channel layout";
  if (d->out_cpitch % 8 != 0 || d->out_c_off != 0 || d->cout > d->out_cpitch) return "dy channel layout";
  const int P = d->in_cpitch / 8, Po = d->out_cpitch / 8;
  if (P != 1 && P != 2 && P != 4) return "input channel pitch above 32";
  // measured (B200, bench shapes): 15-25 % faster than wgrad_mma.cu with 8 output channels, on par / slower with 16
  // (twice the accumulators: register pressure) -- those stay on wgrad_mma.cu
  if (Po != 1) return "more than 8 output channels";
  for (int i = 0; i < 3; ++i)
    if (d->istep[i] != 1 || d->ostep[i] != 1 || d->ooff[i] != 0 || d->out_tsize[i] != d->out_size[i]) return "strided";
  p.N = d->batch; p.IX = d->in_size[0]; p.IY = d->in_size[1]; p.IZ = d->in_size[2];
  p.Cp = d->in_cpitch; p.P = P; p.cin = d->cin;
  p.OX = d->out_size[0]; p.OY = d->out_size[1]; p.OZ = d->out_size[2];
  p.Cop = d->out_cpitch; p.Po = Po; p.cout = d->cout;
  p.KX = d->taps[0]; p.KY = d->taps[1]; p.KZ = d->taps[2];
  p.dx = d->dil[0]; p.dy_ = d->dil[1]; p.dz = d->dil[2];
  p.px = d->pad[0]; p.py = d->pad[1]; p.pz = d->pad[2];
  p.Yv = p.OY + (p.KY - 1) * p.dy_;
  p.Zv = p.OZ + (p.KZ - 1) * p.dz;
  const int span = (p.KX - 1) * p.dx + 1;
  if (span > 6) return "x extent";
  p.E = p.KX * p.KY * p.KZ * P;
  p.MTOT = (p.E + 1) / 2;
  p.NTOT = Po;
  mtc = 9; ntc = Po;
  p.n_mchunk = (p.MTOT + mtc - 1) / mtc;
  const int halo = (p.KY - 1) * p.dy_ * p.Zv + (p.KZ - 1) * p.dz;
  const int plane_q = p.Yv * p.Zv;
  const int m_cands[4] = {512, 256, 128, 64};
  const int want[3] = {4, 2, 1};  // ring slack beyond the span: (D, published slack) = (2, 2), (1, 1), (0, 1)
  for (int wi = 0; wi < 3; ++wi) {
    for (int mi = 0; mi < 4; ++mi) {
      const int M = m_cands[mi];
      if (M > 64 && M / 2 >= plane_q) continue;
      const int run = M + halo;
      if ((run + 128 / P - 1) / (128 / P) > kMaxChunk) continue;   // per-thread chunk tables
      if ((M + 128 / Po - 1) / (128 / Po) > kMaxChunk) continue;
      int ps = run * 16, dps = M * 16;
      if (P > 1) { const int g = 128 / P; ps = round_up(ps, 2 * g) + g; }
      if (Po > 1) { const int g = 128 / Po; dps = round_up(dps, 2 * g) + g; }
      const int slot = ps * P, dslot = dps * Po;
      const int R = span + want[wi];
      const int D = want[wi] >= 4 ? 2 : (want[wi] >= 2 ? 1 : 0);
      const int RD = D + 2;
      const int off_d = round_up(R * slot, 128);
      const int off_bar = round_up(off_d + RD * dslot, 128);
      const int total = std::max(off_bar + 8 * (2 * R + RD) + 16, mtc * 16 * ntc * 8 * 4) + 128;
      if (total > 110 * 1024) continue;   // two CTAs per SM
      p.M = M; p.RUN = run; p.PS = ps; p.SLOT = slot; p.DPS = dps; p.DSLOT = dslot; p.R = R; p.RD = RD; p.D = D;
      p.off_d = off_d; p.off_bar = off_bar; p.smem_bytes = total;
      p.n_runs = (plane_q + M - 1) / M;
      return nullptr;
    }
  }
  return "does not fit in shared memory";
}

template <int MTC, int NTC>
static int launch(const Params& p, cudaStream_t st) {
  auto kern = wgrad_ws_kernel<MTC, NTC>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) { set_error("wgrad_ws: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
    attr = true;
  }
  const long long gx = (long long)p.N * p.n_xseg * p.n_runs;
  HCU_CHECK_ARG(gx <= 0x7fffffffLL && p.n_mchunk <= 65535, "wgrad_ws: grid too large");
  kern<<<dim3((unsigned)gx, (unsigned)p.n_mchunk), kThreads, p.smem_bytes, st>>>(p);
  HCU_CHECK_LAUNCH("wgrad_ws");
  return 0;
}

}  // namespace wgs
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_wgrad_ws_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  wgs::Params p;
  int a, b;
  return wgs::configure(d, p, a, b) == nullptr ? 1 : 0;
}

extern "C" int hcu_conv_wgrad_ws_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                     const void* dy, float* wacc, void* stream) {
  HCU_CHECK_ARG(d && a && dy && wacc, "wgrad_ws: null pointer");
  HCU_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "wgrad_ws: a_scale/a_shift must come together");
  wgs::Params p;
  int mtc, ntc;
  const char* why = wgs::configure(d, p, mtc, ntc);
  if (why != nullptr) {
    set_error("wgrad_ws: unsupported descriptor (%s)", why);
    return HCU_ERR_UNSUPPORTED;
  }
  p.a = (const __half*)a; p.dy = (const __half*)dy; p.wacc = wacc; p.a_scale = a_scale; p.a_shift = a_shift;
  p.in_relu = d->in_relu;
  // x segmentation: about two waves of the 2 CTAs / SM, segments no shorter than 6 planes (each re-reads KX-1 planes)
  const long long base_items = (long long)p.N * p.n_runs * p.n_mchunk;
  const long long target = 4LL * num_sms();
  int nseg = (int)((target + base_items - 1) / base_items);
  nseg = std::max(1, std::min(nseg, (p.OX + 5) / 6));
  p.Lx = (p.OX + nseg - 1) / nseg;
  p.n_xseg = (p.OX + p.Lx - 1) / p.Lx;
  cudaStream_t st = (cudaStream_t)stream;
  if (ntc == 1) return wgs::launch<9, 1>(p, st);
  return wgs::launch<9, 2>(p, st);
}
