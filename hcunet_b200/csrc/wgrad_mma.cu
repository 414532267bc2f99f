// Weight gradient of the gather-convolution on tensor cores (fp16 operands, fp32 accumulate).
//
//   dW[tap][ci][co] = sum_{n, o} act(a[n, o + tap*dil - pad, ci]) * dy[n, o, co]
//
// Same "flat shift" view as conv_tc.cu: a CTA keeps the x-planes a run of M (+halo) flat (y,z) positions needs
// in shared memory as [channel-plane of 8][pixel][8 x fp16] and marches along x.  Here the GEMM's reduction
// dimension is the PIXEL index and the output (taps*Cin x Cout) is tiny, so the A operand would have to be an
// MN-major UMMA tile whose 8-row groups (one per tap) sit at non-uniform strides -- not expressible in a
// tcgen05 shared-memory descriptor, and the transposed problem wastes 7/8 of the M=64 minimum on Cout=8.
// Warp-level mma.sync.m16n8k16 with ldmatrix.trans takes arbitrary per-row addresses, which is exactly what
// the tap shifts need; these layers are HBM/L2-bound, not tensor-bound (DESIGN.md "Kernels and rooflines").
//
// One warp = one slice of the pixel run, all (MTC x NTC) 16x8 output tiles of the CTA's chunk in registers;
// the previous layer's BatchNorm+ReLU is applied while the activation plane is staged (never materialised);
// wrap-around / out-of-range positions are zeroed in the dy tile so they contribute nothing.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace hcu {
namespace wg {
using namespace ptx;

constexpr int kSmemLimit = 227 * 1024;

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
  int M, RUN, PS, SLOT, R, DPS;
  int E, MTOT, NTOT;
  int n_runs, Lx, n_xseg, n_mchunk, n_nchunk;
  int in_relu;
  int off_dy, smem_bytes;
  // dy addressing (elements): coarse strides + stride-phase offsets (HcuConvDesc.ophase on the dy side)
  long long dy_ns, dy_xs;
  int dy_ys, dy_zs, Poc, dps[3];
  long long dy_ph[3];
};



// Warps are arranged WM (tile groups) x WP (pixel slices): warp (wm, wp) owns m-tiles [mt0 + wm*MTC, +MTC) x NTC n-tiles
// and every WP-th 16-pixel block.  Few output tiles (Cout <= 16): WM = 1, the warps split the pixels; many tiles
// (Cout >= 64): WP = 1, the warps split the tiles so the staged planes are re-used by 8x more MMAs.
template <int MTC, int NTC, int WM, int WP>
__global__ void __launch_bounds__(32 * WM * WP) wgrad_mma_kernel(const Params p) {
  constexpr int kThreads = 32 * WM * WP;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % WM, wp = warp / WM;

  int item = blockIdx.x;
  const int run = item % p.n_runs; item /= p.n_runs;
  const int xs = item % p.n_xseg;
  const int n = item / p.n_xseg;
  const int mchunk = blockIdx.y % p.n_mchunk, nchunk = blockIdx.y / p.n_mchunk;
  const int mt0 = mchunk * (MTC * WM) + wm * MTC, nt0 = nchunk * NTC;
  const int x0 = xs * p.Lx;
  const int nout = min(p.Lx, p.OX - x0);
  const int span = (p.KX - 1) * p.dx + 1;
  const int q0 = run * p.M;
  const uint32_t a_base = smem_u32(smem), dy_base = smem_u32(smem + p.off_dy);

  // ---- staging helpers (all 128 threads) -------------------------------------------------------------
  const int plane = tid % p.P, pix0 = tid / p.P, pstep = kThreads / p.P;
  const int nchunk_a = (p.RUN - pix0 + pstep - 1) / pstep;
  float sc[8], sh[8];
  const bool xf = p.a_scale != nullptr;
  if (xf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = p.a_scale[plane * 8 + j];
      sh[j] = p.a_shift[plane * 8 + j];
    }
  }
  const int qf = q0 + pix0;
  const int yv0 = qf / p.Zv, zv0 = qf - yv0 * p.Zv;
  const int ystep = pstep / p.Zv, zstep = pstep - ystep * p.Zv;
  const __half* a_n = p.a + (size_t)n * p.IX * p.IY * p.IZ * p.Cp + plane * 8;

  auto stage_a = [&](int j) {  // virtual input plane j of this item -> ring slot j % R
    const int xm = x0 + j - p.px;
    const bool xok = xm >= 0 && xm < p.IX;
    const __half* a_x = a_n + (size_t)(xok ? xm : 0) * p.IY * p.IZ * p.Cp;
    unsigned char* dst = smem + (j % p.R) * p.SLOT + plane * p.PS + pix0 * 16;
    int yv = yv0, zv = zv0;
    for (int c = 0; c < nchunk_a; c += 4) {
      uint4 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ym = yv - p.py, zm = zv - p.pz;
        ok[u] = xok && (c + u < nchunk_a) && ym >= 0 && ym < p.IY && zm >= 0 && zm < p.IZ;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[u]) v[u] = ldg_nc16(a_x + ((size_t)ym * p.IZ + zm) * p.Cp);
        zv += zstep; yv += ystep;
        if (zv >= p.Zv) { zv -= p.Zv; ++yv; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c + u >= nchunk_a) break;
        if (xf && ok[u]) {
          __half2* h = reinterpret_cast<__half2*>(&v[u]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 f = __half22float2(h[k]);
            f.x = fmaf(f.x, sc[2 * k], sh[2 * k]);
            f.y = fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]);
            if (p.in_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            h[k] = __floats2half2_rn(f.x, f.y);
          }
        }
        *reinterpret_cast<uint4*>(dst + (size_t)(c + u) * pstep * 16) = v[u];
      }
    }
  };

  // dy tile: [Po][M][8]; positions that wrap around a row / fall outside the output are zero
  const int dplane = tid % p.Po, dpix0 = tid / p.Po, dstep = kThreads / p.Po;
  const int dqf = q0 + dpix0;
  const int dy0 = dqf / p.Zv, dz0 = dqf - dy0 * p.Zv;
  const int dystep = dstep / p.Zv, dzstep = dstep - dystep * p.Zv;
  const int nchunk_d = (p.M - dpix0 + dstep - 1) / dstep;
  long long dplane_off = (long long)dplane * 8;
  if (p.dps[0] * p.dps[1] * p.dps[2] > 1) {
    int phi = dplane / p.Poc;
    const int cg = dplane - phi * p.Poc;
    const int fz = phi % p.dps[2]; phi /= p.dps[2];
    const int fy = phi % p.dps[1], fx = phi / p.dps[1];
    dplane_off = (long long)cg * 8 + fx * p.dy_ph[0] + fy * p.dy_ph[1] + fz * p.dy_ph[2];
  }
  const __half* dy_n = p.dy + (size_t)n * p.dy_ns + dplane_off;
  auto stage_dy = [&](int i) {
    const __half* d_x = dy_n + (size_t)(x0 + i) * p.dy_xs;
    unsigned char* dst = smem + p.off_dy + dplane * p.DPS + dpix0 * 16;
    int oy = dy0, oz = dz0;
    for (int c = 0; c < nchunk_d; c += 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (c + u < nchunk_d && oy < p.OY && oz < p.OZ) v[u] = ldg_nc16(d_x + (size_t)oy * p.dy_ys + (size_t)oz * p.dy_zs);
        oz += dzstep; oy += dystep;
        if (oz >= p.Zv) { oz -= p.Zv; ++oy; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < nchunk_d) *reinterpret_cast<uint4*>(dst + (size_t)(c + u) * dstep * 16) = v[u];
    }
  };

  // ---- per-lane fragment addressing ---------------------------------------------------------------------
  // A (x4.trans): matrix mi = lane >> 3: slot (mi & 1) of the m-tile's pair, pixel half (mi >> 1); row = lane & 7
  // slot e = tap * P + plane, tap = (tx*KY + ty)*KZ + tz
  int a_off[MTC];  // byte offset inside a ring slot (excludes the slot base), -1: tile out of range
  int a_tx[MTC];
#pragma unroll
  for (int m = 0; m < MTC; ++m) {
    const int mt = mt0 + m;
    a_off[m] = -1; a_tx[m] = 0;
    if (mt < p.MTOT) {
      int e = 2 * mt + ((lane >> 3) & 1);
      if (e >= p.E) e = p.E - 1;  // odd tail: duplicate the last slot, its rows are discarded
      const int tap = e / p.P, pl = e - tap * p.P;
      const int tz = tap % p.KZ, tq = tap / p.KZ;
      const int ty = tq % p.KY, tx = tq / p.KY;
      a_tx[m] = tx;
      a_off[m] = pl * p.PS + ((ty * p.dy_ * p.Zv + tz * p.dz) + (lane >> 4) * 8 + (lane & 7)) * 16;
    }
  }
  const int b_off = (lane & 15) * 16;

  float acc[MTC][NTC][4];
#pragma unroll
  for (int m = 0; m < MTC; ++m)
#pragma unroll
    for (int nn = 0; nn < NTC; ++nn)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[m][nn][k] = 0.f;

  // ---- march along x ----------------------------------------------------------------------------------------
  for (int j = 0; j < span - 1; ++j) stage_a(j);
  const int nblk = p.M / 16;
  for (int i = 0; i < nout; ++i) {
    stage_a(i + span - 1);
    stage_dy(i);
    __syncthreads();
    uint32_t slot_addr[MTC];
#pragma unroll
    for (int m = 0; m < MTC; ++m) slot_addr[m] = a_base + (uint32_t)(((i + a_tx[m] * p.dx) % p.R) * p.SLOT + a_off[m]);
    for (int blk = wp; blk < nblk; blk += WP) {
      uint32_t bf[NTC][2];
#pragma unroll
      for (int nn = 0; nn < NTC; ++nn) {
        const int nt = min(nt0 + nn, p.NTOT - 1);
        ldsm_x2_t(dy_base + (uint32_t)(nt * p.DPS + blk * 256 + b_off), bf[nn][0], bf[nn][1]);
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
    __syncthreads();
  }

  // ---- WP > 1: reduce the pixel slices in shared memory; then one atomic per element into the global accumulator
  auto flush = [&](int row, int col, float v) {  // row within this warp's MTC*16 rows, col within NTC*8
    const int mt = mt0 + row / 16;
    const int slot = 2 * mt + ((row & 15) >> 3);
    const int co = nt0 * 8 + col;
    if (mt < p.MTOT && slot < p.E && co < p.cout && nt0 + col / 8 < p.NTOT) {
      const int tap = slot / p.P, pl = slot - tap * p.P;
      const int ci = pl * 8 + (row & 7);
      if (ci < p.cin) atomicAdd(&p.wacc[((size_t)tap * p.cin + ci) * p.cout + co], v);
    }
  };
  const int g = lane >> 2, t2 = (lane & 3) * 2;
  if (WP == 1) {
#pragma unroll
    for (int m = 0; m < MTC; ++m)
#pragma unroll
      for (int nn = 0; nn < NTC; ++nn) {
        flush(m * 16 + g, nn * 8 + t2, acc[m][nn][0]);
        flush(m * 16 + g, nn * 8 + t2 + 1, acc[m][nn][1]);
        flush(m * 16 + g + 8, nn * 8 + t2, acc[m][nn][2]);
        flush(m * 16 + g + 8, nn * 8 + t2 + 1, acc[m][nn][3]);
      }
  } else {
    constexpr int RW = NTC * 8, RH = MTC * 16;
    float* red = reinterpret_cast<float*>(smem) + wm * RH * RW;  // [WM][MTC*16][NTC*8]
    for (int e = tid; e < WM * RH * RW; e += kThreads) reinterpret_cast<float*>(smem)[e] = 0.f;
    __syncthreads();
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
    __syncthreads();
    // each warp group flushes its own rows (mt0 is per-wm)
    for (int e = wp * 32 + lane; e < RH * RW; e += WP * 32) flush(e / RW, e % RW, red[e]);
  }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static const char* configure(const HcuConvDesc* d, Params& p, int& mtc, int& ntc, int& wmg, int& wpg) {
  if (d->dtype_in != HCU_F16 || d->dtype_out != HCU_F16) return "fp16 only";
  if (d->groups != 1) return "groups != 1";
  if (d->in_cpitch % 8 != 0 || d->in_c_off != 0 || d->cin > d->in_cpitch) return "input channel layout";
  if (d->out_cpitch % 8 != 0 || d->out_c_off != 0 || d->cout > d->out_cpitch) return "dy channel layout";
  const int P = d->in_cpitch / 8, Po = d->out_cpitch / 8;
  if (P != 1 && P != 2 && P != 4 && P != 8 && P != 16) return "input channel pitch";
  if (Po != 1 && Po != 2 && Po != 4 && Po != 8 && Po != 16 && Po != 32) return "dy channel pitch";
  if (d->iphase) return "iphase";
  for (int i = 0; i < 3; ++i) p.dps[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff);
  const int nph = p.dps[0] * p.dps[1] * p.dps[2];
  if (nph > 1 && (d->cout != d->out_cpitch || d->cout % nph || (d->cout / nph) % 8)) return "ophase needs 8-channel aligned phases";
  for (int i = 0; i < 3; ++i) {
    if (d->istep[i] != 1 || d->ooff[i] != 0) return "strided";
    if (nph == 1 && (d->ostep[i] != 1 || d->out_tsize[i] != d->out_size[i])) return "strided";
    if (nph > 1 && (d->ostep[i] != p.dps[i] || d->out_tsize[i] != d->out_size[i] * p.dps[i])) return "phase geometry";
  }
  p.Poc = Po / nph;
  {
    const long long cr = d->out_cpitch / nph;
    const long long fz = (long long)d->out_size[2] * p.dps[2], fy = (long long)d->out_size[1] * p.dps[1],
                    fx = (long long)d->out_size[0] * p.dps[0];
    p.dy_ph[2] = cr; p.dy_ph[1] = fz * cr; p.dy_ph[0] = fy * fz * cr;
    p.dy_zs = (int)(p.dps[2] * cr);
    if (fz * cr * p.dps[1] * d->out_size[1] >= 0x7fffffffLL) return "x-plane too large";
    p.dy_ys = (int)(p.dps[1] * fz * cr);
    p.dy_xs = p.dps[0] * fy * fz * cr;
    p.dy_ns = fx * fy * fz * cr;
  }
  p.N = d->batch; p.IX = d->in_size[0]; p.IY = d->in_size[1]; p.IZ = d->in_size[2];
  p.Cp = d->in_cpitch; p.P = P; p.cin = d->cin;
  p.OX = d->out_size[0]; p.OY = d->out_size[1]; p.OZ = d->out_size[2];
  p.Cop = d->out_cpitch; p.Po = Po; p.cout = d->cout;
  p.KX = d->taps[0]; p.KY = d->taps[1]; p.KZ = d->taps[2];
  p.dx = d->dil[0]; p.dy_ = d->dil[1]; p.dz = d->dil[2];
  p.px = d->pad[0]; p.py = d->pad[1]; p.pz = d->pad[2];
  p.Yv = p.OY + (p.KY - 1) * p.dy_;
  p.Zv = p.OZ + (p.KZ - 1) * p.dz;
  p.R = (p.KX - 1) * p.dx + 1;
  if (p.R > 8) return "x extent";
  p.E = p.KX * p.KY * p.KZ * P;
  p.MTOT = (p.E + 1) / 2;
  p.NTOT = Po;
  // register tile: all n-tiles up to 8 per CTA, m-tiles so that MTC * NTC <= 24
  if (Po == 1) { mtc = 9; ntc = 1; wmg = 1; wpg = 4; }
  else if (Po == 2) { mtc = 9; ntc = 2; wmg = 1; wpg = 4; }
  else if (Po == 4) { mtc = 3; ntc = 4; wmg = 4; wpg = 2; }  // 48 accumulator registers: two CTAs per SM (6x4 needed 200 registers)
  else { mtc = 3; ntc = 8; wmg = 8; wpg = 1; }
  const int nthreads = 32 * wmg * wpg;
  p.n_mchunk = (p.MTOT + mtc * wmg - 1) / (mtc * wmg);
  p.n_nchunk = (p.NTOT + ntc - 1) / ntc;
  const int halo = (p.KY - 1) * p.dy_ * p.Zv + (p.KZ - 1) * p.dz;
  const int plane_q = p.Yv * p.Zv;
  (void)nthreads;
  const int m_cands[4] = {512, 256, 128, 64};
  for (int pass = 0; pass < 2; ++pass) {
    const int budget = pass == 0 ? 72 * 1024 : kSmemLimit;
    for (int mi = 0; mi < 4; ++mi) {
      const int M = m_cands[mi];
      if (M > 64 && M / 2 >= plane_q) continue;
      const int run = M + halo;
      int ps = run * 16;
      if (P > 1) { const int g = P >= 8 ? 16 : 128 / P; ps = round_up(ps, 2 * g) + g; }
      int dps = M * 16;
      if (Po > 1) { const int g = Po >= 8 ? 16 : 128 / Po; dps = round_up(dps, 2 * g) + g; }
      const int slot = ps * P;
      const int off_dy = round_up(p.R * slot, 128);
      const int total = std::max(off_dy + dps * Po, wpg > 1 ? wmg * mtc * 16 * ntc * 8 * 4 : 0) + 128;
      if (total > budget) continue;
      p.M = M; p.RUN = run; p.PS = ps; p.SLOT = slot; p.DPS = dps; p.off_dy = off_dy; p.smem_bytes = total;
      p.n_runs = (plane_q + M - 1) / M;
      return nullptr;
    }
  }
  return "does not fit in shared memory";
}

template <int MTC, int NTC, int WM, int WP>
static int launch(const Params& p, cudaStream_t st) {
  auto kern = wgrad_mma_kernel<MTC, NTC, WM, WP>;
  constexpr int kThreads = 32 * WM * WP;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) { set_error("wgrad_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
    attr = true;
  }
  const long long gx = (long long)p.N * p.n_xseg * p.n_runs;
  const long long gy = (long long)p.n_mchunk * p.n_nchunk;
  HCU_CHECK_ARG(gx <= 0x7fffffffLL && gy <= 65535, "wgrad_mma: grid too large");
  kern<<<dim3((unsigned)gx, (unsigned)gy), kThreads, p.smem_bytes, st>>>(p);
  HCU_CHECK_LAUNCH("wgrad_mma");
  return 0;
}

}  // namespace wg
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_wgrad_tc_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  wg::Params p;
  int a, b, c, e;
  return wg::configure(d, p, a, b, c, e) == nullptr ? 1 : 0;
}

static int wgrad_tc_impl(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                         const void* dy, float* wacc, void* stream, bool zero) {
  HCU_CHECK_ARG(d && a && dy && wacc, "wgrad_tc: null pointer");
  HCU_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "wgrad_tc: a_scale/a_shift must come together");
  wg::Params p;
  int mtc, ntc, wmg, wpg;
  const char* why = wg::configure(d, p, mtc, ntc, wmg, wpg);
  if (why != nullptr) {
    set_error("wgrad_tc: unsupported descriptor (%s)", why);
    return HCU_ERR_UNSUPPORTED;
  }
  p.a = (const __half*)a; p.dy = (const __half*)dy; p.wacc = wacc; p.a_scale = a_scale; p.a_shift = a_shift;
  p.in_relu = d->in_relu;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero) {
    cudaError_t e = cudaMemsetAsync(wacc, 0, sizeof(float) * (size_t)d->taps[0] * d->taps[1] * d->taps[2] * d->cin * d->cout, st);
    if (e != cudaSuccess) { set_error("wgrad_tc: memset: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
  }
  // x segmentation: ~3 CTAs per SM in flight, several waves
  const long long base_items = (long long)p.N * p.n_runs * p.n_mchunk * p.n_nchunk;
  const int target = 6 * num_sms();
  int nseg = (int)((target + base_items - 1) / base_items);
  nseg = std::max(1, std::min(nseg, (p.OX + 3) / 4));
  p.Lx = (p.OX + nseg - 1) / nseg;
  p.n_xseg = (p.OX + p.Lx - 1) / p.Lx;
  if (mtc == 9 && ntc == 1) return wg::launch<9, 1, 1, 4>(p, st);
  if (mtc == 9 && ntc == 2) return wg::launch<9, 2, 1, 4>(p, st);
  if (mtc == 3 && ntc == 4) return wg::launch<3, 4, 4, 2>(p, st);
  return wg::launch<3, 8, 8, 1>(p, st);
}

extern "C" int hcu_conv_wgrad_tc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                 const void* dy, float* wacc, void* stream) {
  return wgrad_tc_impl(d, a, a_scale, a_shift, dy, wacc, stream, true);
}

extern "C" int hcu_conv_wgrad_tc_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                     const void* dy, float* wacc, void* stream) {
  return wgrad_tc_impl(d, a, a_scale, a_shift, dy, wacc, stream, false);
}
