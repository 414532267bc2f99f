// Weight gradient of the gather-convolution for the CHANNEL-POOR levels (8 ... 32 channels) on the 5th-generation tensor
// cores, fed by TMA tensor-map loads: the "row-stacked" formulation.
//
//   dW[tap][ci][co] = sum_{n, x, y, z} act(a[n, x + tx, y + ty, z + tz, ci]) * dy[n, x, y, z, co]
//
// wgrad_tc5.cu puts the input channels on the M = 128 rows of a tcgen05.mma: with 8 channels 94 % of every MMA is padding, and
// an MMA costs (A bytes + B bytes) / 128 B per clock whatever it computes (profiles/r01_umma_rate.txt) -- which is why the
// 8/16-channel levels stayed on mma.sync (wgrad_ws.cu / wgrad_mma.cu, 100 - 160 us per layer for 15 - 40 us of HBM time).
// Here BOTH operand dimensions carry image rows:
//   M index = (input row r, 8 input channels)   16 rows x 8 channels = 128
//   N index = (dy row r', 8 output channels)    up to 14 rows x 8 channels = 112
//   K index = 16 consecutive z positions of a row
//   D[(r, ci), (r', co)] += sum_z a[y0 + r, z + tz, ci] * dy[y0 + r', z, co]
// so ONE MMA multiplies every input row of a tile with every dy row of the tile, and the block (r, r') of the accumulator
// is the contribution of rows (y0 + r, y0 + r') to the tap ty = r - r': three block diagonals of the 16 x 14 block matrix
// are the KY = 3 taps, the rest is never read.  19 % of the MACs are useful -- and the layer still needs 4x fewer tensor
// cycles than the mma.sync kernels need shared-memory cycles, because one 60-clock MMA covers 14 rows x 16 z = 224 pixels
// of three taps.  A filter tap along z moves the START ADDRESS of the A operand by whole pixels, a tap along x selects
// another x-plane of the ring: one accumulator (N columns of TMEM) per (tx, tz, input plane, dy plane).
//
// Operand layout = memory layout: a row of an x-plane is [z][8 channels] = the canonical no-swizzle MN-major core-matrix
// layout (8 channels in 16 bytes, the 8 z positions of a K group 16 bytes apart, LBO = 128 B to the next K group, SBO =
// row pitch to the next 8-row group of M / N).  A tile (rows x z x 8 channels of one x-plane) is ONE TMA tensor-map box;
// positions outside the tensor (z >= OZ of the last K group, rows >= OY of the last tile) are zero-filled by the TMA
// unit, which is all the masking the formulation needs: a zero dy element kills whatever its partner is.
//
// Interleaved mode (16 / 32 input channels): M groups must be ONE stride apart, so the channel planes of a row become groups
// [row][plane][z][8]: the natural [row][z][C] tile lands in a staging area (one box with a 16 P-byte inner run) and the transform
// warps re-lay it -- M = (16 / P rows) x P planes, N = RP rows x PG planes, one accumulator per (tx, tz).  The same mode takes the
// transposed convolutions' weight gradients: the stride phases folded into the dy channels (HcuConvDesc.ophase) are one tensor map
// per phase, low-side zero padding is a shifted box origin (transformed positions outside the tensor are zeroed by the re-layout).
//
// Roles (448 threads, one CTA per SM -- the accumulators take up to 512 TMEM columns):
//   warps 4-11  operand transforms on a landed tile (generic proxy, then fence.proxy.async): the previous layer's BatchNorm + ReLU
//               on the input (MODE 1), BatchNorm backward's apply pass on (g, y) -> dy (MODE 2, the first layer), the re-layout of
//               interleaved mode (MODE 3, + BatchNorm + ReLU); idle in MODE 0
//   warp 12     TMA producer: the input tile and the dy tile(s) of a step into one ring slot, one mbarrier complete_tx; ~128 KB
//               of loads in flight per SM
//   warp 13     TMEM allocation + MMA issue; the whole warp runs the loop (uniform datapath), x taps and the per-tap MMA list are
//               compile-time for the common shapes; one wait and one tcgen05.commit per step
//   warps 0-11  epilogue, ONCE per CTA (the accumulators run on across row tiles: block (r, r') means the same tap everywhere):
//               three warps per TMEM lane quadrant read the block diagonals, sum them over the rows, and the CTA adds its
//               [tap][ci][co] block to the global accumulator with vector reductions
// Work: the (image, row tile, x) steps of a layer are cut into equal contiguous ranges, one per CTA; a CTA marches along x inside a
// row tile (each input plane serves KX output planes) and re-fills its ring when it crosses into the next tile.  dy planes whose
// accumulators do not fit in TMEM beside each other go to CTA "kinds" on grid.y.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace hcu {
namespace wgr {
using namespace ptx;

constexpr int kThreads = 448;
constexpr int kXfWarps = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxPer = 64;  // MMAs per x tap and step

struct Params {
  float* wacc;  // fp32 [taps][cin][cout], zeroed by the caller
  const float* a_scale;
  const float* a_shift;
  // dy-side fusion (bnb): the dy operand is BatchNorm(+ReLU) backward applied to (g, y) while the tile is staged:
  // dy = c1 * (bn(y) > 0 ? g : 0) + c2 * y + c3, rounded to fp16 exactly like hcu_bn_bwd_apply
  const float* bn_scale;
  const float* bn_shift;
  const float* coef;       // [3][coef_c]
  long long* prof;         // HCU_ROWS_PROF: clock64 stamps of CTA (0, 0) (timing experiments only)
  int N, OX, OY, cin, cout;
  int P, PG;               // input channel planes (all in every CTA), dy channel planes per CTA (kinds = Po / PG on grid.y)
  int KX, KY, KZ, dx, dy_, dz;
  int RP, RA, ZC;          // dy rows / input rows per tile, K groups of 16 z positions
  int ZAP, ZGP;            // row pitch (pixels) of the staged input / dy tiles
  int NCOL, tmem_cols;
  int S;                   // ring depth: slot = [P input planes | pad | PG dy planes]
  int a_plane_bytes, g_plane_bytes, slot_bytes, off_g;   // off_g: the dy tiles inside a slot
  int a_box_bytes, g_box_bytes;
  int off_red, off_bar, smem_bytes;
  int n_ytiles, total_steps, steps_per_cta;
  int in_relu, zero_fill;
  int bnb, off_y, coef_c, OZ;
  // interleaved mode (il, P > 1): the channel planes of a row are M / N groups of ONE MMA.  The natural [row][z][C] tile lands in
  // a staging area (one TMA box with a 16 P-byte inner run) and the transform warps re-lay it as [row][plane][z][8 channels]
  // (+ BatchNorm + ReLU on the input side): M = 128 = (16 / P rows) x P planes x 8 channels, N = RP x PG x 8.
  int il, sbo_a, sbo_b, off_sa, off_sg, ZOA, ZOG, tx_cols;
  // stride phases of a transposed convolution on the dy side (interleaved mode): dy plane q = phase * PPH + plane of the phase; a
  // CTA's PG planes arrive as NB boxes of PB planes, one per phase touched (each phase is a strided sub-lattice: its own tensor map);
  // low-side zero padding of the input (box origin shifted; transformed positions outside the tensor are zeroed by the re-layout)
  int PPH, NB, PB, pad_x, pad_y, pad_z, IX, IY, IZ;
  int dbg;                 // HCU_ROWS_DEBUG (timing experiments only): 1 = no MMA issue
  int nper;                // MMAs per x tap and step: (tz, input plane, dy plane, K group), K group fastest
  // per MMA of an x tap, read through the constant bank with a uniform index (the issuing warp runs on the uniform datapath):
  uint32_t tab_a[kMaxPer];  // input operand: offset from the slot base, 16-byte units
  uint32_t tab_b[kMaxPer];  // dy operand: offset from the slot base, 16-byte units
  uint32_t tab_t[kMaxPer];  // accumulator column offset | (K group != 0) << 31
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
          "r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

// MN-major, no swizzle: LBO = next 8-position K group, SBO = next 8-element M / N group (here: the next image row)
__device__ __forceinline__ uint64_t desc_hi_mn(uint32_t sbo_bytes) {
  return (uint64_t)(((sbo_bytes >> 4) & 0x3FFF) | (1u << 14)) << 32;  // SBO | descriptor version 1 (bit 46)
}

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

// One contiguous piece of a CTA's step range inside one (image, row tile)
struct alignas(64) GMaps {
  CUtensorMap m[8];   // one per stride phase (m[0] alone without phases)
};

struct Segment {
  int n, yt, xb, nout;
};
__device__ __forceinline__ bool next_segment(const Params& p, int& f, int f1, Segment& s) {
  if (f >= f1) return false;
  const int tile = f / p.OX;
  s.xb = f - tile * p.OX;
  s.nout = min(p.OX - s.xb, f1 - f);
  s.n = tile / p.n_ytiles;
  s.yt = tile - s.n * p.n_ytiles;
  f += s.nout;
  return true;
}

// The MMA-issuing warp.  It is bound by its own instruction latency, not by the tensor pipe (one warp, dependent uniform-datapath
// instructions: the first version spent ~1000 clocks of bookkeeping per step beside 6 MMAs of 60): the common shapes get the x
// taps and the per-tap MMA list unrolled at compile time with the descriptor offsets held in registers.
struct MmaCtx {
  uint32_t tmem_base, ring, bar_in, bar_e, bar_accf, bar_acce;
  int f0, f1, span, lane;
  bool prof;
};
// descriptors as 32-bit halves: the high words (SBO, version) are loop invariants, the low words one add per MMA
__device__ __forceinline__ void umma_f16_split(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum)
      : "memory");
}

template <int KXT, int NPT>
__device__ __forceinline__ void mma_role(const Params& p, const MmaCtx& c) {
  const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.NCOL >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t lbo = (128u >> 4) << 16;
  const uint32_t ahi = (((uint32_t)p.sbo_a >> 4) & 0x3FFF) | (1u << 14), bhi = (((uint32_t)p.sbo_b >> 4) & 0x3FFF) | (1u << 14);
  const int S = p.S, span = c.span;
  const int KX = KXT > 0 ? KXT : p.KX, nper = NPT > 0 ? NPT : p.nper;
  const uint32_t ring16 = c.ring >> 4, slot16 = (uint32_t)p.slot_bytes >> 4, wrap16 = (uint32_t)S * slot16, dx16 = (uint32_t)p.dx * slot16;
  const uint32_t end16 = ring16 + wrap16;
  const uint32_t tx_cols = (uint32_t)p.tx_cols;
  const bool no_mma = (p.dbg & 1) != 0;
  constexpr int NP = NPT > 0 ? NPT : 1;
  uint32_t ta[NP], tb[NP], tt[NP], tp[NP];
  if (NPT > 0) {
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      ta[m] = p.tab_a[m] | lbo; tb[m] = p.tab_b[m] | lbo;
      tt[m] = c.tmem_base + (p.tab_t[m] & 0x7fffffffu); tp[m] = p.tab_t[m] >> 31;
    }
  }
  // ring position of the next slot to wait for (w_*) and of output plane i's first input plane (r_*): index + address
  int w_idx = 0, r_idx = 0, seg = 0, f = c.f0;
  uint32_t w_par = 0, w16 = ring16, r16 = ring16;
  long long t_wait = 0, t_issue = 0, t_mark = 0;
  Segment s;
  while (next_segment(p, f, c.f1, s)) {
    for (int k = 1; k < span; ++k) {  // the first output needs `span` input planes
      mbar_wait(c.bar_in + 8 * w_idx, w_par);
      w16 += slot16;
      if (++w_idx == S) { w_idx = 0; w_par ^= 1u; w16 = ring16; }
    }
    if (c.prof) t_mark = clock64();
    for (int i = 0; i < s.nout; ++i) {
      mbar_wait(c.bar_in + 8 * w_idx, w_par);  // slot i + span - 1: the newest input plane and dy plane i
      const uint32_t gslot = w16;
      w16 += slot16;
      if (++w_idx == S) { w_idx = 0; w_par ^= 1u; w16 = ring16; }
      tc_fence_after();
      if (c.prof) { const long long t = clock64(); t_wait += t - t_mark; t_mark = t; }
      const uint32_t acc0 = (uint32_t)(i | seg);  // the accumulators run on across row tiles: (r, r') means the same tap everywhere
      if (elect_one()) {
        if (!no_mma) {
          if (KXT > 0 && NPT > 0) {
            uint32_t blo[NP], acc[NP];
#pragma unroll
            for (int m = 0; m < NP; ++m) { blo[m] = gslot + tb[m]; acc[m] = acc0 | tp[m]; }
            uint32_t a16 = r16;
#pragma unroll
            for (int tx = 0; tx < KXT; ++tx) {
#pragma unroll
              for (int m = 0; m < NP; ++m) umma_f16_split(tt[m] + (uint32_t)tx * tx_cols, a16 + ta[m], ahi, blo[m], bhi, idesc, acc[m]);
              a16 += dx16;
              if (a16 >= end16) a16 -= wrap16;
            }
          } else {
            uint32_t tbase = c.tmem_base, a16 = r16;
            for (int tx = 0; tx < KX; ++tx) {
#pragma unroll 2
              for (int m = 0; m < nper; ++m) {
                const uint32_t t = p.tab_t[m];
                umma_f16_split(tbase + (t & 0x7fffffffu), a16 + (p.tab_a[m] | lbo), ahi, gslot + (p.tab_b[m] | lbo), bhi, idesc, acc0 | (t >> 31));
              }
              tbase += tx_cols;
              a16 += dx16;
              if (a16 >= end16) a16 -= wrap16;
            }
          }
        }
        umma_commit(c.bar_e + 8 * r_idx);  // slot i: its input plane is not needed by later outputs, its dy plane was used earlier
      }
      __syncwarp();
      r16 += slot16;
      if (++r_idx == S) { r_idx = 0; r16 = ring16; }
      if (c.prof) { const long long t = clock64(); t_issue += t - t_mark; t_mark = t; }
    }
    // the segment's last span - 1 input planes
    if (elect_one()) {
      for (int k = 1; k < span; ++k) {
        umma_commit(c.bar_e + 8 * r_idx);
        if (++r_idx == S) r_idx = 0;
      }
    }
    __syncwarp();
    r_idx = w_idx; r16 = w16;  // (the elected lane's r_idx; every lane continues from the next segment's first slot)
    if (c.prof && c.lane == 0 && seg == 0) { p.prof[3] = clock64(); p.prof[10] = s.nout; p.prof[2] = t_wait; p.prof[9] = t_issue; }
    ++seg;
  }
  if (elect_one()) umma_commit(c.bar_accf);  // every MMA of this CTA: the epilogue runs once
  __syncwarp();
  if (c.prof && c.lane == 0) p.prof[4] = clock64();
}

// MODE: 0 = operands as stored, 1 = BatchNorm + ReLU on the input tiles, 2 = BatchNorm backward applied on the dy tiles,
// 3 = interleaved channel planes (re-layout of both tiles, + BatchNorm + ReLU on the input side when a_scale is given)
// (compile-time: the transform warps keep their per-channel vectors in registers, one set per variant)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ GMaps tmGs, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;
  const bool prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (prof && tid == 0) p.prof[0] = clock64();
  // barrier map: full[S] (TMA bytes of the slot), ready[S] (input tile transformed), empty[S], acc_full, acc_empty
  const uint32_t bar_f = smem_u32(smem + p.off_bar), bar_r = bar_f + 8 * S, bar_e = bar_r + 8 * S, bar_accf = bar_e + 8 * S,
                 bar_acce = bar_accf + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * (3 * S + 2));
  const uint32_t ring = smem_u32(smem);
  constexpr bool il = MODE == 3, bnb = MODE == 2;
  const bool xf = MODE == 1 || (il && p.a_scale != nullptr);
  const int span = (p.KX - 1) * p.dx + 1;
  const int kind = blockIdx.y;
  const int f0 = blockIdx.x * p.steps_per_cta, f1 = min(p.total_steps, f0 + p.steps_per_cta);
  const int red_n = p.KY * (MODE == 3 ? p.P * p.PG : 1) * 64;  // one accumulator's (ty, ci, co) block

  if (p.zero_fill) {  // merged rows shorter than the z reach of the last tap: the bytes read past a tile must be finite
    uint4* q = reinterpret_cast<uint4*>(smem);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < p.off_red / 16; i += kThreads) q[i] = z;
  }
  if (warp == 13) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) { mbar_init(bar_f + 8 * i, 1); mbar_init(bar_r + 8 * i, kXfWarps); mbar_init(bar_e + 8 * i, 1); }
      mbar_init(bar_accf, 1);
      mbar_init(bar_acce, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  }
  fence_proxy_async();  // the zero fill above (generic proxy) before any TMA write to the same bytes
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (prof && tid == 0) p.prof[1] = clock64();

  // The single-thread roles (TMA issue, MMA issue) are bound by their own instruction latency: the WHOLE warp runs their
  // loops so that every index stays on the uniform datapath (no R2UR per operand), ring positions and parities are kept
  // incrementally, and only the issuing instructions sit behind elect_one().
  if (warp == 12) {
    // =========================================== TMA PRODUCER ========================================
    int idx = 0, f = f0;
    uint32_t par = 1, dst = ring;
    const uint32_t a_bytes = il ? (uint32_t)p.a_box_bytes : (uint32_t)(p.P * p.a_box_bytes);
    const uint32_t ag_bytes = a_bytes + (il ? (uint32_t)(p.NB * p.g_box_bytes) : (uint32_t)((bnb ? 2 : 1) * p.PG * p.g_box_bytes));
    const int c0g = 8 * kind * p.PG;
    Segment s;
    while (next_segment(p, f, f1, s)) {
      const int y0 = s.yt * p.RP;
      const int nplanes = s.nout + span - 1;
      for (int j = 0; j < nplanes; ++j) {  // slot j: input plane xb + j and (from j = span - 1 on) dy plane xb + j - (span - 1)
        mbar_wait(bar_e + 8 * idx, par);
        if (elect_one()) {
          const uint32_t bar = bar_f + 8 * idx;
          const bool with_g = j >= span - 1;
          mbar_expect_tx(bar, with_g ? ag_bytes : a_bytes);
          if (il) {  // one box per tile: all channel planes of the rows, [row][z][C]
            tma_load_5d(dst + (uint32_t)p.off_sa, &tmA, 0, -p.pad_z, y0 - p.pad_y, s.xb + j - p.pad_x, s.n, bar);
            if (with_g)
              for (int bx = 0; bx < p.NB; ++bx) {   // one box per stride phase touched
                const int q = kind * p.PG + bx * p.PB, ph = q / p.PPH;
                tma_load_5d(dst + (uint32_t)(p.off_sg + bx * p.g_box_bytes), &tmGs.m[ph], 8 * (q - ph * p.PPH), 0, y0, s.xb + j - (span - 1), s.n,
                            bar);
              }
          } else {
          for (int pl = 0; pl < p.P; ++pl) tma_load_5d(dst + (uint32_t)(pl * p.a_plane_bytes), &tmA, 8 * pl, 0, y0, s.xb + j, s.n, bar);
          if (with_g) {
            for (int q = 0; q < p.PG; ++q)
              tma_load_5d(dst + (uint32_t)(p.off_g + q * p.g_plane_bytes), &tmGs.m[0], c0g + 8 * q, 0, y0, s.xb + j - (span - 1), s.n, bar);
            if (bnb)
              for (int q = 0; q < p.PG; ++q)
                tma_load_5d(dst + (uint32_t)(p.off_y + q * p.g_plane_bytes), &tmY, c0g + 8 * q, 0, y0, s.xb + j - (span - 1), s.n, bar);
          }
          }
        }
        __syncwarp();
        dst += (uint32_t)p.slot_bytes;
        if (++idx == S) { idx = 0; par ^= 1u; dst = ring; }
      }
    }
  } else if (warp == 13) {
    // =========================================== MMA ISSUER ==========================================
    const MmaCtx c{tmem_base, ring, (xf || bnb || il) ? bar_r : bar_f, bar_e, bar_accf, bar_acce, f0, f1, span, lane, prof};
    if (p.KX == 3 && p.nper == 2) mma_role<3, 2>(p, c);
    else if (p.KX == 3 && p.nper == 4) mma_role<3, 4>(p, c);
    else if (p.KX == 3 && p.nper == 8) mma_role<3, 8>(p, c);
    else mma_role<0, 0>(p, c);
  } else if (warp >= 4) {
    // =========================================== OPERAND TRANSFORMS ===================================
    if (il) {
      // staging [row][z][plane] (chunk e = (row * Z + z) * planes + plane) -> operand [row][plane][z]: thread t owns the chunks
      // e = t, t + 256, ...: its plane is fixed (256 % planes == 0), (row, z) walk incrementally
      const int xt = tid - 128;
      const int P = p.P, PG = p.PG;
      const int pla = xt % P;
      const int na = p.RA * p.ZAP * P;
      const int sa = 256 / P;                                      // pixels advanced per 256 chunks
      const int a_r0 = (xt / P) / p.ZAP, a_z0 = (xt / P) - a_r0 * p.ZAP, a_rs = sa / p.ZAP, a_zs = sa - a_rs * p.ZAP;
      float sc[8], sh[8];
      if (xf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = p.a_scale[pla * 8 + j]; sh[j] = p.a_shift[pla * 8 + j]; }
      }
      const int relu = p.in_relu;
      int sl = 0, f = f0;
      uint32_t par = 0;
      Segment s;
      while (next_segment(p, f, f1, s)) {
        const int nplanes = s.nout + span - 1;
        for (int j = 0; j < nplanes; ++j) {
          mbar_wait(bar_f + 8 * sl, par);
          unsigned char* slot = smem + sl * p.slot_bytes;
          {
            const uint4* src = reinterpret_cast<const uint4*>(slot + p.off_sa);
            uint4* dst = reinterpret_cast<uint4*>(slot);
            int row = a_r0, z = a_z0;
            // transformed positions outside the tensor (zero padding) must stay zero: relu(shift) would leak into the taps
            const bool padded = xf && (p.pad_x | p.pad_y | p.pad_z) != 0;
            const int xt_in = s.xb + j - p.pad_x, y_in0 = s.yt * p.RP - p.pad_y;
            const bool x_in = xt_in >= 0 && xt_in < p.IX;
            for (int e = xt; e < na; e += 256) {
              uint4 v = src[e];
              if (xf) {
                v = bn_relu8(v, sc, sh, relu);
                if (padded) {
                  const int zi = z - p.pad_z, yi = y_in0 + row;
                  if (!x_in || zi < 0 || zi >= p.IZ || yi < 0 || yi >= p.IY) v = make_uint4(0u, 0u, 0u, 0u);
                }
              }
              dst[(row * P + pla) * p.ZOA + z] = v;
              z += a_zs; row += a_rs;
              if (z >= p.ZAP) { z -= p.ZAP; ++row; }
            }
          }
          if (j >= span - 1) {
            const uint4* src = reinterpret_cast<const uint4*>(slot + p.off_sg);
            uint4* dst = reinterpret_cast<uint4*>(slot + p.off_g);
            // staging: NB boxes of [row][z][PB planes]; chunk e of a box -> dy plane bx * PB + (e mod PB)
            const int PB = p.PB, nb1 = p.RP * p.ZGP * PB;
            const int plb = xt % PB, sb = 256 / PB;
            const int b_r0 = (xt / PB) / p.ZGP, b_z0 = (xt / PB) - b_r0 * p.ZGP, b_rs = sb / p.ZGP, b_zs = sb - b_rs * p.ZGP;
            for (int bx = 0; bx < p.NB; ++bx) {
              int row = b_r0, z = b_z0;
              for (int e = xt; e < nb1; e += 256) {
                dst[(row * PG + bx * PB + plb) * p.ZOG + z] = src[bx * nb1 + e];
                z += b_zs; row += b_rs;
                if (z >= p.ZGP) { z -= p.ZGP; ++row; }
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_r + 8 * sl);
          if (++sl == S) { sl = 0; par ^= 1u; }
        }
      }
    } else if (xf || bnb) {
      const int xt = tid - 128;
      // input side: thread -> (channel plane, chunk); dy side: thread -> (dy plane q, chunk)
      const int pln = xf ? xt % p.P : xt % p.PG, c0 = xf ? xt / p.P : xt / p.PG, cstep = (32 * kXfWarps) / (xf ? p.P : p.PG);
      const int nchunk = xf ? p.RA * p.ZAP : p.RP * p.ZGP;
      float sc[8], sh[8], c1[bnb ? 8 : 1], c2[bnb ? 8 : 1], c3[bnb ? 8 : 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (xf) { sc[j] = p.a_scale[pln * 8 + j]; sh[j] = p.a_shift[pln * 8 + j]; }
        if (bnb) {
          const int ch = (kind * p.PG + pln) * 8 + j;
          sc[j] = p.bn_scale[ch]; sh[j] = p.bn_shift[ch];
          c1[j] = p.coef[ch]; c2[j] = p.coef[p.coef_c + ch]; c3[j] = p.coef[2 * p.coef_c + ch];
        }
      }
      const int relu = p.in_relu;
      const int row0 = c0 / p.ZGP, z0 = c0 - row0 * p.ZGP, rstep = cstep / p.ZGP, zstep = cstep - rstep * p.ZGP;  // dy side walk
      int sl = 0, f = f0;
      uint32_t par = 0;
      Segment s;
      while (next_segment(p, f, f1, s)) {
        const int nplanes = s.nout + span - 1;
        const int rows_left = p.OY - s.yt * p.RP;  // dy rows of this tile inside the tensor
        for (int j = 0; j < nplanes; ++j) {
          mbar_wait(bar_f + 8 * sl, par);
          unsigned char* slot = smem + sl * p.slot_bytes;
          if (xf) {
            uint4* tile = reinterpret_cast<uint4*>(slot + pln * p.a_plane_bytes);
            for (int c = c0; c < nchunk; c += cstep) tile[c] = bn_relu8(tile[c], sc, sh, relu);
          }
          if (bnb && j >= span - 1) {
            uint4* gt = reinterpret_cast<uint4*>(slot + p.off_g + pln * p.g_plane_bytes);
            const uint4* yt = reinterpret_cast<const uint4*>(slot + p.off_y + pln * p.g_plane_bytes);
            int row = row0, z = z0;
            for (int c = c0; c < nchunk; c += cstep) {
              uint4 o = make_uint4(0u, 0u, 0u, 0u);  // positions outside the tensor stay zero (c3 must not leak into them)
              if (z < p.OZ && row < rows_left) {
                const uint4 gr = gt[c], yr = yt[c];
                const __half2* gh = reinterpret_cast<const __half2*>(&gr);
                const __half2* yh = reinterpret_cast<const __half2*>(&yr);
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 yv = __half22float2(yh[k]);
                  float2 gv = __half22float2(gh[k]);
                  if (fmaf(yv.x, sc[2 * k], sh[2 * k]) <= 0.f) gv.x = 0.f;
                  if (fmaf(yv.y, sc[2 * k + 1], sh[2 * k + 1]) <= 0.f) gv.y = 0.f;
                  oh[k] = __floats2half2_rn(fmaf(c1[bnb ? 2 * k : 0], gv.x, fmaf(c2[bnb ? 2 * k : 0], yv.x, c3[bnb ? 2 * k : 0])),
                                            fmaf(c1[bnb ? 2 * k + 1 : 0], gv.y, fmaf(c2[bnb ? 2 * k + 1 : 0], yv.y, c3[bnb ? 2 * k + 1 : 0])));
                }
              }
              gt[c] = o;
              z += zstep; row += rstep;
              if (z >= p.ZGP) { z -= p.ZGP; ++row; }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_r + 8 * sl);
          if (++sl == S) { sl = 0; par ^= 1u; }
        }
      }
    }
  }

  if (warp < 12 && f0 < f1) {
    // =========================================== EPILOGUE ============================================
    // Once per CTA, by the twelve warps whose roles are over (the four idle ones and the eight transform warps; a warp reads the
    // TMEM lanes 32 * (warp % 4) ...: three warps share a quadrant and split its rounds).
    // TMEM lane = 8 * g + ci with M group g = r * PL + plane (PL = P planes per row in interleaved mode, else 1: one plane per
    // MMA); quadrant w holds the groups 4w .. 4w + 3 and writes its sums over its rows into its own copy of the accumulator's
    // [ty][ci][co] block (plain stores: every entry once).  Accumulator columns of one (tx, tz[, plane, dy plane]) combination:
    // 8 * (r' * QL + q) + co, QL = PG dy planes per row in interleaved mode, else 1.
    const int qd = warp & 3, helper = warp >> 2;
    const int et = tid;                                  // 0 .. 383
    const int ci = lane & 7;
    const int PL = il ? p.P : 1, QL = il ? p.PG : 1;
    const int pl_lane = il ? ((4 * qd + (lane >> 3)) % PL) : 0;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float* red = reinterpret_cast<float*>(smem + p.off_red);
    float* mine = red + (size_t)qd * red_n;
    const int npl = il ? 1 : p.P, nq = il ? 1 : p.PG;
    const int cin_b = (il ? p.P : 1) * 8, cout_b = (il ? p.PG : 1) * 8;   // channel block of one accumulator (powers of two)
    const int lg_ci = 31 - __clz(cin_b), lg_co = 31 - __clz(cout_b);
    const int blk_n = p.KY * cin_b * cout_b;                              // floats of one accumulator's taps
    const bool vec4 = (p.cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.wacc) & 15) == 0);
    mbar_wait(bar_accf, 0u);
    tc_fence_after();
    if (prof && tid == 0) p.prof[5] = clock64();
    // Every CTA adds its accumulators to the same [taps][cin][cout] block: CTA b starts with accumulator b mod ncomb and walks its
    // flush from a different element, so that the CTAs' reductions spread over the addresses.
    const int ncomb = p.KX * p.KZ * npl * nq;
    const int NR = 4 / PL, QC = min(QL, 4 / NR);
    const int jr_me = (lane >> 3) / PL;
    for (int cidx = 0; cidx < ncomb; ++cidx) {
      const int comb = (cidx + (int)blockIdx.x) % ncomb;
      int rem = comb;
      const int q = rem % nq; rem /= nq;
      const int pl = rem % npl; rem /= npl;
      const int tz = rem % p.KZ, tx = rem / p.KZ;
      // A round = one ty and up to four 8-column loads issued back to back: NR = 4 / PL distinct rows per quadrant x QC dy planes
      // (NR * QC <= 4).  The column offset of a load is warp-uniform, every lane keeps the loads of its own row.
      int round = 0;
      for (int ty = 0; ty < p.KY; ++ty)
        for (int q0 = 0; q0 < QL; q0 += QC, ++round) {
          if (round % 3 != helper) continue;   // warp-uniform
          uint32_t u[4][8];
          bool any = false;
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            const int jr = sl / QC, qc = sl - jr * QC;
            const int r = (4 * qd) / PL + jr, rp = r - ty * p.dy_;
            if (jr < NR && r < p.RA && rp >= 0 && rp < p.RP) {
              tmem_ld8_nowait(tmem_base + lane_base + (uint32_t)(comb * p.NCOL + 8 * (rp * QL + q0 + qc)), u[sl]);
              any = true;
            }
          }
          if (any) tmem_wait_ld();  // warp-uniform
          const int r_me = (4 * qd) / PL + jr_me, rp_me = r_me - ty * p.dy_;
          const bool ok_me = r_me < p.RA && rp_me >= 0 && rp_me < p.RP;
          for (int qc = 0; qc < QC; ++qc) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              const int jr = sl / QC, qcs = sl - jr * QC;
              const int r = (4 * qd) / PL + jr, rp = r - ty * p.dy_;
              if (jr < NR && r < p.RA && rp >= 0 && rp < p.RP) {   // the load was issued (warp-uniform)
                tmem_pin8(u[sl]);
                if (ok_me && jr == jr_me && qcs == qc) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(u[sl][k]);
                }
              }
            }
            for (int off = 8 * PL; off < 32; off <<= 1) {  // the rows of this quadrant that share (plane, channel)
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
            }
            if (lane < 8 * PL) {
              float4* o = reinterpret_cast<float4*>(mine + ((size_t)(ty * cin_b + pl_lane * 8 + ci)) * cout_b + (q0 + qc) * 8);
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
        }
      // this accumulator's taps: sum of the four quadrants' copies -> global accumulator
      named_bar_sync(1, 384);
      const int ci0 = il ? 0 : pl * 8, co0 = kind * p.PG * 8 + (il ? 0 : q * 8);
      const int rot = (int)((blockIdx.x * 131u) % (unsigned)(blk_n / 4)) * 4;
      for (int e0 = et * 4; e0 < blk_n; e0 += 384 * 4) {
        const int e = e0 + rot < blk_n ? e0 + rot : e0 + rot - blk_n;
        const int col = e & (cout_b - 1), row = e >> lg_co;  // row = ty * cin_b + input channel of the block
        const int ty = row >> lg_ci, cc = ci0 + (row & (cin_b - 1)), co = co0 + col;
        const int tap = (tx * p.KY + ty) * p.KZ + tz;
        const float4 a0 = *reinterpret_cast<const float4*>(red + e), a1 = *reinterpret_cast<const float4*>(red + red_n + e),
                     a2 = *reinterpret_cast<const float4*>(red + 2 * red_n + e), a3 = *reinterpret_cast<const float4*>(red + 3 * red_n + e);
        const float w0 = (a0.x + a1.x) + (a2.x + a3.x), w1 = (a0.y + a1.y) + (a2.y + a3.y), w2 = (a0.z + a1.z) + (a2.z + a3.z),
                    w3 = (a0.w + a1.w) + (a2.w + a3.w);
        if (cc < p.cin) {
          float* o = p.wacc + ((size_t)tap * p.cin + cc) * p.cout + co;
          if (vec4 && co + 3 < p.cout) red_add_v4(o, w0, w1, w2, w3);
          else {
            if (co < p.cout) atomicAdd(o, w0);
            if (co + 1 < p.cout) atomicAdd(o + 1, w1);
            if (co + 2 < p.cout) atomicAdd(o + 2, w2);
            if (co + 3 < p.cout) atomicAdd(o + 3, w3);
          }
        }
      }
      named_bar_sync(1, 384);
    }
    tc_fence_before();
    if (prof && tid == 0) p.prof[7] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (prof && tid == 0) p.prof[8] = clock64();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Config {
  Params p;
  int kinds, gx;
  bool merged_a, merged_g;
  int za;  // z extent the taps reach in an input row
};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static const char* configure(const HcuConvDesc* d, Config& c, bool bnb = false) {
  Params& p = c.p;
  memset(&p, 0, sizeof(p));
  if (d->dtype_in != HCU_F16 || d->dtype_out != HCU_F16) return "fp16 only";
  if (d->groups != 1) return "groups != 1";
  if (d->iphase) return "input stride phases";
  int dps[3];
  for (int i = 0; i < 3; ++i) dps[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff);
  const int nph = dps[0] * dps[1] * dps[2];
  if (nph > 8) return "more than 8 stride phases";
  if (nph > 1 && (d->cout != d->out_cpitch || d->cout % nph || (d->cout / nph) % 8)) return "ophase needs 8-channel aligned phases";
  if (d->in_cpitch % 8 != 0 || d->in_c_off != 0 || d->cin > d->in_cpitch) return "input channel layout";
  if (d->out_cpitch % 8 != 0 || d->out_c_off != 0 || d->cout > d->out_cpitch) return "dy channel layout";
  const int P = d->in_cpitch / 8, Po = d->out_cpitch / 8;
  static const int maxp = env_int("HCU_ROWS_MAXP", 4);
  if (P != 1 && P != 2 && P != 4) return "input channel pitch above 32";
  if (Po != 1 && Po != 2 && Po != 4 && !(Po == 8 && P > 1)) return "dy channel pitch above 32 (64 with interleaved input planes)";
  if (P > maxp || Po > 2 * maxp) return "channel pitch above HCU_ROWS_MAXP";
  bool padded = false;
  for (int i = 0; i < 3; ++i) {
    if (d->istep[i] != 1 || d->ostep[i] != dps[i] || d->ooff[i] != 0 || d->out_tsize[i] != d->out_size[i] * dps[i]) return "strided";
    if (d->pad[i] < 0 || d->pad[i] > (d->taps[i] - 1) * d->dil[i]) return "padding beyond the taps' reach";
    padded = padded || d->pad[i] != 0;
    if (d->pad[i] == 0 && d->in_size[i] != d->out_size[i] + (d->taps[i] - 1) * d->dil[i]) return "not a valid convolution";
  }
  // padding and stride phases (the transposed convolutions' weight gradients) are handled by the interleaved mode only
  if ((padded || nph > 1) && (P == 1 || bnb)) return "padding / stride phases need interleaved input planes";
  p.N = d->batch; p.OX = d->out_size[0]; p.OY = d->out_size[1];
  const int OZ = d->out_size[2];
  p.cin = d->cin; p.cout = d->cout; p.P = P;
  p.KX = d->taps[0]; p.KY = d->taps[1]; p.KZ = d->taps[2];
  p.dx = d->dil[0]; p.dy_ = d->dil[1]; p.dz = d->dil[2];
  const int span = (p.KX - 1) * p.dx + 1, halo_y = (p.KY - 1) * p.dy_, halo_z = (p.KZ - 1) * p.dz;
  if (span > 6) return "x extent";
  if (halo_y > 12) return "y extent";
  if (OZ < 8) return "rows too short for a K group";  // K = 16 z positions per MMA: mostly padding
  p.ZC = (OZ + 15) / 16;
  const int ZG = 16 * p.ZC;
  c.za = ZG + halo_z;
  if (c.za > 256) return "rows too long for one box";
  // Cp == 8: (z, channel) is one contiguous run -> box rows of up to 256 elements instead of 16-byte pieces
  c.merged_g = d->out_cpitch == 8 && ZG <= 32;
  // (a merged input row of 32 positions may be shorter than the z reach of the last tap: what is read past it pairs with
  // zero-filled dy positions only while the whole input row fits, IZ <= 32)
  c.merged_a = d->in_cpitch == 8 && ZG <= 32 && d->in_size[2] <= 32 && halo_z <= 8;
  p.ZGP = ZG;
  p.ZAP = c.merged_a ? std::min(c.za, 32) : c.za;
  p.zero_fill = p.ZAP < c.za ? 1 : 0;
  // dy planes per CTA and rows per tile: all (tx, tz, input plane, dy plane) accumulators of a CTA live in TMEM
  static const int tmem_max = env_int("HCU_ROWS_TMEM", 512);
  static const int il_on = env_int("HCU_ROWS_IL", 1);
  const int oy_even = round_up(p.OY, 2);
  p.il = (P > 1 && il_on && !bnb) ? 1 : 0;
  if ((padded || nph > 1) && !p.il) return "padding / stride phases need interleaved input planes";
  p.PPH = Po / nph;   // dy planes per stride phase
  p.pad_x = d->pad[0]; p.pad_y = d->pad[1]; p.pad_z = d->pad[2];
  p.IX = d->in_size[0]; p.IY = d->in_size[1]; p.IZ = d->in_size[2];
  double best = 1e30;
  int best_pg = 0, best_rp = 0;
  if (p.il) {
    // interleaved: M = (16 / P rows) x P planes, N = RP rows x PG planes; one accumulator per (tx, tz)
    const int ra_max = 16 / P, ncomb = p.KX * p.KZ;
    if (ra_max - halo_y < 1) return "y extent of the taps above the rows of an interleaved tile";
    for (int pg = Po; pg >= 1; pg >>= 1)
      for (int rp = std::min(ra_max - halo_y, oy_even); rp >= 1; --rp) {
        if ((rp * pg) % 2 != 0 || ncomb * 8 * rp * pg > tmem_max || 8 * rp * pg > 256) continue;
        const int tiles = (p.OY + rp - 1) / rp;
        const double cost = (double)ncomb * (32 + 2 * rp * pg) / (16.0 * rp) * (Po / pg) * ((double)tiles * rp / p.OY);
        if (cost < best) { best = cost; best_pg = pg; best_rp = rp; }
      }
  } else {
    for (int pg = Po; pg >= 1; pg >>= 1) {
      const int ncomb = p.KX * p.KZ * P * pg;
      int rp = std::min(16 - halo_y, tmem_max / (8 * ncomb));
      rp = std::min(rp & ~1, oy_even);
      if (rp < 2) continue;
      const int tiles = (p.OY + rp - 1) / rp;
      const double cost = (double)ncomb * (32 + 2 * rp) / (16.0 * rp) * (Po / pg) * ((double)tiles * rp / p.OY);
      if (cost < best) { best = cost; best_pg = pg; best_rp = rp; }
    }
  }
  if (best_pg == 0) return "accumulators do not fit in TMEM";
  p.PG = best_pg; p.RP = best_rp; p.RA = best_rp + halo_y;
  c.kinds = Po / p.PG;
  p.NCOL = p.il ? 8 * p.RP * p.PG : 8 * p.RP;
  p.tx_cols = p.il ? p.KZ * p.NCOL : p.KZ * P * p.PG * p.NCOL;
  {
    int cols = p.KX * p.tx_cols, t = 32;
    while (t < cols) t <<= 1;
    p.tmem_cols = t;
  }
  const int red_bytes = 4 * p.KY * (p.il ? P * p.PG : 1) * 64 * 4;  // one accumulator's (ty, ci, co) block, one copy per epilogue warp
  int a_reach;  // bytes an M = 128 tile may read from a slot's base (16 row groups whatever RA is)
  if (p.il) {
    c.merged_a = c.merged_g = false;
    p.ZAP = c.za; p.zero_fill = 0;
    // operand row pitch = 1 (mod 8) positions: the P groups of a row land in different banks for the re-layout's stores
    p.ZOA = round_up(p.ZAP, 8) + 1; p.ZOG = round_up(p.ZGP, 8) + 1;
    p.sbo_a = p.ZOA * 16; p.sbo_b = p.ZOG * 16;
    p.a_box_bytes = p.RA * p.ZAP * 16 * P;       // one box: all planes
    p.PB = std::min(p.PG, p.PPH); p.NB = p.PG / p.PB;   // dy boxes: one per stride phase touched by this CTA's planes
    p.g_box_bytes = p.RP * p.ZGP * 16 * p.PB;
    if (p.g_box_bytes % 128 != 0) return "dy box not 128-byte aligned";   // (ZGP is a multiple of 16: always true)
    const int opa = round_up(16 * p.ZOA * 16, 128);                 // operand: 16 groups
    const int opg = round_up(p.RP * p.PG * p.ZOG * 16, 128);
    p.off_g = opa;
    p.off_sa = p.off_g + opg;
    p.off_sg = p.off_sa + round_up(p.a_box_bytes, 128);
    p.slot_bytes = p.off_sg + round_up(p.NB * p.g_box_bytes, 128);
    p.off_y = 0; p.bnb = 0;
    p.a_plane_bytes = 0; p.g_plane_bytes = 0;
    p.nper = p.KZ * p.ZC;
    if (p.nper > kMaxPer) return "too many MMAs per step";
    for (int m = 0; m < p.nper; ++m) {
      const int zc = m % p.ZC, tz = m / p.ZC;
      p.tab_a[m] = (uint32_t)(tz * p.dz * 16 + zc * 256) >> 4;
      p.tab_b[m] = (uint32_t)(p.off_g + zc * 256) >> 4;
      p.tab_t[m] = (uint32_t)(tz * p.NCOL) | (zc ? 0x80000000u : 0u);
    }
    a_reach = opa;
  } else {
    p.sbo_a = p.ZAP * 16; p.sbo_b = p.ZGP * 16;
    p.a_box_bytes = p.RA * p.ZAP * 16;
    p.g_box_bytes = p.RP * p.ZGP * 16;
    p.a_plane_bytes = round_up(p.a_box_bytes, 128);
    p.g_plane_bytes = round_up(p.g_box_bytes, 128);
    p.off_g = P * p.a_plane_bytes + 128;  // + what the last tap reads past a merged row
    p.off_y = p.off_g + p.PG * p.g_plane_bytes;
    p.slot_bytes = p.off_y + (bnb ? p.PG * p.g_plane_bytes : 0);
    p.bnb = bnb ? 1 : 0;
    p.nper = p.KZ * P * p.PG * p.ZC;
    if (p.nper > kMaxPer) return "too many MMAs per step";
    for (int m = 0; m < p.nper; ++m) {
      int r = m;
      const int zc = r % p.ZC; r /= p.ZC;
      const int q = r % p.PG; r /= p.PG;
      const int pl = r % P;
      const int tz = r / P;
      p.tab_a[m] = (uint32_t)(pl * p.a_plane_bytes + tz * p.dz * 16 + zc * 256) >> 4;
      p.tab_b[m] = (uint32_t)(p.off_g + q * p.g_plane_bytes + zc * 256) >> 4;
      p.tab_t[m] = (uint32_t)(((tz * P + pl) * p.PG + q) * p.NCOL) | (zc ? 0x80000000u : 0u);
    }
    a_reach = (P - 1) * p.a_plane_bytes + 16 * p.ZAP * 16 + 512;
  }
  p.OZ = OZ;
  p.coef_c = d->out_cpitch;
  // look-ahead: ~128 KB of loads in flight per SM (Little's law at ~2 us of loaded-HBM latency: with 64 KB the HBM-bound first-level
  // layers ran at 17 B per clock and SM; d0.conv2 73.7 -> 65.6 us, the fused first layer 150 -> 131 us with 8 slots ahead)
  static const int la_env = env_int("HCU_ROWS_LA", 0);
  static const int il_kb = env_int("HCU_ROWS_IL_KB", 128);   // in-flight target of the interleaved (16 / 32-channel) layers
  const int fly = (p.il ? il_kb : 128) * 1024;
  int la = la_env > 0 ? la_env : std::max(2, std::min(10, (fly + p.slot_bytes - 1) / p.slot_bytes));
  for (;; --la) {
    if (la < 1) return "does not fit in shared memory";
    p.S = span + la;
    // an M = 128 tile reads 16 row groups from a plane's base whatever RA is: keep those reads inside the allocation
    const int a_end = (p.S - 1) * p.slot_bytes + a_reach;
    p.off_red = round_up(p.S * p.slot_bytes, 128);
    p.off_bar = round_up(p.off_red + red_bytes, 128);
    p.smem_bytes = std::max(p.off_bar + 8 * (3 * p.S + 2) + 16, a_end) + 128;
    if (p.smem_bytes <= 200 * 1024) break;
  }
  p.n_ytiles = (p.OY + p.RP - 1) / p.RP;
  const long long total = (long long)p.N * p.n_ytiles * p.OX;
  if (total >= 0x7fffffffLL) return "too many steps";
  p.total_steps = (int)total;
  // equal contiguous step ranges, one CTA per SM (and kind); no range shorter than 16 planes (pipeline fill + flush)
  int gx = std::max(1, num_sms() / c.kinds);
  static const int min_steps = std::max(1, env_int("HCU_ROWS_MINSTEPS", 16));
  gx = (int)std::max(1LL, std::min<long long>(gx, total / min_steps));
  {
    // (fewer, longer CTAs to save end-of-kernel reductions were measured and lose: the epilogue is bound by its own instructions
    // per CTA, not by the L2's reduction rate -- 32 -> 64 layer 77 us with 143 CTAs, 106 us with 50)
    static const int gx_env = env_int("HCU_ROWS_GX", 0);
    if (gx_env > 0) gx = std::min(gx, gx_env);
  }
  p.steps_per_cta = (int)((total + gx - 1) / gx);
  c.gx = (int)((total + p.steps_per_cta - 1) / p.steps_per_cta);
  return nullptr;
}

// rank-5 map of a channels-last fp16 tensor [N][X][Y][Z][cpitch]; box = rows x boxz z positions x 8 channels of one x-plane
static const char* encode_map(CUtensorMap* tm, const void* base, int cpitch, int Z, int Y, int X, int N, bool merged, int boxz,
                              int rows, int boxc = 8) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  const cuuint64_t px = (cuuint64_t)cpitch * 2;
  if (merged) {
    gdim[0] = (cuuint64_t)Z * 8; gdim[1] = 1;
    gstr[0] = (cuuint64_t)Z * px;
    box[0] = (cuuint32_t)boxz * 8; box[1] = 1;
  } else {
    gdim[0] = (cuuint64_t)cpitch; gdim[1] = (cuuint64_t)Z;
    gstr[0] = px;
    box[0] = (cuuint32_t)boxc; box[1] = (cuuint32_t)boxz;
  }
  gdim[2] = (cuuint64_t)Y; gdim[3] = (cuuint64_t)X; gdim[4] = (cuuint64_t)N;
  gstr[1] = (cuuint64_t)Z * px; gstr[2] = gstr[1] * Y; gstr[3] = gstr[2] * X;
  box[2] = (cuuint32_t)rows; box[3] = 1; box[4] = 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return "tensor not 16-byte aligned";
  // the driver entry point is looked up through the runtime: the library keeps no link-time dependency on libcuda.so
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return "cuTensorMapEncodeTiled not available from this driver";
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return msg;
  }
  return nullptr;
}

// dy of a transposed convolution: stride phase (fx, fy, fz) is the sub-lattice o * s + f of the full-resolution tensor
// [N][TX][TY][TZ][cr]; one map per phase, coordinates in coarse positions
static const char* encode_phase_map(CUtensorMap* tm, const void* base, int cr, const int* tsize, const int* osize, const int* dps,
                                    const int* f, int N, int boxc, int boxz, int rows) {
  const long long ez = cr, ey = ez * tsize[2], ex = ey * tsize[1], en = ex * tsize[0];   // element strides of the fine tensor
  const char* b = reinterpret_cast<const char*>(base) + 2 * (f[0] * ex + f[1] * ey + f[2] * ez);
  cuuint64_t gdim[5] = {(cuuint64_t)cr, (cuuint64_t)osize[2], (cuuint64_t)osize[1], (cuuint64_t)osize[0], (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)(2 * ez * dps[2]), (cuuint64_t)(2 * ey * dps[1]), (cuuint64_t)(2 * ex * dps[0]), (cuuint64_t)(2 * en)};
  cuuint32_t box[5] = {(cuuint32_t)boxc, (cuuint32_t)boxz, (cuuint32_t)rows, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(b) & 15) != 0) return "dy phase not 16-byte aligned";
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return "cuTensorMapEncodeTiled not available from this driver";
  }
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<char*>(b), gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed (dy phase)";
  return nullptr;
}

}  // namespace wgr
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_wgrad_rows_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  wgr::Config c;
  const char* why = wgr::configure(d, c);
  if (why != nullptr && (wgr::env_int("HCU_TC_DEBUG", 0) & 8)) fprintf(stderr, "wgrad_rows: not taken (%s)\n", why);
  return why == nullptr ? 1 : 0;
}

extern "C" int hcu_conv_wgrad_rows_bnb_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  wgr::Config c;
  return wgr::configure(d, c, true) == nullptr ? 1 : 0;
}

static int wgrad_rows_launch(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                             const void* y, const float* bn_scale, const float* bn_shift, const float* coef, float* wacc, void* stream) {
  HCU_CHECK_ARG(d && a && dy && wacc, "wgrad_rows: null pointer");
  HCU_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr), "wgrad_rows: a_scale/a_shift must come together");
  const bool bnb = y != nullptr;
  HCU_CHECK_ARG(!bnb || (bn_scale && bn_shift && coef), "wgrad_rows: the fused BatchNorm backward needs scale, shift and coef");
  HCU_CHECK_ARG(!bnb || a_scale == nullptr, "wgrad_rows: the fused BatchNorm backward takes an untransformed input (the first layer)");
  wgr::Config c;
  const char* why = wgr::configure(d, c, bnb);
  if (why != nullptr) {
    set_error("wgrad_rows: unsupported descriptor (%s)", why);
    return HCU_ERR_UNSUPPORTED;
  }
  wgr::Params& p = c.p;
  p.wacc = wacc; p.a_scale = a_scale; p.a_shift = a_shift; p.in_relu = d->in_relu;
  p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.coef = coef;
  static const int dbg_env = wgr::env_int("HCU_ROWS_DEBUG", 0);
  p.dbg = dbg_env;
  static const int prof_env = wgr::env_int("HCU_ROWS_PROF", 0);
  static long long* prof_buf = nullptr;
  if (prof_env && prof_buf == nullptr) cudaMalloc(&prof_buf, 16 * sizeof(long long));
  p.prof = prof_env ? prof_buf : nullptr;
  CUtensorMap tmA, tmY;
  wgr::GMaps tmGs;
  why = wgr::encode_map(&tmA, a, d->in_cpitch, d->in_size[2], d->in_size[1], d->in_size[0], d->batch, c.merged_a, p.ZAP, p.RA,
                        p.il ? 8 * p.P : 8);
  {
    int dps[3], nph = 1;
    for (int i = 0; i < 3; ++i) { dps[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff); nph *= dps[i]; }
    if (why == nullptr && nph == 1) {
      why = wgr::encode_map(&tmGs.m[0], dy, d->out_cpitch, d->out_size[2], d->out_size[1], d->out_size[0], d->batch, c.merged_g, p.ZGP, p.RP,
                            p.il ? 8 * p.PB : 8);
      for (int i = 1; i < 8; ++i) tmGs.m[i] = tmGs.m[0];
    } else if (why == nullptr) {
      for (int ph = 0; ph < 8 && why == nullptr; ++ph) {
        int r = ph % nph;
        const int fz = r % dps[2]; r /= dps[2];
        const int f[3] = {r / dps[1], r % dps[1], fz};
        why = wgr::encode_phase_map(&tmGs.m[ph], dy, d->out_cpitch / nph, d->out_tsize, d->out_size, dps, f, d->batch, 8 * p.PB, p.ZGP, p.RP);
      }
    }
  }
  if (why == nullptr)
    why = wgr::encode_map(&tmY, bnb ? y : dy, d->out_cpitch, d->out_size[2], d->out_size[1], d->out_size[0], d->batch, c.merged_g && !d->ophase,
                          p.ZGP, p.RP);
  if (why != nullptr) {
    set_error("wgrad_rows: %s", why);
    return HCU_ERR_CUDA;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgr::wgrad_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgr::kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgr::wgrad_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgr::kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgr::wgrad_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgr::kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgr::wgrad_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgr::kSmemLimit);
    if (e != cudaSuccess) { set_error("wgrad_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HCU_ERR_CUDA; }
    attr = true;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("HCU_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (dbg & 8)
      fprintf(stderr, "wgrad_rows: P %d PG %d kinds %d RP %d RA %d ZC %d ZAP %d ZGP %d N %d tmem %d S %d smem %d grid %d x %d steps/cta %d merged %d/%d bnb %d il %d\n",
              p.P, p.PG, c.kinds, p.RP, p.RA, p.ZC, p.ZAP, p.ZGP, p.NCOL, p.tmem_cols, p.S, p.smem_bytes, c.gx, c.kinds,
              p.steps_per_cta, (int)c.merged_a, (int)c.merged_g, p.bnb, p.il);
  }
  const dim3 grid((unsigned)c.gx, (unsigned)c.kinds);
  if (p.il) wgr::wgrad_rows_kernel<3><<<grid, wgr::kThreads, p.smem_bytes, (cudaStream_t)stream>>>(tmA, tmGs, tmY, p);
  else if (bnb) wgr::wgrad_rows_kernel<2><<<grid, wgr::kThreads, p.smem_bytes, (cudaStream_t)stream>>>(tmA, tmGs, tmY, p);
  else if (a_scale != nullptr) wgr::wgrad_rows_kernel<1><<<grid, wgr::kThreads, p.smem_bytes, (cudaStream_t)stream>>>(tmA, tmGs, tmY, p);
  else wgr::wgrad_rows_kernel<0><<<grid, wgr::kThreads, p.smem_bytes, (cudaStream_t)stream>>>(tmA, tmGs, tmY, p);
  HCU_CHECK_LAUNCH("wgrad_rows");
  if (prof_env) {  // timing experiments: synchronous read-back of the stamps
    long long h[16];
    cudaDeviceSynchronize();
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "wgrad_rows prof (clk from entry): setup %lld | mma seg0: waiting %lld issuing %lld, done at %lld (%lld steps) role-end %lld | epi accfull %lld tmem-read %lld flushed %lld | end %lld\n",
            h[1] - h[0], h[2], h[9], h[3] - h[0], h[10], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0]);
  }
  return 0;
}

extern "C" int hcu_conv_wgrad_rows_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                       const void* dy, float* wacc, void* stream) {
  return wgrad_rows_launch(d, a, a_scale, a_shift, dy, nullptr, nullptr, nullptr, nullptr, wacc, stream);
}

extern "C" int hcu_conv_wgrad_rows_bnb_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                                           const void* g, const void* y, const float* bn_scale, const float* bn_shift,
                                           const float* coef, float* wacc, void* stream) {
  HCU_CHECK_ARG(y != nullptr, "wgrad_rows_bnb: null pointer");
  return wgrad_rows_launch(d, a, a_scale, a_shift, g, y, bn_scale, bn_shift, coef, wacc, stream);
}
