// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Formulation ("flat shift"): with channels-last fp16 activations [n][x][y][z][C], one x-plane is a flat
// array of Yv*Zv pixels of C channels.  For a stride-1 convolution the output at flat position q of plane
// ox is   out[ox][q] = sum_{tx,ty,tz} in[ox + tx*dx][q + ty*dy*Zv + tz*dz] . W[tx,ty,tz]
// i.e. every filter tap is the SAME [pixels x C] matrix shifted by a constant number of pixels.  Positions
// whose (y, z) wrap around the row / plane end are computed and discarded (<= 5 % waste on the big layers).
// So the A operand of the GEMM is never expanded (no im2col, not even in shared memory): a CTA keeps a ring
// of R input x-planes for a run of M + halo flat positions in shared memory, laid out as
// [channel-plane of 8][pixel][8 x fp16] = the canonical no-swizzle K-major UMMA core-matrix layout, and the
// MMA issuer just points shared-memory descriptors at (pixel offset of the tap, channel plane).  Each input
// element is read from L2/HBM ~(1 + halo/M)(1 + 2/Lx) times instead of prod(kernel) times.
//
// The CTA marches along x: producer warps load plane x+KX-1 (applying the previous layer's BatchNorm scale /
// shift + ReLU on the fly, so the normalised activation never exists in HBM) while the single MMA thread
// issues the taps of plane x into one of two TMEM accumulator buffers and the epilogue warps drain the other
// (bias, per-channel sum / sum-of-squares for train-mode BN, optional affine + ReLU, fp16/fp32 store).
//
// Roles (288 threads): warps 0-3 epilogue (thread = accumulator row / TMEM lane), warps 4-7 producers,
// warp 8 = TMEM allocation + barrier init + weight bulk copy (cp.async.bulk) + MMA issue (lane 0).
//
// Two kernels share this file, the packed weight layout and the C entry points (configure() picks per descriptor,
// hcu_conv_tc_describe() says which): conv_tc_kernel below (the weight slice of its column chunk resident in shared
// memory, x-march ring of planes: every input element fetched ~once -- the 8..64-channel levels) and conv_ks_kernel
// (K-streamed: activations AND weights streamed through rings, accumulators resident in TMEM -- from 64 input channels up,
// when the weights no longer fit beside the planes).  Both are launched with programmatic dependent launch: their prologue
// (barrier init, TMEM allocation, index tables) overlaps the predecessor's tail.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_tc_kernel.cuh"
#include "ptx.cuh"

namespace hcu {

namespace tc {
using namespace ptx;

// ===================================================================================================
// K-streamed variant for the CHANNEL-RICH levels (>= 64 input channels: the classic 2D U-Net's 64..1024-channel
// layers, the bottom of the 3D model).  conv_tc_kernel keeps ALL weights of its column chunk in shared memory beside
// the ring of x-planes; with >= 256 input channels neither the weights nor one x-plane of all channels fit.  Here the
// GEMM's K dimension is streamed instead: a CTA owns a tile of M = MB * 128 flat output positions of ONE output
// x-plane (the planes of all images of the batch are stacked into one flat index: rows = N * Yv, so tiny planes still
// fill the M = 128 rows of an MMA) and Nc output channels, with the fp32 accumulators resident in TMEM for the whole
// K loop.  K is walked as (tx, channel chunk of PC 8-channel planes, (ty, tz)):
//   A stage  = the flat run [q0, q0 + M + halo) of input plane ox + tx*dx, PC channel planes, staged ONCE per (tx, chunk)
//              by the producer warps (cp.async with zero fill, previous layer's BatchNorm + ReLU applied in place) in
//              the same no-swizzle K-major core-matrix layout [plane of 8 ch][pixel][8]: the KY*KZ taps are descriptor
//              start-address shifts of that one buffer ("flat shift", as in conv_tc_kernel);
//   B tile   = weights of (tx, ty, tz, chunk): PC * Nc * 16 contiguous bytes of the SAME packed layout
//              [nsplit][tx][K8 slab = tap * P + plane][Nc][8] conv_tc_kernel uses, fetched by one cp.async.bulk each
//              into a ring by a dedicated loader warp.
// Per (A stage, tap): PC / 2 K16 steps x MB tcgen05.mma.  The epilogue runs once per CTA.
// Roles (320 threads): warps 0-3 epilogue, 4-7 A producers, warp 8 TMEM + MMA issue, warp 9 B-tile loader.
// ===================================================================================================
constexpr int kKsThreads = 320;

__global__ void __launch_bounds__(kKsThreads, 1) conv_ks_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RA = p.RA, RB = p.RB, MB = p.MB, Nc = p.Nc, PC = p.PC;
  const int KYZ = p.KY * p.KZ;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // barrier map: full_a[RA], empty_a[RA], full_b[RB], empty_b[RB], tfull
  const uint32_t bar_fa = smem_u32(bars), bar_ea = bar_fa + 8 * RA, bar_fb = bar_ea + 8 * RA, bar_eb = bar_fb + 8 * RB,
                 bar_t = bar_eb + 8 * RB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * (2 * RA + 2 * RB + 1));
  float* sstat = reinterpret_cast<float*>(smem + p.off_stat);  // [4 warps][2][Nc]
  float* sbias = sstat + 8 * Nc;                               // [3][Nc]
  int* soff = reinterpret_cast<int*>(smem + p.off_tab);        // [RUN]: element offset of a staged pixel, -1 = zero fill
  const uint32_t a_base = smem_u32(smem + p.off_a), b_base = smem_u32(smem + p.off_w);

  // ---- work item: (output x-plane, run of M flat positions over the stacked images, column chunk) ----
  int item = blockIdx.x;
  const int ns = item % p.nsplit; item /= p.nsplit;
  const int run = item % p.n_runs;
  const int ox = item / p.n_runs;
  const int q0 = run * p.M;

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < RA; ++i) { mbar_init(bar_fa + 8 * i, 4); mbar_init(bar_ea + 8 * i, 1); }
      for (int i = 0; i < RB; ++i) { mbar_init(bar_fb + 8 * i, 1); mbar_init(bar_eb + 8 * i, 1); }
      mbar_init(bar_t, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  }
  for (int i = threadIdx.x; i < 8 * Nc; i += kKsThreads) sstat[i] = 0.f;
  for (int i = threadIdx.x; i < p.RUN; i += kKsThreads) {
    const int v = q0 + i;
    const int r = v / p.Zv, vz = v - r * p.Zv;
    const int nn = r / p.Yv, vy = r - nn * p.Yv;
    const int iy = vy - p.py, iz = vz - p.pz;
    const bool ok = nn < p.N && iy >= 0 && iy < p.IY && iz >= 0 && iz < p.IZ;
    soff[i] = ok ? (int)(nn * p.in_ns) + iy * p.in_ys + iz * p.in_zs : -1;
  }
  // ---- everything above touched no global memory: it overlapped the previous kernel's tail (PDL) ----
  pdl_wait();
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < Nc; i += kKsThreads) {
    const int ch = ns * Nc + i;
    const bool in = ch < p.cout;
    sbias[i] = (in && p.bias != nullptr) ? p.bias[ch % p.cpp] : 0.f;
    sbias[Nc + i] = (in && p.out_scale != nullptr) ? p.out_scale[ch] : 1.f;
    sbias[2 * Nc + i] = (in && p.out_shift != nullptr) ? p.out_shift[ch] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // =========================================== A PRODUCERS =========================================
    const int ptid = threadIdx.x - 128;
    const int pl = ptid % PC, pix0 = ptid / PC, pstep = 128 / PC;
    const bool xf = p.in_scale != nullptr;
    const int relu = p.in_relu;
    const uint32_t sstep = (uint32_t)pstep * 16u;
    const int nstage = p.KX * p.NCH;
    const bool phased_in = p.ips[0] * p.ips[1] * p.ips[2] > 1;
    int slot_i = 0, slot_f = 0;
    uint32_t par = 1;
    auto finish = [&](int st) {  // stage st has landed: transform in place, publish
      const int tx = st / p.NCH, c = st - tx * p.NCH;
      const int xm = ox + tx * p.dx - p.px;
      if (xf && xm >= 0 && xm < p.IX) {
        BnH8 bn;
        bn_h8_setup(bn, p.in_scale + (c * PC + pl) * 8, p.in_shift + (c * PC + pl) * 8);
        unsigned char* dp = smem + p.off_a + slot_f * p.SLOT + pl * p.PS + pix0 * 16;
        for (int i = pix0; i < p.RUN; i += pstep, dp += sstep) {
          if (soff[i] >= 0) {
            uint4* q = reinterpret_cast<uint4*>(dp);
            *q = bn_relu8(*q, bn, relu);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fa + 8 * slot_f);
      if (++slot_f == RA) slot_f = 0;
    };
    for (int st = 0; st < nstage; ++st) {
      const int tx = st / p.NCH, c = st - tx * p.NCH;
      mbar_wait(bar_ea + 8 * slot_i, par);
      const int xm = ox + tx * p.dx - p.px;
      const bool xok = xm >= 0 && xm < p.IX;
      long long plane_off = (long long)(c * PC + pl) * 8;
      if (phased_in) {  // channel plane -> (stride phase, 8-channel group): the phase selects a sub-lattice of the input
        int phi = (c * PC + pl) / p.Pc;
        const int cg = (c * PC + pl) - phi * p.Pc;
        const int fz = phi % p.ips[2]; phi /= p.ips[2];
        const int fy = phi % p.ips[1], fx = phi / p.ips[1];
        plane_off = (long long)cg * 8 + fx * p.in_ph[0] + fy * p.in_ph[1] + fz * p.in_ph[2];
      }
      const __half* src = p.in + (size_t)(xok ? xm : 0) * (size_t)p.in_xs + plane_off;
      uint32_t dst = a_base + (uint32_t)(slot_i * p.SLOT + pl * p.PS + pix0 * 16);
      for (int i = pix0; i < p.RUN; i += pstep, dst += sstep) {
        const int o = soff[i];
        const bool ok = xok && o >= 0;
        cp_async16(dst, ok ? src + o : p.in, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (++slot_i == RA) { slot_i = 0; par ^= 1; }
      // ONE stage in flight behind the one just issued.  Two (measured, RA = 3): 645 -> 503 TFLOP/s on the 256-channel 2D
      // level -- the producer blocks on the ring with a landed stage still unpublished, so the MMA sees one stage less
      if (st >= 1) {
        cp_async_wait<1>();
        finish(st - 1);
      }
    }
    cp_async_wait<0>();
    finish(nstage - 1);
  }
  if (warp == 8) {
    // =========================================== MMA ISSUER ==========================================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(Nc >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, descriptor version 1
    const uint32_t lbo_a = ((uint32_t)p.PS >> 4) << 16;                    // next K8 slab = next channel plane
    const uint32_t lbo_b = (((uint32_t)(Nc * 16)) >> 4) << 16;
    const uint32_t kstep_a = 2u * ((uint32_t)p.PS >> 4), kstep_b = 2u * (uint32_t)Nc;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t acc = 0;
    const int nstage = p.KX * p.NCH;
    for (int st = 0; st < nstage; ++st) {
      mbar_wait(bar_fa + 8 * sa, pa);
      tc_fence_after();
      const uint32_t abase = (a_base + (uint32_t)(sa * p.SLOT)) >> 4;
      int ty = 0, tz = 0;
      for (int t = 0; t < KYZ; t += p.TB) {
        mbar_wait(bar_fb + 8 * sb, pb);
        tc_fence_after();
        for (int tt = 0; tt < p.TB; ++tt) {
          const uint32_t bbase = (b_base + (uint32_t)((sb * p.TB + tt) * p.BT)) >> 4;
          const uint32_t tap = (uint32_t)(ty * p.dy * p.Zv + tz * p.dz);  // pixels == 16-byte units
          for (int k = 0; k < PC / 2; ++k) {
            const uint64_t ad = desc_hi | (uint64_t)((abase + tap + (uint32_t)k * kstep_a) | lbo_a);
            const uint64_t bd = desc_hi | (uint64_t)((bbase + (uint32_t)k * kstep_b) | lbo_b);
            if (elect_one()) {
              umma_f16(tmem_base, ad, bd, idesc, acc);
              if (MB > 1) umma_f16(tmem_base + (uint32_t)Nc, ad + 128u, bd, idesc, acc);
              if (MB > 2) umma_f16(tmem_base + (uint32_t)(2 * Nc), ad + 256u, bd, idesc, acc);
              if (MB > 3) umma_f16(tmem_base + (uint32_t)(3 * Nc), ad + 384u, bd, idesc, acc);
            }
            __syncwarp();
            acc = 1u;
          }
          if (++tz == p.KZ) { tz = 0; ++ty; }
        }
        if (elect_one()) umma_commit(bar_eb + 8 * sb);  // these B tiles are consumed
        __syncwarp();
        if (++sb == RB) { sb = 0; pb ^= 1; }
      }
      if (elect_one()) umma_commit(bar_ea + 8 * sa);    // this A stage is consumed
      __syncwarp();
      if (++sa == RA) { sa = 0; pa ^= 1; }
    }
    if (elect_one()) umma_commit(bar_t);
    __syncwarp();
  } else if (warp == 9) {
    // =========================================== B LOADER ============================================
    if (lane == 0) {
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wp) + (size_t)ns * p.E * Nc * 16;
      const uint32_t bt = (uint32_t)p.BT;
      int sb = 0;
      uint32_t pb = 1;
      for (int tx = 0; tx < p.KX; ++tx)
        for (int c = 0; c < p.NCH; ++c)
          for (int t = 0; t < KYZ; t += p.TB) {
            mbar_wait(bar_eb + 8 * sb, pb);
            mbar_expect_tx(bar_fb + 8 * sb, bt * (uint32_t)p.TB);
            for (int tt = 0; tt < p.TB; ++tt)
              bulk_g2s(b_base + (uint32_t)(sb * p.TB + tt) * bt, wsrc + ((size_t)(tx * p.E_tx + (t + tt) * p.P + c * PC)) * Nc * 16, bt,
                       bar_fb + 8 * sb);
            if (++sb == RB) { sb = 0; pb ^= 1; }
          }
    }
  }
  if (warp < 8) {
    // =========================================== EPILOGUE ============================================
    // Warps 0-3 take the lower half of the CTA's column chunks; the producer warps 4-7 (done staging by now) the upper half:
    // warp w reads TMEM lanes 32 * (w % 4) .. +31, so the two groups split the COLUMNS of the same accumulator rows.
    const int row = threadIdx.x & 127;  // accumulator row within an M-block == TMEM lane
    const int ewarp = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(ewarp * 32) << 16;
    const bool do_stats = p.stats != nullptr;
    const bool affine = p.out_scale != nullptr;
    const bool has_bias = p.bias != nullptr;
    const int out_relu = p.out_relu, cout = p.cout;
    const int nch = min(Nc, cout - ns * Nc);
    const bool phased = p.ops[0] * p.ops[1] * p.ops[2] > 1;
    long long poff[4];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int q = q0 + mb * 128 + row;
      const int r = q / p.Zv, oz = q - r * p.Zv;
      const int nn = r / p.Yv, oy = r - nn * p.Yv;
      poff[mb] = (mb < MB && nn < p.N && oy < p.OY && oz < p.OZ) ? nn * p.out_sn + oy * p.out_sy + oz * p.out_sz : -1;
    }
    __half* obase = reinterpret_cast<__half*>(p.out) + p.out_base + (long long)ox * p.out_sx + p.out_c_off;
    mbar_wait(bar_t, 0);
    tc_fence_after();
    // column chunk outermost: the per-channel sums of the MB M-blocks are added per thread first, ONE shuffle reduction
    // per 16 columns (it was one per M-block: the reductions were most of the epilogue's instructions)
    const int nchunks16 = (nch + 15) >> 4, csplit = ((nchunks16 + 1) >> 1) << 4;
    const int c_lo = half ? csplit : 0, c_hi = half ? nch : min(nch, csplit);
#pragma unroll 1
    for (int cc = c_lo; cc < c_hi; cc += 16) {
      float s1[16], s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
      float bs[16], oa[16], ob[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { bs[j] = sbias[cc + j]; oa[j] = sbias[Nc + cc + j]; ob[j] = sbias[2 * Nc + cc + j]; }
#pragma unroll 1
      for (int mb = 0; mb < MB; ++mb) {
        const bool valid = poff[mb] >= 0;
        float v[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)(mb * Nc + cc), v);
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += bs[j];
        }
        if (do_stats && valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s1[j] += v[j];
            s2[j] = fmaf(v[j], v[j], s2[j]);
          }
        }
        if (valid) {
          if (affine) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], oa[j], ob[j]);
          }
          if (out_relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          __half2 h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          if (phased) {
            // each 8-channel half of the chunk belongs to one stride phase: its own spatial offset
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int ch = ns * Nc + cc + 8 * hh;
              if (ch < cout) {
                int phi = ch / p.cpp;
                const int co = ch - phi * p.cpp;
                const int fz = phi % p.ops[2]; phi /= p.ops[2];
                const int fy = phi % p.ops[1], fx = phi / p.ops[1];
                __half* o = obase + poff[mb] + fx * p.out_ph[0] + fy * p.out_ph[1] + fz * p.out_ph[2] + co;
                *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(&h[4 * hh]);
              }
            }
          } else {
            __half* o = obase + poff[mb] + ns * Nc + cc;
            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(&h[0]);
            if (cc + 8 < nch) *reinterpret_cast<uint4*>(o + 8) = *reinterpret_cast<uint4*>(&h[4]);
          }
        }
      }
      if (do_stats) {
        const float r1 = reduce16(s1, lane);
        const float r2 = reduce16(s2, lane);
        if ((lane & 1) == 0) {
          sstat[ewarp * 2 * Nc + cc + (lane >> 1)] = r1;   // this lane is the slot's only writer
          sstat[ewarp * 2 * Nc + Nc + cc + (lane >> 1)] = r2;
        }
      }
    }
    tc_fence_before();
    if (do_stats) stats_tail(p, sstat, (int)threadIdx.x, ns, Nc, 256);
  }

  // ---- teardown --------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// weights: fp32 [taps][cin][cout] (hcu_weight_gather layout, one group) -> fp16 [nsplit][E][Nc][8]
// packed weight layouts: normal [nsplit][tx][K8 slab e][Nc][8]; wide (x-fused MMAs) [nsplit][e][KX-1-tx][Nc][8]
__device__ __forceinline__ void unpack_index(uint32_t r, int KX, int E_tx, int Nc, int wide, int& nn, int& e, int& tx, int& ns) {
  nn = (int)(r % (uint32_t)Nc); r /= (uint32_t)Nc;
  if (wide) {
    tx = KX - 1 - (int)(r % (uint32_t)KX); r /= (uint32_t)KX;
    e = (int)(r % (uint32_t)E_tx);
    ns = (int)(r / (uint32_t)E_tx);
  } else {
    e = (int)(r % (uint32_t)E_tx); r /= (uint32_t)E_tx;
    tx = (int)(r % (uint32_t)KX);
    ns = (int)(r / (uint32_t)KX);
  }
}

__global__ void pack_tc_kernel(const float* __restrict__ w, __half* __restrict__ out, int KX, int KYZ, int P, int E_tx,
                               int Nc, int nsplit, int cin, int cout, int wide) {
  const long long total = (long long)nsplit * KX * E_tx * Nc * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    int nn, e, tx, ns;
    unpack_index((uint32_t)(i >> 3), KX, E_tx, Nc, wide, nn, e, tx, ns);
    float v = 0.f;
    if (e < KYZ * P) {
      const int t = e / P, pl = e % P;
      const int ci = pl * 8 + j, co = ns * Nc + nn;
      if (ci < cin && co < cout) v = w[((long long)(tx * KYZ + t) * cin + ci) * cout + co];
    }
    out[i] = __float2half_rn(v);
  }
}

// weights straight from the reference-layout parameter (HcuWeightMap) -> fp16 [nsplit][E][Nc][8]
__global__ void pack_tc_ref_kernel(HcuWeightMap m, const float* __restrict__ ref, __half* __restrict__ out, int KX, int KYZ,
                                   int P, int E_tx, int Nc, int nsplit, int cin, int cout, int wide) {
  const long long total = (long long)nsplit * KX * E_tx * Nc * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    int nn, e, tx, ns;
    unpack_index((uint32_t)(i >> 3), KX, E_tx, Nc, wide, nn, e, tx, ns);
    float v = 0.f;
    if (e < KYZ * P) {
      const int t = e / P, pl = e % P;
      const int ci = pl * 8 + j, co = ns * Nc + nn;
      if (ci < cin && co < cout) {
        const long long idx = wm_index(m, ((long long)(tx * KYZ + t) * cin + ci) * cout + co);
        if (idx >= 0) {
          v = ref[idx];
          if (m.fold) v += ref[idx + m.fold_stride];
        }
      }
    }
    out[i] = __float2half_rn(v);
  }
}


// ---- batched weight packing: one launch for every tensor-core conv of a step ------------------------------------
struct PackJob {
  HcuWeightMap m;
  long long ref_off, out_off, total;
  int KX, KYZ, P, E_tx, Nc, nsplit, cin, cout;
  int block0, nblocks, wide, pad_;
};
static_assert(sizeof(PackJob) <= HCU_BATCH_JOB_BYTES, "PackJob does not fit its table slot");
constexpr int kPackPerBlock = 2048;  // elements per block (256 threads x 8)

__global__ void __launch_bounds__(256) pack_tc_batch_kernel(const unsigned char* __restrict__ jobs, int n,
                                                            const float* __restrict__ params, unsigned char* __restrict__ packed) {
  __shared__ PackJob J;
  __shared__ int jidx;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n - 1;  // last job whose block0 <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const PackJob* pj = reinterpret_cast<const PackJob*>(jobs + (size_t)mid * HCU_BATCH_JOB_BYTES);
      if (pj->block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    jidx = lo;
  }
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + (size_t)jidx * HCU_BATCH_JOB_BYTES);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = threadIdx.x; i < (int)(sizeof(PackJob) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const float* ref = params + J.ref_off;
  __half* out = reinterpret_cast<__half*>(packed + J.out_off);
  const long long base = (long long)(blockIdx.x - J.block0) * kPackPerBlock;
  // One thread per 16-byte output unit (8 consecutive input channels of one K8 slab row): the index decomposition -- a
  // dozen 32-bit divisions -- is done once per unit instead of once per element (the per-element kernel took 0.63 ms per
  // step for the 31 M parameters of the classic 2D U-Net: 100 G elements/s, instruction bound).
  for (int k = threadIdx.x; k < kPackPerBlock / 8; k += 256) {
    const long long i = base + 8 * k;
    if (i >= J.total) break;
    int nn, e, tx, ns;  // 32-bit index arithmetic: job sizes are < 2^31 (checked at build time)
    unpack_index((uint32_t)(i >> 3), J.KX, J.E_tx, J.Nc, J.wide, nn, e, tx, ns);
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(0.f);
    const int co = ns * J.Nc + nn;
    if (e < J.KYZ * J.P && co < J.cout) {
      const int t = e / J.P, pl = e - t * J.P;
      // tap (tx, t) -> (jx, jy, jz) of the map; output channel b -> (phase, b) when the phases sit on the b side
      const HcuWeightMap& m = J.m;
      const int tl = tx * J.KYZ + t;
      const int jz = tl % m.j[2], tq = tl / m.j[2];
      const int jy = tq % m.j[1], jx = tq / m.j[1];
      long long idx0 = m.base + (long long)(m.t0[0] + jx * m.tstep[0]) * m.st[0] + (long long)(m.t0[1] + jy * m.tstep[1]) * m.st[1] +
                       (long long)(m.t0[2] + jz * m.tstep[2]) * m.st[2];
      int b = co;
      int gb = 0;
      if (m.bdiag > 1) { gb = b / m.nb; b -= gb * m.nb; idx0 += (long long)gb * m.sg; }
      if (m.phase_on == 2) {
        int phi = b / m.nb;
        b -= phi * m.nb;
        const int pz = phi % m.ph[2]; phi /= m.ph[2];
        const int py = phi % m.ph[1], px = phi / m.ph[1];
        idx0 += px * m.pst[0] + py * m.pst[1] + pz * m.pst[2];
      }
      idx0 += (long long)b * m.sb;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int a = pl * 8 + j;
        if (a < J.cin) {
          long long idx = idx0;
          if (m.bdiag > 1) {
            if (a / m.na != gb) continue;  // structural zero of the block-diagonal weight
            a -= gb * m.na;
          }
          if (m.phase_on == 1) {
            int phi = a / m.na;
            a -= phi * m.na;
            const int pz = phi % m.ph[2]; phi /= m.ph[2];
            const int py = phi % m.ph[1], px = phi / m.ph[1];
            idx += px * m.pst[0] + py * m.pst[1] + pz * m.pst[2];
          }
          idx += (long long)a * m.sa;
          float v = ref[idx];
          if (m.fold) v += ref[idx + m.fold_stride];
          h[j] = __float2half_rn(v);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(h);
  }
}

// ---------------------------------------------------------------------------------------------------
// host-side configuration
// ---------------------------------------------------------------------------------------------------
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// returns 0 and fills p (geometry part) when the TC kernel takes this descriptor, else a reason string
static const char* configure_classic(const HcuConvDesc* d, Params& p) {
  p.ks = 0;
  if (d->dtype_in != HCU_F16) return "input must be fp16";
  if (d->dtype_out != HCU_F16 && d->dtype_out != HCU_F32) return "output must be fp16 or fp32";
  if (d->groups != 1) return "groups != 1";
  if (d->in_cpitch % 8 != 0 || d->in_c_off != 0) return "input channel pitch must be a multiple of 8, offset 0";
  const int P = d->in_cpitch / 8;
  if (P != 1 && P != 2 && P != 4 && P != 8 && P != 16 && P != 32) return "input channel pitch must be 8..256, power of two";
  if (d->cin > d->in_cpitch) return "cin > pitch";
  for (int i = 0; i < 3; ++i) {
    p.ops[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff);
    p.ips[i] = std::max(1, (d->iphase >> (8 * i)) & 0xff);
  }
  const int nph_o = p.ops[0] * p.ops[1] * p.ops[2], nph_i = p.ips[0] * p.ips[1] * p.ips[2];
  if (nph_o > 1) {
    if (d->dtype_out != HCU_F16 || d->cout % nph_o || (d->cout / nph_o) % 8 || d->out_c_off % 8 || d->out_cpitch % 8)
      return "ophase needs fp16 output and 8-channel aligned phases";
    for (int i = 0; i < 3; ++i)
      if (d->ostep[i] < p.ops[i]) return "ophase larger than the output step";
  }
  if (nph_i > 1 && (d->cin != d->in_cpitch || d->cin % nph_i || (d->cin / nph_i) % 8)) return "iphase needs 8-channel aligned phases";
  p.cpp = d->cout / nph_o;
  p.Pc = P / nph_i;
  {
    // element strides of the (full-resolution) input tensor per COARSE step, and per phase step
    const long long cr = d->in_cpitch / nph_i;
    const long long fz = (long long)d->in_size[2] * p.ips[2], fy = (long long)d->in_size[1] * p.ips[1],
                    fx = (long long)d->in_size[0] * p.ips[0];
    p.in_ph[2] = cr; p.in_ph[1] = fz * cr; p.in_ph[0] = fy * fz * cr;
    p.in_zs = (int)(p.ips[2] * cr);
    if (fz * cr * p.ips[1] * d->in_size[1] >= 0x7fffffffLL) return "x-plane too large";
    p.in_ys = (int)(p.ips[1] * fz * cr);
    p.in_xs = p.ips[0] * fy * fz * cr;
    p.in_ns = fx * fy * fz * cr;
  }
  for (int i = 0; i < 3; ++i)
    if (d->istep[i] != 1) return "strided gather";
  if (d->cout < 2 && d->dtype_out == HCU_F16) return "single output channel";
  p.N = d->batch; p.IX = d->in_size[0]; p.IY = d->in_size[1]; p.IZ = d->in_size[2];
  p.Cp = d->in_cpitch; p.P = P;
  p.OX = d->out_size[0]; p.OY = d->out_size[1]; p.OZ = d->out_size[2];
  p.KX = d->taps[0]; p.KY = d->taps[1]; p.KZ = d->taps[2];
  p.dx = d->dil[0]; p.dy = d->dil[1]; p.dz = d->dil[2];
  p.px = d->pad[0]; p.py = d->pad[1]; p.pz = d->pad[2];
  p.Yv = p.OY + (p.KY - 1) * p.dy;
  p.Zv = p.OZ + (p.KZ - 1) * p.dz;
  {
    static int zal = -1;
    if (zal < 0) { const char* e = getenv("HCU_TC_ZALIGN"); zal = e ? atoi(e) : 1; }
    if (zal > 1 && p.Zv > zal) p.Zv = round_up(p.Zv, zal);
  }
  const int span = (p.KX - 1) * p.dx + 1;
  if (span > kMaxRing) return "x extent of the filter too large";
  const int per_tx = p.KY * p.KZ * P;
  p.E_tx = round_up(per_tx, 2);
  p.npairs = p.E_tx / 2;
  if (p.npairs > kMaxPairs) return "too many taps per x-plane";
  p.E = p.KX * p.E_tx;
  p.cout = d->cout;
  const int npad = round_up(d->cout, 16);
  const int halo = (p.KY - 1) * p.dy * p.Zv + (p.KZ - 1) * p.dz;
  const int plane_q = p.Yv * p.Zv;
  // candidates: big M first; Nc as large as fits
  const int m_cands[4] = {512, 384, 256, 128};
  if (p.KX * p.npairs > kMaxTab) return "too many K16 steps per output plane";
  // Ring depth R = span (planes the MMA of one output needs) + D (planes in flight, unpublished) + slack (published planes
  // the MMA has not consumed yet: lets producer and MMA overlap instead of alternating).  Sweeps: want D = 2 + slack 3,
  // then D = 2 + slack 2, D = 1 + slack 1, anything.  Within a sweep aim for 2 CTAs / SM (the register file allows no
  // more), then whatever fits; big M first; Nc as large as fits.
  const int want[4] = {5, 4, 2, 0};
  static int wide_on = -1;
  if (wide_on < 0) { const char* e = getenv("HCU_TC_WIDE"); wide_on = e ? atoi(e) : 1; }
  // x-fused ("wide") MMAs for the 8..32-channel levels: the tensor pipe's time there is the shared-memory fetch of the A
  // operand (profiles/r01_umma_rate.txt), which this mode does once per input plane instead of once per (plane, tx)
  const bool can_wide = wide_on && p.KX >= 2 && p.KX <= 3 && p.dx == 1;
  const bool small_wide = can_wide && npad <= 32;  // the byte-bound levels: insist on the wide mode
  auto search = [&](int pass, bool full_n) -> bool {
    const int budget = pass == 0 ? 112 * 1024 : kSmemLimit;
    for (int sweep = 0; sweep < (full_n ? 3 : 4); ++sweep) {
      for (int mi = 0; mi < 4; ++mi) {
        const int M = m_cands[mi];
        if (M > 128 && M - 128 >= plane_q) continue;  // do not use a longer run than the plane needs
        const int MB = M / 128;
        if (small_wide && 4 * MB * npad > 256 && MB > 1) continue;  // prefer a shorter run that can go wide at 2 CTAs / SM
        int run = M + halo;
        int ps = run * 16;
        if (P > 1) {  // spread the channel planes over the banks: PS = g (mod 2g), g = max(16, 128 / P)
          const int g = P >= 8 ? 16 : 128 / P;
          ps = round_up(ps, 2 * g) + g;
        }
        const int slot = ps * P;
        for (int nc = npad > 128 ? 128 : npad; nc >= 16; nc -= 16) {
          if (npad % nc != 0) continue;
          if (full_n && nc != npad) break;  // this search wants every output channel in one CTA
          // wide: 4 accumulator slots per M-block in TMEM, an MMA spans up to KX slots (N = KX * nc <= 256)
          const bool wide = can_wide && p.KX * nc <= 256 && 4 * MB * nc <= 512 && (!small_wide || 4 * MB * nc <= 256);
          if ((wide ? 4 : 2) * MB * nc > 512) continue;
          // the specialised variants (Nc = 16, wide) keep KX rotated copies of the weights (rotating accumulator window)
          const int wbytes = p.E * nc * 16 * ((wide && nc == 16 && npad == 16) ? p.KX : 1);
          const int R = (wide ? 1 : span) + want[sweep];  // wide: every input plane is consumed by ONE step
          if (R > kMaxRing) continue;
          const int off_w = 0;
          const int off_a = round_up(wbytes, 128);
          const int off_bar = off_a + R * slot;
          const int off_stat = round_up(off_bar + 8 * (2 * R + 9) + 8, 16);
          const int total = off_stat + (nc == 16 ? 19 : 11) * nc * 4 + 128;  // bulk mode (Nc = 16 only): 8 statistics slots
          if (total > budget) continue;
          p.M = M; p.MB = MB; p.RUN = run; p.PS = ps; p.SLOT = slot; p.R = R; p.wide = wide ? 1 : 0;
          p.D = want[sweep] >= 4 ? 2 : (want[sweep] >= 2 ? 1 : 0);
          p.Nc = nc; p.nsplit = npad / nc;
          p.off_w = off_w; p.off_a = off_a; p.off_bar = off_bar; p.off_tab = 0; p.off_stat = off_stat;
          p.smem_bytes = total;
          int cols = (wide ? 4 : 2) * MB * nc, t = 32;
          while (t < cols) t <<= 1;
          p.tmem_cols = t;
          p.n_runs = (plane_q + M - 1) / M;
          // K16 steps of one tx group: entries 2e, 2e+1 of its (ty, tz, channel-plane) list
          for (int tx = 0; tx < p.KX; ++tx)
            for (int e = 0; e < p.npairs; ++e) {
              const int e0 = 2 * e, e1 = 2 * e + 1;
              const int t0 = e0 / P, c0 = e0 % P;
              const int off0 = ((t0 / p.KZ) * p.dy * p.Zv + (t0 % p.KZ) * p.dz) * 16 + c0 * ps;
              int off1 = off0;  // odd tail: the second K8 half re-reads the same rows against zero weights
              if (e1 < per_tx) {
                const int t1 = e1 / P, c1 = e1 % P;
                off1 = ((t1 / p.KZ) * p.dy * p.Zv + (t1 % p.KZ) * p.dz) * 16 + c1 * ps;
              }
              if (off1 < off0 || ((off1 - off0) >> 4) > 0x3fff) return false;
              p.tab[tx * p.npairs + e].x = ((uint32_t)off0 >> 4) | (((uint32_t)(off1 - off0) >> 4) << 16);
              // wide: [K8 slab][KX-1-tx][Nc] -- the tx block is added by the issuer; normal: [tx][K8 slab][Nc]
              p.tab[tx * p.npairs + e].y = wide ? (uint32_t)(e0 * p.KX * nc) : (uint32_t)((tx * p.E_tx + e0) * nc * 16) >> 4;
            }
          return true;
        }
      }
    }
    return false;
  };
  // Two candidates: the best configuration that leaves room for two CTAs per SM, and the best one using the whole SM.
  // The deep levels' weights force a split of the output channels over CTAs (nsplit) at the small budget, and every
  // split re-fetches the A operand: tensor-pipe time per output ~ nsplit * (32 + Nc/4) clk per K16 step
  // (profiles/r01_umma_rate.txt).  Take the whole-SM configuration when it cuts that by more than a quarter.
  Params p0 = p, p1 = p;
  bool ok0, ok1;
  { ok0 = search(0, false); p0 = p; }
  if (ok0 && p0.nsplit > 1) {
    // The deepest ring only fitted beside a SLICE of the weights: a shallower (still pipelined) ring with all output
    // channels in the CTA keeps two CTAs per SM without re-staging the A operand per slice.  (Classic 2D U-Net, 32 -> 32
    // channels on 568-pixel rows: the whole-SM configuration ran one 9-warp CTA per SM at 11 % tensor-pipe activity.)
    Params keep = p0;
    if (search(0, true)) p0 = p; else p0 = keep;
  }
  { ok1 = search(1, false); p1 = p; }
  if (ok1 && p1.nsplit > 1) {
    Params keep = p1;
    if (search(1, true)) p1 = p; else p1 = keep;
  }
  if (!ok0 && !ok1) return "does not fit in shared memory";
  auto cost = [](const Params& q) { return (double)q.nsplit * (32.0 + q.Nc / 4.0) * (q.wide ? 0.55 : 1.0); };
  if (ok0 && (!ok1 || cost(p1) > 0.75 * cost(p0))) p = p0; else p = p1;
  return nullptr;
}

// ---- K-streamed kernel: geometry, tile search --------------------------------------------------------------------
// Work per CTA: (output x-plane, run of M = MB * 128 flat positions over the stacked images, Nc output channels).  The
// search minimises waves * max(tensor time, L2 streaming time) + fixed cost over (M, Nc, PC) under the shared-memory and
// TMEM budgets.  A 2D problem (Z == 1) is re-read as ONE x-plane of Y x Z = image rows x columns, so that the filter
// rows are flat shifts as well (a 2D image row as an "x-plane" would leave an MMA's 128 rows mostly empty on the deep
// levels: 28 pixels wide at the bottom of the classic U-Net).
static bool ks_flat2d(const HcuConvDesc* d) {
  return d->in_size[2] == 1 && d->out_size[2] == 1 && d->taps[2] == 1 && d->pad[2] == 0;
}

static const char* configure_ks(const HcuConvDesc* d, Params& p) {
  p.ks = 1;
  if (d->dtype_in != HCU_F16 || d->dtype_out != HCU_F16) return "fp16 in and out";
  if (d->groups != 1) return "groups != 1";
  if (d->in_cpitch % 32 != 0 || d->in_c_off != 0 || d->cin != d->in_cpitch) return "input channels: dense multiple of 32";
  if (d->cin < 64) return "fewer than 64 input channels";
  if (d->cout % 8 != 0 || d->out_cpitch % 8 != 0 || d->out_c_off % 8 != 0) return "output channels: multiples of 8";
  for (int i = 0; i < 3; ++i) {
    p.ops[i] = std::max(1, (d->ophase >> (8 * i)) & 0xff);
    p.ips[i] = std::max(1, (d->iphase >> (8 * i)) & 0xff);
    if (d->istep[i] != 1) return "strided gather";
  }
  const int nph_o = p.ops[0] * p.ops[1] * p.ops[2], nph_i = p.ips[0] * p.ips[1] * p.ips[2];
  if (nph_o > 1) {
    if (d->cout % nph_o || (d->cout / nph_o) % 8) return "ophase needs 8-channel aligned phases";
    for (int i = 0; i < 3; ++i)
      if (d->ostep[i] < p.ops[i]) return "ophase larger than the output step";
  }
  if (nph_i > 1 && (d->cin % nph_i || (d->cin / nph_i) % 8)) return "iphase needs 8-channel aligned phases";
  p.cpp = d->cout / nph_o;
  const int P = d->in_cpitch / 8;
  p.P = P; p.Pc = P / nph_i; p.Cp = d->in_cpitch;
  p.N = d->batch;
  p.cout = d->cout;
  // axis roles: a (march / separate planes), b, c (flat plane).  2D: a is a dummy axis of extent 1.
  const bool flat2d = ks_flat2d(d);
  const int A = flat2d ? -1 : 0, B = flat2d ? 0 : 1, Cc = flat2d ? 1 : 2;
  auto get = [&](const int32_t* v, int ax, int dflt) { return ax < 0 ? dflt : v[ax]; };
  p.IX = get(d->in_size, A, 1); p.IY = d->in_size[B]; p.IZ = d->in_size[Cc];
  p.OX = get(d->out_size, A, 1); p.OY = d->out_size[B]; p.OZ = d->out_size[Cc];
  p.KX = get(d->taps, A, 1); p.KY = d->taps[B]; p.KZ = d->taps[Cc];
  p.dx = get(d->dil, A, 1); p.dy = d->dil[B]; p.dz = d->dil[Cc];
  p.px = get(d->pad, A, 0); p.py = d->pad[B]; p.pz = d->pad[Cc];
  {
    // element strides of the (full-resolution) input tensor per COARSE step of each descriptor axis, and per phase step
    const long long cr = d->in_cpitch / nph_i;
    const long long fz = cr, fy = fz * d->in_size[2] * p.ips[2], fx = fy * d->in_size[1] * p.ips[1],
                    fn = fx * d->in_size[0] * p.ips[0];
    const long long st[3] = {fx * p.ips[0], fy * p.ips[1], fz * p.ips[2]};
    if (fn * d->batch >= 0x7fffffffLL) return "input tensor too large for 32-bit offsets";
    p.in_ns = fn;
    p.in_xs = A < 0 ? 0 : st[A];
    p.in_ys = (int)st[B];
    p.in_zs = (int)st[Cc];
    p.in_ph[0] = fx; p.in_ph[1] = fy; p.in_ph[2] = fz;
  }
  p.Yv = p.OY + (p.KY - 1) * p.dy;
  p.Zv = p.OZ + (p.KZ - 1) * p.dz;
  const int KYZ = p.KY * p.KZ;
  if (p.KX > 8 || KYZ > 64) return "too many taps";
  p.E_tx = KYZ * P;
  p.npairs = p.E_tx / 2;
  p.E = p.KX * p.E_tx;
  p.wide = 0;
  const int npad = round_up(d->cout, 16);
  const int halo = (p.KY - 1) * p.dy * p.Zv + (p.KZ - 1) * p.dz;
  const long long n_last = ((long long)(p.N - 1) * p.Yv + p.OY - 1) * p.Zv + p.OZ;
  if (n_last + 512 + halo >= 0x7fffffffLL) return "plane too large";
  p.n_last = (int)n_last;
  const int sms = 148;
  double best = 1e300;
  Params b = p;
  bool found = false;
  const int m_cands[3] = {512, 256, 128};
  // reserved[1] (tests): forced tile, MB | Nc << 8 | PC << 20 (0 fields = free)
  const int f_mb = d->reserved[1] & 0xf, f_nc = (d->reserved[1] >> 8) & 0xfff, f_pc = (d->reserved[1] >> 20) & 0xf;
  for (int mi = 0; mi < 3; ++mi) {
    const int M = m_cands[mi], MB = M / 128;
    if (f_mb ? MB != f_mb : (M > 128 && M - 128 >= n_last)) continue;
    const int run = M + halo;
    for (int nc = std::min(npad, 256); nc >= 16; nc -= 16) {
      if (npad % nc != 0 || MB * nc > 512) continue;
      if (f_nc && nc != f_nc) continue;
      // The chunk width fixes the ORDER in which (channel chunk, tap) partial products enter the fp32 accumulators, and the
      // tile search depends on the batch size: one width (4 planes = 32 channels) for every launch keeps an output pixel's
      // arithmetic independent of how the batch is tiled (an image alone == the same image inside a batch, bit for bit).
      // Measured: 8-plane chunks are no faster ((256, 256) tile 733 vs 733 TFLOP/s).  Tests may force 8.
      for (int pc = f_pc ? f_pc : 4; pc >= 4; pc >>= 1) {
        if (P % pc != 0 || (f_pc && pc != f_pc)) continue;
        const int bt = pc * nc * 16;
        if (bt > 32768) continue;
        // one barrier round per filter row of a chunk (KZ taps) while the slot stays <= 24 KB: fewer, larger rounds
        static int tb_on = -1;
        if (tb_on < 0) { const char* e = getenv("HCU_KS_TB"); tb_on = e ? atoi(e) : 1; }
        const int tb = (tb_on && p.KZ > 1 && p.KZ * bt <= 24576) ? p.KZ : 1;
        const int bslot = tb * bt;
        int ps = run * 16;
        { const int g = pc >= 8 ? 16 : 128 / pc; ps = round_up(ps, 2 * g) + g; }
        if ((ps >> 4) > 0x3fff) continue;
        const int slot = ps * pc;
        // ring depths: a third A stage first (measured on the (256, 256) tile: RA = 3 / RB = 2 733 TFLOP/s, RA = 2 / RB = 3
        // 668), then weight tiles worth up to 64 KB in flight behind the one being consumed; among equals the smaller footprint
        static int f_ra = -1, f_rb = -1;
        if (f_ra < 0) { const char* e = getenv("HCU_KS_RA"); f_ra = e ? atoi(e) : 0; e = getenv("HCU_KS_RB"); f_rb = e ? atoi(e) : 0; }
        int best_ra = 0, best_rb = 0, best_total = 0;
        double best_score = -1.0;
        int offs[5] = {0, 0, 0, 0, 0};
        for (int ra = (f_ra > 3 ? f_ra : 3); ra >= 2; --ra) {
          for (int rb = 8; rb >= 2; --rb) {
            if ((f_ra && ra != f_ra) || (f_rb && rb != f_rb)) continue;
            const int off_a = round_up(rb * bslot, 128);
            const int off_tab = off_a + ra * slot;
            const int off_bar = round_up(off_tab + run * 4, 8);
            const int off_stat = round_up(off_bar + 8 * (2 * ra + 2 * rb + 1) + 8, 16);
            const int total = off_stat + 11 * nc * 4 + 128;
            if (total > kSmemLimit) continue;
            const double score = 2.0 * (ra - 1) + std::min(1.0, (double)(rb - 1) * bslot / 65536.0) - 1e-7 * total;
            if (score > best_score) {
              best_score = score; best_ra = ra; best_rb = rb; best_total = total;
              offs[0] = off_a; offs[1] = off_tab; offs[2] = off_bar; offs[3] = off_stat;
            }
          }
        }
        if (best_ra == 0) continue;
        {
          const int ra = best_ra, rb = best_rb, total = best_total;
          const int off_w = 0, off_a = offs[0], off_tab = offs[1], off_bar = offs[2], off_stat = offs[3];
          const int nch = P / pc;
          // one M = 128, K = 16 MMA: shared-memory operand fetch (32 + N/4 clk, profiles/r01_umma_rate.txt) vs tensor rate
          // (N/2 clk), times a measured efficiency of the issue / barrier machinery per column-chunk width (2D U-Net 256- and
          // 512-channel levels: (M, Nc) = (256, 256) runs at 730 TFLOP/s = 174 clk per MMA, (512, 128) at 620 - 645 = 100 clk, (256, 128) at 510 - 520)
          const double mma1 = std::max(32.0 + nc / 4.0, nc / 2.0) * (nc >= 256 ? 1.36 : nc >= 128 ? 1.56 : 1.8);
          const double t_mma = (double)p.KX * nch * KYZ * (pc / 2) * MB * mma1;
          const double bytes = (double)p.KX * nch * ((double)run * pc * 16 + (double)KYZ * bt);
          const double t_mem = bytes / 28.0;  // L2 -> shared memory per SM: 27 B/clk measured on the (256, 256) tile
          const double t_cta = std::max(t_mma, t_mem) + 4000.0 + MB * (nc / 16) * 150.0 + (ra < 3 ? 0.10 * t_mma : 0.0) +
                               ((rb - 1) * bslot < 65536 ? 0.03 * t_mma : 0.0);
          const long long n_runs = (n_last + M - 1) / M;
          const long long items = (long long)p.OX * n_runs * (npad / nc);
          const double cost = (double)((items + sms - 1) / sms) * t_cta;
          if (cost < best) {
            best = cost; found = true;
            b = p;
            b.M = M; b.MB = MB; b.RUN = run; b.PS = ps; b.SLOT = slot; b.PC = pc; b.NCH = nch; b.RA = ra; b.RB = rb; b.BT = bt; b.TB = tb;
            b.R = ra; b.D = 1;
            b.Nc = nc; b.nsplit = npad / nc;
            b.off_w = off_w; b.off_a = off_a; b.off_tab = off_tab; b.off_bar = off_bar; b.off_stat = off_stat;
            b.smem_bytes = total;
            int t = 32;
            while (t < MB * nc) t <<= 1;
            b.tmem_cols = t;
            b.n_runs = (int)n_runs;
            b.Lx = 1; b.n_xseg = p.OX;
          }
        }
      }
    }
  }
  if (!found) return "does not fit in shared memory";
  p = b;
  return nullptr;
}

// Which kernel takes a descriptor.  reserved[0] is a hint (tests / experiments): 0 auto, 1 classic only, 2 K-streamed only.
// Auto: the classic kernel (whole weight slice resident, x-march ring: every input element fetched ~once) unless it cannot
// take the descriptor or has to split the output channels over CTAs (each split re-stages the A operand).
static const char* configure(const HcuConvDesc* d, Params& p) {
  static int ks_mode = -1;
  if (ks_mode < 0) { const char* e = getenv("HCU_TC_KS"); ks_mode = e ? atoi(e) : 1; }
  const int hint = d->reserved[0];
  if (hint == 2) return configure_ks(d, p);
  Params pc;
  memset(&pc, 0, sizeof(pc));
  const char* why_c = configure_classic(d, pc);
  if (hint == 1 || ks_mode == 0) { p = pc; return why_c; }
  if (why_c == nullptr && pc.nsplit == 1 && ks_mode != 2) { p = pc; return nullptr; }
  Params pk;
  memset(&pk, 0, sizeof(pk));
  const char* why_k = configure_ks(d, pk);
  if (why_k == nullptr) { p = pk; return nullptr; }
  p = pc;
  return why_c;
}

}  // namespace tc
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_conv_tc_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  tc::Params p;
  return tc::configure(d, p) == nullptr ? 1 : 0;
}

extern "C" int hcu_conv_tc_describe(const HcuConvDesc* d, char* buf, int32_t n) {
  if (d == nullptr || buf == nullptr || n <= 0) return HCU_ERR_INVALID;
  tc::Params p;
  const char* why = tc::configure(d, p);
  if (why != nullptr) { snprintf(buf, (size_t)n, "unsupported: %s", why); return 0; }
  if (p.ks)
    snprintf(buf, (size_t)n, "ks M=%d Nc=%d nsplit=%d PC=%d RA=%d RB=%d runs=%d grid=%lld smem=%d tmem=%d", p.M, p.Nc, p.nsplit, p.PC,
             p.RA, p.RB, p.n_runs, (long long)p.OX * p.n_runs * p.nsplit, p.smem_bytes, p.tmem_cols);
  else
    snprintf(buf, (size_t)n, "classic M=%d Nc=%d nsplit=%d wide=%d R=%d D=%d runs=%d smem=%d tmem=%d", p.M, p.Nc, p.nsplit, p.wide,
             p.R, p.D, p.n_runs, p.smem_bytes, p.tmem_cols);
  return 0;
}

extern "C" long long hcu_conv_tc_packed_bytes(const HcuConvDesc* d) {
  tc::Params p;
  if (d == nullptr || tc::configure(d, p) != nullptr) return -1;
  return (long long)p.nsplit * p.E * p.Nc * 16;
}

extern "C" int hcu_conv_tc_pack(const HcuConvDesc* d, const float* w, void* packed, void* stream) {
  HCU_CHECK_ARG(d && w && packed, "conv_tc_pack: null pointer");
  tc::Params p;
  const char* why = tc::configure(d, p);
  HCU_CHECK_ARG(why == nullptr, "conv_tc_pack: unsupported descriptor (%s)", why);
  const long long total = (long long)p.nsplit * p.E * p.Nc * 8;
  int grid = (int)std::min<long long>((total + 255) / 256, 4096);
  tc::pack_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, (__half*)packed, p.KX, p.KY * p.KZ, p.P, p.E_tx, p.Nc,
                                                             p.nsplit, d->cin, d->cout, p.wide);
  HCU_CHECK_LAUNCH("pack_tc");
  return 0;
}

extern "C" int hcu_conv_tc_pack_ref(const HcuConvDesc* d, const HcuWeightMap* m, const float* ref, void* packed,
                                    void* stream) {
  HCU_CHECK_ARG(d && m && ref && packed, "conv_tc_pack_ref: null pointer");
  tc::Params p;
  const char* why = tc::configure(d, p);
  HCU_CHECK_ARG(why == nullptr, "conv_tc_pack_ref: unsupported descriptor (%s)", why);
  {
    const int nph = m->phase_on ? m->ph[0] * m->ph[1] * m->ph[2] : 1;
    const int bd = m->bdiag > 1 ? m->bdiag : 1;
    HCU_CHECK_ARG(bd == 1 || m->phase_on == 0, "conv_tc_pack_ref: block-diagonal maps carry no stride phases");
    HCU_CHECK_ARG(m->groups == 1 && m->j[0] == d->taps[0] && m->j[1] == d->taps[1] && m->j[2] == d->taps[2] &&
                      m->na * bd * (m->phase_on == 1 ? nph : 1) == d->cin && m->nb * bd * (m->phase_on == 2 ? nph : 1) == d->cout,
                  "conv_tc_pack_ref: weight map does not match the descriptor");
  }
  const long long total = (long long)p.nsplit * p.E * p.Nc * 8;
  int grid = (int)std::min<long long>((total + 255) / 256, 4096);
  tc::pack_tc_ref_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*m, ref, (__half*)packed, p.KX, p.KY * p.KZ, p.P, p.E_tx,
                                                                 p.Nc, p.nsplit, d->cin, d->cout, p.wide);
  HCU_CHECK_LAUNCH("pack_tc_ref");
  return 0;
}


extern "C" int hcu_conv_tc_pack_batch_build(const HcuConvDesc* descs, const HcuWeightMap* maps, const int64_t* ref_off,
                                            const int64_t* out_off, int32_t n, void* host_jobs, int32_t* blocks) {
  HCU_CHECK_ARG(descs && maps && ref_off && out_off && host_jobs && blocks && n > 0, "conv_tc_pack_batch_build: bad arguments");
  int b0 = 0;
  for (int i = 0; i < n; ++i) {
    tc::Params p;
    const char* why = tc::configure(&descs[i], p);
    HCU_CHECK_ARG(why == nullptr, "conv_tc_pack_batch_build: job %d: unsupported descriptor (%s)", i, why);
    const HcuWeightMap* m = &maps[i];
    const int nph = m->phase_on ? m->ph[0] * m->ph[1] * m->ph[2] : 1;
    const int bd = m->bdiag > 1 ? m->bdiag : 1;
    HCU_CHECK_ARG(bd == 1 || m->phase_on == 0, "conv_tc_pack_batch_build: job %d: block-diagonal maps carry no stride phases", i);
    HCU_CHECK_ARG(m->groups == 1 && m->j[0] == descs[i].taps[0] && m->j[1] == descs[i].taps[1] && m->j[2] == descs[i].taps[2] &&
                      m->na * bd * (m->phase_on == 1 ? nph : 1) == descs[i].cin && m->nb * bd * (m->phase_on == 2 ? nph : 1) == descs[i].cout,
                  "conv_tc_pack_batch_build: job %d: weight map does not match the descriptor", i);
    HCU_CHECK_ARG(out_off[i] % 16 == 0, "conv_tc_pack_batch_build: job %d: packed offset not 16-byte aligned", i);
    tc::PackJob j;
    memset(&j, 0, sizeof(j));
    j.m = *m; j.ref_off = ref_off[i]; j.out_off = out_off[i];
    j.total = (long long)p.nsplit * p.E * p.Nc * 8;
    HCU_CHECK_ARG(j.total < 0x7fffffffLL, "conv_tc_pack_batch_build: job %d too large", i);
    j.KX = p.KX; j.KYZ = p.KY * p.KZ; j.P = p.P; j.E_tx = p.E_tx; j.Nc = p.Nc; j.nsplit = p.nsplit;
    j.cin = descs[i].cin; j.cout = descs[i].cout; j.wide = p.wide;
    j.block0 = b0;
    j.nblocks = (int)((j.total + tc::kPackPerBlock - 1) / tc::kPackPerBlock);
    b0 += j.nblocks;
    unsigned char* slot = reinterpret_cast<unsigned char*>(host_jobs) + (size_t)i * HCU_BATCH_JOB_BYTES;
    memset(slot, 0, HCU_BATCH_JOB_BYTES);
    memcpy(slot, &j, sizeof(j));
  }
  *blocks = b0;
  return 0;
}

extern "C" int hcu_conv_tc_pack_batch(const void* dev_jobs, int32_t n, int32_t blocks, const float* params, void* packed,
                                      void* stream) {
  HCU_CHECK_ARG(dev_jobs && params && packed && n > 0 && blocks > 0, "conv_tc_pack_batch: bad arguments");
  tc::pack_tc_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)dev_jobs, n, params,
                                                                     (unsigned char*)packed);
  HCU_CHECK_LAUNCH("pack_tc_batch");
  return 0;
}

static int conv_tc_fwd_impl(const HcuConvDesc* d, const void* in, const void* packed, const float* bias,
                            const float* in_scale, const float* in_shift, const float* out_scale,
                            const float* out_shift, void* out, double* stats, const HcuBnFin* fin, void* stream,
                            const tc::Params::BnBwdFuse* bnb = nullptr, bool query_only = false) {
  HCU_CHECK_ARG(d && in && packed && out, "conv_tc_fwd: null pointer");
  HCU_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "conv_tc_fwd: in_scale/in_shift must come together");
  HCU_CHECK_ARG((out_scale == nullptr) == (out_shift == nullptr), "conv_tc_fwd: out_scale/out_shift must come together");
  tc::Params p;
  const char* why = tc::configure(d, p);
  if (why != nullptr) {
    set_error("conv_tc_fwd: unsupported descriptor (%s)", why);
    return HCU_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < 3; ++i)
    HCU_CHECK_ARG((long long)(d->out_size[i] - 1) * d->ostep[i] + d->ooff[i] + (((d->ophase >> (8 * i)) & 0xff) > 1 ? ((d->ophase >> (8 * i)) & 0xff) - 1 : 0) <
                      d->out_tsize[i],
                  "conv_tc_fwd: output grid exceeds output tensor in dim %d", i);
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("HCU_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  HCU_CHECK_ARG(d->out_c_off >= 0 && d->out_c_off + p.cpp <= d->out_cpitch, "conv_tc_fwd: output channel slice");
  HCU_CHECK_ARG((long long)d->in_size[1] * d->in_size[2] * d->in_cpitch < 0x7fffffffLL, "conv_tc_fwd: x-plane too large");
  p.in = (const __half*)in; p.wp = (const __half*)packed; p.out = out;
  p.bias = bias; p.in_scale = in_scale; p.in_shift = in_shift; p.out_scale = out_scale; p.out_shift = out_shift;
  p.stats = stats; p.stats_pitch = d->out_cpitch;
  memset(&p.fin, 0, sizeof(p.fin));
  if (fin != nullptr) {
    HCU_CHECK_ARG(stats && fin->gamma && fin->beta && fin->mean && fin->invstd && fin->scale && fin->shift && fin->counter &&
                      fin->count > 0 && d->out_c_off == 0 && d->out_cpitch == d->cout,
                  "conv_tc_fwd_bn: bad finalize arguments");
    p.fin = *fin;
  }
  const long long tz = d->out_cpitch, ty = tz * d->out_tsize[2], tx = ty * d->out_tsize[1], tn = tx * d->out_tsize[0];
  p.out_sn = tn; p.out_sx = tx * d->ostep[0]; p.out_sy = ty * d->ostep[1]; p.out_sz = tz * d->ostep[2];
  p.out_base = tx * d->ooff[0] + ty * d->ooff[1] + tz * d->ooff[2];
  p.out_c_off = d->out_c_off; p.out_f32 = d->dtype_out == HCU_F32; p.in_relu = d->in_relu; p.out_relu = d->out_relu;
  p.out_ph[0] = tx; p.out_ph[1] = ty; p.out_ph[2] = tz;
  HCU_CHECK_ARG(tx * d->ostep[0] < 0x7fffffffLL, "conv_tc_fwd: output x-plane too large");
  p.epi_fast = d->dtype_out == HCU_F16 && p.ops[0] * p.ops[1] * p.ops[2] == 1 && d->cout % 8 == 0 && d->out_cpitch % 8 == 0 &&
               d->out_c_off % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  HCU_CHECK_ARG(p.ops[0] * p.ops[1] * p.ops[2] == 1 || stats == nullptr, "conv_tc_fwd: no statistics with ophase");

  if (p.ks && bnb != nullptr) {
    if (!query_only) set_error("conv_tc_fwd_bnbwd: the K-streamed kernel has no fused statistics");
    return HCU_ERR_UNSUPPORTED;
  }
  if (p.ks) {
    HCU_CHECK_ARG(d->dtype_out == HCU_F16 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv_tc_fwd: K-streamed kernel needs a 16-byte aligned fp16 output");
    if (tc::ks_flat2d(d)) {  // (x, y) of the descriptor are the flat plane's (row, column); no march axis
      p.out_sx = 0; p.out_sy = tx * d->ostep[0]; p.out_sz = ty * d->ostep[1];
    }
    static int ks_attr = 0;
    if (!ks_attr) {
      cudaError_t e = cudaFuncSetAttribute(tc::conv_ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemLimit);
      if (e != cudaSuccess) {
        set_error("conv_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return HCU_ERR_CUDA;
      }
      ks_attr = 1;
    }
    const long long grid = (long long)p.OX * p.n_runs * p.nsplit;
    HCU_CHECK_ARG(grid <= 0x7fffffffLL, "conv_tc_fwd: grid too large");
    if (p.debug & 8)
      fprintf(stderr, "conv_ks: grid %lld smem %d tmem %d M %d Nc %d PC %d RA %d RB %d runs %d\n", grid, p.smem_bytes, p.tmem_cols, p.M,
              p.Nc, p.PC, p.RA, p.RB, p.n_runs);
    launch_pdl(1, tc::conv_ks_kernel, dim3((unsigned)grid), dim3(tc::kKsThreads), (size_t)p.smem_bytes, (cudaStream_t)stream, p);
    HCU_CHECK_LAUNCH("conv_ks");
    return 0;
  }
  // ---- kernel variant (see Var): a specialised instantiation when the launch matches one, else the generic kernel ----
  using KernelFn = tc::ConvTcFn;
  struct VariantSlot { KernelFn fn; int regs; bool ready; };
  static VariantSlot variants[1 + 4 * 2 * 4 * 2] = {};
  static bool variants_init = false;
  if (!variants_init) {
    variants[0].fn = tc::conv_tc_kernel<tc::VarGeneric>;
    for (int epi = 1; epi <= 2; ++epi)
      for (int fi = 0; fi < 4; ++fi)
        for (int bulk = 0; bulk < 2; ++bulk) {
          variants[1 + ((0 * 2 + (epi - 1)) * 4 + fi) * 2 + bulk].fn = tc::conv_tc_variant_mb1(epi, fi, bulk);
          variants[1 + ((1 * 2 + (epi - 1)) * 4 + fi) * 2 + bulk].fn = tc::conv_tc_variant_mb2(epi, fi, bulk);
          variants[1 + ((2 * 2 + (epi - 1)) * 4 + fi) * 2 + bulk].fn = tc::conv_tc_variant_mb3(epi, fi, bulk);
          variants[1 + ((3 * 2 + (epi - 1)) * 4 + fi) * 2 + bulk].fn = tc::conv_tc_variant_mb4(epi, fi, bulk);
        }
    variants_init = true;
  }
  int vi = 0;
  {
    static int spec_on = -1;
    if (spec_on < 0) { const char* e = getenv("HCU_TC_SPEC"); spec_on = e ? atoi(e) : 1; }
    const int nch = d->cout;
    int fi = -1;
    if (bias != nullptr && stats != nullptr && out_scale == nullptr && !d->out_relu) fi = 0;         // training forward
    else if (bias == nullptr && stats == nullptr && out_scale == nullptr && !d->out_relu) fi = 1;    // data gradient
    else if (bias == nullptr && stats == nullptr && out_scale != nullptr && d->out_relu) fi = 2;     // inference (BN folded)
    if (bnb != nullptr) fi = fi == 1 ? 3 : -1;   // data gradient + fused BatchNorm-backward statistics of the previous layer
    if (spec_on && p.debug == 0 && p.wide && p.Nc == 16 && p.nsplit == 1 && p.epi_fast && (nch == 8 || nch == 16) && fi >= 0 &&
        p.MB >= 1 && p.MB <= 4)
    {
      // bulk-copy producer: 8-channel pixels that need no transform, rows of the x-plane contiguous in memory
      static int bulk_on = -1;
      if (bulk_on < 0) { const char* e = getenv("HCU_TC_BULK"); bulk_on = e ? atoi(e) : 1; }
      const bool bulk = bulk_on && p.P == 1 && in_scale == nullptr && p.ips[0] * p.ips[1] * p.ips[2] == 1 && p.in_zs == 8 &&
                        p.in_ys == p.IZ * p.in_zs && p.RUN * 16 <= 32768 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
      vi = 1 + (((p.MB - 1) * 2 + (nch == 8 ? 0 : 1)) * 4 + fi) * 2 + (bulk ? 1 : 0);
      // rotating accumulator window: KX slots of Nc columns per M-block
      int cols = p.MB * p.KX * p.Nc, t = 32;
      while (t < cols) t <<= 1;
      p.tmem_cols = t;
    }
  }
  {
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("HCU_TC_TRACE"); trace = e ? atoi(e) : 0; }
    if (trace)
      fprintf(stderr, "conv_tc: variant %d (MB %d Nc %d cout %d wide %d nsplit %d epi_fast %d P %d bias %d stats %d affine %d relu %d xf %d)\n", vi,
              p.MB, p.Nc, d->cout, p.wide, p.nsplit, p.epi_fast, p.P, bias != nullptr, stats != nullptr, out_scale != nullptr,
              d->out_relu, in_scale != nullptr);
  }
  if (bnb != nullptr) {
    // only the specialised variants carry the fused statistics; the output must be the dense [pixels][cout] tensor y has
    // The generic kernel's vector epilogue can do it too (32 / 64-channel levels, every output channel in the CTA), but there
    // the y loads are not prefetched and the sums need a shuffle reduction per 16-channel chunk and plane: measured SLOWER than
    // the separate pass (d2.conv2.dgrad 37 -> 63 us, d3.conv2.dgrad 28 -> 60 us for 28 + 19 us of statistics saved; bench step
    // 2.86 -> 2.93 ms).  Opt-in only (HCU_TC_BNB_GENERIC=1).
    static int bnb_generic = -1;
    if (bnb_generic < 0) { const char* e = getenv("HCU_TC_BNB_GENERIC"); bnb_generic = e ? atoi(e) : 0; }
    const bool generic_ok = bnb_generic && vi == 0 && p.debug == 0 && p.nsplit == 1 && p.epi_fast && d->cout % 16 == 0 &&
                            d->cout == p.Nc && bias == nullptr && stats == nullptr && out_scale == nullptr && !d->out_relu;
    if ((vi == 0 && !generic_ok) || d->out_cpitch != d->cout || d->out_c_off != 0 || p.ks) {
      if (!query_only) set_error("conv_tc_fwd_bnbwd: no specialised variant takes this descriptor");
      return HCU_ERR_UNSUPPORTED;
    }
    p.bnb = *bnb;
  } else {
    memset(&p.bnb, 0, sizeof(p.bnb));
  }
  if (query_only) return 0;
  VariantSlot& var = variants[vi];
  if (!var.ready) {
    cudaError_t e = cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemLimit);
    if (e != cudaSuccess) {
      set_error("conv_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return HCU_ERR_CUDA;
    }
    cudaFuncAttributes fa;
    var.regs = cudaFuncGetAttributes(&fa, var.fn) == cudaSuccess ? fa.numRegs : 128;
    var.ready = true;
  }
  const int regs_per_thread = var.regs;
  // x segmentation: the grid should fill whole waves of the CTAs that are really co-resident (registers, shared
  // memory, TMEM columns), with segments no shorter than ~6 planes (each segment re-reads KX-1 planes)
  const int warps = tc::kThreads / 32;
  int per_sm = 65536 / (((regs_per_thread * 32 + 255) / 256 * 256) * warps);   // register file
  per_sm = std::min(per_sm, 233472 / (p.smem_bytes + 1024));                    // shared memory
  per_sm = std::min(per_sm, 2048 / tc::kThreads);                               // threads
  per_sm = std::max(1, std::min(per_sm, 512 / p.tmem_cols));                    // TMEM columns
  if (p.debug & 8) fprintf(stderr, "conv_tc: per_sm %d smem %d tmem %d M %d R %d Nc %d\n", per_sm, p.smem_bytes, p.tmem_cols, p.M, p.R, p.Nc);
  const long long slots = (long long)per_sm * num_sms();
  const long long base_items = (long long)p.N * p.n_runs * p.nsplit;
  const int max_seg = p.OX;
  int best_seg = 1;
  double best_cost = 1e30;
  for (int nseg = 1; nseg <= max_seg; ++nseg) {
    const int lx = (p.OX + nseg - 1) / nseg;
    const int segs = (p.OX + lx - 1) / lx;
    const long long items = base_items * segs;
    const long long waves = (items + slots - 1) / slots;
    // time ~ waves * (planes per item incl. the KX-1 warm-up planes + fixed per-CTA overhead of ~3 planes)
    const double cost = (double)waves * (lx + (p.KX - 1) * p.dx + 3);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_seg = nseg; }
  }
  p.Lx = (p.OX + best_seg - 1) / best_seg;
  p.n_xseg = (p.OX + p.Lx - 1) / p.Lx;
  const long long grid = base_items * p.n_xseg;
  HCU_CHECK_ARG(grid <= 0x7fffffffLL, "conv_tc_fwd: grid too large");
  launch_pdl(1, var.fn, dim3((unsigned)grid), dim3(tc::kThreads), (size_t)p.smem_bytes, (cudaStream_t)stream, p);
  HCU_CHECK_LAUNCH("conv_tc");
  return 0;
}

extern "C" int hcu_conv_tc_fwd(const HcuConvDesc* d, const void* in, const void* packed, const float* bias,
                               const float* in_scale, const float* in_shift, const float* out_scale,
                               const float* out_shift, void* out, double* stats, void* stream) {
  return conv_tc_fwd_impl(d, in, packed, bias, in_scale, in_shift, out_scale, out_shift, out, stats, nullptr, stream);
}

extern "C" int hcu_conv_tc_fwd_bn(const HcuConvDesc* d, const void* in, const void* packed, const float* bias,
                                  const float* in_scale, const float* in_shift, const float* out_scale,
                                  const float* out_shift, void* out, double* stats, const HcuBnFin* fin, void* stream) {
  return conv_tc_fwd_impl(d, in, packed, bias, in_scale, in_shift, out_scale, out_shift, out, stats, fin, stream);
}

extern "C" int hcu_conv_tc_fwd_bnbwd(const HcuConvDesc* d, const void* in, const void* packed, void* out, const void* y,
                                     const float* scale, const float* shift, const float* mean, const float* invstd,
                                     double* sums, const HcuBnBwdFin* fin, void* stream) {
  HCU_CHECK_ARG(y && scale && shift && mean && invstd && sums && fin && fin->gamma && fin->coef && fin->count > 0,
                "conv_tc_fwd_bnbwd: null pointer");
  tc::Params::BnBwdFuse f;
  f.y = (const __half*)y; f.scale = scale; f.shift = shift; f.mean = mean; f.invstd = invstd; f.sums = sums; f.fin = *fin;
  return conv_tc_fwd_impl(d, in, packed, nullptr, nullptr, nullptr, nullptr, nullptr, out, nullptr, nullptr, stream, &f);
}

extern "C" int hcu_conv_tc_bnbwd_supported(const HcuConvDesc* d) {
  if (d == nullptr) return 0;
  tc::Params::BnBwdFuse f;
  memset(&f, 0, sizeof(f));
  alignas(16) static char dummy[16];
  return conv_tc_fwd_impl(d, dummy, dummy, nullptr, nullptr, nullptr, nullptr, nullptr, dummy, nullptr, nullptr, nullptr, &f,
                          true) == 0 ? 1 : 0;
}
