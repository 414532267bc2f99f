// Overlap-tile inference driver on the device: the data movement of the reference's
// `predict_segmentation_mask` (hcat/segment.py:21-136) around the network call, as three memory-bound kernels.
//
//   hcu_tile_gather   replaces  NaN / inf scrub (segment.py:66-67)  +  pad_image_with_reflections (utils.py:33-74; numpy
//                     flips + three torch.cat of the whole stack)  +  the tile slice and `.float().to(device)`
//                     (segment.py:86)  +  the NCDHW -> channels-last layout pass of the engine: ONE pass from the ORIGINAL
//                     stack to the dense tile the first convolution reads.  The padded stack never exists.
//   hcu_tile_flags    the "everything is -1, skip for speed" test (segment.py:89-93) of a tile, on the device.
//   hcu_sigmoid_paste replaces  the centre crop (segment.py:99-103), the in-place sigmoid (segment.py:107-110:
//                     mul_(-1).exp_().add_(1).pow_(-1)), the threshold + uint8 cast (segment.py:113-117) and the paste
//                     into the full mask (segment.py:121-123): one pass from the logits to the mask volume.
//
// Reflection rule of the reference (utils.py:48-52, per dimension with pad p and extent n): padded index u < p reads
// original p-1-u; u >= p+n reads original n-1-(u-p-n); else u-p.  (The first / last p voxels mirrored INCLUDING the edge.)
#include "common.cuh"

namespace hcu {
namespace tiler {

struct Geom {
  int C;
  int size[3];     // original stack extent
  int pad[3];      // reflection padding per side
  int origin[3];   // tile origin in padded coordinates
  int extent[3];   // tile extent
  int sorg[3];     // original coordinates of the first voxel of the resident (sub-)stack
  int ssize[3];    // extent of the resident (sub-)stack
};

__device__ __forceinline__ int reflect(int u, int p, int n) {
  if (u < p) return p - 1 - u;
  if (u >= p + n) return n - 1 - (u - p - n);
  return u - p;
}

template <typename T>
__device__ __forceinline__ float load_scrubbed(const T* p, bool scrub = true) {
  const float v = to_f(*p);
  if (!scrub) return v;
  if (v != v) return 0.f;                 // image[np.isnan(image)] = 0
  if (fabsf(v) > 3.4028234e38f) return 1.f;  // image[np.isinf(image)] = 1 (either sign)
  return v;
}

// one thread per tile voxel, z fastest (source order); layout 1: channels-last, one 16-byte store per voxel when cpitch == 8
template <typename TS, typename TD, int LAYOUT>
__global__ void __launch_bounds__(256) tile_gather_kernel(const Geom g, const TS* __restrict__ src, TD* __restrict__ dst, int cpitch,
                                                          bool scrub) {
  const long long nvox = (long long)g.extent[0] * g.extent[1] * g.extent[2];
  const long long splane = (long long)g.ssize[0] * g.ssize[1] * g.ssize[2];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nvox; e += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(e % g.extent[2]);
    const long long r = e / g.extent[2];
    const int y = (int)(r % g.extent[1]), x = (int)(r / g.extent[1]);
    const int ox = reflect(g.origin[0] + x, g.pad[0], g.size[0]) - g.sorg[0];
    const int oy = reflect(g.origin[1] + y, g.pad[1], g.size[1]) - g.sorg[1];
    const int oz = reflect(g.origin[2] + z, g.pad[2], g.size[2]) - g.sorg[2];
    const long long so = ((long long)ox * g.ssize[1] + oy) * g.ssize[2] + oz;
    if (LAYOUT == 1) {
      if (sizeof(TD) == 2 && cpitch == 8) {
        __align__(16) __half h[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) h[c] = c < g.C ? __float2half_rn(load_scrubbed(src + c * splane + so, scrub)) : __float2half_rn(0.f);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(dst) + e * 8) = *reinterpret_cast<const uint4*>(h);
      } else {
        for (int c = 0; c < cpitch; ++c) dst[e * cpitch + c] = from_f<TD>(c < g.C ? load_scrubbed(src + c * splane + so, scrub) : 0.f);
      }
    } else {
      for (int c = 0; c < g.C; ++c) dst[c * nvox + e] = from_f<TD>(load_scrubbed(src + c * splane + so, scrub));
    }
  }
}

// flag[0] += number of (scrubbed) tile values that are not exactly -1 (saturating at 2^30): 0 <=> the reference skips the tile
template <typename TS>
__global__ void __launch_bounds__(256) tile_flags_kernel(const Geom g, const TS* __restrict__ src, unsigned int* __restrict__ flag) {
  const long long nvox = (long long)g.extent[0] * g.extent[1] * g.extent[2];
  const long long splane = (long long)g.ssize[0] * g.ssize[1] * g.ssize[2];
  unsigned int cnt = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nvox; e += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(e % g.extent[2]);
    const long long r = e / g.extent[2];
    const int y = (int)(r % g.extent[1]), x = (int)(r / g.extent[1]);
    const int ox = reflect(g.origin[0] + x, g.pad[0], g.size[0]) - g.sorg[0];
    const int oy = reflect(g.origin[1] + y, g.pad[1], g.size[1]) - g.sorg[1];
    const int oz = reflect(g.origin[2] + z, g.pad[2], g.size[2]) - g.sorg[2];
    const long long so = ((long long)ox * g.ssize[1] + oy) * g.ssize[2] + oz;
    for (int c = 0; c < g.C; ++c) cnt += load_scrubbed(src + c * splane + so) != -1.f ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt != 0u) {
    const unsigned int old = atomicAdd(flag, min(cnt, 1u << 20));
    if (old > (1u << 30)) atomicExch(flag, 1u << 30);
  }
}

struct Paste {
  int lsize[3];    // logits extent (one channel)
  int crop[3];     // first logit voxel taken
  int ext[3];      // extent pasted
  int msize[3];    // mask volume extent
  int morg[3];     // where the block goes
  float threshold;
  int probability;  // 1: float probabilities, 0: uint8 (p > threshold)
};

__global__ void __launch_bounds__(256) sigmoid_paste_kernel(const Paste q, const float* __restrict__ logits, void* __restrict__ mask,
                                                            uint8_t* __restrict__ written) {
  const long long n = (long long)q.ext[0] * q.ext[1] * q.ext[2];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(e % q.ext[2]);
    const long long r = e / q.ext[2];
    const int y = (int)(r % q.ext[1]), x = (int)(r / q.ext[1]);
    const float v = logits[((long long)(q.crop[0] + x) * q.lsize[1] + (q.crop[1] + y)) * q.lsize[2] + (q.crop[2] + z)];
    // the reference's in-place chain in fp32: x * -1, exp, + 1, ^ -1
    const float pr = 1.f / (expf(-v) + 1.f);
    const long long mo = ((long long)(q.morg[0] + x) * q.msize[1] + (q.morg[1] + y)) * q.msize[2] + (q.morg[2] + z);
    if (q.probability) reinterpret_cast<float*>(mask)[mo] = pr;
    else reinterpret_cast<uint8_t*>(mask)[mo] = pr > q.threshold ? 1 : 0;
    if (written != nullptr) written[mo] = 1;
  }
}

static inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)std::max(1LL, std::min(b, cap));
}

}  // namespace tiler
}  // namespace hcu

using namespace hcu;

static int check_geom(const HcuTileGeom* g, const char* who) {
  HCU_CHECK_ARG(g != nullptr && g->channels > 0 && g->channels <= 64, "%s: bad channel count", who);
  for (int i = 0; i < 3; ++i) {
    HCU_CHECK_ARG(g->size[i] > 0 && g->pad[i] >= 0 && g->pad[i] <= g->size[i] && g->extent[i] > 0 && g->origin[i] >= 0 &&
                      g->origin[i] + g->extent[i] <= g->size[i] + 2 * g->pad[i],
                  "%s: tile [%d, %d) outside the padded extent %d (+ 2 x %d) in dim %d", who, g->origin[i],
                  g->origin[i] + g->extent[i], g->size[i], g->pad[i], i);
    HCU_CHECK_ARG(g->stack_size[i] > 0 && g->stack_origin[i] >= 0 && g->stack_origin[i] + g->stack_size[i] <= g->size[i],
                  "%s: resident sub-stack outside the stack in dim %d", who, i);
    // every original index the tile touches must be resident
    int lo = g->size[i], hi = -1;
    const int ends[2] = {g->origin[i], g->origin[i] + g->extent[i] - 1};
    for (int k = 0; k < 2; ++k) {
      const int u = ends[k];
      const int o = u < g->pad[i] ? g->pad[i] - 1 - u : (u >= g->pad[i] + g->size[i] ? g->size[i] - 1 - (u - g->pad[i] - g->size[i]) : u - g->pad[i]);
      lo = std::min(lo, o); hi = std::max(hi, o);
    }
    if (g->origin[i] < g->pad[i] && g->origin[i] + g->extent[i] > g->pad[i]) lo = 0;                      // crosses the low edge
    if (g->origin[i] < g->pad[i] + g->size[i] && g->origin[i] + g->extent[i] > g->pad[i] + g->size[i]) hi = g->size[i] - 1;
    HCU_CHECK_ARG(lo >= g->stack_origin[i] && hi < g->stack_origin[i] + g->stack_size[i],
                  "%s: the tile reads original voxels [%d, %d] of dim %d, resident are [%d, %d)", who, lo, hi, i, g->stack_origin[i],
                  g->stack_origin[i] + g->stack_size[i]);
  }
  return 0;
}

static tiler::Geom to_geom(const HcuTileGeom* g) {
  tiler::Geom o;
  o.C = g->channels;
  for (int i = 0; i < 3; ++i) {
    o.size[i] = g->size[i]; o.pad[i] = g->pad[i]; o.origin[i] = g->origin[i]; o.extent[i] = g->extent[i];
    o.sorg[i] = g->stack_origin[i]; o.ssize[i] = g->stack_size[i];
  }
  return o;
}

extern "C" int hcu_tile_gather(const HcuTileGeom* g, const void* stack, int32_t dtype_stack, void* tile, int32_t dtype_tile,
                               int32_t layout, int32_t cpitch, void* stream) {
  HCU_CHECK_ARG(stack && tile, "tile_gather: null pointer");
  if (int rc = check_geom(g, "tile_gather")) return rc;
  const bool scrub = (layout & 2) == 0;
  layout &= 1;
  HCU_CHECK_ARG(layout == 0 || cpitch >= g->channels, "tile_gather: channels-last layout needs pitch >= C");
  const tiler::Geom q = to_geom(g);
  const long long nvox = (long long)q.extent[0] * q.extent[1] * q.extent[2];
  const int grid = tiler::grid_for(nvox);
  cudaStream_t st = (cudaStream_t)stream;
  HCU_DISPATCH_ACT(dtype_stack, TS, HCU_DISPATCH_ACT(dtype_tile, TD, {
    if (layout == 1) tiler::tile_gather_kernel<TS, TD, 1><<<grid, 256, 0, st>>>(q, (const TS*)stack, (TD*)tile, cpitch, scrub);
    else tiler::tile_gather_kernel<TS, TD, 0><<<grid, 256, 0, st>>>(q, (const TS*)stack, (TD*)tile, cpitch, scrub);
  }));
  HCU_CHECK_LAUNCH("tile_gather");
  return 0;
}

extern "C" int hcu_tile_flags(const HcuTileGeom* g, const void* stack, int32_t dtype_stack, uint32_t* flag, void* stream) {
  HCU_CHECK_ARG(stack && flag, "tile_flags: null pointer");
  if (int rc = check_geom(g, "tile_flags")) return rc;
  const tiler::Geom q = to_geom(g);
  const long long nvox = (long long)q.extent[0] * q.extent[1] * q.extent[2];
  const int grid = tiler::grid_for(nvox);
  HCU_DISPATCH_ACT(dtype_stack, TS, tiler::tile_flags_kernel<TS><<<grid, 256, 0, (cudaStream_t)stream>>>(q, (const TS*)stack, flag));
  HCU_CHECK_LAUNCH("tile_flags");
  return 0;
}

extern "C" int hcu_sigmoid_paste(const float* logits, const int32_t* lsize, const int32_t* crop, const int32_t* ext, void* mask,
                                 int32_t probability, const int32_t* msize, const int32_t* morigin, float threshold,
                                 uint8_t* written, void* stream) {
  HCU_CHECK_ARG(logits && lsize && crop && ext && mask && msize && morigin, "sigmoid_paste: null pointer");
  tiler::Paste q;
  for (int i = 0; i < 3; ++i) {
    HCU_CHECK_ARG(ext[i] > 0 && crop[i] >= 0 && crop[i] + ext[i] <= lsize[i] && morigin[i] >= 0 && morigin[i] + ext[i] <= msize[i],
                  "sigmoid_paste: block does not fit (dim %d: crop %d + %d of %d logits, origin %d + %d of %d mask voxels)", i, crop[i],
                  ext[i], lsize[i], morigin[i], ext[i], msize[i]);
    q.lsize[i] = lsize[i]; q.crop[i] = crop[i]; q.ext[i] = ext[i]; q.msize[i] = msize[i]; q.morg[i] = morigin[i];
  }
  q.threshold = threshold;
  q.probability = probability;
  const long long n = (long long)ext[0] * ext[1] * ext[2];
  tiler::sigmoid_paste_kernel<<<tiler::grid_for(n), 256, 0, (cudaStream_t)stream>>>(q, logits, mask, written);
  HCU_CHECK_LAUNCH("sigmoid_paste");
  return 0;
}
