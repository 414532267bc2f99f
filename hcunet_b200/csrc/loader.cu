// Input path on the device: the deterministic part of the reference dataloader,
//   Stack.__getitem__ (dataloader.py:68-92) -> to_float (transforms.py:94-113) -> reshape (transforms.py:139-157)
//   -> normalize (transforms.py:257-282) -> to_tensor (transforms.py:118-136),
// applied to the RAW stack as skimage.io.imread yields it ([Z][Y][X][C] uint8 / uint16), producing the fp16
// channels-last [X][Y][Z][8] tensor the first convolution consumes.  The reference does this with numpy on the host in
// float64 and ships an fp16 [1][C][X][Y][Z] tensor; here the raw bytes are what crosses PCIe (half / a quarter of the fp16
// bytes) and one transposing pass through shared memory replaces to_float + swapaxes + normalize + the NCDHW -> NDHWC pass.
// Arithmetic is the reference's, bit for bit: v / 2^bits, += -mean[c], /= std[c] in float64, then torch's double -> half
// conversion (which goes through float: half(float(d))).
//
// hcu_load_labels does the same for mask / pwl ([Z][Y][X] uint8 / uint16 with to_float's scaling, or floating point as is),
// but only for the origin crop the loss will read (loss.py:51-56 crops them to the prediction's shape from the origin;
// nothing else of them is ever used) and straight into the [B][1][x][y][z] order the loss kernels index.  The source may be
// pinned host memory (UVA): the crop is read where it lies, never copied whole.
#include "common.cuh"

namespace hcu {
namespace ld {

constexpr int kTile = 32;
constexpr int kYG = 8;   // y rows per CTA: the per-CTA lookup table (1024 float64 conversions) cost a quarter of a one-row CTA's work

struct StackParams {
  const void* src;
  __half* dst;
  int B, Z, Y, X, C, Cp;
  int nzt;
  double neg_mean[8], std_[8];
};

template <typename TS>
__device__ __forceinline__ __half convert(TS v, double inv_range, double neg_mean, double sd) {
  double d = (double)v * inv_range;   // exact: v / 2^bits
  d += neg_mean;
  d /= sd;
  return __float2half_rn((float)d);   // torch.as_tensor(float64 array, dtype=torch.half): double -> float -> half
}

// block = (32 x-values) x (32 z-values) of one (image, y): coalesced reads along x (source order), coalesced writes along z
template <typename TS, bool U8LUT>
__global__ void __launch_bounds__(256) load_stack_kernel(const StackParams p) {
  __shared__ uint4 tile[kTile][kTile + 1];        // [z][x] -> 8 halfs
  __shared__ __half lut[U8LUT ? 8 * 256 : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kTile, y_lo = blockIdx.y * kYG, y_hi = min(p.Y, y_lo + kYG);
  const int b = blockIdx.z / p.nzt, z0 = (blockIdx.z - b * p.nzt) * kTile;
  const double inv_range = sizeof(TS) == 1 ? 1.0 / 256.0 : 1.0 / 65536.0;
  if (U8LUT) {
    for (int i = threadIdx.x; i < p.C * 256; i += 256) {
      const int c = i >> 8;
      lut[i] = convert<int>(i & 255, inv_range, p.neg_mean[c], p.std_[c]);
    }
    __syncthreads();
  }
  const TS* src = reinterpret_cast<const TS*>(p.src);
  const int x = x0 + lane;
  for (int y = y_lo; y < y_hi; ++y) {
  for (int zz = warp; zz < kTile; zz += 8) {
    const int z = z0 + zz;
    __align__(16) __half h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = __float2half_rn(0.f);
    if (z < p.Z && x < p.X) {
      const TS* px = src + ((((size_t)b * p.Z + z) * p.Y + y) * p.X + x) * p.C;
      if (p.C == 4) {   // one aligned 4- / 8-byte load per voxel
        TS v[4];
        if (sizeof(TS) == 1) *reinterpret_cast<uint32_t*>(v) = *reinterpret_cast<const uint32_t*>(px);
        else *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(px);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          h[c] = U8LUT ? lut[c * 256 + (int)v[c]] : convert<TS>(v[c], inv_range, p.neg_mean[c], p.std_[c]);
      } else {
        for (int c = 0; c < p.C; ++c)
          h[c] = U8LUT ? lut[c * 256 + (int)px[c]] : convert<TS>(px[c], inv_range, p.neg_mean[c], p.std_[c]);
      }
    }
    tile[zz][lane] = *reinterpret_cast<const uint4*>(h);
  }
  __syncthreads();
  const int z = z0 + lane;
  for (int xx = warp; xx < kTile; xx += 8) {
    const int xo = x0 + xx;
    if (xo < p.X && z < p.Z) {
      __half* o = p.dst + ((((size_t)b * p.X + xo) * p.Y + y) * p.Z + z) * p.Cp;
      *reinterpret_cast<uint4*>(o) = tile[lane][xx];
    }
  }
  __syncthreads();   // the tile is re-used by the next row
  }
}

struct LabelParams {
  const void* src;
  __half* dst;
  long long total;
  int B, Z, Y, X, ox, oy, oz;
};

template <typename TS> __device__ __forceinline__ __half label_value(TS v);
template <> __device__ __forceinline__ __half label_value<uint8_t>(uint8_t v) { return __float2half_rn((float)v * (1.f / 256.f)); }
template <> __device__ __forceinline__ __half label_value<uint16_t>(uint16_t v) { return __float2half_rn((float)v * (1.f / 65536.f)); }
template <> __device__ __forceinline__ __half label_value<__half>(__half v) { return v; }
template <> __device__ __forceinline__ __half label_value<float>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __half label_value<double>(double v) { return __float2half_rn((float)v); }

// one thread per voxel of the crop, in SOURCE order (x fastest): the (possibly host-resident) source is read in contiguous
// rows of ox elements, the small device-side result is written scattered
template <typename TS>
__global__ void __launch_bounds__(256) load_labels_kernel(const LabelParams p) {
  const TS* src = reinterpret_cast<const TS*>(p.src);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < p.total; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % p.ox);
    long long r = e / p.ox;
    const int y = (int)(r % p.oy); r /= p.oy;
    const int z = (int)(r % p.oz);
    const long long b = r / p.oz;
    const TS v = src[((b * p.Z + z) * p.Y + y) * p.X + x];
    p.dst[((b * p.ox + x) * p.oy + y) * (long long)p.oz + z] = label_value<TS>(v);
  }
}

}  // namespace ld
}  // namespace hcu

using namespace hcu;

extern "C" int hcu_load_stack(const void* src, int32_t dtype_src, int64_t b, int32_t z, int32_t y, int32_t x, int32_t c,
                              const double* mean, const double* stdv, void* dst, int32_t dst_cpitch, void* stream) {
  HCU_CHECK_ARG(src && dst && mean && stdv, "load_stack: null pointer");
  HCU_CHECK_ARG(dtype_src == HCU_U8 || dtype_src == HCU_U16, "load_stack: the raw stack must be uint8 or uint16 (to_float, transforms.py:104-112)");
  HCU_CHECK_ARG(b > 0 && z > 0 && y > 0 && x > 0 && c > 0 && c <= 8, "load_stack: bad shape (1..8 channels)");
  HCU_CHECK_ARG(dst_cpitch == 8, "load_stack: the destination channel pitch is 8 (16-byte voxels)");
  HCU_CHECK_ARG((y + ld::kYG - 1) / ld::kYG <= 65535 && b * ((z + 31) / 32) <= 65535, "load_stack: grid too large");
  HCU_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0, "load_stack: unaligned pointer");
  ld::StackParams p;
  memset(&p, 0, sizeof(p));
  p.src = src; p.dst = (__half*)dst; p.B = (int)b; p.Z = z; p.Y = y; p.X = x; p.C = c; p.Cp = dst_cpitch;
  p.nzt = (z + ld::kTile - 1) / ld::kTile;
  for (int i = 0; i < c; ++i) {
    HCU_CHECK_ARG(stdv[i] != 0.0, "load_stack: std[%d] == 0", i);
    p.neg_mean[i] = -mean[i]; p.std_[i] = stdv[i];
  }
  dim3 grid((x + ld::kTile - 1) / ld::kTile, (y + ld::kYG - 1) / ld::kYG, (unsigned)(b * p.nzt));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_src == HCU_U8) ld::load_stack_kernel<uint8_t, true><<<grid, 256, 0, st>>>(p);
  else ld::load_stack_kernel<uint16_t, false><<<grid, 256, 0, st>>>(p);
  HCU_CHECK_LAUNCH("load_stack");
  return 0;
}

extern "C" int hcu_load_labels(const void* src, int32_t dtype_src, int64_t b, int32_t z, int32_t y, int32_t x, int32_t ox,
                               int32_t oy, int32_t oz, void* dst, void* stream) {
  HCU_CHECK_ARG(src && dst, "load_labels: null pointer");
  HCU_CHECK_ARG(b > 0 && ox > 0 && oy > 0 && oz > 0 && ox <= x && oy <= y && oz <= z, "load_labels: the crop must lie inside the source");
  ld::LabelParams p;
  p.src = src; p.dst = (__half*)dst; p.B = (int)b; p.Z = z; p.Y = y; p.X = x; p.ox = ox; p.oy = oy; p.oz = oz;
  p.total = (long long)b * ox * oy * oz;
  const int grid = (int)std::min<long long>((p.total + 255) / 256, (long long)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype_src) {
    case HCU_U8: ld::load_labels_kernel<uint8_t><<<grid, 256, 0, st>>>(p); break;
    case HCU_U16: ld::load_labels_kernel<uint16_t><<<grid, 256, 0, st>>>(p); break;
    case HCU_F16: ld::load_labels_kernel<__half><<<grid, 256, 0, st>>>(p); break;
    case HCU_F32: ld::load_labels_kernel<float><<<grid, 256, 0, st>>>(p); break;
    case HCU_F64: ld::load_labels_kernel<double><<<grid, 256, 0, st>>>(p); break;
    default: set_error("load_labels: unsupported source dtype %d", (int)dtype_src); return HCU_ERR_INVALID;
  }
  HCU_CHECK_LAUNCH("load_labels");
  return 0;
}
