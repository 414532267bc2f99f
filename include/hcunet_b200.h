/*
 * hcunet_b200.h -- C ABI of libhcunet_b200.so (hand-written sm_100a kernels for the HcUnet hot path).
 *
 * The reference (wisamreid/HcUnet) is pure Python with NO FFI: every arithmetic call on its hot
 * path is a torch.nn module (hcat/unet.py:49-51) or torch.nn.functional (hcat/loss.py:65).  The
 * entry points below are therefore what a maintainer binds *instead of* those ATen/cuDNN calls;
 * each one names the reference call site it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All pointers are DEVICE pointers unless
 *     the name says host.  The caller owns every buffer (PyTorch caching allocator); the library
 *     allocates nothing persistent.
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately: no implicit synchronisation, no default-stream use, CUDA-graph capturable.
 *   - return value: 0 on success, negative HcuStatus on failure; hcu_last_error() gives a
 *     thread-local message.  Nothing throws, nothing exits.
 *   - activations are CHANNELS-LAST: [N][X][Y][Z][C] (2D: Z == 1), element type HcuDType.
 *     The reference layout [N][C][X][Y][Z] only exists at the boundary (hcu_nc_to_cl /
 *     hcu_cl_to_nc).  Parameters stay in the reference (PyTorch) layout in fp32 and are
 *     gathered into GEMM-B layouts by hcu_weight_gather.
 *   - the tensor-core conv kernels are launched with PROGRAMMATIC DEPENDENT LAUNCH (cudaLaunchKernelEx,
 *     programmaticStreamSerialization): their prologue may overlap the previous kernel of the stream, they touch global
 *     memory only after griddepcontrol.wait.  A caller needs to do nothing: stream order is preserved, and the launches are
 *     capturable (programmatic edges).  HCU_PDL=0 in the environment launches them normally.
 */
#ifndef HCUNET_B200_H
#define HCUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCU_ABI_VERSION 24

/* Per-channel reduction buffers (`stats` of the convolutions, `sums` of hcu_bn_bwd_stats) are fp64 and BINNED:
 * HCU_STAT_BINS consecutive [2][c] blocks; a producing CTA adds its partial sums into ONE bin (same-address L2 atomics
 * serialise: thousands of CTAs on 2c addresses cost more than the memory pass itself), the finalize kernels add the
 * bins in a fixed order.  The caller zeroes the whole [HCU_STAT_BINS][2][c] buffer. */
#define HCU_STAT_BINS 4

typedef enum HcuStatus {
  HCU_OK = 0,
  HCU_ERR_INVALID = -1,   /* bad descriptor / unsupported combination */
  HCU_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has the CUDA string) */
  HCU_ERR_UNSUPPORTED = -3
} HcuStatus;

/* activations / gradients are HCU_F32 or HCU_F16; HCU_BF16 is accepted for loss masks / weights only */
/* HCU_U8 / HCU_U16 / HCU_F64: source types of the input path only (hcu_load_stack / hcu_load_labels) */
typedef enum HcuDType { HCU_F32 = 0, HCU_BF16 = 1, HCU_F16 = 2, HCU_U8 = 3, HCU_U16 = 4, HCU_F64 = 5 } HcuDType;

/* ---- library ------------------------------------------------------------------------------ */
int hcu_abi_version(void);
const char* hcu_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long hcu_launch_count(void);
int hcu_zero(void* ptr, size_t bytes, void* stream);
/* Strided host -> device copy of one overlap tile: a gather kernel reading the pinned (UVA-mapped) host memory in 16-byte
 * units when everything is 16-byte aligned, else `planes` cudaMemcpy2DAsync calls; enqueued on `stream`.  Plane p
 * is `rows` rows of `row_bytes` contiguous bytes, `src_row_pitch` bytes apart in (pinned) host memory, packed densely on
 * the device; planes are `src_plane_pitch` bytes apart in the source.  For an NCDHW stack tile [C][nx][ny][Z] cut out of
 * [C][X][Y][Z]: planes = C, rows = nx, row_bytes = ny*Z*esz, src_row_pitch = Y*Z*esz, src_plane_pitch = X*Y*Z*esz.
 * Replaces the tile slice + `.to(device)` of the reference's tiler (segment.py:91), which materialises a contiguous copy
 * of the tile on the host first. */
int hcu_h2d_tile(const void* src, int64_t planes, int64_t src_plane_pitch, int64_t rows, int64_t src_row_pitch,
                 int64_t row_bytes, void* dst, void* stream);

/* ---- overlap-tile inference driver (hcat/segment.py:21-136 `predict_segmentation_mask`) ---------------
 * The reference pads the WHOLE stack with reflections on the host (utils.py:33-74: numpy flips + torch.cat per dimension),
 * slices one tile at a time, `.float().to(device)`, runs the net, crops the centre, applies an in-place sigmoid +
 * threshold and pastes into a host mask.  Here the stack stays resident in HBM and the padded stack never exists:
 * a tile is described in PADDED coordinates and gathered straight from the original stack. */
typedef struct HcuTileGeom {
  int32_t channels;         /* C of the stack [C][X][Y][Z] (one image) */
  int32_t size[3];          /* original stack extent X, Y, Z */
  int32_t pad[3];           /* reflection padding per side (utils.py:33: pad_size) */
  int32_t origin[3];        /* tile origin in padded coordinates (segment.py:86: x[0], y[0], z[0]) */
  int32_t extent[3];        /* tile extent (x[1]-x[0], ...) */
  int32_t stack_origin[3];  /* original coordinates of the first voxel of the RESIDENT sub-stack (0,0,0: whole stack) */
  int32_t stack_size[3];    /* extent of the resident sub-stack [C][sx][sy][sz] that `stack` points at */
} HcuTileGeom;
/* tile = scrubbed (NaN -> 0, +-inf -> 1: segment.py:66-67), reflection-padded (utils.py:48-73) slice of the stack.
 * layout bit 0: 0 = [C][x][y][z] (the reference's tile), 1 = channels-last [x][y][z][cpitch], channels >= C zero (the layout the
 * first convolution reads: no NCDHW -> NDHWC pass); bit 1: no scrub (plain pad_image_with_reflections).
 * dtype_stack / dtype_tile: HCU_F32 | HCU_F16. */
int hcu_tile_gather(const HcuTileGeom* g, const void* stack, int32_t dtype_stack, void* tile, int32_t dtype_tile,
                    int32_t layout, int32_t cpitch, void* stream);
/* flag[0] += number of tile values != -1 after the scrub (saturating): 0 <=> the reference skips the tile
 * (segment.py:89-93 `(padded_image_slice.float() == -1).all()`).  The caller zeroes flag. */
int hcu_tile_flags(const HcuTileGeom* g, const void* stack, int32_t dtype_stack, uint32_t* flag, void* stream);
/* mask[morigin + i] = sigmoid(logits[crop + i]) (probability != 0: fp32 mask) or sigmoid(...) > threshold (uint8 mask) for
 * i < ext: centre crop + in-place sigmoid + threshold + paste of segment.py:99-123 in one pass.  logits: fp32 [lsize] of ONE
 * channel; written (optional, uint8 [msize]): set to 1 where the mask was written (merging the shares of several ranks). */
int hcu_sigmoid_paste(const float* logits, const int32_t* lsize, const int32_t* crop, const int32_t* ext, void* mask,
                      int32_t probability, const int32_t* msize, const int32_t* morigin, float threshold, uint8_t* written,
                      void* stream);

/* ---- input path ---------------------------------------------------------------------------------
 * hcu_load_stack: the RAW stack as skimage.io.imread yields it, src = [b][z][y][x][c] uint8 / uint16 (device memory), ->
 * dst = fp16 channels-last [b][x][y][z][dst_cpitch = 8] (channels >= c zero), value half((v / 2^bits - mean[c]) / std[c])
 * computed in float64 and converted like torch does (double -> float -> half): bit-identical to the reference chain.
 * mean / std: HOST arrays of c doubles.
 * Replaces: Stack.__getitem__ (dataloader.py:68-92) with to_float (transforms.py:94-113), reshape (transforms.py:139-157),
 * normalize (transforms.py:257-282), to_tensor (transforms.py:118-136) -- and the engine's own NCDHW -> NDHWC pass.
 * hcu_load_labels: mask / pwl, src = [b][z][y][x] (uint8 / uint16: to_float's 1/2^bits scaling; fp16 / fp32 / fp64: as is;
 * device memory OR pinned host memory read in place), -> dst = fp16 [b][1][ox][oy][oz], the ORIGIN CROP of extent
 * (ox, oy, oz) <= (x, y, z): the only part of them the loss reads (loss.py:51-56).
 * Replaces: the same chain for mask / pwl + the origin crop of loss.py:51-56. */
int hcu_load_stack(const void* src, int32_t dtype_src, int64_t b, int32_t z, int32_t y, int32_t x, int32_t c,
                   const double* mean, const double* stdv, void* dst, int32_t dst_cpitch, void* stream);
int hcu_load_labels(const void* src, int32_t dtype_src, int64_t b, int32_t z, int32_t y, int32_t x, int32_t ox, int32_t oy,
                    int32_t oz, void* dst, void* stream);

/* ---- generic gather-convolution descriptor -------------------------------------------------
 * One launch computes, for every output position o = (n, ox, oy, oz) of an OX*OY*OZ grid and
 * every group g:
 *   out[n, o*ostep + ooff, out_c_off + g*cout + co] =
 *       epilogue( sum_{t, ci} A(in[n, o*istep - pad + t*dil, in_c_off + g*in_c_gstep + ci]) * W[g][t][ci][co] )
 * with zero fill outside the input.  With the right (istep, pad, ostep, ooff, W) this one engine is
 *   - Conv{2,3}d forward, any kernel/dilation/groups, padding 0       (hcat/unet.py:246-257,281-292,120)
 *   - its data gradient (pad = (k-1)*dil, flipped W)                  (autograd of the same calls)
 *   - ConvTranspose{2,3}d forward as stride-phase sub-convolutions    (hcat/unet.py:294-298,310)
 *   - ConvTranspose data gradient (istep = stride)
 * A() is an optional per-input-channel affine+ReLU applied on load (fuses the producer's
 * BatchNorm+ReLU, hcat/unet.py:264-265); the epilogue optionally adds bias, applies a per-channel
 * affine (+ReLU) (eval-mode BN folded in) and accumulates per-channel sum / sum-of-squares in
 * fp64 for training-mode BatchNorm (hcat/unet.py:259-260).
 */
typedef struct HcuConvDesc {
  int32_t dtype_in;        /* HcuDType of `in`  */
  int32_t dtype_out;       /* HcuDType of `out` */
  int32_t batch;
  int32_t in_size[3];      /* IX, IY, IZ */
  int32_t in_cpitch;       /* channels per input voxel in memory */
  int32_t in_c_off;        /* first input channel of group 0 */
  int32_t in_c_gstep;      /* input-channel offset between consecutive groups */
  int32_t cin;             /* input channels reduced over, per group */
  int32_t out_size[3];     /* OX, OY, OZ: output positions computed by this launch */
  int32_t out_tsize[3];    /* TX, TY, TZ: spatial size of the output TENSOR */
  int32_t out_cpitch;
  int32_t out_c_off;
  int32_t cout;            /* output channels per group */
  int32_t groups;
  int32_t taps[3];
  int32_t dil[3];
  int32_t pad[3];          /* low-side zero padding (in input voxels) */
  int32_t istep[3];        /* input step per output position (conv stride) */
  int32_t ostep[3];        /* output step per output position (transposed-conv stride) */
  int32_t ooff[3];         /* output offset (transposed-conv phase) */
  int32_t in_relu;         /* apply ReLU after the input affine */
  int32_t out_relu;        /* apply ReLU in the epilogue */
  /* Stride phases of a transposed convolution folded into the channel dimension (tensor-core kernels only;
   * packed as sx | sy << 8 | sz << 16, 0 = none).  With nph = sx*sy*sz:
   *  ophase: the `cout` output channels are [nph][cout / nph]; channel block phi = (phix, phiy, phiz) (z fastest) of
   *          output position o is written to spatial position o*ostep + ooff + phi, channel out_c_off + co.
   *          (ConvTranspose forward with kernel % stride == 0: ONE launch instead of one per phase.)
   *  iphase: the `cin` input channels are [nph][cin / nph]; channel block phi of input position i is READ from
   *          spatial position i*s + phi of a tensor of spatial size in_size*s and channel pitch in_cpitch / nph.
   *          (ConvTranspose data gradient; hcu_conv_wgrad_tc applies it to its `b` = dy side via `ophase`.) */
  int32_t ophase;
  int32_t iphase;
  int32_t reserved[2];     /* [0]: tensor-core kernel hint (see hcu_conv_tc_describe); [1]: 0, or a forced
                            * K-streamed tile for tests: MB | Nc << 8 | PC << 20 */
} HcuConvDesc;

/* Optional fused tails ("last CTA done" pattern: the CTA that takes the last ticket of `counter` -- a uint32 the caller
 * zeroes -- runs the tiny per-channel finalize inside the producing kernel, saving one dependent launch per layer):
 *  HcuBnFin    : what hcu_bn_finalize computes, run at the end of hcu_conv_tc_fwd_bn;
 *  HcuBnBwdFin : what hcu_bn_bwd_finalize computes, run at the end of hcu_bn_bwd_stats_fin. */
typedef struct HcuBnFin {
  double count;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;   /* may be NULL together with running_var */
  float* running_var;
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
  uint32_t* counter;
} HcuBnFin;
typedef struct HcuBnBwdFin {
  double count;
  const float* gamma;
  int32_t training;
  float grad_scale;
  const float* dscale;   /* optional device scalar multiplied into grad_scale */
  float* dgamma;
  float* dbeta;
  float* dbias;          /* may be NULL */
  float* coef;           /* [3][c] */
  uint32_t* counter;
} HcuBnBwdFin;

struct HcuWeightMap;

/* W: fp32 [groups][taps][cin][cout].  bias/out_scale/out_shift: fp32 [groups*cout] or NULL.
 * in_scale/in_shift: fp32 [in_cpitch] or NULL.  stats: fp64 [2][out_cpitch] (sum, sumsq) or NULL;
 * stats see the value after bias, before out_scale/out_shift/ReLU.
 * Replaces: nn.Conv3d/Conv2d.forward (unet.py:246-257), conv backward-data, ConvTranspose3d (unet.py:294). */
int hcu_conv_fwd(const HcuConvDesc* d, const void* in, const float* W, const float* bias,
                 const float* in_scale, const float* in_shift, const float* out_scale,
                 const float* out_shift, void* out, double* stats, void* stream);

/* Tensor-core flavour of hcu_conv_fwd (tcgen05.mma, TMEM accumulators; conv_tc.cu): fp16 activations, groups == 1,
 * istep == 1, input channel pitch in {8,16,32,64,128}.  hcu_conv_tc_supported() says whether a descriptor is taken
 * (else use hcu_conv_fwd).  Weights: hcu_weight_gather's fp32 [taps][cin][cout] re-packed by hcu_conv_tc_pack into
 * hcu_conv_tc_packed_bytes(d) bytes of fp16 UMMA core matrices.  Same semantics as hcu_conv_fwd otherwise.
 * Replaces: nn.Conv3d/Conv2d.forward (unet.py:246-257), conv backward-data, ConvTranspose3d phases (unet.py:294). */
int hcu_conv_tc_supported(const HcuConvDesc* d);
/* Which kernel and tile configuration a descriptor gets (text, for logs / tests): "classic ..." = conv_tc_kernel (weight
 * slice resident in shared memory, x-march ring), "ks ..." = conv_ks_kernel (K-streamed, >= 64 input channels: A stages
 * and weight tiles streamed through shared memory, accumulators resident in TMEM), "unsupported: why".
 * HcuConvDesc.reserved[0] is a kernel hint: 0 auto, 1 classic only, 2 K-streamed only. */
int hcu_conv_tc_describe(const HcuConvDesc* d, char* buf, int32_t n);
long long hcu_conv_tc_packed_bytes(const HcuConvDesc* d);
int hcu_conv_tc_pack(const HcuConvDesc* d, const float* w, void* packed, void* stream);
/* same, straight from the reference-layout parameter through a HcuWeightMap (gather + fold + fp16 pack in one launch) */
int hcu_conv_tc_pack_ref(const HcuConvDesc* d, const struct HcuWeightMap* m, const float* ref, void* packed, void* stream);
int hcu_conv_tc_fwd(const HcuConvDesc* d, const void* in, const void* packed, const float* bias,
                    const float* in_scale, const float* in_shift, const float* out_scale,
                    const float* out_shift, void* out, double* stats, void* stream);
/* Same, plus the BatchNorm finalize of the produced statistics as a fused tail (fin may be NULL). */
int hcu_conv_tc_fwd_bn(const HcuConvDesc* d, const void* in, const void* packed, const float* bias,
                    const float* in_scale, const float* in_shift, const float* out_scale,
                    const float* out_shift, void* out, double* stats, const HcuBnFin* fin,
                       void* stream);
/* Data gradient of a convolution (hcu_conv_tc_fwd with pad = (k-1)*dil and the flipped packed weights, no bias / statistics /
 * affine) whose EPILOGUE also computes the BatchNorm(+ReLU)-backward statistics of the layer its output is the gradient of:
 * out = g [pixels][c] (fp16, dense); y = that layer's raw convolution output, same shape; scale / shift / mean / invstd its
 * BatchNorm vectors; sums = binned fp64 [HCU_STAT_BINS][2][c], zeroed by the caller; fin as for hcu_bn_bwd_stats_fin (the
 * finalize runs in the last CTA).  Equivalent to hcu_conv_tc_fwd followed by hcu_bn_bwd_stats_fin(g, y, relu = 1) without
 * the second pass over g and y.  Only the specialised 8 / 16-channel variants carry it: hcu_conv_tc_bnbwd_supported(). */
int hcu_conv_tc_fwd_bnbwd(const HcuConvDesc* d, const void* in, const void* packed, void* out, const void* y,
                          const float* scale, const float* shift, const float* mean, const float* invstd, double* sums,
                          const HcuBnBwdFin* fin, void* stream);
int hcu_conv_tc_bnbwd_supported(const HcuConvDesc* d);


/* Weight gradient of the same gather-convolution:
 *   R[g][t][ca][cb] = sum_{n,o} A(a[n, o*istep - pad + t*dil, a_c_off + g*a_c_gstep + ca]) * b[n, o, b_c_off + g*cb_n + cb]
 * `d` describes the gather side exactly like hcu_conv_fwd (in_* = a, cin = ca count, cout = cb count,
 * out_size = the o grid, out_cpitch/out_c_off/dtype_out describe `b`, ostep/ooff must be 1/0).
 * The reduction over (n,o) is split over `nsplit` CTAs rows; partial[nsplit][groups*taps*cin*cout] fp32.
 * Replaces: convolution_backward's weight gradient (autograd of unet.py:246-257,294-298). */
int hcu_conv_wgrad_partial(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift,
                           const void* b, float* partial, int32_t nsplit, void* stream);

/* Tensor-core flavour of the weight gradient (mma.sync m16n8k16, fp16 operands, fp32 accumulate; wgrad_mma.cu):
 * wacc: fp32 [taps][cin][cout] (groups == 1), zeroed and accumulated by this call; feed it to hcu_weight_scatter with
 * nsplit = 1.  hcu_conv_wgrad_tc_supported() says whether the descriptor is taken (else hcu_conv_wgrad_partial). */
int hcu_conv_wgrad_tc_supported(const HcuConvDesc* d);
int hcu_conv_wgrad_tc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                      float* wacc, void* stream);
/* Same, but ACCUMULATES into a `wacc` the caller has zeroed (one memset for every layer of a step). */
int hcu_conv_wgrad_tc_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                          float* wacc, void* stream);
/* hcu_conv_wgrad_ws_*: warp-specialised mma.sync pipeline for the 8/16-channel levels (wgrad_ws.cu; producer warps
 * stage planes with cp.async while consumer warps run ldmatrix + MMA); same contract as hcu_conv_wgrad_tc_acc.
 * tcgen05 flavour for the channel-rich levels (wgrad_tc5.cu): the staged planes are MN-major UMMA operands with the
 * pixel index as K, one M = 128 (Cin rows) x N = Cout MMA per tap and 16 pixels, taps split over CTAs by TMEM capacity.
 * Same contract as hcu_conv_wgrad_tc_acc (accumulates into a zeroed fp32 [taps][cin][cout]). */
int hcu_conv_wgrad_ws_supported(const HcuConvDesc* d);
int hcu_conv_wgrad_ws_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                          float* wacc, void* stream);
int hcu_conv_wgrad_tc5_supported(const HcuConvDesc* d);
int hcu_conv_wgrad_tc5_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                           float* wacc, void* stream);
/* hcu_conv_wgrad_rows_*: tcgen05 flavour for the channel-POOR levels (8 ... 32 channels; wgrad_rows.cu), fed by TMA
 * tensor-map loads: image rows are stacked on BOTH operand dimensions (M = 16 input rows x 8 channels, N = up to 14 dy rows
 * x 8 channels, K = 16 z positions), so one MMA covers 224 pixels of three taps and the KY taps are block diagonals of the
 * accumulator.  Valid 3D convolutions with >= 8 output z positions, no stride phases.
 * Same contract as hcu_conv_wgrad_tc_acc (accumulates into a zeroed fp32 [taps][cin][cout]).
 * Replaces: the weight-gradient half of autograd's conv backward for unet.py:246-257 on the first levels. */
int hcu_conv_wgrad_rows_supported(const HcuConvDesc* d);
int hcu_conv_wgrad_rows_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* dy,
                            float* wacc, void* stream);
/* Same kernel with hcu_bn_bwd_apply fused into the staging of its dy operand: dy = c1 * (bn(y) > 0 ? g : 0) + c2 * y + c3 is
 * computed (fp32, rounded to fp16 like the stand-alone pass) on the landed (g, y) tiles and never written to memory.  For a
 * layer whose data gradient nobody needs (the FIRST conv of the net) this removes the apply pass: 2 reads + 1 write of the
 * layer's gradient tensor and the re-read by the weight gradient.  bn_scale / bn_shift: the layer's BatchNorm as scale / shift
 * (the ReLU mask), coef: hcu_bn_bwd_finalize's [3][out_cpitch].
 * Replaces: BatchNorm + ReLU backward followed by the conv weight gradient (autograd of unet.py:259-265 / 246-250). */
int hcu_conv_wgrad_rows_bnb_supported(const HcuConvDesc* d);
int hcu_conv_wgrad_rows_bnb_acc(const HcuConvDesc* d, const void* a, const float* a_scale, const float* a_shift, const void* g,
                                const void* y, const float* bn_scale, const float* bn_shift, const float* coef, float* wacc,
                                void* stream);

/* ---- optimiser -------------------------------------------------------------------------------
 * Adam on ONE flat fp32 parameter buffer (hcunet_b200.FlatParameters: every nn.Parameter is a view of it, its gradient is the
 * engine's flat gradient buffer), one launch per step, with the skip-step check of fp16-storage training inside: if any gradient
 * is inf / NaN nothing is touched and `step` is not advanced.  torch.optim.Adam's arithmetic (amsgrad = False, maximize = False),
 * `weight_decay` as its L2 term.  step: device int32 (steps taken so far); scratch: 2 device words the call zeroes, scratch[0]
 * reads 1.0f afterwards when the step was skipped.
 * Replaces: torch.optim.Adam.step() of the reference's training loop (tests/r_unet_test.py:24, hcat/train scripts). */
int hcu_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t* step, float* scratch, void* stream);

/* ---- weight layout transforms -------------------------------------------------------------
 * Generic strided gather between a reference-layout parameter and a packed GEMM-B tensor
 * P[g][jx][jy][jz][a][b]:
 *   ref_index = base + g*sg + a*sa + b*sb + sum_d (t0[d] + j_d*tstep[d]) * st[d]   (+ fold_stride if fold)
 *               (+ sum_d phi_d * pst[d] when a phase is folded into a / b)
 * gather : P = ref[idx] (+ ref[idx+fold_stride] when fold)      -- forward / dgrad / phase weights
 * scatter: ref[idx] (and ref[idx+fold_stride]) = scale * sum_{s<nsplit} P_s   -- wgrad results
 * `dscale` (every function that has it): optional DEVICE fp32 scalar multiplied into `scale`; it carries
 * the 1/S of the fp16 backward's loss scaling (hcu_grad_scale) without a host round trip.
 */
typedef struct HcuWeightMap {
  int32_t groups, j[3], na, nb;
  int64_t base, sg, sa, sb, st[3];
  int32_t t0[3], tstep[3];
  int32_t fold;            /* 0/1 */
  int64_t fold_stride;
  /* stride phases folded into a channel index (see HcuConvDesc.ophase): phase_on = 0 none, 1: the `a` index is
   * [nph][na], 2: the `b` index is [nph][nb]; nph = ph[0]*ph[1]*ph[2]; phase (phix,phiy,phiz) adds sum phi_d*pst[d] */
  int32_t phase_on, ph[3];
  /* block-diagonal: bdiag = G > 1 (groups must be 1, no phases) makes the packed tensor DENSE [taps][G*na][G*nb] over a
   * grouped reference weight (`groups=G` of nn.Conv, main.py:46-55): entry (a, b) maps to group a/na's block when
   * a/na == b/nb (sg = group stride) and is ZERO otherwise -- gathers/packs write 0 there, scatters skip it.  The grouped
   * convolution then runs as one dense tensor-core convolution. */
  int32_t bdiag;
  int64_t pst[3];
} HcuWeightMap;
int hcu_weight_gather(const HcuWeightMap* m, const float* ref, float* packed, void* stream);
/*
 * Batched forms (one launch for every layer of a step instead of one per layer; the per-layer calls stay available).
 * The caller owns a job table: hcu_*_batch_build fills `n` jobs of HCU_BATCH_JOB_BYTES bytes each into HOST memory,
 * the caller copies them to the device once (they only hold geometry and OFFSETS, no pointers) and passes the device
 * copy to the launch together with the base pointers.  `blocks` (build output) is the grid size of the launch.
 *  pack   : job i packs the reference-layout parameter at params + ref_off[i] (float elements) into
 *           packed + out_off[i] (bytes; hcu_conv_tc_packed_bytes(&descs[i]) bytes, 16-byte aligned offsets).
 *  scatter: job i reduces nsplit[i] partial results at partial + part_off[i] (float elements, split stride =
 *           the job's element count) and writes scale * sum into grads + ref_off[i] (float elements) through maps[i]
 *           (same semantics as hcu_weight_scatter with accumulate = 0).
 */
#define HCU_BATCH_JOB_BYTES 256
int hcu_conv_tc_pack_batch_build(const HcuConvDesc* descs, const HcuWeightMap* maps, const int64_t* ref_off,
                                 const int64_t* out_off, int32_t n, void* host_jobs, int32_t* blocks);
int hcu_conv_tc_pack_batch(const void* dev_jobs, int32_t n, int32_t blocks, const float* params, void* packed,
                           void* stream);
int hcu_weight_scatter_batch_build(const HcuWeightMap* maps, const int32_t* nsplit, const int64_t* part_off,
                                   const int64_t* ref_off, int32_t n, void* host_jobs, int32_t* blocks);
int hcu_weight_scatter_batch(const void* dev_jobs, int32_t n, int32_t blocks, const float* partial, float scale,
                             const float* dscale, float* grads, void* stream);

int hcu_weight_scatter(const HcuWeightMap* m, const float* partial, int32_t nsplit, int64_t split_stride,
                       float scale, const float* dscale, int32_t accumulate, float* ref, void* stream);

/* ---- boundary layout ------------------------------------------------------------------------
 * [N][C][S] (reference, S = X*Y*Z) <-> [N][S][cpitch] channels-last.  Extra channels are zero.
 * Replaces nothing in the reference; it is the cost of keeping its NCDHW API (unet.py:125). */
int hcu_nc_to_cl(const void* src, int32_t dtype_src, void* dst, int32_t dtype_dst, int64_t n, int32_t c,
                 int64_t s, int32_t cpitch, const float* dscale, void* stream);
int hcu_cl_to_nc(const void* src, int32_t dtype_src, void* dst, int32_t dtype_dst, int64_t n, int32_t c,
                 int64_t s, int32_t cpitch, const float* dscale, void* stream);
/* Loss scaling of the fp16 backward: scales[0] = S = 2^k such that max|g| * S is in (target/2, target],
 * scales[1] = 1/S; g: fp32 [n]; scratch: one uint32.  No host synchronisation. */
int hcu_grad_scale(const float* g, int64_t n, float target, unsigned int* scratch, float* scales, void* stream);

/* ---- BatchNorm (+ReLU, +MaxPool) --------------------------------------------------------------
 * Training statistics come from hcu_conv_fwd's fp64 `stats`.  hcu_bn_finalize turns them into
 * mean / invstd / (scale, shift) and updates running stats exactly like nn.BatchNorm3d
 * (momentum, unbiased running var, eps) -- unet.py:259-260,305-306.
 * hcu_bn_eval_affine builds (scale, shift) from running stats for eval(). */
int hcu_bn_finalize(const double* stats, int32_t c, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* mean, float* invstd, float* scale, float* shift, void* stream);
int hcu_bn_eval_affine(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, const float* conv_bias, float* scale, float* shift,
                       void* stream);
/* a = relu?(y*scale[c] + shift[c]) elementwise over [npix][c].  Replaces batchN + relu_ (unet.py:264-265). */
int hcu_bn_relu_apply(const void* y, int32_t dtype_y, void* a, int32_t dtype_a, int64_t npix, int32_t c,
                      const float* scale, const float* shift, int32_t relu, void* stream);
/* Fused (optional affine+ReLU) + MaxPool with kernel == stride, floor mode (unet.py:123,131).
 * in [n][ix][iy][iz][c] -> pooled [n][ix/px][iy/py][iz/pz][c]; argmax: uint8 window index
 * ((wx*py + wy)*pz + wz) of the FIRST maximum in PyTorch scan order, NaN propagates. */
int hcu_bn_relu_maxpool(const void* y, int32_t dtype_y, void* pooled, int32_t dtype_p, uint8_t* argmax,
                        int32_t n, int32_t ix, int32_t iy, int32_t iz, int32_t c, int32_t px, int32_t py,
                        int32_t pz, const float* scale, const float* shift, int32_t relu, void* stream);
/* dfull[n][ix][iy][iz][c] = dpooled routed to the argmax voxel, zero elsewhere (max_pool backward). */
int hcu_maxpool_bwd(const void* dpooled, int32_t dtype_dp, const uint8_t* argmax, void* dfull, int32_t dtype_df,
                    int32_t n, int32_t ix, int32_t iy, int32_t iz, int32_t c, int32_t px, int32_t py, int32_t pz,
                    void* stream);
/* geometry of the max-pool that followed a BN+ReLU block (kernel == stride): [n][ix][iy][iz] -> [n][ix/px][iy/py][iz/pz] */
typedef struct HcuPoolGeom { int32_t n, ix, iy, iz, px, py, pz; } HcuPoolGeom;
/* BatchNorm+ReLU backward, pass 1: g = da * [y*scale+shift > 0 or !relu];
 * sums[0][c] += sum g, sums[1][c] += sum g*(y-mean)*invstd   (fp64).
 * Fused max-pool backward (fp16, c % 8 == 0): when `argmax` != NULL, `da` is the gradient of the POOLED tensor and
 * `pool` its geometry; the full-resolution gradient (hcu_maxpool_bwd's output) is formed on the fly. */
int hcu_bn_bwd_stats(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, int64_t npix, int32_t c,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     int32_t relu, const uint8_t* argmax, const HcuPoolGeom* pool, double* sums, void* stream);
/* pass 1 + pass 2 in one call: fused tail on the fp16 vector path (one launch), two launches otherwise. */
int hcu_bn_bwd_stats_fin(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, int64_t npix, int32_t c,
                         const float* scale, const float* shift, const float* mean, const float* invstd,
                         int32_t relu, const uint8_t* argmax, const HcuPoolGeom* pool, double* sums,
                         const HcuBnBwdFin* fin, void* stream);
/* pass 2 (tiny): dgamma, dbeta, conv-bias grad and the coefficients of dy = c1*g + c2*y + c3.
 * training != 0: batch-stat backward; training == 0: running-stat (eval) backward. */
int hcu_bn_bwd_finalize(const double* sums, int32_t c, double count, const float* gamma, const float* mean,
                        const float* invstd, int32_t training, float grad_scale, const float* dscale, float* dgamma,
                        float* dbeta, float* dbias, float* coef, void* stream);
/* pass 3: dy = c1[c]*g + c2[c]*y + c3[c]  (same optional fused max-pool backward). */
int hcu_bn_bwd_apply(const void* da, int32_t dtype_da, const void* y, int32_t dtype_y, void* dy, int32_t dtype_dy,
                     int64_t npix, int32_t c, const float* scale, const float* shift, int32_t relu,
                     const float* coef, const uint8_t* argmax, const HcuPoolGeom* pool, void* stream);
/* out[c] = scale * dscale[0] * sum over pixels of x[pix][c_off + c]  (bias gradients of convT / out_conv).  scratch: 4096 doubles
 * (binned partial sums, zeroed by the call); c <= 4096. */
int hcu_colsum(const void* x, int32_t dtype_x, int64_t npix, int32_t cpitch, int32_t c_off, int32_t c,
               float scale, const float* dscale, double* scratch, float* out, void* stream);

/* ---- losses (hcat/loss.py) ----------------------------------------------------------------------
 * pred: fp32 contiguous [B][C][x][y][z] (2D: z == 1).  mask / pwl: any HcuDType, [B][C][X][Y][Z]
 * with X>=x..., cropped from the origin (loss.py:51-56); pwl may be NULL (weight 2, loss.py:46-48).
 * mode 0 'pixel' (loss.py:70-72); mode 1 'sigmoid' (loss.py:38-40,97-99: BCE-with-logits of sigmoid(pred)).
 * out_sum: fp64 scalar accumulator (caller zeroes it); if zsums != NULL also per-z sums fp64 [z]
 * (the 'worst_z' reduction, loss.py:74-80). */
typedef struct HcuLossDesc {
  int32_t b, c, x, y, z;      /* pred shape */
  int32_t mx, my, mz;         /* mask / pwl spatial shape (>= pred's) */
  int32_t dtype_mask, dtype_pwl;
  int32_t mode;
  int32_t reserved[3];
} HcuLossDesc;
int hcu_wbce_fwd(const HcuLossDesc* d, const float* pred, const void* mask, const void* pwl, double* out_sum,
                 double* zsums, void* stream);
/* dpred = gout[0] * mult * zscale[z] * dBCE/dpred * (pwl+1); gout is a DEVICE fp32 scalar (the upstream
 * gradient), mult a host scalar (1/N for the mean), zscale: DEVICE fp32 [z] or NULL. */
int hcu_wbce_bwd(const HcuLossDesc* d, const float* pred, const void* mask, const void* pwl, const float* gout,
                 float mult, const float* zscale, float* dpred, void* stream);
/* dice / L1 / MSE reductions (loss.py:104-177): sums[0..2] fp64 =
 *   kind 0 (dice): sum sigmoid(p)*m, sum sigmoid(p), sum m;  kind 1 (L1): sum |p-m|;  kind 2 (MSE): sum (p-m)^2 */
int hcu_pair_reduce(const HcuLossDesc* d, int32_t kind, const float* pred, const void* mask, double* sums,
                    void* stream);
/* dpred for the same three; coef: DEVICE fp32 [2]:
 *   dice: dpred = s(1-s) * (coef[0]*m + coef[1]);  L1: coef[0]*sign(p-m);  MSE: coef[0]*(p-m) */
int hcu_pair_bwd(const HcuLossDesc* d, int32_t kind, const float* pred, const void* mask, const float* coef,
                 float* dpred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HCUNET_B200_H */
