"""Planner + executor for the U-Net hot path on the C ABI (``include/hcunet_b200.h``).

``plan_unet`` turns the reference's constructor kwargs (`hcat/unet.py:16-27`) plus an input shape into
a static list of layer geometries, raising the same ``RuntimeError`` the reference raises for inputs
that are too small (`unet.py:246-257` via ATen's "Kernel size can't be greater than actual input
size", and the ``torch.cat`` size mismatch at `unet.py:312`).  ``UnetEngine`` executes that plan:
every arithmetic step is one of the library's hand-written kernels, enqueued on the current CUDA
stream.  Activations are channels-last ``[N][X][Y][Z][C]``; the reference layout only exists at the
boundary.  The dead skip connection (`unet.py:309-312`: ``conv1(cat(x_up, x_up))``) is reproduced by
folding the two K-halves of ``Up.conv1``'s weight, so no concatenated tensor is ever built.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import HcuConvDesc, HcuWeightMap

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
_ACT_DTYPE = {"fp32": torch.float32, "mixed": torch.float16}
GRAD_SCALE_TARGET = 64.0  # fp16 backward: max|dlogits| * S lands in (32, 64]


def _t3(v, dims: int, fill: int = 1) -> Tuple[int, int, int]:
    """int / tuple of len dims -> 3-tuple (2D gets a trailing Z entry of ``fill``)."""
    if isinstance(v, int):
        v = (v,) * dims
    v = tuple(int(e) for e in v)
    if len(v) != dims:
        raise RuntimeError(f"expected {dims} values, got {v}")
    return v + (fill,) * (3 - dims)


@dataclass
class ConvGeom:
    name: str            # e.g. down_steps.0.conv1
    bn: Optional[str]    # e.g. down_steps.0.batch1 (None for out_conv)
    cin_t: int           # channels of the tensor this conv reads
    cout_t: int
    groups: int          # groups as launched (after the cat(x,x) rewrite)
    cin_g: int
    cout_g: int
    ref_cin_g: int       # second dim of the reference weight tensor
    fold: bool           # Up.conv1 with groups == 1: W_eff = W[:, :C] + W[:, C:]
    taps: Tuple[int, int, int]
    dil: Tuple[int, int, int]
    in_sz: Tuple[int, int, int]
    out_sz: Tuple[int, int, int]
    pool: Optional[Tuple[int, int, int]] = None      # max-pool applied to this block's output
    pool_sz: Optional[Tuple[int, int, int]] = None
    first: bool = False


@dataclass
class UpGeom:
    name: str            # up_steps.i.up_conv
    cin: int
    cout: int
    k: Tuple[int, int, int]
    s: Tuple[int, int, int]
    in_sz: Tuple[int, int, int]
    out_sz: Tuple[int, int, int]


@dataclass
class Plan:
    dims: int
    batch: int
    in_channels: int
    out_channels: int
    in_sz: Tuple[int, int, int]
    steps: List[object] = field(default_factory=list)   # ConvGeom / UpGeom in execution order
    out_sz: Tuple[int, int, int] = (0, 0, 0)


def _conv_out(sz, taps, dil, what):
    out = []
    for i in range(3):
        eff = (taps[i] - 1) * dil[i] + 1
        if sz[i] < eff:
            raise RuntimeError(
                f"Calculated padded input size per channel: {tuple(sz)}. Kernel size: "
                f"{tuple((t - 1) * d + 1 for t, d in zip(taps, dil))}. Kernel size can't be greater than actual "
                f"input size ({what})")
        out.append(sz[i] - eff + 1)
    return tuple(out)


def plan_unet(spec: dict, xshape) -> Plan:
    dims = spec["image_dimensions"]
    if len(xshape) != dims + 2:
        raise RuntimeError(f"Expected {dims + 2}D input to a {dims}D U-Net, but got input of size: {list(xshape)}")
    feats = list(spec["feature_sizes"])
    if xshape[1] != spec["in_channels"]:
        raise RuntimeError(f"Given groups={spec['groups']['conv1']}, expected input{list(xshape)} to have "
                           f"{spec['in_channels']} channels, but got {xshape[1]} channels instead")
    sz = tuple(int(v) for v in xshape[2:]) + (1,) * (3 - dims)
    plan = Plan(dims=dims, batch=int(xshape[0]), in_channels=spec["in_channels"],
                out_channels=spec["out_channels"], in_sz=sz)
    taps = {k: _t3(spec["kernel"][k], dims) for k in ("conv1", "conv2")}
    dil = {k: _t3(spec["dilation"][k], dims) for k in ("conv1", "conv2")}
    grp = {k: int(spec["groups"][k]) for k in ("conv1", "conv2")}
    pool = _t3(spec["max_pool_kernel"], dims)
    up_k = _t3(spec["upsample_kernel"], dims)
    up_s = _t3(spec["upsample_stride"], dims)

    def block(prefix, cin, cout, sz, up):
        for idx in ("1", "2"):
            key = "conv" + idx
            c_in = cin if idx == "1" else cout
            g = grp[key]
            ref_cin = (2 * c_in if (up and idx == "1") else c_in)
            if ref_cin % g or cout % g:
                raise ValueError("in_channels and out_channels must be divisible by groups")
            fold = False
            groups, cin_g, cout_g = g, c_in // g, cout // g
            if up and idx == "1":
                # conv1(cat(x, x)) -- unet.py:311-312.  g == 1: fold the K halves.  g == 2: group j sees
                # cat channels [j*C, (j+1)*C) == all of x, i.e. a dense conv with the weight as stored.
                if g == 1:
                    fold, groups, cin_g, cout_g = True, 1, c_in, cout
                elif g == 2:
                    groups, cin_g, cout_g = 1, c_in, cout
                else:
                    raise NotImplementedError("groups > 2 on Up.conv1 (cat(x, x) slices wrap) is not supported yet")
            osz = _conv_out(sz, taps[key], dil[key], f"{prefix}.{key}")
            plan.steps.append(ConvGeom(name=f"{prefix}.{key}", bn=f"{prefix}.batch{idx}", cin_t=c_in, cout_t=cout,
                                       groups=groups, cin_g=cin_g, cout_g=cout_g, ref_cin_g=ref_cin // g, fold=fold,
                                       taps=taps[key], dil=dil[key], in_sz=sz, out_sz=osz))
            sz = osz
        return sz

    skips = []
    cin = spec["in_channels"]
    for i, f in enumerate(feats):
        sz = block(f"down_steps.{i}", cin, f, sz, up=False)
        cin = f
        if i < len(feats) - 1:
            skips.append(sz)
            psz = tuple(sz[d] // pool[d] for d in range(3))
            if min(psz) < 1:
                raise RuntimeError(f"Given input size: ({f}x{'x'.join(map(str, sz[:dims]))}). Calculated output size: "
                                   f"({f}x{'x'.join(map(str, psz[:dims]))}). Output size is too small")
            last = plan.steps[-1]
            last.pool, last.pool_sz = pool, psz
            sz = psz
    plan.steps[0].first = True
    for i in range(len(feats) - 1):
        f_in, f_out = feats[-1 - i], feats[-2 - i]
        for d in range(3):
            if up_k[d] < up_s[d]:
                raise NotImplementedError("upsample_kernel smaller than upsample_stride is not supported")
        osz = tuple((sz[d] - 1) * up_s[d] + up_k[d] for d in range(3))
        plan.steps.append(UpGeom(name=f"up_steps.{i}.up_conv", cin=f_in, cout=f_out, k=up_k, s=up_s, in_sz=sz,
                                 out_sz=osz))
        skip = skips.pop()
        # crop(x_up, skip) then cat((x_up, cropped)) -- unet.py:311-312: fails when the skip is smaller anywhere
        for d in range(dims):
            if skip[d] < osz[d]:
                raise RuntimeError(f"Sizes of tensors must match except in dimension 1. Expected size {osz[d]} but got "
                                   f"size {skip[d]} for tensor number 1 in the list.")
        sz = block(f"up_steps.{i}", f_out, f_out, osz, up=True)
    plan.steps.append(ConvGeom(name="out_conv", bn=None, cin_t=feats[0], cout_t=spec["out_channels"], groups=1,
                               cin_g=feats[0], cout_g=spec["out_channels"], ref_cin_g=feats[0], fold=False,
                               taps=(1, 1, 1), dil=(1, 1, 1), in_sz=sz, out_sz=sz))
    plan.out_sz = sz
    return plan


# ---------------------------------------------------------------------------------------------
# descriptor helpers
# ---------------------------------------------------------------------------------------------

def _i3(v):
    return (C.c_int32 * 3)(*[int(e) for e in v])


def conv_desc(dt_in, dt_out, batch, in_size, in_cpitch, in_c_off, in_c_gstep, cin, out_size, out_tsize, out_cpitch,
              out_c_off, cout, groups, taps, dil=(1, 1, 1), pad=(0, 0, 0), istep=(1, 1, 1), ostep=(1, 1, 1),
              ooff=(0, 0, 0), in_relu=0, out_relu=0) -> HcuConvDesc:
    d = HcuConvDesc()
    d.dtype_in, d.dtype_out, d.batch = dt_in, dt_out, batch
    d.in_size, d.in_cpitch, d.in_c_off, d.in_c_gstep, d.cin = _i3(in_size), in_cpitch, in_c_off, in_c_gstep, cin
    d.out_size, d.out_tsize, d.out_cpitch, d.out_c_off = _i3(out_size), _i3(out_tsize), out_cpitch, out_c_off
    d.cout, d.groups = cout, groups
    d.taps, d.dil, d.pad, d.istep, d.ostep, d.ooff = _i3(taps), _i3(dil), _i3(pad), _i3(istep), _i3(ostep), _i3(ooff)
    d.in_relu, d.out_relu = in_relu, out_relu
    return d


def weight_map(groups, j, na, nb, sg, sa, sb, st, t0=(0, 0, 0), tstep=(1, 1, 1), base=0, fold=0,
               fold_stride=0, phase_on=0, ph=(1, 1, 1), pst=(0, 0, 0), bdiag=0) -> HcuWeightMap:
    m = HcuWeightMap()
    m.phase_on, m.ph, m.bdiag = phase_on, _i3(ph), bdiag
    m.pst = (C.c_int64 * 3)(*[int(e) for e in pst])
    m.groups, m.j, m.na, m.nb = groups, _i3(j), na, nb
    m.base, m.sg, m.sa, m.sb = base, sg, sa, sb
    m.st = (C.c_int64 * 3)(*[int(e) for e in st])
    m.t0, m.tstep, m.fold, m.fold_stride = _i3(t0), _i3(tstep), fold, fold_stride
    return m


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _nsplit(m: int, roles: int) -> int:
    """How many CTAs rows split the pixel reduction of a weight gradient."""
    target = 148 * 8 * 3
    n = max(1, -(-target // max(1, roles)))
    n = min(n, max(1, m // 256))
    return int(min(n, 2048))


class _StepCache:
    """What one (input shape, precision, mode) step needs every time, recorded on its first execution:

    * the tensor-core weight packs of every conv launch (forward, data gradient, transposed-conv phases) -- replayed
      as ONE ``hcu_conv_tc_pack_batch`` launch into a persistent buffer at the start of the forward;
    * the weight-gradient scatters -- every weight gradient accumulates into a slice of a persistent fp32 workspace
      (one memset per backward) and ONE ``hcu_weight_scatter_batch`` launch writes all of them, in the reference
      layouts, into a freshly allocated flat gradient buffer the returned gradients are views of.
    The job tables hold geometry and offsets only; the parameter addresses they were built against are re-checked on
    every step (``ptr_sig``) and the cache re-records itself when they move."""

    def __init__(self):
        self.ready = False
        self.fwd_done = False
        self.bwd_done = False
        self.pack_jobs: Dict[str, tuple] = {}      # key -> (HcuConvDesc, HcuWeightMap, param name, packed bytes)
        self.scatter_jobs: Dict[str, tuple] = {}   # weight name -> (HcuWeightMap, nsplit, element count)
        self.ptr_sig = None
        self.pack_base = 0
        self.packed = None
        self.pack_off: Dict[str, int] = {}
        self.pack_table = None
        self.pack_blocks = 0
        self.wacc = None
        self.part_off: Dict[str, int] = {}
        self.g_off: Dict[str, int] = {}
        self.g_total = 0
        self.scatter_table = None
        self.scatter_blocks = 0
        # data-parallel overlap: the scatter jobs recorded before / after the bucket split (UnetEngine.grad_ready_hook)
        self.split_at = None            # number of scatter jobs that belong to the EARLY bucket (None: one bucket)
        self.scatter_tables2 = None     # [(table, njobs, blocks)] for the early and the late part

    @staticmethod
    def _sig(params, names):
        return tuple(params[n].data_ptr() for n in names)

    def finalize(self, lib, params, device):
        n = len(self.pack_jobs)
        if n:
            names = [j[2] for j in self.pack_jobs.values()]
            self.pack_base = min(params[nm].data_ptr() for nm in names)
            descs = (HcuConvDesc * n)(*[j[0] for j in self.pack_jobs.values()])
            maps = (HcuWeightMap * n)(*[j[1] for j in self.pack_jobs.values()])
            ref_off = (C.c_int64 * n)(*[(params[nm].data_ptr() - self.pack_base) // 4 for nm in names])
            offs, cur = [], 0
            for key, j in self.pack_jobs.items():
                self.pack_off[key] = cur
                offs.append(cur)
                cur += -(-j[3] // 256) * 256
            out_off = (C.c_int64 * n)(*offs)
            host = C.create_string_buffer(n * _lib.BATCH_JOB_BYTES)
            blocks = C.c_int32(0)
            _lib.check(lib.hcu_conv_tc_pack_batch_build(descs, maps, ref_off, out_off, n, host, C.byref(blocks)),
                       "conv_tc_pack_batch_build")
            self.pack_blocks = blocks.value
            self.pack_table = torch.frombuffer(bytearray(host.raw), dtype=torch.uint8).to(device)
            self.packed = torch.empty(cur, dtype=torch.uint8, device=device)
        m = len(self.scatter_jobs)
        if m:
            maps = (HcuWeightMap * m)(*[j[0] for j in self.scatter_jobs.values()])
            nsp = (C.c_int32 * m)(*[j[1] for j in self.scatter_jobs.values()])
            p_off, acc = {}, 0   # offsets of every parameter in the step's ONE flat gradient buffer (parameter order)
            for name, p in params.items():
                p_off[name] = acc
                acc += p.numel()
            po, go, pc = [], [], 0
            for name, j in self.scatter_jobs.items():
                self.part_off[name], self.g_off[name] = pc, p_off[name]
                po.append(pc)
                go.append(p_off[name])
                pc += j[1] * j[2]
            self.g_total = acc
            host = C.create_string_buffer(m * _lib.BATCH_JOB_BYTES)
            blocks = C.c_int32(0)
            _lib.check(lib.hcu_weight_scatter_batch_build(maps, nsp, (C.c_int64 * m)(*po), (C.c_int64 * m)(*go), m, host,
                                                          C.byref(blocks)), "weight_scatter_batch_build")
            self.scatter_blocks = blocks.value
            self.scatter_table = torch.frombuffer(bytearray(host.raw), dtype=torch.uint8).to(device)
            self.wacc = torch.empty(pc, dtype=torch.float32, device=device)
            if self.split_at is not None and 0 < self.split_at < m:
                self.scatter_tables2 = []
                for a, b in ((0, self.split_at), (self.split_at, m)):
                    k = b - a
                    host2 = C.create_string_buffer(k * _lib.BATCH_JOB_BYTES)
                    blk = C.c_int32(0)
                    _lib.check(lib.hcu_weight_scatter_batch_build((HcuWeightMap * k)(*list(maps)[a:b]), (C.c_int32 * k)(*list(nsp)[a:b]),
                                                                  (C.c_int64 * k)(*po[a:b]), (C.c_int64 * k)(*go[a:b]), k, host2,
                                                                  C.byref(blk)), "weight_scatter_batch_build")
                    self.scatter_tables2.append((torch.frombuffer(bytearray(host2.raw), dtype=torch.uint8).to(device), k, blk.value))
        self.ptr_sig = self._sig(params, [j[2] for j in self.pack_jobs.values()])
        self.ready = True


class UnetEngine:
    """Executes forward / backward of one ``Plan`` on the CUDA library."""

    def __init__(self, spec: dict):
        self.spec = spec
        self._plans: Dict[tuple, Plan] = {}
        self._inv = None
        self.use_tc = os.environ.get("HCUNET_TC", "1") != "0"
        self.use_ws = os.environ.get("HCUNET_WGRADWS", "1") != "0"      # warp-specialised weight gradient (8/16-channel levels)
        self.use_tc5 = os.environ.get("HCUNET_WGRAD5", "1") != "0"      # tcgen05 weight gradient on the channel-rich levels
        # row-stacked tcgen05 weight gradient fed by TMA (channel-poor levels); input channel pitch up to ROWS_MAXCP
        self.use_rows = os.environ.get("HCUNET_WGRADROWS", "1") != "0"
        self.rows_maxcp = int(os.environ.get("HCUNET_WGRADROWS_MAXCP", "32"))
        self.rows_maxcp_out = int(os.environ.get("HCUNET_WGRADROWS_MAXCP_OUT", "64"))
        self.fuse_apply = os.environ.get("HCUNET_FUSE_APPLY", "1") != "0"   # BN-backward apply inside the first layer's weight gradient
        self.use_batch = os.environ.get("HCUNET_BATCH", "1") != "0"      # batched packs / scatters (_StepCache)
        self.overlap_wgrad = os.environ.get("HCUNET_OVERLAP", "1") != "0"  # weight gradients on a side stream
        # BatchNorm-backward statistics of a block's conv1 computed in the epilogue of conv2's data gradient (its producer)
        self.fuse_bnbwd = os.environ.get("HCUNET_FUSE_BNBWD", "1") != "0"
        self._caches: Dict[tuple, _StepCache] = {}
        self._cache: Optional[_StepCache] = None   # cache of the call in progress
        self._side = None
        self._side2 = None
        self.n_side = int(os.environ.get("HCUNET_SIDE_STREAMS", "2"))  # two gradient streams: 2.560 -> 2.540 ms (no gain before wgrad_rows)
        self._keep: List[torch.Tensor] = []
        self.last_grad_flat: Optional[torch.Tensor] = None   # the flat gradient buffer of the latest backward
        self._goff: Dict[str, tuple] = {}
        # Test instrumentation (tests/test_gpu_teacher.py): ``tap(tag, tensor, channels, spatial)`` is called on the current
        # stream right after every stored tensor of a step has been produced (channels-last [B, S, C pitch]); it may read
        # the tensor or overwrite it in place.  None in normal operation.
        self.tap = None
        # Data-parallel overlap (hcunet_b200.parallel.GradSync.attach): called twice per backward with (flat, lo, hi, events) -- the
        # range [lo, hi) of the flat gradient buffer `flat` that is final once `events` have completed.  First when the deep levels
        # and the whole up path are done (most of the parameters: their all-reduce runs while the first levels' backward --
        # most of the time -- is still computing), then for the rest at the end.  Returns an event the backward's stream
        # waits for before it returns (or None).
        self.grad_ready_hook = None

    @property
    def lib(self):
        """The ctypes library, loaded on first use (raises when it is missing: no fallback)."""
        return _lib.load()

    def plan(self, xshape) -> Plan:
        key = tuple(xshape)
        p = self._plans.get(key)
        if p is None:
            p = plan_unet(self.spec, xshape)
            self._plans[key] = p
        return p

    def _tap(self, tag, t, c, sz):
        if self.tap is not None and t is not None:
            self.tap(tag, t, int(c), tuple(int(v) for v in sz))

    # ---- small wrappers -----------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _gather_w(self, wm: HcuWeightMap, ref: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.float32, device=ref.device)
        _lib.check(self.lib.hcu_weight_gather(C.byref(wm), _ptr(ref), _ptr(out), self._stream()), "weight_gather")
        return out

    def _conv(self, d, x, wspec, bias=None, out=None, stats=None, in_scale=None, in_shift=None, out_scale=None,
              out_shift=None, layer=None, key=None, bn_fin=None):
        """One gather-convolution launch.  wspec = (HcuWeightMap, reference-layout parameter, packed element count).
        fp16 activations take the tcgen05 kernel whenever it supports the descriptor (weights gathered, folded and
        packed to fp16 UMMA tiles in one launch), everything else the FFMA kernel (fp32 [g][taps][cin][cout])."""
        lib = self.lib
        wm, ref, nw = wspec[:3]
        m = d.batch * d.out_size[0] * d.out_size[1] * d.out_size[2]
        esz_i = 4 if d.dtype_in == _lib.F32 else 2
        esz_o = 4 if d.dtype_out == _lib.F32 else 2
        nin = d.batch * d.in_size[0] * d.in_size[1] * d.in_size[2] * d.cin * d.groups
        nbytes = nin * esz_i + m * d.cout * d.groups * esz_o
        flops = 2 * m * d.cout * d.groups * d.cin * d.taps[0] * d.taps[1] * d.taps[2]
        if self.use_tc and d.dtype_in == _lib.F16 and lib.hcu_conv_tc_supported(C.byref(d)):
            cache, key = self._cache, key or layer
            if cache is not None and cache.ready and key in cache.pack_off:
                pk = C.c_void_p(cache.packed.data_ptr() + cache.pack_off[key])  # packed by this step's batch launch
            else:
                nb = lib.hcu_conv_tc_packed_bytes(C.byref(d))
                packed = torch.empty(nb, dtype=torch.uint8, device=x.device)
                _lib.check(lib.hcu_conv_tc_pack_ref(C.byref(d), C.byref(wm), _ptr(ref), _ptr(packed), self._stream()),
                           "conv_tc_pack_ref")
                pk = _ptr(packed)
                if cache is not None and not cache.ready and isinstance(wspec[3] if len(wspec) > 3 else None, str):
                    cache.pack_jobs[key] = (HcuConvDesc.from_buffer_copy(d), HcuWeightMap.from_buffer_copy(wm), wspec[3], nb)
            _lib.note(layer, nbytes, flops)
            if bn_fin is not None:  # BatchNorm finalize as the kernel's fused tail (returns True: no separate launch)
                _lib.check(lib.hcu_conv_tc_fwd_bn(C.byref(d), _ptr(x), pk, _ptr(bias), _ptr(in_scale), _ptr(in_shift),
                                                  _ptr(out_scale), _ptr(out_shift), _ptr(out), _ptr(stats),
                                                  C.byref(bn_fin), self._stream()), "conv_tc_fwd_bn")
                return True
            _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), _ptr(x), pk, _ptr(bias), _ptr(in_scale), _ptr(in_shift),
                                           _ptr(out_scale), _ptr(out_shift), _ptr(out), _ptr(stats), self._stream()),
                       "conv_tc_fwd")
            return False
        w = self._gather_w(wm, ref, nw)
        if d.dtype_in == _lib.F16:
            # fp16 path on the FFMA kernel (channel pitches that are not a multiple of 8): same arithmetic as the tensor-core
            # kernels -- fp16 operands (the packed weights are fp16 there), fp32 accumulate
            w = w.half().float()
        _lib.note(layer, nbytes, flops)
        _lib.check(lib.hcu_conv_fwd(C.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(in_scale), _ptr(in_shift),
                                    _ptr(out_scale), _ptr(out_shift), _ptr(out), _ptr(stats), self._stream()),
                   "conv_fwd")

    # ---- weight maps (reference layouts: Conv [Cout, Cin/g, kx, ky, kz]; ConvT [Cin, Cout, kx, ky, kz]) ----
    def _eff(self, g: ConvGeom, f16: bool):
        """(groups, cin, cout, bdiag) a conv is LAUNCHED with.  fp16 path: a grouped convolution (`groups=2` of the
        production model, `main.py:46-55`) runs as ONE dense tensor-core convolution over block-diagonal packed weights
        (HcuWeightMap.bdiag): its groups have 2 .. 64 channels, far below a tensor-core tile."""
        if f16 and self.use_tc and g.groups > 1:
            return 1, g.cin_t, g.cout_t, g.groups
        return g.groups, g.cin_g, g.cout_g, 0

    @staticmethod
    def _wm_conv_fwd(g: ConvGeom, bdiag: int = 0) -> HcuWeightMap:
        T = g.taps[0] * g.taps[1] * g.taps[2]
        rc = g.ref_cin_g
        return weight_map(1 if bdiag else g.groups, g.taps, g.cin_g, g.cout_g, sg=g.cout_g * rc * T, sa=T, sb=rc * T,
                          st=(g.taps[1] * g.taps[2], g.taps[2], 1), fold=int(g.fold), fold_stride=g.cin_g * T, bdiag=bdiag)

    @staticmethod
    def _wm_conv_dgrad(g: ConvGeom, bdiag: int = 0) -> HcuWeightMap:
        T = g.taps[0] * g.taps[1] * g.taps[2]
        rc = g.ref_cin_g
        return weight_map(1 if bdiag else g.groups, g.taps, g.cout_g, g.cin_g, sg=g.cout_g * rc * T, sa=rc * T, sb=T,
                          st=(g.taps[1] * g.taps[2], g.taps[2], 1), t0=tuple(t - 1 for t in g.taps),
                          tstep=(-1, -1, -1), fold=int(g.fold), fold_stride=g.cin_g * T, bdiag=bdiag)

    @staticmethod
    def _wm_up_phase(u: UpGeom, phi, J) -> HcuWeightMap:
        T = u.k[0] * u.k[1] * u.k[2]
        return weight_map(1, J, u.cin, u.cout, sg=0, sa=u.cout * T, sb=T, st=(u.k[1] * u.k[2], u.k[2], 1),
                          t0=tuple(phi[d] + u.s[d] * (J[d] - 1) for d in range(3)),
                          tstep=tuple(-u.s[d] for d in range(3)))

    @staticmethod
    def _up_fusable(u: UpGeom) -> bool:
        """All stride phases share one tap count (kernel % stride == 0) and every phase is a whole number of 8-channel
        groups: the transposed convolution runs as ONE stride-1 convolution with the phases folded into the channels."""
        return all(u.k[d] % u.s[d] == 0 for d in range(3)) and u.cout % 8 == 0 and u.cin % 8 == 0 and \
            u.s[0] * u.s[1] * u.s[2] > 1

    @staticmethod
    def _wm_up_fused(u: UpGeom, J) -> HcuWeightMap:
        """GEMM-B [taps J][cin][(phase, cout)]: W[ci][co][phi + s*(J-1-t)] (forward and weight gradient)."""
        T = u.k[0] * u.k[1] * u.k[2]
        st = (u.k[1] * u.k[2], u.k[2], 1)
        return weight_map(1, J, u.cin, u.cout, sg=0, sa=u.cout * T, sb=T, st=st,
                          t0=tuple(u.s[d] * (J[d] - 1) for d in range(3)), tstep=tuple(-u.s[d] for d in range(3)),
                          phase_on=2, ph=u.s, pst=st)

    @staticmethod
    def _wm_up_fused_dgrad(u: UpGeom, J) -> HcuWeightMap:
        """GEMM-B [taps J][(phase, cout)][cin]: W[ci][co][phi + s*t] (data gradient: a valid correlation over the phases)."""
        T = u.k[0] * u.k[1] * u.k[2]
        st = (u.k[1] * u.k[2], u.k[2], 1)
        return weight_map(1, J, u.cout, u.cin, sg=0, sa=T, sb=u.cout * T, st=st, tstep=u.s, phase_on=1, ph=u.s, pst=st)

    @staticmethod
    def _phase_word(s3) -> int:
        return int(s3[0]) | (int(s3[1]) << 8) | (int(s3[2]) << 16)

    @staticmethod
    def _wm_up_dgrad(u: UpGeom) -> HcuWeightMap:
        T = u.k[0] * u.k[1] * u.k[2]
        return weight_map(1, u.k, u.cout, u.cin, sg=0, sa=T, sb=u.cout * T, st=(u.k[1] * u.k[2], u.k[2], 1))

    # ---- forward ------------------------------------------------------------------------------
    # Data flow (training): every conv writes its RAW output y (+ per-channel sum / sum of squares from the fp32
    # accumulators); BatchNorm + ReLU of layer L are never materialised -- they are applied by whoever READS y_L
    # (the next conv's operand load, the max-pool pass, the weight-gradient's operand load).  Only y tensors, pooled
    # activations and the up-convolution outputs exist in HBM.
    def forward(self, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor], x: torch.Tensor,
                training: bool, save: bool, precision: str = "fp32", prelaid: bool = False):
        """Returns (logits [B, Cout, *spatial] fp32, saved-state or None).  ``prelaid``: ``x`` is the channels-last fp16 view
        ``hcunet_b200.loader.StackLoader.image`` produced (pitch 8, zero padding): the mixed path reads its storage as is."""
        lib, st = self.lib, self._stream()
        plan = self.plan(x.shape)
        dev = x.device
        act_dtype = _ACT_DTYPE[precision]
        cache = None
        if self.use_batch and self.use_tc and act_dtype == torch.float16:
            ckey = (tuple(x.shape), precision, bool(training), bool(save), dev.index)
            cache = self._caches.get(ckey)
            if cache is None:
                cache = self._caches[ckey] = _StepCache()
            if cache.ready and cache.ptr_sig != cache._sig(params, [j[2] for j in cache.pack_jobs.values()]):
                cache = self._caches[ckey] = _StepCache()   # parameters were re-allocated: record again
            if not cache.ready and cache.fwd_done and (cache.bwd_done or not save):
                cache.finalize(lib, params, dev)
            if cache.ready and cache.pack_table is not None:
                _lib.check(lib.hcu_conv_tc_pack_batch(_ptr(cache.pack_table), len(cache.pack_jobs), cache.pack_blocks,
                                                      C.c_void_p(cache.pack_base), _ptr(cache.packed), st),
                           "conv_tc_pack_batch")
        self._cache = cache
        adt = _DT[act_dtype]
        esz = 2 if act_dtype == torch.float16 else 4
        B = plan.batch
        S = plan.in_sz[0] * plan.in_sz[1] * plan.in_sz[2]
        # fp16 path: pad the input channels to a multiple of 8 (16-byte pixels) so the first conv is tensor-core too
        cp = -(-plan.in_channels // 8) * 8 if act_dtype == torch.float16 else plan.in_channels
        want = (S * cp, 1) + tuple(s * cp for s in (plan.in_sz[1] * plan.in_sz[2], plan.in_sz[2], 1)[:plan.dims])
        if prelaid and act_dtype == torch.float16 and x.dtype == torch.float16 and cp == 8 and tuple(x.stride()) == want \
                and x.data_ptr() % 16 == 0:
            cur = x.as_strided((B, S, cp), (S * cp, cp, 1))     # the loader's storage, no layout pass
        else:
            x = x.contiguous()
            if x.dtype not in _DT:
                x = x.float()
            cur = torch.empty((B, S, cp), dtype=act_dtype, device=dev)
            _lib.note("input", x.numel() * x.element_size() + cur.numel() * esz, 0)
            _lib.check(lib.hcu_nc_to_cl(_ptr(x), _DT[x.dtype], _ptr(cur), adt, B, plan.in_channels, S, cp, None, st),
                       "nc_to_cl")
        self._tap("input", cur, plan.in_channels, plan.in_sz)
        xf = None  # pending (scale, shift) + ReLU to apply when `cur` is read
        saved = [] if save else None
        fold_eval = not save and not training  # inference: BN folded into the conv epilogue
        SB = _lib.STAT_BINS  # binned fp64 reductions (include/hcunet_b200.h "HCU_STAT_BINS")
        nstat = sum(2 * g.cout_t * SB for g in plan.steps if isinstance(g, ConvGeom) and g.bn is not None)
        nbn = sum(1 for g in plan.steps if isinstance(g, ConvGeom) and g.bn is not None)
        # one memset per forward: the binned statistics + one 8-byte "last CTA" ticket counter per BatchNorm
        zero_ws = torch.zeros(nstat + nbn, dtype=torch.float64, device=dev) if training else None
        kbn = 0
        zoff = 0
        logits = None
        for g in plan.steps:
            if isinstance(g, UpGeom):
                out = self._up_forward(g, params, cur, cp, xf, B, act_dtype)
                if save:
                    saved.append(("up", g, cur, cp, xf))
                self._tap(g.name + ".out", out, g.cout, g.out_sz)
                cur, cp, xf = out, g.cout, None
                continue
            npix = B * g.out_sz[0] * g.out_sz[1] * g.out_sz[2]
            eg, ecin, ecout, bd = self._eff(g, act_dtype == torch.float16)
            w = (self._wm_conv_fwd(g, bd), params[g.name + ".weight"],
                 eg * g.taps[0] * g.taps[1] * g.taps[2] * ecin * ecout, g.name + ".weight")
            bias = params[g.name + ".bias"]
            isc, ish = (xf[0], xf[1]) if xf is not None else (None, None)
            if g.bn is None:  # out_conv: logits, fp32
                y = torch.empty((B, npix // B, g.cout_t), dtype=torch.float32, device=dev)
                d = conv_desc(adt, _lib.F32, B, g.in_sz, cp, 0, ecin, ecin, g.out_sz, g.out_sz, g.cout_t, 0,
                              ecout, eg, g.taps, g.dil, in_relu=int(xf is not None))
                self._conv(d, cur, w, bias, y, in_scale=isc, in_shift=ish, layer=g.name)
                self._tap("logits", y, g.cout_t, g.out_sz)
                if save:
                    saved.append(("out", g, cur, cp, xf))
                if g.cout_t == 1:
                    logits = y.view((B, 1) + tuple(g.out_sz[:plan.dims]))
                else:
                    logits = torch.empty((B, g.cout_t) + tuple(g.out_sz[:plan.dims]), dtype=torch.float32, device=dev)
                    _lib.check(lib.hcu_cl_to_nc(_ptr(y), _lib.F32, _ptr(logits), _lib.F32, B, g.cout_t, npix // B,
                                                g.cout_t, None, st), "cl_to_nc")
                break
            gamma, beta = params[g.bn + ".weight"], params[g.bn + ".bias"]
            rm, rv = buffers[g.bn + ".running_mean"], buffers[g.bn + ".running_var"]
            d = conv_desc(adt, adt, B, g.in_sz, cp, 0, ecin, ecin, g.out_sz, g.out_sz, g.cout_t, 0,
                          ecout, eg, g.taps, g.dil, in_relu=int(xf is not None))
            vec = torch.empty((4, g.cout_t), dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
            if fold_eval:
                _lib.check(lib.hcu_bn_eval_affine(g.cout_t, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), BN_EPS,
                                                  _ptr(bias), _ptr(vec[2]), _ptr(vec[3]), st), "bn_eval_affine")
                d.out_relu = 1
                a = torch.empty((B, npix // B, g.cout_t), dtype=act_dtype, device=dev)
                self._conv(d, cur, w, None, a, out_scale=vec[2], out_shift=vec[3], layer=g.name)
                self._tap(g.name + ".a", a, g.cout_t, g.out_sz)
                if g.pool is not None:
                    a, _ = self._pool(a, g, B, act_dtype, None, None, 0, want_argmax=False)
                    self._tap(g.name + ".pool", a, g.cout_t, g.pool_sz)
                cur, cp, xf = a, g.cout_t, None
                continue
            y = torch.empty((B, npix // B, g.cout_t), dtype=act_dtype, device=dev)
            if training:
                stats = zero_ws[zoff:zoff + 2 * g.cout_t * SB]
                zoff += 2 * g.cout_t * SB
                fin = _lib.HcuBnFin(float(npix), gamma.data_ptr(), beta.data_ptr(), BN_EPS, BN_MOMENTUM, rm.data_ptr(),
                                    rv.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(),
                                    vec[3].data_ptr(), zero_ws.data_ptr() + 8 * (nstat + kbn))
                kbn += 1
                fused = self._conv(d, cur, w, bias, y, stats=stats, in_scale=isc, in_shift=ish, layer=g.name, bn_fin=fin)
                if not fused:
                    _lib.check(lib.hcu_bn_finalize(_ptr(stats), g.cout_t, float(npix), _ptr(gamma), _ptr(beta), BN_EPS,
                                                   BN_MOMENTUM, _ptr(rm), _ptr(rv), _ptr(vec[0]), _ptr(vec[1]),
                                                   _ptr(vec[2]), _ptr(vec[3]), st), "bn_finalize")
            else:
                # eval-mode forward that must be differentiable: running statistics
                self._conv(d, cur, w, bias, y, in_scale=isc, in_shift=ish, layer=g.name)
                _lib.check(lib.hcu_bn_eval_affine(g.cout_t, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), BN_EPS, None,
                                                  _ptr(vec[2]), _ptr(vec[3]), st), "bn_eval_affine")
                vec[0].copy_(rm)
                vec[1].copy_(torch.rsqrt(rv + BN_EPS))
            self._tap(g.name + ".y", y, g.cout_t, g.out_sz)
            self._tap(g.name + ".bn", vec, g.cout_t, (1, 1, 1))
            argmax = None
            a_in, a_cp, a_xf = cur, cp, xf
            if g.pool is not None:
                cur, argmax = self._pool(y, g, B, act_dtype, vec[2], vec[3], 1, want_argmax=save)
                self._tap(g.name + ".pool", cur, g.cout_t, g.pool_sz)
                cp, xf = g.cout_t, None
            else:
                cur, cp, xf = y, g.cout_t, (vec[2], vec[3])
            if save:
                saved.append(("conv", g, a_in, a_cp, a_xf, y, vec, argmax))
        if training:
            nbt = [buffers[g.bn + ".num_batches_tracked"] for g in plan.steps
                   if isinstance(g, ConvGeom) and g.bn is not None]
            torch._foreach_add_(nbt, 1)
        if cache is not None:
            cache.fwd_done = True
        self._cache = None
        # the packed weights / job tables of the step cache are shared by every forward with the same key: a backward must
        # run against the parameters its forward saw (autograd's version check, which the detached views bypass)
        versions = tuple(p._version for p in params.values())
        return logits, (plan, saved, act_dtype, training, cache, versions)

    def _pool(self, y, g: ConvGeom, B, act_dtype, scale, shift, relu, want_argmax=True):
        adt = _DT[act_dtype]
        esz = 2 if act_dtype == torch.float16 else 4
        ps = g.pool_sz
        pooled = torch.empty((B, ps[0] * ps[1] * ps[2], g.cout_t), dtype=act_dtype, device=y.device)
        argmax = torch.empty((B, ps[0] * ps[1] * ps[2], g.cout_t), dtype=torch.uint8, device=y.device)
        _lib.note(g.name, y.numel() * esz + pooled.numel() * (esz + 1), 0)
        _lib.check(self.lib.hcu_bn_relu_maxpool(_ptr(y), adt, _ptr(pooled), adt, _ptr(argmax), B, g.out_sz[0],
                                                g.out_sz[1], g.out_sz[2], g.cout_t, g.pool[0], g.pool[1], g.pool[2],
                                                _ptr(scale), _ptr(shift), relu, self._stream()), "bn_relu_maxpool")
        return pooled, argmax

    @staticmethod
    def _phases(u: UpGeom):
        for px in range(u.s[0]):
            for py in range(u.s[1]):
                for pz in range(u.s[2]):
                    phi = (px, py, pz)
                    J = tuple(-(-(u.k[d] - phi[d]) // u.s[d]) for d in range(3))
                    Q = tuple(-(-(u.out_sz[d] - phi[d]) // u.s[d]) for d in range(3))
                    yield phi, J, Q

    def _up_forward(self, u: UpGeom, params, cur, cp, xf, B, act_dtype):
        """ConvTranspose (`unet.py:294-298,310`) as prod(stride) stride-1 sub-convolutions, one per output phase."""
        adt = _DT[act_dtype]
        wt, bias = params[u.name + ".weight"], params[u.name + ".bias"]
        out = torch.empty((B, u.out_sz[0] * u.out_sz[1] * u.out_sz[2], u.cout), dtype=act_dtype, device=cur.device)
        isc, ish = (xf[0], xf[1]) if xf is not None else (None, None)
        if self.use_tc and act_dtype == torch.float16 and self._up_fusable(u):
            nph = u.s[0] * u.s[1] * u.s[2]
            J = tuple(u.k[d] // u.s[d] for d in range(3))
            Q = tuple(u.out_sz[d] // u.s[d] for d in range(3))
            d = conv_desc(adt, adt, B, u.in_sz, cp, 0, u.cin, u.cin, Q, u.out_sz, u.cout, 0, nph * u.cout, 1, J,
                          pad=tuple(j - 1 for j in J), ostep=u.s, in_relu=int(xf is not None))
            d.ophase = self._phase_word(u.s)
            if self.lib.hcu_conv_tc_supported(C.byref(d)):
                w = (self._wm_up_fused(u, J), wt, nph * J[0] * J[1] * J[2] * u.cin * u.cout, u.name + ".weight")
                self._conv(d, cur, w, bias, out, in_scale=isc, in_shift=ish, layer=u.name, key=u.name + ".fused")
                return out
        for phi, J, Q in self._phases(u):
            w = (self._wm_up_phase(u, phi, J), wt, J[0] * J[1] * J[2] * u.cin * u.cout, u.name + ".weight")
            d = conv_desc(adt, adt, B, u.in_sz, cp, 0, u.cin, u.cin, Q, u.out_sz, u.cout, 0, u.cout, 1, J,
                          pad=tuple(j - 1 for j in J), ostep=u.s, ooff=phi, in_relu=int(xf is not None))
            self._conv(d, cur, w, bias, out, in_scale=isc, in_shift=ish, layer=u.name,
                       key=f"{u.name}.phase{phi[0]}{phi[1]}{phi[2]}")
        return out

    # ---- backward -----------------------------------------------------------------------------
    def backward(self, params: Dict[str, torch.Tensor], state, dlogits: torch.Tensor, need_dx: bool):
        """Returns ({param name: grad}, dx or None)."""
        lib, st = self.lib, self._stream()
        plan, saved, act_dtype, training, cache, versions = state
        if versions != tuple(p._version for p in params.values()):
            raise RuntimeError("hcunet_b200: a parameter was modified in place (e.g. optimizer.step()) between this forward and "
                               "its backward; the step's packed weights no longer match what the forward computed with")
        self._cache = cache
        batched = cache is not None and cache.ready and cache.scatter_table is not None
        # ONE flat fp32 buffer holds every parameter gradient of the step, in parameter order; the gradients handed to autograd
        # are views of it, so a data-parallel all-reduce runs on it in place (hcunet_b200.parallel.GradSync), no gather copies
        self._goff, acc = {}, 0
        for name, p in params.items():
            self._goff[name] = (acc, p.shape)
            acc += p.numel()
        self._gflat = torch.empty(acc, dtype=torch.float32, device=dlogits.device)
        self.last_grad_flat = self._gflat
        side = None
        if batched and self.overlap_wgrad and (_lib._ProfState.profiler is None or os.environ.get("HCUNET_PROFILE_OVERLAP") == "1"):
            if self._side is None or self._side.device != dlogits.device:
                self._side = torch.cuda.Stream(device=dlogits.device)
                self._side2 = torch.cuda.Stream(device=dlogits.device)
            side = self._side
            side.wait_stream(torch.cuda.current_stream())
        self._wstream = side
        self._wflip = 0
        if batched:
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                cache.wacc.zero_()
            if side is not None and self.n_side > 1:
                self._side2.wait_stream(side)   # the second gradient stream starts behind the workspace memset
        adt = _DT[act_dtype]
        esz = 2 if act_dtype == torch.float16 else 4
        dev = dlogits.device
        B = plan.batch
        grads: Dict[str, torch.Tensor] = {}
        scratch = torch.empty(4096, dtype=torch.float64, device=dev)
        dlogits = dlogits.contiguous().float()
        So = plan.out_sz[0] * plan.out_sz[1] * plan.out_sz[2]
        co = plan.out_channels
        # fp16 backward: scale dlogits by a device-computed power of two S (no host sync); every parameter
        # gradient is multiplied by 1/S where it is written (inv), so nothing downstream sees S
        scl = inv = None
        if act_dtype == torch.float16:
            scales = torch.empty(2, dtype=torch.float32, device=dev)
            _lib.check(lib.hcu_grad_scale(_ptr(dlogits), dlogits.numel(), GRAD_SCALE_TARGET,
                                          _ptr(scratch.view(torch.int32)), _ptr(scales), st), "grad_scale")
            scl, inv = scales[0:1], scales[1:2]
        dcur_cp = co
        if co == 1 and scl is None:
            dcur = dlogits.view(B, So, 1)
            dcur_dt = _lib.F32
        else:
            # fp16: channel pitch padded to 8 (zeros) so the out_conv gradients take the tensor-core kernels too
            dcur_cp = -(-co // 8) * 8 if act_dtype == torch.float16 else co
            dcur = torch.empty((B, So, dcur_cp), dtype=act_dtype, device=dev)
            _lib.check(lib.hcu_nc_to_cl(_ptr(dlogits), _lib.F32, _ptr(dcur), adt, B, co, So, dcur_cp, _ptr(scl), st),
                       "nc_to_cl")
            dcur_dt = adt
        self._tap("dlogits", dcur, co, plan.out_sz)
        self._inv = inv
        dx = None
        split_done = False
        SB = _lib.STAT_BINS
        nstat = sum(2 * it[1].cout_t * SB for it in saved if it[0] == "conv")
        nbn = sum(1 for it in saved if it[0] == "conv")
        # one memset for every BN-backward reduction + one 8-byte ticket counter per BatchNorm
        zero_ws = torch.zeros(nstat + nbn, dtype=torch.float64, device=dev)
        kbn = 0
        zoff = 0
        # bucket split of the data-parallel overlap: the down levels below the two deepest are the LATE bucket (the flat
        # buffer is in parameter order out_conv | down_steps.* | up_steps.*, so late = [0, offset of the first early level))
        hook = self.grad_ready_hook
        split_name, split_off, hook_events = None, 0, []
        if hook is not None:
            nlev = len(self.spec["feature_sizes"])
            first_early = f"down_steps.{max(0, nlev - 2)}."
            offs = [o for n_, (o, _) in self._goff.items() if n_.startswith(first_early)]
            late_names = [n_ for n_, (o, _) in self._goff.items() if offs and o < min(offs)]
            if offs and nlev > 2 and all(n_.startswith("out_conv") or n_.startswith("down_steps.") for n_ in late_names):
                split_off = min(offs)
                split_name = f"down_steps.{nlev - 3}.conv2"      # the first late layer the backward reaches
        items = list(reversed(saved))
        prefused: Dict[str, tuple] = {}      # conv name -> BN-backward context whose statistics the producer already computed

        def bn_ctx(g_, npix_):
            """(sums, coef, fin, dgamma, dbeta, dbias) of one BatchNorm backward; slices of this backward's zeroed workspace."""
            nonlocal zoff, kbn
            sums_ = zero_ws[zoff:zoff + 2 * g_.cout_t * SB]
            zoff += 2 * g_.cout_t * SB
            dgamma_, dbeta_ = self._gview(g_.bn + ".weight"), self._gview(g_.bn + ".bias")
            dbias_ = self._gview(g_.name + ".bias")
            coef_ = torch.empty((3, g_.cout_t), dtype=torch.float32, device=dev)
            fin_ = _lib.HcuBnBwdFin(float(npix_), params[g_.bn + ".weight"].data_ptr(), 1 if training else 0, 1.0,
                                    inv.data_ptr() if inv is not None else None, dgamma_.data_ptr(), dbeta_.data_ptr(),
                                    dbias_.data_ptr(), coef_.data_ptr(), zero_ws.data_ptr() + 8 * (nstat + kbn))
            kbn += 1
            return sums_, coef_, fin_, dgamma_, dbeta_, dbias_

        for idx, item in enumerate(items):
            kind = item[0]
            if split_name is not None and kind == "conv" and item[1].name == split_name:
                # ---- early bucket complete: up path + the two deepest levels ----
                if cache is not None and not cache.ready and cache.split_at is None:
                    cache.split_at = len(cache.scatter_jobs)
                evs = []
                if batched and cache.scatter_tables2 is not None:
                    tab, k, blk = cache.scatter_tables2[0]
                    with torch.cuda.stream(side) if side is not None else _NullCtx():
                        _lib.check(lib.hcu_weight_scatter_batch(_ptr(tab), k, blk, _ptr(cache.wacc), 1.0, _ptr(inv),
                                                                _ptr(self._gflat), self._stream()), "weight_scatter_batch")
                        if side is not None:
                            e_s = torch.cuda.Event()
                            e_s.record()
                            evs.append(e_s)
                    e_m = torch.cuda.Event()
                    e_m.record()
                    evs.append(e_m)
                    done = hook(self._gflat, split_off, self._gflat.numel(), evs)
                    if done is not None:
                        hook_events.append(done)
                    split_done = True
                elif not batched:
                    e_m = torch.cuda.Event()
                    e_m.record()
                    done = hook(self._gflat, split_off, self._gflat.numel(), [e_m])
                    if done is not None:
                        hook_events.append(done)
                    split_done = True
            if kind == "out":
                _, g, a_in, a_cp, a_xf = item
                npix = B * So
                grads[g.name + ".bias"] = self._colsum(dcur, dcur_dt, npix, co, scratch, cpitch=dcur_cp,
                                                       out=self._gview(g.name + ".bias"))
                grads[g.name + ".weight"] = self._wgrad_conv(g, a_in, a_cp, a_xf, adt, dcur, dcur_dt, B,
                                                             params[g.name + ".weight"], dy_cp=dcur_cp)
                dcur = self._dgrad_conv(g, dcur, dcur_dt, B, params[g.name + ".weight"], act_dtype, dy_cp=dcur_cp)
                dcur_dt = adt
                self._tap(g.name + ".dgrad", dcur, g.cin_t, g.in_sz)
            elif kind == "conv":
                _, g, a_in, a_cp, a_xf, y, vec, argmax = item
                npix = B * g.out_sz[0] * g.out_sz[1] * g.out_sz[2]
                pool_arg, pool_geom = None, None
                if (argmax is not None and act_dtype == torch.float16 and g.cout_t % 8 == 0 and 256 % (g.cout_t // 8) == 0
                        and g.cout_t <= 512 and dcur_dt == _lib.F16):  # the h8 kernels take up to 512 channels
                    # fp16: the max-pool backward is fused into the two BN-backward passes (no full-resolution dA)
                    pool_arg = argmax
                    pool_geom = _lib.HcuPoolGeom(B, g.out_sz[0], g.out_sz[1], g.out_sz[2], g.pool[0], g.pool[1], g.pool[2])
                elif argmax is not None:
                    dfull = torch.empty((B, npix // B, g.cout_t), dtype=act_dtype, device=dev)
                    _lib.note(g.name, dcur.numel() * (esz + 1) + dfull.numel() * esz, 0)
                    _lib.check(lib.hcu_maxpool_bwd(_ptr(dcur), dcur_dt, _ptr(argmax), _ptr(dfull), adt, B, g.out_sz[0],
                                                   g.out_sz[1], g.out_sz[2], g.cout_t, g.pool[0], g.pool[1],
                                                   g.pool[2], st), "maxpool_bwd")
                    dcur, dcur_dt = dfull, adt
                ctx = prefused.pop(g.name, None)
                if ctx is None:
                    sums, coef, fin, dgamma, dbeta, dbias = bn_ctx(g, npix)
                    _lib.note(g.name, 2 * npix * g.cout_t * esz, 0)
                    _lib.check(lib.hcu_bn_bwd_stats_fin(_ptr(dcur), dcur_dt, _ptr(y), adt, npix, g.cout_t, _ptr(vec[2]),
                                                        _ptr(vec[3]), _ptr(vec[0]), _ptr(vec[1]), 1, _ptr(pool_arg),
                                                        C.byref(pool_geom) if pool_geom is not None else None, _ptr(sums),
                                                        C.byref(fin), st), "bn_bwd_stats_fin")
                else:   # the data gradient that produced `dcur` computed these statistics in its epilogue
                    sums, coef, fin, dgamma, dbeta, dbias = ctx
                grads[g.bn + ".weight"], grads[g.bn + ".bias"], grads[g.name + ".bias"] = dgamma, dbeta, dbias
                if (g.first and not need_dx and argmax is None and self.tap is None and self.fuse_apply
                        and dcur_dt == _lib.F16 and act_dtype == torch.float16
                        and self._wgrad_conv(g, a_in, a_cp, a_xf, adt, dcur, adt, B, params[g.name + ".weight"], probe=True)):
                    # nobody needs this layer's data gradient: BatchNorm backward's apply pass is fused into the staging of
                    # the weight gradient's dy operand (wgrad_rows.cu), the gradient tensor is never written
                    grads[g.name + ".weight"] = self._wgrad_conv(g, a_in, a_cp, a_xf, adt, dcur, adt, B, params[g.name + ".weight"],
                                                                 bnb=(y, vec[2], vec[3], coef))
                    dcur = None
                    continue
                dy = torch.empty((B, npix // B, g.cout_t), dtype=act_dtype, device=dev)
                _lib.note(g.name, 3 * npix * g.cout_t * esz, 0)
                _lib.check(lib.hcu_bn_bwd_apply(_ptr(dcur), dcur_dt, _ptr(y), adt, _ptr(dy), adt, npix, g.cout_t,
                                                _ptr(vec[2]), _ptr(vec[3]), 1, _ptr(coef), _ptr(pool_arg),
                                                C.byref(pool_geom) if pool_geom is not None else None, st), "bn_bwd_apply")
                self._tap(g.name + ".dy", dy, g.cout_t, g.out_sz)
                grads[g.name + ".weight"] = self._wgrad_conv(g, a_in, a_cp, a_xf, adt, dy, adt, B,
                                                             params[g.name + ".weight"])
                if g.first and not need_dx:
                    dcur = None
                else:
                    # the next item is the conv whose output this conv reads (conv1 of the same block, no pool in between):
                    # its BatchNorm-backward statistics are sums over exactly the tensor this data gradient writes
                    fuse = None
                    nxt = items[idx + 1] if idx + 1 < len(items) else None
                    if (self.fuse_bnbwd and self.tap is None and act_dtype == torch.float16 and nxt is not None and nxt[0] == "conv"
                            and nxt[5] is a_in and nxt[7] is None and a_xf is not None and not g.first):
                        g1, y1, vec1 = nxt[1], nxt[5], nxt[6]
                        npix1 = B * g1.out_sz[0] * g1.out_sz[1] * g1.out_sz[2]
                        fuse = (g1, y1, vec1, npix1)
                    dcur = self._dgrad_conv(g, dy, adt, B, params[g.name + ".weight"], act_dtype,
                                            out_cp=a_cp if g.first else None, fuse=fuse, bn_ctx=bn_ctx, prefused=prefused)
                    dcur_dt = adt
                    self._tap(g.name + ".dgrad", dcur, g.cin_t, g.in_sz)
                if g.first and need_dx:
                    S = plan.in_sz[0] * plan.in_sz[1] * plan.in_sz[2]
                    dx = torch.empty((B, plan.in_channels) + tuple(plan.in_sz[:plan.dims]), dtype=torch.float32,
                                     device=dev)
                    _lib.check(lib.hcu_cl_to_nc(_ptr(dcur), adt, _ptr(dx), _lib.F32, B, plan.in_channels, S,
                                                a_cp, _ptr(inv), st), "cl_to_nc")
            else:  # "up"
                _, u, a_in, a_cp, a_xf = item
                npix_out = B * u.out_sz[0] * u.out_sz[1] * u.out_sz[2]
                grads[u.name + ".bias"] = self._colsum(dcur, dcur_dt, npix_out, u.cout, scratch, out=self._gview(u.name + ".bias"))
                T = u.k[0] * u.k[1] * u.k[2]
                m = B * u.in_sz[0] * u.in_sz[1] * u.in_sz[2]
                wt = params[u.name + ".weight"]
                dprev = torch.empty((B, m // B, u.cin), dtype=act_dtype, device=dev)
                done = False
                if self.use_tc and act_dtype == torch.float16 and dcur_dt == _lib.F16 and self._up_fusable(u):
                    # stride phases folded into the channels: both gradients on the tensor-core kernels, one launch each
                    nph = u.s[0] * u.s[1] * u.s[2]
                    J = tuple(u.k[d] // u.s[d] for d in range(3))
                    Q = tuple(u.out_sz[d] // u.s[d] for d in range(3))
                    isc, ish = (a_xf[0], a_xf[1]) if a_xf is not None else (None, None)
                    dw = conv_desc(adt, adt, B, u.in_sz, a_cp, 0, u.cin, u.cin, Q, u.out_sz, nph * u.cout, 0, nph * u.cout,
                                   1, J, pad=tuple(j - 1 for j in J), ostep=u.s, in_relu=int(a_xf is not None))
                    dw.ophase = self._phase_word(u.s)
                    dd = conv_desc(adt, adt, B, Q, nph * u.cout, 0, nph * u.cout, nph * u.cout, u.in_sz, u.in_sz, u.cin, 0,
                                   u.cin, 1, J)
                    dd.iphase = self._phase_word(u.s)
                    if (lib.hcu_conv_wgrad_tc_supported(C.byref(dw)) or lib.hcu_conv_wgrad_tc5_supported(C.byref(dw))) and \
                            lib.hcu_conv_tc_supported(C.byref(dd)):
                        total = nph * J[0] * J[1] * J[2] * u.cin * u.cout
                        note = (u.name, (npix_out * u.cout + m * u.cin) * esz, 2 * m * T * u.cin * u.cout)
                        grads[u.name + ".weight"] = self._wgrad_dispatch(u.name + ".weight", wt, dw, a_in, isc, ish, dcur,
                                                                         self._wm_up_fused(u, J), total, 1, note)
                        w = (self._wm_up_fused_dgrad(u, J), wt, total, u.name + ".weight")
                        self._conv(dd, dcur, w, None, dprev, layer=u.name + ".dgrad")
                        done = True
                if not done:
                    # weight gradient: R[t][co][ci] = sum_i dOut[i*s + t][co] * act(In[i])[ci]
                    # the gather side (a) is dOut, the dense side (b) is the up-conv's input: its pending BN+ReLU has to
                    # be materialised once for the dense side
                    a_act = self._materialise(a_in, a_cp, a_xf, m, u.cin, act_dtype)
                    d = conv_desc(dcur_dt, adt, B, u.out_sz, u.cout, 0, u.cout, u.cout, u.in_sz, u.in_sz, u.cin, 0, u.cin, 1,
                                  u.k, istep=u.s)
                    roles = T * (-(-u.cout // 8)) * (-(-u.cin // 8))
                    ns = _nsplit(m, roles)
                    total = T * u.cout * u.cin
                    wm = self._wm_up_dgrad(u)
                    note = (u.name, (npix_out * u.cout + m * u.cin) * esz, 2 * m * T * u.cin * u.cout)
                    grads[u.name + ".weight"] = self._wgrad_dispatch(u.name + ".weight", wt, d, dcur, None, None, a_act, wm,
                                                                     total, ns, note)
                    # data gradient: strided gather convolution over dOut
                    w = (wm, wt, total, u.name + ".weight")
                    d2 = conv_desc(dcur_dt, adt, B, u.out_sz, u.cout, 0, u.cout, u.cout, u.in_sz, u.in_sz, u.cin, 0, u.cin,
                                   1, u.k, istep=u.s)
                    self._conv(d2, dcur, w, None, dprev, layer=u.name + ".dgrad")
                self._tap(u.name + ".dgrad", dprev, u.cin, u.in_sz)
                dcur, dcur_dt = dprev, adt
        if batched:
            if side is not None and self.n_side > 1:
                side.wait_stream(self._side2)
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                if split_done and cache.scatter_tables2 is not None:   # the early part was scattered at the split
                    tab, k, blk = cache.scatter_tables2[1]
                    _lib.check(lib.hcu_weight_scatter_batch(_ptr(tab), k, blk, _ptr(cache.wacc), 1.0, _ptr(inv),
                                                            _ptr(self._gflat), self._stream()), "weight_scatter_batch")
                else:
                    _lib.check(lib.hcu_weight_scatter_batch(_ptr(cache.scatter_table), len(cache.scatter_jobs),
                                                            cache.scatter_blocks, _ptr(cache.wacc), 1.0, _ptr(inv),
                                                            _ptr(self._gflat), self._stream()), "weight_scatter_batch")
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
        if hook is not None:
            e_m = torch.cuda.Event()
            e_m.record()
            done = hook(self._gflat, 0, split_off if split_done else self._gflat.numel(), [e_m])
            if done is not None:
                hook_events.append(done)
            for ev in hook_events:
                torch.cuda.current_stream().wait_event(ev)
        if cache is not None:
            cache.bwd_done = True
        self._keep.clear()
        self._cache, self._gflat, self._wstream = None, None, None
        self.last_grad_names = list(grads.keys())
        return grads, dx

    def _materialise(self, t, cp, xf, npix, c, act_dtype):
        """relu(t * scale + shift) as a tensor (only where a kernel cannot apply the pending transform on load)."""
        if xf is None:
            return t
        adt = _DT[act_dtype]
        a = torch.empty_like(t)
        _lib.check(self.lib.hcu_bn_relu_apply(_ptr(t), adt, _ptr(a), adt, npix, cp, _ptr(xf[0]), _ptr(xf[1]), 1,
                                              self._stream()), "bn_relu_apply")
        return a

    def _gview(self, name):
        """The slice of the step's flat gradient buffer that belongs to parameter ``name``."""
        off, shape = self._goff[name]
        n = 1
        for s in shape:
            n *= s
        return self._gflat[off:off + n].view(shape)

    def _colsum(self, x, dt, npix, c, scratch, cpitch=None, out=None):
        """Per-channel sum (bias gradients of the layers without a BatchNorm behind them).  Only a gradient comes out of
        it, so with a side stream it leaves the data-gradient chain like the weight gradients do."""
        if out is None:
            out = torch.empty(c, dtype=torch.float32, device=x.device)
        side = getattr(self, "_wstream", None)
        if side is not None:
            ev = torch.cuda.Event()
            ev.record()
            side.wait_event(ev)
            self._keep.append(x)
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            _lib.check(self.lib.hcu_colsum(_ptr(x), dt, npix, cpitch or c, 0, c, 1.0, _ptr(self._inv), _ptr(scratch),
                                           _ptr(out), self._stream()), "colsum")
        return out

    def _wgrad_conv(self, g: ConvGeom, a_in, a_cp, a_xf, a_dt, dy, dy_dt, B, wref, dy_cp=None, bnb=None, probe=False):
        T = g.taps[0] * g.taps[1] * g.taps[2]
        eg, ecin, ecout, bd = self._eff(g, a_dt == _lib.F16 and dy_dt == _lib.F16)
        d = conv_desc(a_dt, dy_dt, B, g.in_sz, a_cp, 0, ecin, ecin, g.out_sz, g.out_sz, dy_cp or g.cout_t, 0, ecout,
                      eg, g.taps, g.dil, in_relu=int(a_xf is not None))
        m = B * g.out_sz[0] * g.out_sz[1] * g.out_sz[2]
        roles = eg * T * (-(-ecin // 8)) * (-(-ecout // 8))
        ns = _nsplit(m, roles)
        total = eg * T * ecin * ecout
        esz = 4 if a_dt == _lib.F32 else 2
        nin = B * g.in_sz[0] * g.in_sz[1] * g.in_sz[2] * g.cin_t
        note = (g.name, (nin + m * g.cout_t) * esz, 2 * m * T * g.cin_g * g.cout_g * g.groups)
        isc, ish = (a_xf[0], a_xf[1]) if a_xf is not None else (None, None)
        if probe:   # would the row-stacked kernel with the fused BatchNorm-backward apply take this layer?
            return self._rows_takes(d) and a_xf is None and bool(self.lib.hcu_conv_wgrad_rows_bnb_supported(C.byref(d)))
        if bnb is not None:  # the fused launch also reads y: the apply pass it replaces read g and y and wrote dy
            note = (note[0], note[1] + m * g.cout_t * esz, note[2])
        return self._wgrad_dispatch(g.name + ".weight", wref, d, a_in, isc, ish, dy, self._wm_conv_fwd(g, bd), total, ns, note,
                                    bnb=bnb)

    def _rows_takes(self, d) -> bool:
        f16 = self.use_tc and d.dtype_in == _lib.F16 and d.dtype_out == _lib.F16
        if d.ophase and d.in_cpitch > 16:   # transposed convs from 32 input channels: wgrad_tc5 is faster (35 vs 24 us on up_steps.2)
            return False
        return bool(f16 and self.use_rows and d.in_cpitch <= self.rows_maxcp and d.out_cpitch <= self.rows_maxcp_out and
                    self.lib.hcu_conv_wgrad_rows_supported(C.byref(d)))

    def _wgrad_dispatch(self, wname, wref, d, a, isc, ish, b, wm, total, ns, note, bnb=None):
        """Weight gradient of one conv: tensor-core kernel when it takes the descriptor, else the FFMA split-K kernel;
        result scattered into the reference layout.  With a ready step cache the kernel runs on the side stream
        (overlapping the data-gradient chain), accumulates into the persistent workspace and the scatter is left to
        the one batched launch at the end of the backward."""
        lib, cache = self.lib, self._cache
        f16 = self.use_tc and d.dtype_in == _lib.F16 and d.dtype_out == _lib.F16
        # channel-rich levels: tcgen05 kernel (M = 128 rows of Cin would be mostly padding below 32 input channels)
        # channel-poor levels: rows of the image stacked on both MMA dimensions, TMA-fed (wgrad_rows.cu)
        rows = self._rows_takes(d)
        assert rows or bnb is None
        tc5 = bool(f16 and not rows and self.use_tc5 and d.in_cpitch >= 32 and
                   lib.hcu_conv_wgrad_tc5_supported(C.byref(d)))
        # 8/16-channel levels the row kernel does not take: warp-specialised mma.sync pipeline
        ws = bool(f16 and not rows and not tc5 and self.use_ws and lib.hcu_conv_wgrad_ws_supported(C.byref(d)))
        tc = bool(f16 and (rows or tc5 or ws or lib.hcu_conv_wgrad_tc_supported(C.byref(d))))
        acc_fn, acc_name = ((lib.hcu_conv_wgrad_rows_acc, "wgrad_rows") if rows else
                            (lib.hcu_conv_wgrad_tc5_acc, "wgrad_tc5") if tc5 else (lib.hcu_conv_wgrad_ws_acc, "wgrad_ws"))
        if bnb is not None:
            by, bsc, bsh, bcoef = bnb

            def acc_fn(dd, pa, psc, psh, pb, pacc, stream):   # noqa: F811 -- same call shape as the plain kernels
                return lib.hcu_conv_wgrad_rows_bnb_acc(dd, pa, psc, psh, pb, _ptr(by), _ptr(bsc), _ptr(bsh), _ptr(bcoef), pacc, stream)
            acc_name = "wgrad_rows_bnb"
        nsplit = 1 if tc else ns
        if cache is not None and cache.ready and wname in cache.part_off and cache.scatter_jobs[wname][1:] == (nsplit, total):
            off = cache.part_off[wname]
            part = cache.wacc[off:off + nsplit * total]
            gw = self._gview(wname)
            side = self._wstream
            if side is not None:
                if self.n_side > 1:  # alternate two gradient streams: the small deep-level kernels overlap each other too
                    self._wflip ^= 1
                    side = self._side2 if self._wflip else side
                ev = torch.cuda.Event()
                ev.record()
                side.wait_event(ev)
                self._keep.extend(t for t in (a, b, isc, ish) if t is not None)  # alive until the streams join
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                _lib.note(*note)
                if rows or tc5 or ws:
                    _lib.check(acc_fn(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(part), self._stream()), acc_name)
                elif tc:
                    _lib.check(lib.hcu_conv_wgrad_tc_acc(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(part),
                                                         self._stream()), "wgrad_tc")
                else:
                    _lib.check(lib.hcu_conv_wgrad_partial(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(part),
                                                          nsplit, self._stream()), "wgrad")
            return gw
        partial = torch.empty((nsplit, total), dtype=torch.float32, device=wref.device)
        gw = self._gview(wname)
        _lib.note(*note)
        if rows or tc5 or ws:
            partial.zero_()
            _lib.check(acc_fn(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(partial), self._stream()), acc_name)
        elif tc:
            _lib.check(lib.hcu_conv_wgrad_tc(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(partial),
                                             self._stream()), "wgrad_tc")
        else:
            _lib.check(lib.hcu_conv_wgrad_partial(C.byref(d), _ptr(a), _ptr(isc), _ptr(ish), _ptr(b), _ptr(partial),
                                                  nsplit, self._stream()), "wgrad")
        _lib.check(lib.hcu_weight_scatter(C.byref(wm), _ptr(partial), nsplit, total, 1.0, _ptr(self._inv), 0, _ptr(gw),
                                          self._stream()), "weight_scatter")
        if cache is not None and not cache.ready:
            cache.scatter_jobs[wname] = (HcuWeightMap.from_buffer_copy(wm), nsplit, total)
        return gw

    def _dgrad_conv(self, g: ConvGeom, dy, dy_dt, B, wref, act_dtype, out_cp=None, dy_cp=None, fuse=None, bn_ctx=None,
                    prefused=None):
        adt = _DT[act_dtype]
        T = g.taps[0] * g.taps[1] * g.taps[2]
        eg, ecin, ecout, bd = self._eff(g, dy_dt == _lib.F16 and act_dtype == torch.float16)
        w = (self._wm_conv_dgrad(g, bd), wref, eg * T * ecin * ecout, g.name + ".weight")
        cpo = out_cp or g.cin_t
        alloc = torch.zeros if cpo != g.cin_t else torch.empty
        dprev = alloc((B, g.in_sz[0] * g.in_sz[1] * g.in_sz[2], cpo), dtype=act_dtype, device=dy.device)
        pad = tuple((g.taps[i] - 1) * g.dil[i] for i in range(3))
        d = conv_desc(dy_dt, adt, B, g.out_sz, dy_cp or g.cout_t, 0, ecout, ecout, g.in_sz, g.in_sz, cpo, 0, ecin,
                      eg, g.taps, g.dil, pad=pad)
        lib, cache, key = self.lib, self._cache, g.name + ".dgrad"
        if (fuse is not None and self.use_tc and cache is not None and cache.ready and key in cache.pack_off
                and lib.hcu_conv_tc_bnbwd_supported(C.byref(d))):
            # BatchNorm-backward statistics of the layer `dprev` is the gradient of, computed in this launch's epilogue
            g1, y1, vec1, npix1 = fuse
            ctx = bn_ctx(g1, npix1)
            sums, coef, fin = ctx[0], ctx[1], ctx[2]
            pk = C.c_void_p(cache.packed.data_ptr() + cache.pack_off[key])
            esz = 2
            nin = B * g.out_sz[0] * g.out_sz[1] * g.out_sz[2] * g.cout_t
            _lib.note(key, (nin + 2 * npix1 * g1.cout_t) * esz, 2 * npix1 * g1.cout_t * T * g.cout_t)
            _lib.check(lib.hcu_conv_tc_fwd_bnbwd(C.byref(d), _ptr(dy), pk, _ptr(dprev), _ptr(y1), _ptr(vec1[2]), _ptr(vec1[3]),
                                                 _ptr(vec1[0]), _ptr(vec1[1]), _ptr(sums), C.byref(fin), self._stream()),
                       "conv_tc_fwd_bnbwd")
            prefused[g1.name] = ctx
            self._keep.append(coef)
            return dprev
        self._conv(d, dy, w, None, dprev, layer=key)
        return dprev
