"""ctypes binding of the C ABI declared in ``include/hcunet_b200.h``.

There is NO fallback: if ``libhcunet_b200.so`` is missing or was built for another ABI version the
import of any compute path raises.  ``hcunet_b200.build.build()`` (or ``python __graft_entry__.py``)
produces the library in-tree.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HCUNET_LIB: an alternative build of the SAME library (A/B experiments: tools/build_variant.py); never a fallback
LIB_PATH = os.environ.get("HCUNET_LIB") or os.path.join(HERE, "libhcunet_b200.so")
ABI_VERSION = 24

F32, BF16, F16, U8, U16, F64 = 0, 1, 2, 3, 4, 5
BATCH_JOB_BYTES = 256
STAT_BINS = 4


class HcuConvDesc(C.Structure):
    _fields_ = [
        ("dtype_in", C.c_int32), ("dtype_out", C.c_int32), ("batch", C.c_int32),
        ("in_size", C.c_int32 * 3), ("in_cpitch", C.c_int32), ("in_c_off", C.c_int32),
        ("in_c_gstep", C.c_int32), ("cin", C.c_int32),
        ("out_size", C.c_int32 * 3), ("out_tsize", C.c_int32 * 3), ("out_cpitch", C.c_int32),
        ("out_c_off", C.c_int32), ("cout", C.c_int32), ("groups", C.c_int32),
        ("taps", C.c_int32 * 3), ("dil", C.c_int32 * 3), ("pad", C.c_int32 * 3),
        ("istep", C.c_int32 * 3), ("ostep", C.c_int32 * 3), ("ooff", C.c_int32 * 3),
        ("in_relu", C.c_int32), ("out_relu", C.c_int32), ("ophase", C.c_int32), ("iphase", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


class HcuWeightMap(C.Structure):
    _fields_ = [
        ("groups", C.c_int32), ("j", C.c_int32 * 3), ("na", C.c_int32), ("nb", C.c_int32),
        ("base", C.c_int64), ("sg", C.c_int64), ("sa", C.c_int64), ("sb", C.c_int64), ("st", C.c_int64 * 3),
        ("t0", C.c_int32 * 3), ("tstep", C.c_int32 * 3), ("fold", C.c_int32), ("fold_stride", C.c_int64),
        ("phase_on", C.c_int32), ("ph", C.c_int32 * 3), ("bdiag", C.c_int32), ("pst", C.c_int64 * 3),
    ]


class HcuBnFin(C.Structure):
    _fields_ = [("count", C.c_double), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
                ("momentum", C.c_float), ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("mean", C.c_void_p),
                ("invstd", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("counter", C.c_void_p)]


class HcuBnBwdFin(C.Structure):
    _fields_ = [("count", C.c_double), ("gamma", C.c_void_p), ("training", C.c_int32), ("grad_scale", C.c_float),
                ("dscale", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
                ("coef", C.c_void_p), ("counter", C.c_void_p)]


class HcuPoolGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n", "ix", "iy", "iz", "px", "py", "pz")]


class HcuTileGeom(C.Structure):
    _fields_ = [("channels", C.c_int32), ("size", C.c_int32 * 3), ("pad", C.c_int32 * 3), ("origin", C.c_int32 * 3),
                ("extent", C.c_int32 * 3), ("stack_origin", C.c_int32 * 3), ("stack_size", C.c_int32 * 3)]


class HcuLossDesc(C.Structure):
    _fields_ = [
        ("b", C.c_int32), ("c", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32),
        ("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32),
        ("dtype_mask", C.c_int32), ("dtype_pwl", C.c_int32), ("mode", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


P = C.c_void_p
I32, I64, F, D = C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> argtypes; every function returns int (0 == ok) unless listed in _RESTYPES
SIGNATURES = {
    "hcu_abi_version": [],
    "hcu_last_error": [],
    "hcu_launch_count": [],
    "hcu_zero": [P, C.c_size_t, P],
    "hcu_h2d_tile": [P, I64, I64, I64, I64, I64, P, P],
    "hcu_conv_tc_fwd_bnbwd": [C.POINTER(HcuConvDesc), P, P, P, P, P, P, P, P, P, C.POINTER(HcuBnBwdFin), P],
    "hcu_conv_tc_bnbwd_supported": [C.POINTER(HcuConvDesc)],
    "hcu_tile_gather": [C.POINTER(HcuTileGeom), P, I32, P, I32, I32, I32, P],
    "hcu_tile_flags": [C.POINTER(HcuTileGeom), P, I32, P, P],
    "hcu_sigmoid_paste": [P, C.POINTER(I32), C.POINTER(I32), C.POINTER(I32), P, I32, C.POINTER(I32), C.POINTER(I32), F, P, P],
    "hcu_load_stack": [P, I32, I64, I32, I32, I32, I32, C.POINTER(D), C.POINTER(D), P, I32, P],
    "hcu_load_labels": [P, I32, I64, I32, I32, I32, I32, I32, I32, P, P],
    "hcu_conv_fwd": [C.POINTER(HcuConvDesc), P, P, P, P, P, P, P, P, P, P],
    "hcu_conv_tc_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_tc_describe": [C.POINTER(HcuConvDesc), C.c_char_p, I32],
    "hcu_conv_tc_packed_bytes": [C.POINTER(HcuConvDesc)],
    "hcu_conv_tc_pack": [C.POINTER(HcuConvDesc), P, P, P],
    "hcu_conv_tc_pack_ref": [C.POINTER(HcuConvDesc), C.POINTER(HcuWeightMap), P, P, P],
    "hcu_conv_tc_fwd": [C.POINTER(HcuConvDesc), P, P, P, P, P, P, P, P, P, P],
    "hcu_conv_tc_fwd_bn": [C.POINTER(HcuConvDesc), P, P, P, P, P, P, P, P, P, C.POINTER(HcuBnFin), P],
    "hcu_conv_wgrad_partial": [C.POINTER(HcuConvDesc), P, P, P, P, P, I32, P],
    "hcu_conv_wgrad_tc_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_wgrad_tc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P],
    "hcu_conv_wgrad_tc_acc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P],
    "hcu_conv_wgrad_ws_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_wgrad_ws_acc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P],
    "hcu_conv_wgrad_tc5_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_wgrad_tc5_acc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P],
    "hcu_conv_wgrad_rows_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_wgrad_rows_acc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P],
    "hcu_conv_wgrad_rows_bnb_supported": [C.POINTER(HcuConvDesc)],
    "hcu_conv_wgrad_rows_bnb_acc": [C.POINTER(HcuConvDesc), P, P, P, P, P, P, P, P, P, P],
    "hcu_adam_flat": [P, P, P, P, I64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, P, P, P],
    "hcu_conv_tc_pack_batch_build": [C.POINTER(HcuConvDesc), C.POINTER(HcuWeightMap), C.POINTER(I64), C.POINTER(I64), I32, P,
                                     C.POINTER(I32)],
    "hcu_conv_tc_pack_batch": [P, I32, I32, P, P, P],
    "hcu_weight_scatter_batch_build": [C.POINTER(HcuWeightMap), C.POINTER(I32), C.POINTER(I64), C.POINTER(I64), I32, P,
                                       C.POINTER(I32)],
    "hcu_weight_scatter_batch": [P, I32, I32, P, F, P, P, P],
    "hcu_weight_gather": [C.POINTER(HcuWeightMap), P, P, P],
    "hcu_weight_scatter": [C.POINTER(HcuWeightMap), P, I32, I64, F, P, I32, P, P],
    "hcu_nc_to_cl": [P, I32, P, I32, I64, I32, I64, I32, P, P],
    "hcu_cl_to_nc": [P, I32, P, I32, I64, I32, I64, I32, P, P],
    "hcu_grad_scale": [P, I64, F, P, P, P],
    "hcu_bn_finalize": [P, I32, D, P, P, F, F, P, P, P, P, P, P, P],
    "hcu_bn_eval_affine": [I32, P, P, P, P, F, P, P, P, P],
    "hcu_bn_relu_apply": [P, I32, P, I32, I64, I32, P, P, I32, P],
    "hcu_bn_relu_maxpool": [P, I32, P, I32, P, I32, I32, I32, I32, I32, I32, I32, I32, P, P, I32, P],
    "hcu_maxpool_bwd": [P, I32, P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, P],
    "hcu_bn_bwd_stats": [P, I32, P, I32, I64, I32, P, P, P, P, I32, P, C.POINTER(HcuPoolGeom), P, P],
    "hcu_bn_bwd_stats_fin": [P, I32, P, I32, I64, I32, P, P, P, P, I32, P, C.POINTER(HcuPoolGeom), P,
                             C.POINTER(HcuBnBwdFin), P],
    "hcu_bn_bwd_finalize": [P, I32, D, P, P, P, I32, F, P, P, P, P, P, P],
    "hcu_bn_bwd_apply": [P, I32, P, I32, P, I32, I64, I32, P, P, I32, P, P, C.POINTER(HcuPoolGeom), P],
    "hcu_colsum": [P, I32, I64, I32, I32, I32, F, P, P, P, P],
    "hcu_wbce_fwd": [C.POINTER(HcuLossDesc), P, P, P, P, P, P],
    "hcu_wbce_bwd": [C.POINTER(HcuLossDesc), P, P, P, P, F, P, P, P],
    "hcu_pair_reduce": [C.POINTER(HcuLossDesc), I32, P, P, P, P],
    "hcu_pair_bwd": [C.POINTER(HcuLossDesc), I32, P, P, P, P, P],
}
_RESTYPES = {"hcu_last_error": C.c_char_p, "hcu_launch_count": C.c_longlong, "hcu_conv_tc_packed_bytes": C.c_longlong}

_lib = None


class _ProfState:
    profiler = None   # hcunet_b200.profiler.KernelProfile while one is active
    note = None       # (layer, algorithmic bytes, flops) for the next call, set by the engine


def note(layer=None, nbytes=0, flops=0):
    """Annotate the next C-ABI call (only looked at while a KernelProfile is active)."""
    if _ProfState.profiler is not None:
        _ProfState.note = (layer, nbytes, flops)


def _wrap(name, fn):
    def call(*args):
        prof = _ProfState.profiler
        if prof is None:
            return fn(*args)
        import torch

        n, _ProfState.note = _ProfState.note, None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        prof.records.append((name, n, e0, e1))
        return rc

    call.__name__ = name
    return call


class _Lib:
    """Attribute access returns the (profiling-aware) C entry points."""


def load():
    """Load (once) and return the ctypes library; raises RuntimeError when it is unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"hcunet_b200: CUDA library not built ({LIB_PATH} missing). There is no CPU / PyTorch fallback: "
            "run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc) first.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"hcunet_b200: {LIB_PATH} does not export {name}; rebuild it") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    ver = lib.hcu_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"hcunet_b200: library ABI {ver} != binding ABI {ABI_VERSION}; rebuild it")
    out = _Lib()
    out._cdll = lib
    for name in SIGNATURES:
        raw = getattr(lib, name)
        setattr(out, name, raw if name in ("hcu_abi_version", "hcu_last_error", "hcu_launch_count", "hcu_h2d_tile", "hcu_conv_tc_supported", "hcu_conv_wgrad_tc_supported", "hcu_conv_wgrad_tc5_supported", "hcu_conv_wgrad_ws_supported", "hcu_conv_wgrad_rows_supported", "hcu_conv_wgrad_rows_bnb_supported",
                                         "hcu_conv_tc_packed_bytes", "hcu_conv_tc_describe", "hcu_conv_tc_bnbwd_supported", "hcu_conv_tc_pack_batch_build", "hcu_weight_scatter_batch_build") else _wrap(name, raw))
    _lib = out
    return out


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().hcu_last_error()
        raise RuntimeError(f"hcunet_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().hcu_launch_count())
