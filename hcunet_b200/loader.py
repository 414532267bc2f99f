"""GPU input path: the deterministic chain of the reference dataloader on the device.

The reference (`hcat/dataloader.py:68-92` with `hcat/transforms.py`) turns the raw TIFF stack -- ``[Z, Y, X, C]`` uint8 /
uint16 as ``skimage.io.imread`` yields it -- into an fp16 ``[1, C, X, Y, Z]`` tensor with numpy on the host:
``to_float`` (`transforms.py:94-113`: v / 2^bits in float64) -> ``reshape`` (`:139-157`: swapaxes -> ``[X, Y, Z, C]``) ->
``normalize`` (`:257-282`: ``+= -mean[c]; /= std[c]``) -> ``to_tensor`` (`:118-136`: ``torch.half``, channels first); mask and
pwl go through ``to_float`` / ``reshape`` / ``to_tensor`` with it.  (The random augmentations between them are outside this
path.)

``StackLoader`` does the same on the device with the library's ``hcu_load_stack`` / ``hcu_load_labels`` kernels:

    loader = StackLoader(model)                       # mean / std default to normalize()'s [0.5] * C
    x, m, w = loader(image_u8_zyxc, mask_zyx, pwl_zyx)  # raw arrays: pinned host or device tensors (or numpy)
    loss = cross_entropy(model(x), m, w)

* the raw bytes cross PCIe (4 B / voxel for a 4-channel uint8 stack instead of 8 B of fp16) and ONE transposing pass writes
  the channels-last fp16 layout the first convolution consumes -- ``x`` is a ``[B, C, X, Y, Z]`` view of it that the engine
  recognises (no NCDHW tensor, no layout pass); values are bit-identical to the reference chain;
* of mask / pwl only the origin crop the loss reads (`loss.py:51-56` crops them to the prediction's shape) is gathered --
  straight from pinned host memory when that is where they live -- into the ``[B, 1, x, y, z]`` fp16 tensors the loss takes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

_SRC_DT = {torch.uint8: _lib.U8, torch.float16: _lib.F16, torch.float32: _lib.F32, torch.float64: _lib.F64}
if hasattr(torch, "uint16"):
    _SRC_DT[torch.uint16] = _lib.U16
_SRC_DT[torch.int16] = _lib.U16   # numpy uint16 viewed as int16 (older torch builds have no uint16)

CL_PITCH = 8   # channel pitch of the layout the first convolution reads (16-byte voxels)


def as_tensor(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a
    import numpy as np

    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16 and not hasattr(torch, "uint16"):
        a = a.view(np.int16)
    return torch.from_numpy(a)


def is_prelaid(x: torch.Tensor) -> bool:
    """True for the ``[B, C, *spatial]`` fp16 view ``StackLoader.image`` returns (channels-last storage, pitch 8, zero padding)."""
    return bool(getattr(x, "_hcu_cl8", False))


class StackLoader:
    def __init__(self, model, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None):
        self.model = model
        c = model.model_specification["in_channels"]
        if c > CL_PITCH:
            raise ValueError(f"StackLoader supports up to {CL_PITCH} input channels, the model has {c}")
        self.mean = list(mean) if mean is not None else [0.5] * c     # normalize.__init__ defaults (transforms.py:258-266)
        self.std = list(std) if std is not None else [0.5] * c
        if len(self.mean) < c or len(self.std) < c:
            raise ValueError("mean / std need one entry per input channel")
        self._mean = (C.c_double * CL_PITCH)(*(self.mean[:c] + [0.0] * (CL_PITCH - c)))
        self._std = (C.c_double * CL_PITCH)(*(self.std[:c] + [1.0] * (CL_PITCH - c)))

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _device(self):
        return next(self.model.parameters()).device

    def label_extent(self, image_shape) -> Tuple[int, ...]:
        """Spatial extent of the model's prediction for a raw ``[B, Z, Y, X, C]`` (or ``[B, Y, X, C]``) stack."""
        dims = self.model.model_specification["image_dimensions"]
        sp = tuple(image_shape[1:1 + dims])[::-1]                       # (X, Y, Z) / (X, Y)
        plan = self.model._engine.plan((image_shape[0], image_shape[-1]) + sp)
        return tuple(plan.out_sz[:dims])

    def image(self, raw, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """raw ``[B, Z, Y, X, C]`` / ``[Z, Y, X, C]`` (2D: ``[B, Y, X, C]`` / ``[Y, X, C]``) uint8 | uint16 -> the model input."""
        raw = as_tensor(raw)
        dims = self.model.model_specification["image_dimensions"]
        if raw.dim() == dims + 1:
            raw = raw.unsqueeze(0)
        if raw.dim() != dims + 2:
            raise ValueError(f"expected a raw [B, {'Z, ' if dims == 3 else ''}Y, X, C] stack, got shape {tuple(raw.shape)}")
        if raw.dtype not in (torch.uint8, torch.int16, getattr(torch, "uint16", torch.uint8)):
            raise TypeError("Expected image datatype of uint8 or uint16 ")     # to_float, transforms.py:112
        dev = self._device()
        if not raw.is_cuda:
            raw = raw.to(dev, non_blocking=True)
        raw = raw.contiguous()
        if dims == 3:
            b, z, y, x, c = raw.shape
        else:
            b, y, x, c = raw.shape
            z = 1
        if c != self.model.model_specification["in_channels"]:
            raise RuntimeError(f"expected a stack with {self.model.model_specification['in_channels']} channels, got {c}")
        if out is None:
            out = torch.empty((b, x * y * z, CL_PITCH), dtype=torch.float16, device=dev)
        _lib.check(_lib.load().hcu_load_stack(C.c_void_p(raw.data_ptr()), _SRC_DT[raw.dtype], b, z, y, x, c, self._mean,
                                              self._std, C.c_void_p(out.data_ptr()), CL_PITCH, self._stream()), "load_stack")
        s = x * y * z * CL_PITCH
        if dims == 3:
            view = out.as_strided((b, c, x, y, z), (s, 1, y * z * CL_PITCH, z * CL_PITCH, CL_PITCH))
        else:
            view = out.as_strided((b, c, x, y), (s, 1, y * CL_PITCH, CL_PITCH))
        view._hcu_cl8 = True
        view._hcu_raw = raw      # keeps the staging buffer alive until the kernel has consumed it
        return view

    def labels(self, t, extent: Sequence[int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mask / pwl ``[B, Z, Y, X]`` / ``[Z, Y, X]`` (2D: ``[B, Y, X]`` / ``[Y, X]``) -> fp16 ``[B, 1, *extent]`` (origin crop)."""
        t = as_tensor(t)
        dims = self.model.model_specification["image_dimensions"]
        if t.dim() == dims:
            t = t.unsqueeze(0)
        if t.dim() != dims + 1:
            raise ValueError(f"expected [B, {'Z, ' if dims == 3 else ''}Y, X] labels, got shape {tuple(t.shape)}")
        if t.dtype not in _SRC_DT:
            raise TypeError(f"unsupported label dtype {t.dtype}")
        dev = self._device()
        if not t.is_cuda and not t.is_pinned():
            t = t.to(dev, non_blocking=True)     # pageable host memory cannot be read in place
        t = t.contiguous()
        if dims == 3:
            b, z, y, x = t.shape
            ox, oy, oz = extent
        else:
            b, y, x = t.shape
            z, (ox, oy), oz = 1, extent, 1
        if out is None:
            out = torch.empty((b, 1) + tuple(extent), dtype=torch.float16, device=dev)
        _lib.check(_lib.load().hcu_load_labels(C.c_void_p(t.data_ptr()), _SRC_DT[t.dtype], b, z, y, x, ox, oy, oz,
                                               C.c_void_p(out.data_ptr()), self._stream()), "load_labels")
        out._hcu_src = t
        return out

    def __call__(self, image, mask=None, pwl=None):
        image = as_tensor(image)
        dims = self.model.model_specification["image_dimensions"]
        shape = tuple(image.shape) if image.dim() == dims + 2 else (1,) + tuple(image.shape)
        x = self.image(image)
        ext = self.label_extent(shape)
        m = self.labels(mask, ext) if mask is not None else None
        w = self.labels(pwl, ext) if pwl is not None else None
        return x, m, w
