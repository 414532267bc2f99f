"""GPU-resident overlap-tile inference: the reference's `predict_segmentation_mask` (`hcat/segment.py:21-136`) with
its helpers `pad_image_with_reflections` (`hcat/utils.py:33-74`) and `calculate_indexes` (`hcat/utils.py:77-124`),
same names, arguments, tile arithmetic, loop order and error messages.

What the reference does per call, on the host: scrub NaN / inf in place, build a reflection-padded copy of the whole
stack (numpy flips + three `torch.cat`), then for every tile: slice, `.float().to(device)`, test for all -1, run the net,
crop the centre, in-place sigmoid, threshold, paste into a host mask (an implicit device -> host copy per tile).  Its tile
size comes from a table keyed by the GPU's memory in GB (`segment.py:48-57`) that has no entry for a 180 GB B200 (KeyError).

Here the stack is moved to HBM once (or only this rank's footprint of it), and three kernels of `csrc/tiler.cu` do the rest:
`hcu_tile_flags` (the skip test for every tile, one read-back for all of them), `hcu_tile_gather` (scrub + reflection +
slice + channels-last layout: the padded stack never exists) and `hcu_sigmoid_paste` (crop + sigmoid + threshold + paste
into the device-resident mask).  Tiles are independent (valid convolutions), so `world` ranks take contiguous ranges of the
reference's tile order with no communication; `merge_rank_masks` reproduces the reference's later-tile-wins overwrite.

The reference's arithmetic quirks are kept because they define its output: tiles are `eval + 2 * pad - 1` wide
(`utils.py:108`), the crop `[pad, pad + eval)` of the logits is taken as if the network did not shrink its input (it is
shifted by half the network's margin), and +-inf becomes +1.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .engine import plan_unet
from .parallel import shard_range

# segment.py:48-57
_EVAL_IM_SIZE = {"4": [128, 128, 6], "6": [300, 300, 6], "8": [300, 300, 10], "11": [350, 350, 15]}


def calculate_indexes(pad_size: int, eval_image_size: int, image_shape: int, padded_image_shape: int) -> List[List[int]]:
    """`hcat/utils.py:77-124`, integer for integer (including the `- 1` that makes a tile `eval + 2 * pad - 1` wide)."""
    if eval_image_size > image_shape:
        return [[0, image_shape]]
    if eval_image_size <= 0:
        raise RuntimeError(f"Calculate_indexes has incorrect values {pad_size} | {image_shape} | {eval_image_size}:\n"
                           "You are likely trying to have a chunk smaller than the set evaluation image size. "
                           "Please decrease number of chunks.")
    ind_list = list(range(0, image_shape, eval_image_size))
    ind = []
    for i, z in enumerate(ind_list):
        if i == 0:
            continue
        z1 = int(ind_list[i - 1])
        z2 = int(z - 1) + (2 * pad_size)
        if z2 < padded_image_shape:
            ind.append([z1, z2])
        else:
            break
    if not ind:
        ind.append([0, eval_image_size + pad_size * 2])
        ind.append([padded_image_shape - (eval_image_size + pad_size * 2), padded_image_shape])
    else:
        ind.append([padded_image_shape - (eval_image_size + pad_size * 2), padded_image_shape - 1])
    return ind


def _i3(v):
    return (C.c_int32 * 3)(*[int(e) for e in v])


def _geom(channels, size, pad, origin, extent, sorg=(0, 0, 0), ssize=None) -> _lib.HcuTileGeom:
    g = _lib.HcuTileGeom()
    g.channels = int(channels)
    g.size, g.pad, g.origin, g.extent = _i3(size), _i3(pad), _i3(origin), _i3(extent)
    g.stack_origin, g.stack_size = _i3(sorg), _i3(ssize if ssize is not None else size)
    return g


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float16:
        return _lib.F16
    raise TypeError(f"expected a float32 or float16 stack, got {t.dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad_image_with_reflections(image: torch.Tensor, pad_size: Sequence[int] = (30, 30, 6), device=None) -> torch.Tensor:
    """`hcat/utils.py:33-74` on the device: [B, C, X, Y, Z] -> [B, C, X + 2 px, Y + 2 py, Z + 2 pz], the first / last
    `pad` voxels of every dimension mirrored including the edge voxel.  Same checks (`TypeError` for a non-tensor,
    `ValueError` for an odd pad).  Unlike the tile path (`hcu_tile_gather`) nothing is scrubbed: the reference scrubs in
    `predict_segmentation_mask`, not here.  The result lives where `image` lives."""
    if not isinstance(image, torch.Tensor):
        raise TypeError(f"Expected image to be of type torch.tensor not {type(image)}")
    for pad in pad_size:
        if pad % 2 != 0:
            raise ValueError("Padding must be divisible by 2")
    if image.dim() != 5:
        raise RuntimeError(f"expected [B, C, X, Y, Z], got {tuple(image.shape)}")
    dev = torch.device(device) if device is not None else (image.device if image.is_cuda else torch.device("cuda", 0))
    src = image.to(dev).contiguous()
    if src.dtype not in (torch.float32, torch.float16):
        src = src.float()
    B, Cn, X, Y, Z = (int(v) for v in src.shape)
    pad = [int(p) for p in pad_size]
    if pad[0] > X or pad[1] > Y or pad[2] > Z:
        raise RuntimeError("padding larger than the image")
    ext = (X + 2 * pad[0], Y + 2 * pad[1], Z + 2 * pad[2])
    out = torch.empty((B, Cn) + ext, dtype=src.dtype, device=dev)
    lib = _lib.load()
    g = _geom(Cn, (X, Y, Z), pad, (0, 0, 0), ext)
    with torch.cuda.device(dev):
        for b in range(B):   # hcu_tile_gather with the whole padded extent as the "tile", layout 0, no scrub
            _lib.check(lib.hcu_tile_gather(C.byref(g), C.c_void_p(src[b].data_ptr()), _dt(src), C.c_void_p(out[b].data_ptr()),
                                           _dt(src), 2, Cn, _stream()), "tile_gather")
    out = out.to(image.dtype) if out.dtype != image.dtype else out
    return out if image.is_cuda else out.to(image.device)


def tile_list(im_shape: Sequence[int], pad_size: Sequence[int], eval_image_size: Sequence[int]):
    """The reference's tile order (`segment.py:73-84`: z outermost, then x, then y) as [(x, y, z)] with each of x, y, z a
    `calculate_indexes` pair in padded coordinates."""
    X, Y, Z = (int(v) for v in im_shape)
    x_ind = calculate_indexes(pad_size[0], eval_image_size[0], X, X + 2 * pad_size[0])
    y_ind = calculate_indexes(pad_size[1], eval_image_size[1], Y, Y + 2 * pad_size[1])
    z_ind = calculate_indexes(pad_size[2], eval_image_size[2], Z, Z + 2 * pad_size[2])
    return [(x, y, z) for z in z_ind for x in x_ind for y in y_ind]


def _orig_range(lo: int, hi: int, p: int, n: int) -> Tuple[int, int]:
    """Original-coordinate interval [a, b) that the padded indices [lo, hi) read through the reflection
    (u < p -> p-1-u; u >= p+n -> n-1-(u-p-n); else u-p)."""
    segs = []
    if lo < min(hi, p):                       # low mirror
        segs.append((p - min(hi, p), p - 1 - lo))
    if max(lo, p) < min(hi, p + n):           # interior
        segs.append((max(lo, p) - p, min(hi, p + n) - p - 1))
    if max(lo, p + n) < hi:                   # high mirror
        segs.append((n - 1 - (hi - 1 - p - n), n - 1 - (max(lo, p + n) - p - n)))
    return max(0, min(a for a, _ in segs)), min(n, max(b for _, b in segs) + 1)


def _sizes(cuda_mem, eval_image_size, pad_size, z_extent):
    if eval_image_size is None or pad_size is None:
        if cuda_mem:
            key = str(int(math.floor(cuda_mem / 1e9)))
            if key not in _EVAL_IM_SIZE:
                raise KeyError(key)    # segment.py:54: the table has no entry for other memory sizes
            pad, ev = [128, 128, 10], list(_EVAL_IM_SIZE[key])
        else:
            pad, ev = [128, 128, 10], [300, 300, 15]
        pad_size = pad if pad_size is None else list(pad_size)
        eval_image_size = ev if eval_image_size is None else list(eval_image_size)
    pad_size, eval_image_size = [int(v) for v in pad_size], [int(v) for v in eval_image_size]
    if z_extent < eval_image_size[2]:
        eval_image_size[2] = int(z_extent)      # segment.py:62-63
    return pad_size, eval_image_size


@torch.no_grad()
def predict_segmentation_mask(unet, image: torch.Tensor, device="cuda", use_probability_map: bool = False,
                              mask_cell_prob_threshold: float = 0.5, *, cuda_mem: Optional[float] = None,
                              eval_image_size: Optional[Sequence[int]] = None, pad_size: Optional[Sequence[int]] = None,
                              world: int = 1, rank: int = 0, return_written: bool = False, keep_on_device: bool = False):
    """`hcat/segment.py:21-136`.  `image` [1, C, X, Y, Z] float32 / float16 with the transforms applied, on the host
    (pinned or not) or on the device; returns the mask [1, 1, X, Y, Z] (uint8, or float32 probabilities with
    `use_probability_map`) on the host like the reference does, or on the device with `keep_on_device`.

    Keyword-only extensions: `cuda_mem` (bytes) selects the reference's tile table (`hcat.__CUDA_MEM__`; None = its
    defaults PAD [128, 128, 10], EVAL [300, 300, 15]); `eval_image_size` / `pad_size` override them; `world` / `rank`
    shard the tile list (this rank's blocks are written, `return_written` also returns the uint8 map of written voxels)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hcunet_b200.segment runs on a CUDA device: there is no CPU path")
    if image.dim() != 5 or image.shape[0] != 1:
        raise RuntimeError(f"expected an image of shape [1, C, X, Y, Z], got {tuple(image.shape)}")
    if unet.training:
        raise RuntimeError("predict_segmentation_mask needs unet.eval() (segment.py runs under no_grad on an eval model)")
    lib = _lib.load()
    Cn, X, Y, Z = (int(v) for v in image.shape[1:])
    PAD, EVAL = _sizes(cuda_mem, eval_image_size, pad_size, Z)
    tiles = tile_list((X, Y, Z), PAD, EVAL)
    lo, hi = shard_range(len(tiles), world, rank)
    mine = tiles[lo:hi]
    size = (X, Y, Z)
    padded = tuple(size[d] + 2 * PAD[d] for d in range(3))
    for (x, y, z) in mine:
        for d, pr in enumerate((x, y, z)):
            if pr[0] < 0 or pr[1] > padded[d] or pr[1] <= pr[0]:
                raise RuntimeError(f"tile {pr} outside the padded image ({padded[d]}) in dim {d}: the image is smaller than the "
                                   f"evaluation size plus padding")
    with torch.cuda.device(dev):
        st = _stream()
        # ---- residency: the original voxels this rank's tiles read (whole stack for one rank) ----
        if mine:
            box = []
            for d in range(3):
                rs = [_orig_range(t[d][0], t[d][1], PAD[d], size[d]) for t in mine]
                box.append((min(r[0] for r in rs), max(r[1] for r in rs)))
        else:
            box = [(0, 1)] * 3
        sorg = tuple(b[0] for b in box)
        ssize = tuple(b[1] - b[0] for b in box)
        sub = image[0, :, box[0][0]:box[0][1], box[1][0]:box[1][1], box[2][0]:box[2][1]]
        stack = sub.to(dev, non_blocking=True).contiguous()       # [C][sx][sy][sz]
        sdt = _dt(stack)
        mixed = getattr(unet, "precision", "fp32") == "mixed"
        # ---- skip test of every tile, one read-back (segment.py:89-93) ----
        flags = torch.zeros(max(1, len(mine)), dtype=torch.int32, device=dev)
        geoms = []
        for k, (x, y, z) in enumerate(mine):
            g = _geom(Cn, size, PAD, (x[0], y[0], z[0]), (x[1] - x[0], y[1] - y[0], z[1] - z[0]), sorg, ssize)
            geoms.append(g)
            _lib.check(lib.hcu_tile_flags(C.byref(g), C.c_void_p(stack.data_ptr()), sdt,
                                          C.c_void_p(flags.data_ptr() + 4 * k), st), "tile_flags")
        live = flags.cpu().tolist() if mine else []
        mask = torch.zeros((1, 1, X, Y, Z), dtype=torch.float32, device=dev)
        written = torch.zeros((X, Y, Z), dtype=torch.uint8, device=dev) if return_written else None
        spec = unet.model_specification
        for k, (x, y, z) in enumerate(mine):
            if live[k] == 0:
                continue      # "Occasionally everything is just -1 in the whole mat. Skip for speed"
            ext = (x[1] - x[0], y[1] - y[0], z[1] - z[0])
            # ---- scrub + reflection + slice + layout in one pass ----
            if mixed:
                tile = torch.empty((1,) + ext + (8,), dtype=torch.float16, device=dev)
                _lib.check(lib.hcu_tile_gather(C.byref(geoms[k]), C.c_void_p(stack.data_ptr()), sdt,
                                               C.c_void_p(tile.data_ptr()), _lib.F16, 1, 8, st), "tile_gather")
                xin = tile.as_strided((1, Cn) + ext, (tile.numel(), 1, ext[1] * ext[2] * 8, ext[2] * 8, 8))
                xin._hcu_cl8 = True     # the engine reads this storage as is (no NCDHW -> NDHWC pass), like StackLoader.image
            else:
                xin = torch.empty((1, Cn) + ext, dtype=torch.float32, device=dev)
                _lib.check(lib.hcu_tile_gather(C.byref(geoms[k]), C.c_void_p(stack.data_ptr()), sdt,
                                               C.c_void_p(xin.data_ptr()), _lib.F32, 0, Cn, st), "tile_gather")
            valid_out = unet(xin)                                   # [1, 1, ox, oy, oz] fp32 logits
            if valid_out.shape[1] != 1:
                raise RuntimeError("predict_segmentation_mask expects a one-channel model (mask has one channel)")
            osz = tuple(int(v) for v in valid_out.shape[2:])
            # valid_out[:, :, PAD:EVAL+PAD, ...] (slices clip) pasted at mask[x0:x0+EVAL, ...] (slices clip): shapes must agree
            got = tuple(max(0, min(osz[d], EVAL[d] + PAD[d]) - PAD[d]) for d in range(3))
            org = (x[0], y[0], z[0])
            want = tuple(max(0, min(size[d], org[d] + EVAL[d]) - org[d]) for d in range(3))
            if got != want or min(got) <= 0:
                raise RuntimeError(f"Amount of padding is not sufficient.\nvalid_out.shape: {(1, 1) + got}\n"
                                   f"eval_image_size: {EVAL} \npadded_image_slice.shape{(1, Cn) + ext} ")
            if not use_probability_map and mask.dtype != torch.uint8:
                mask = mask.to(torch.uint8)                          # segment.py:116-117
            _lib.check(lib.hcu_sigmoid_paste(C.c_void_p(valid_out.data_ptr()), _i3(osz), _i3(PAD), _i3(got),
                                             C.c_void_p(mask.data_ptr()), 1 if use_probability_map else 0, _i3(size),
                                             _i3(org), float(mask_cell_prob_threshold),
                                             C.c_void_p(written.data_ptr()) if written is not None else None, st),
                       "sigmoid_paste")
        out = mask if keep_on_device else mask.cpu()
        if return_written:
            return out, (written if keep_on_device else written.cpu())
        return out


def merge_rank_masks(parts: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """Masks of ranks 0 .. world-1 (each with its `written` map) -> the single-process mask: ranks hold contiguous ranges of
    the reference's tile order, so applying them in rank order reproduces its later-tile-wins overwrite."""
    any_u8 = any(m.dtype == torch.uint8 for m, _ in parts)
    out = torch.zeros_like(parts[0][0], dtype=torch.uint8 if any_u8 else torch.float32)
    for m, w in parts:
        sel = w.to(torch.bool).view(out.shape)
        out = torch.where(sel, m.to(out.dtype), out)
    return out


def tile_plan_summary(unet, im_shape, cuda_mem=None, eval_image_size=None, pad_size=None):
    """(PAD, EVAL, number of tiles, tile input extent, logits extent) without touching the GPU."""
    PAD, EVAL = _sizes(cuda_mem, eval_image_size, pad_size, im_shape[2])
    tiles = tile_list(im_shape, PAD, EVAL)
    x, y, z = tiles[0]
    ext = (x[1] - x[0], y[1] - y[0], z[1] - z[0])
    spec = unet.model_specification
    out = plan_unet(spec, (1, spec["in_channels"]) + ext).out_sz
    return PAD, EVAL, len(tiles), ext, out
