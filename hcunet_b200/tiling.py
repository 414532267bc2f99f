"""Overlap-tile inference sharded over GPUs with no communication (SURVEY.md sections 3.4, 8e; BASELINE config 5).

Valid convolutions make output blocks independent: output voxels ``[a, b)`` (``a`` a multiple of the pooling
stride ``align``, 16 for the README model) depend on input voxels ``[a, b + margin)`` only (``margin`` = 184 for the
README model).  The tile grid over (X, Y) is partitioned across ranks; each rank runs its tiles through the model in
``eval()`` mode and writes its blocks of the logits.  Z is never tiled: the Z up-convolution is a full correlation
whose edge taps see implicit zeros (`unet.py:294-298`).

This replaces the *sharding arithmetic* of the reference's serial tiler (`segment.py:73-126`: one tile at a time, host
round trip per tile, GPU-memory table that has no entry for a B200); the reference's reflection padding / thresholding
around it are out of scope this round ("next" item 1).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .engine import plan_unet
from .parallel import shard_range


_GEOMETRY_CACHE: dict = {}


def _geometry(spec: dict) -> Tuple[int, int, int, int]:
    """Cached per model specification (the search below plans the network a few hundred times on the host)."""
    key = repr(sorted((k, repr(v)) for k, v in spec.items()))
    if key not in _GEOMETRY_CACHE:
        _GEOMETRY_CACHE[key] = _geometry_search(spec)
    return _GEOMETRY_CACHE[key]


def _geometry_search(spec: dict) -> Tuple[int, int, int, int]:
    """(align, margin_xy, margin_z, residue): an XY input extent ``n`` with ``n % align == residue`` loses no voxel to a
    pooling floor at any level, and then the logits are exactly ``n - margin_xy`` wide (README model: 16, 184, 5, 12 --
    SURVEY.md section 8d "in = 16 b + 124").  Derived from the planner, not hard-coded."""
    dims = spec["image_dimensions"]
    pool = spec["max_pool_kernel"]
    pool = (pool,) * dims if isinstance(pool, int) else tuple(pool)
    stride = spec["upsample_stride"]
    stride = (stride,) * dims if isinstance(stride, int) else tuple(stride)
    levels = len(spec["feature_sizes"]) - 1
    if pool[0] != pool[1] or stride[0] != stride[1] or pool[0] != stride[0]:
        raise NotImplementedError("tiling needs equal pooling / up-sampling strides in X and Y")
    align = pool[0] ** levels
    zin = 8 * 4

    def out_of(n):
        shape = (1, spec["in_channels"], n, n) + ((zin,) if dims == 3 else ())
        try:
            return plan_unet(spec, shape).out_sz
        except RuntimeError:
            return None

    # The margin n - out(n) is smallest on the residue class where no pooling level floors; every other class loses voxels.
    best = None
    for n in range(align * 8, align * 40):
        o, o2 = out_of(n), out_of(n + align)
        if o is None or o2 is None or o2[0] - o[0] != align:
            continue
        m = n - o[0]
        if best is None or m < best[1]:
            best = (n % align, m, (zin - o[2]) if dims == 3 else 0)
    if best is None:
        raise RuntimeError("could not find a tile size the model accepts")
    return align, best[1], best[2], best[0]


def tile_geometry(spec: dict) -> Tuple[int, int, int]:
    """(align, margin_xy, margin_z): outputs at multiples of ``align`` need ``margin_xy`` more input voxels per XY
    dim; the logits are ``margin_z`` shorter than the input in Z."""
    return _geometry(spec)[:3]


def tile_input_extent(spec: dict, want_out: int, avail: int) -> int:
    """Smallest XY input extent that yields >= ``want_out`` logit rows without a pooling floor, at most ``avail``."""
    align, margin, _, res = _geometry(spec)
    n = want_out + margin
    n += (res - n) % align
    while n > avail:
        n -= align
    return n


def tiled_output_extent(spec: dict, x: int) -> int:
    """Logit rows an XY stack extent ``x`` produces under tiling (the largest floor-free input extent minus the margin)."""
    align, margin, _, res = _geometry(spec)
    return x - ((x - res) % align) - margin


def tile_grid(out_xy: Tuple[int, int], tile_out: int, align: int) -> List[Tuple[int, int, int, int]]:
    """Output blocks (x0, x1, y0, y1) with origins at multiples of ``align`` covering an ``out_xy`` logit plane."""
    if tile_out % align:
        raise ValueError(f"tile_out must be a multiple of {align}")
    tiles = []
    for x0 in range(0, out_xy[0], tile_out):
        for y0 in range(0, out_xy[1], tile_out):
            tiles.append((x0, min(out_xy[0], x0 + tile_out), y0, min(out_xy[1], y0 + tile_out)))
    return tiles


def shard_tiles(tiles, world: int, rank: int):
    lo, hi = shard_range(len(tiles), world, rank)
    return tiles[lo:hi]


def _tile_to_device(stack: torch.Tensor, x0: int, nx: int, y0: int, ny: int, device) -> torch.Tensor:
    """stack[:, :, x0:x0+nx, y0:y0+ny, :] on the device.  From pinned host memory the tile is copied by strided DMA
    (`hcu_h2d_tile`: one cudaMemcpy2DAsync per channel) -- `tensor[...].to(device)` of a strided host view first gathers the
    tile into a contiguous host buffer on one CPU thread (0.2 s per 700 x 700 x 128 tile, 10x the forward pass)."""
    xin = stack[:, :, x0:x0 + nx, y0:y0 + ny, :]
    if stack.device.type != "cpu" or not stack.is_pinned() or not stack.is_contiguous() or stack.shape[0] != 1:
        return xin.to(device, non_blocking=True)
    import ctypes as C

    from . import _lib

    _, ch, X, Y, Z = stack.shape
    esz = stack.element_size()
    dst = torch.empty((1, ch, nx, ny, Z), dtype=stack.dtype, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().hcu_h2d_tile(C.c_void_p(xin.data_ptr()), ch, X * Y * Z * esz, nx, Y * Z * esz, ny * Z * esz,
                                            C.c_void_p(dst.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "h2d_tile")
    return dst


@torch.no_grad()
def predict_tiled(model, stack: torch.Tensor, tile_out: int = 256, world: int = 1, rank: int = 0,
                  out: Optional[torch.Tensor] = None, device=None):
    """Logits of this rank's share of the overlap tiles.  ``stack`` [1, C, X, Y, Z] may live on the host (pinned) or
    on the device; returns (out, tiles) where ``out`` [1, Cout, X - margin, Y - margin, Z - mz] holds this rank's
    blocks (zeros elsewhere; concatenating / summing the ranks' outputs needs no reduction of overlapping data)."""
    spec = model.model_specification
    if spec["image_dimensions"] != 3:
        raise NotImplementedError("predict_tiled handles 3D stacks")
    if model.training:
        raise RuntimeError("tiled inference needs model.eval(): batch statistics differ per tile")
    align, margin, mz = tile_geometry(spec)
    X, Y, Z = stack.shape[2:]
    # A tile's input extent must sit on the residue class where no pooling level floors (README model: 16 b + 124), so an
    # interior tile reads a few voxels more than tile_out + margin and the surplus logit rows (identical to the
    # neighbour's first rows) are dropped.  A ragged stack produces the rows its largest floor-free extent allows; the
    # reference pads every stack with reflections before tiling (`segment.py:70`).
    ox, oy = tiled_output_extent(spec, X), tiled_output_extent(spec, Y)
    if ox <= 0 or oy <= 0:
        raise RuntimeError(f"stack {tuple(stack.shape)} is smaller than the receptive field ({margin} + {align})")
    device = device or next(model.parameters()).device
    tiles = shard_tiles(tile_grid((ox, oy), tile_out, align), world, rank)
    if out is None:
        out = torch.zeros((1, spec["out_channels"], ox, oy, Z - mz), dtype=torch.float32, device=device)
    # Tiles are copied and run one after the other on the caller's stream.  (Prefetching the next tile on a copy stream
    # was measured slower, 37.8 vs 34.8 ms for two 700 x 700 x 128 tiles: the gather competes with the network for SMs.)
    for (x0, x1, y0, y1) in tiles:
        nx = tile_input_extent(spec, x1 - x0, X - x0)
        ny = tile_input_extent(spec, y1 - y0, Y - y0)
        logits = model(_tile_to_device(stack, x0, nx, y0, ny, device))
        out[:, :, x0:x1, y0:y1] = logits[:, :, : x1 - x0, : y1 - y0]
    return out, tiles
