"""CPU tests of the multi-GPU host logic: patch / tile sharding and the gradient all-reduce on gloo (world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import unet_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    from hcunet_b200.parallel import shard_range

    for n in (0, 1, 7, 8, 9, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_tile_grid_covers_output_once_across_ranks():
    from hcunet_b200.tiling import shard_tiles, tile_geometry, tile_grid

    spec = O.normalise_spec(O.README_3D)
    align, margin, mz = tile_geometry(spec)
    assert (align, margin, mz) == (16, 184, 5)
    ox = oy = 2048 - margin                                # BASELINE config 5: 4x2048x2048x128
    tiles = tile_grid((ox, oy), 512, align)
    cover = torch.zeros(ox, oy, dtype=torch.int32)
    for r in range(8):
        for (x0, x1, y0, y1) in shard_tiles(tiles, 8, r):
            assert x0 % align == 0 and y0 % align == 0
            cover[x0:x1, y0:y1] += 1
    assert int(cover.min()) == 1 and int(cover.max()) == 1
    with pytest.raises(ValueError):
        tile_grid((100, 100), 24, align)


def h3(m, lo, hi):
    # first element of the first three gradients after the GENERIC all-reduce (recomputed: the in-place phase overwrote .grad)
    return [(3 + 7) / 2 + i for i in range(3)]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hcunet_b200 as H
        from hcunet_b200.parallel import GradSync, shard_range

        torch.manual_seed(100 + rank)  # ranks start from DIFFERENT weights: the broadcast must fix that
        m = H.Unet_Constructor(**dict(O.README_3D, feature_sizes=[4, 8]))
        sync = GradSync(m, world)
        w0 = torch.cat([p.detach().flatten() for p in m.parameters()])
        # fake per-rank gradients of a sharded batch: rank r holds patches shard_range(5, world, r)
        lo, hi = shard_range(5, world, rank)
        for i, p in enumerate(m.parameters()):
            p.grad = torch.full_like(p, float(sum(range(lo, hi))) + i)
        sync.allreduce()
        g = torch.cat([p.grad.flatten() for p in m.parameters()])
        assert sync.in_place_steps == 0          # foreign .grad tensors: the generic (packing) path
        # the engine's layout: every .grad is a view of ONE flat buffer in parameter order -> reduced in place, no copies
        n = sum(p.numel() for p in m.parameters())
        flat = torch.empty(n)
        off = 0
        for i, p in enumerate(m.parameters()):
            p.grad = flat[off:off + p.numel()].view_as(p)
            p.grad.fill_(float(rank) + i)
            off += p.numel()
        m._engine.last_grad_flat = flat
        ptr = flat.data_ptr()
        sync.allreduce()
        assert sync.in_place_steps == 1 and flat.data_ptr() == ptr
        first = [float(p.grad.flatten()[0]) for p in m.parameters()][:3]
        assert first == [0.5 + i for i in range(3)], first      # mean over ranks {0, 1} of (rank + i)
        # overlapped exchange (GradSync.attach): the engine reports two ranges of its flat buffer, each is averaged in place
        sync.attach()
        flat2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
        split = n // 3
        assert sync._on_grads_ready(flat2, split, n, []) is None and sync._on_grads_ready(flat2, 0, split, []) is None
        assert torch.equal(flat2, torch.arange(n, dtype=torch.float32) * 1.5) and sync.overlapped_steps == 0 and sync._bucket_calls == 2
        sync.allreduce()                          # attached: a no-op
        assert torch.equal(flat2, torch.arange(n, dtype=torch.float32) * 1.5)
        ret[rank] = (w0, g, [float(v) for v in g[:1]] and [float(x) for x in h3(m, lo, hi)])
    finally:
        dist.destroy_process_group()


def test_gradsync_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (w_a, g_a, h_a), (w_b, g_b, h_b) = ret[0], ret[1]
    assert torch.equal(w_a, w_b), "parameters were not broadcast from rank 0"
    assert torch.equal(g_a, g_b), "all-reduced gradients differ between ranks"
    # mean over ranks of (sum of the rank's patch ids + i): ranks hold {0,1,2} and {3,4} -> (3 + 7) / 2 = 5
    assert h_a == [5.0, 6.0, 7.0]


def test_tile_extents_avoid_pooling_floors_for_the_readme_model():
    """README model: an XY extent loses no voxel to a pooling floor only at 16 b + 124, so tiles read
    a few voxels more than tile_out + margin; the whole-cochlea stack (2048) yields 1860 logit rows."""
    import hcunet_b200 as H
    from hcunet_b200 import tiling
    from hcunet_b200.engine import plan_unet

    kw = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
              kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
              upsample_stride=(2, 2, 1), dilation=1, groups=1)
    spec = H.Unet_Constructor(**kw).model_specification
    align, margin, mz = tiling.tile_geometry(spec)
    assert (align, margin, mz) == (16, 184, 5)
    assert tiling.tiled_output_extent(spec, 2048) == 1860
    for want, avail in ((512, 2048), (256, 2048), (324, 512), (516, 700)):
        n = tiling.tile_input_extent(spec, want, avail)
        assert n <= avail and n % 16 == 12
        got = plan_unet(spec, (1, 4, n, n, 32)).out_sz[0]
        assert got == n - margin and got >= want
    # a 2 x 4 split of the stack into 1212 x 700 inputs gives 1028 x 516 logits per rank
    assert plan_unet(spec, (1, 4, 1212, 700, 128)).out_sz == (1028, 516, 123)
