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
        ret[rank] = (w0, g, [float(p.grad.flatten()[0]) for p in m.parameters()][:3])
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
