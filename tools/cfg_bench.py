#!/usr/bin/env python
"""Train-step timing + per-layer CUDA-event profile of the BASELINE.json parity configurations that are not the bench
line (bench.py measures configs[1]):

    python tools/cfg_bench.py cfg3 [--batch 16] [--steps 5] [--out gpurun_out/cfg3_layers.txt]   # 2D classic U-Net, 572^2
    python tools/cfg_bench.py cfg4 [--batch 4]                                                   # README 3D, 256x256x64
    python tools/cfg_bench.py cfg5 [--world 8] [--tile-out 512]   # one rank's share of the 4x2048x2048x128 overlap-tile inference

One JSON line on stdout (same keys as bench.py where they apply), the per-layer table in --out.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hcunet_b200 as H  # noqa: E402
from hcunet_b200 import _lib, profiler  # noqa: E402
from hcunet_b200.graph import GraphedTrainStep  # noqa: E402

README_3D = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                 kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
                 upsample_stride=(2, 2, 1), dilation=1, groups=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cfg1", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--tile-out", type=int, default=512)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if args.config == "cfg5":
        return cfg5(args, dev)
    if args.config == "cfg1":
        return cfg1(args, dev)
    if args.config == "cfg3":
        B = args.batch or 16
        model = H.Unet_Constructor()   # the reference's defaults: 2D, in 3, out 2, features 32..1024, 3x3, up 2x2 stride 2
        xs, ms = (B, 3, 572, 572), (B, 2, 572, 572)
        workload = "classic 2D U-Net [32..1024] k(3,3), ConvTranspose2d s2: train step on 572x572x3 tiles"
        flops_fwd = 169471 * 572 * 572 * B
    else:
        B = args.batch or 4
        model = H.Unet_Constructor(**README_3D)
        xs, ms = (B, 4, 256, 256, 64), (B, 1, 256, 256, 64)
        workload = "README 3D U-Net: train step on 256x256x64 patches"
        flops_fwd = 10769 * 256 * 256 * 64 * B
    model.precision = args.precision
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(xs, generator=g).half().to(dev)
    msk = (torch.rand(ms, generator=g) > 0.7).half().to(dev)
    pwl = (torch.rand(ms, generator=g) * 3).half().to(dev)
    loss_fn = lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel")  # noqa: E731

    def eager():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(img), msk, pwl)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        l_first = float(eager())
    l0 = _lib.launch_count()
    eager()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    step = eager if args.no_graph else GraphedTrainStep(model, opt, loss_fn, (img, msk, pwl))
    run = (lambda: step()) if args.no_graph else (lambda: step.run())
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for _ in range(args.warmup):
        flush.zero_(); run()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    for _ in range(args.steps):
        flush.zero_(); run()
    e[1].record()
    e[2].record()
    for _ in range(args.steps):
        flush.zero_()
    e[3].record()
    torch.cuda.synchronize()
    ms_step = (e[0].elapsed_time(e[1]) - e[2].elapsed_time(e[3])) / args.steps
    l_last = float(step.loss if not args.no_graph else eager())
    nvox = B
    for v in xs[2:]:
        nvox *= v
    prof = profiler.KernelProfile()
    with prof:
        for _ in range(2):
            torch.cuda._sleep(int(40e-3 * 1.9e9))
            eager()
    tab = prof.table()
    tot = sum(v["ms"] for v in tab.values())
    line = {"config": args.config, "workload": workload, "batch": B, "precision": args.precision, "ms_per_step": ms_step,
            "voxels_per_s": nvox / (ms_step / 1e3), "train_tflops": 3 * flops_fwd / (ms_step / 1e3) / 1e12,
            "fwd_gflop": flops_fwd / 1e9, "launches_per_step": int(launches), "loss_first_last": [l_first, l_last],
            "kernel_ms_per_step": {k: round(v["ms"] / 2, 4) for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"])},
            "kernel_ms_sum": tot / 2}
    print(json.dumps(line))
    if args.out:
        rows = [dict(kernel=k[0], layer=k[1], ms=v["ms"] / 2, calls=v["calls"] / 2, bytes=v["bytes"] / 2, flops=v["flops"] / 2)
                for k, v in prof.by_layer().items()]
        rows.sort(key=lambda r: -r["ms"])
        with open(args.out, "w") as f:
            for r in rows:
                gbs = r["bytes"] / (r["ms"] * 1e6) if r["ms"] > 0 else 0
                tfs = r["flops"] / (r["ms"] * 1e9) if r["ms"] > 0 else 0
                f.write(f"{r['ms']:9.4f} ms  {r['calls']:5.1f}x  {gbs:8.1f} GB/s {tfs:8.2f} TF/s  {r['kernel']:28s} {r['layer']}\n")


def cfg1(args, dev):
    """BASELINE.json configs[0] (with the shape correction of SURVEY.md section 0.4): README 3D model, eval-mode forward of
    one 1 x 4 x 256 x 256 x 32 stack (fp32 NCDHW in, fp32 logits 1 x 1 x 68 x 68 x 27 out), through model(x)."""
    model = H.Unet_Constructor(**README_3D)
    model.precision = args.precision
    model = model.to(dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, 4, 256, 256, 32), generator=g)
    xh = x.pin_memory()
    xd = x.to(dev)
    with torch.no_grad():
        model.train()
        model(xd)           # populate the BatchNorm running statistics
        model.eval()
        for _ in range(5):
            y = model(xd)
        torch.cuda.synchronize()
        ts, te = [], []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y = model(xd); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        for _ in range(20):   # end to end: pinned host input -> device, forward, logits back to the host
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); yh = model(xh.to(dev, non_blocking=True)).cpu(); e1.record()
            torch.cuda.synchronize()
            te.append(e0.elapsed_time(e1))
    ms, mse = sorted(ts)[len(ts) // 2], sorted(te)[len(te) // 2]
    from hcunet_b200.graph import GraphedForward
    gf = GraphedForward(model, xd)
    yg = gf(xd)
    torch.cuda.synchronize()
    same = bool(torch.equal(yg, y))
    tg, tge = [], []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gf(xd); e1.record()
        torch.cuda.synchronize()
        tg.append(e0.elapsed_time(e1))
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); yh = gf(xh).cpu(); e1.record()
        torch.cuda.synchronize()
        tge.append(e0.elapsed_time(e1))
    msg, msge = sorted(tg)[len(tg) // 2], sorted(tge)[len(tge) // 2]
    nvox = 256 * 256 * 32
    print(json.dumps({"config": "cfg1", "workload": "README 3D U-Net eval forward, 1x4x256x256x32", "precision": args.precision,
                      "out_shape": list(y.shape), "ms_forward": ms, "voxels_per_s": nvox / (ms / 1e3), "ms_e2e": mse,
                      "e2e_voxels_per_s": nvox / (mse / 1e3), "ms_forward_graph": msg, "graph_voxels_per_s": nvox / (msg / 1e3),
                      "ms_e2e_graph": msge, "graph_equals_eager": same, "fwd_gflop": 10205 * nvox / 1e9,
                      "finite": bool(torch.isfinite(y).all())}))


def cfg5(args, dev):
    """BASELINE.json configs[4]: whole-cochlea synthetic stack [1, 4, 2048, 2048, 128] (fp16, pinned host memory), overlap
    tiles sharded over `world` ranks with no collective (hcunet_b200.tiling); this process runs rank 0's share on one
    GPU: tiles copied host -> device one by one, eval-mode forward (BatchNorm folded into the conv epilogues), logits
    written into the rank's output volume on the device."""
    from hcunet_b200 import tiling

    model = H.Unet_Constructor(**README_3D)
    model.precision = args.precision
    model = model.to(dev)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():   # populate the running statistics on one small training-mode forward
        model.train()
        model(torch.randn((1, 4, 256, 256, 32), generator=g).half().to(dev))
        model.eval()
    X = Y = 2048
    Z = 128
    stack = torch.empty((1, 4, X, Y, Z), dtype=torch.float16).pin_memory()
    for c in range(4):          # cheap synthetic content (randn of 2 G elements on the host takes minutes)
        stack[0, c] = torch.randn((X, 1, 1), generator=g).half() * torch.randn((1, Y, Z), generator=g).half()
    align, margin, mz = tiling.tile_geometry(model.model_specification)
    torch.cuda.synchronize()
    out, tiles = tiling.predict_tiled(model, stack, tile_out=args.tile_out, world=args.world, rank=0)   # warm-up (step cache)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, tiles = tiling.predict_tiled(model, stack, tile_out=args.tile_out, world=args.world, rank=0, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the forward pass of one interior tile alone (input resident on the device)
    x0, x1, y0, y1 = tiles[0]
    nx, ny = tiling.tile_input_extent(model.model_specification, x1 - x0, X - x0), tiling.tile_input_extent(model.model_specification, y1 - y0, Y - y0)
    xin = tiling._tile_to_device(stack, x0, nx, y0, ny, dev)
    fw = []
    with torch.no_grad():
        for _ in range(4):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(); model(xin); f1.record()
            torch.cuda.synchronize()
            fw.append(f0.elapsed_time(f1))
    ntiles_all = len(tiling.tile_grid(((X - margin) // align * align, (Y - margin) // align * align), args.tile_out, align))
    out_vox = sum((x1 - x0) * (y1 - y0) for x0, x1, y0, y1 in tiles) * (Z - mz)
    in_vox = sum((x1 - x0 + margin) * (y1 - y0 + margin) for x0, x1, y0, y1 in tiles) * Z
    line = {"config": "cfg5", "workload": "README 3D U-Net eval: overlap-tile inference of a 4x2048x2048x128 stack, rank 0's share",
            "world": args.world, "tile_out": args.tile_out, "tiles_this_rank": len(tiles), "tiles_total": ntiles_all,
            "align": align, "margin": margin, "ms_this_rank": ms, "output_voxels_this_rank": out_vox,
            "input_voxels_this_rank": in_vox, "output_voxels_per_s": out_vox / (ms / 1e3),
            "input_voxels_per_s": in_vox / (ms / 1e3), "finite": bool(torch.isfinite(out).all()),
            "tile_input": [4, nx, ny, Z], "tile_forward_ms": sorted(fw)[len(fw) // 2],
            "tile_forward_input_voxels_per_s": nx * ny * Z / (sorted(fw)[len(fw) // 2] / 1e3),
            "whole_stack_seconds_at_world": ms / 1e3 * ntiles_all / max(1, len(tiles)) / args.world,
            "note": "tiles copied pinned host -> device inside the timed region; no inter-rank communication"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
