#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200: 3D U-Net training voxels/s.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # our arm (one JSON line on stdout)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                     # N ranks, data-parallel over patches
    python bench.py --impl reference --steps 2 --warmup 1          # the reference's CPU path (oracle port)

Workload (`config.workload`): BASELINE.json configs[1] with the shape correction of SURVEY.md section 0.4 --
README 3D Unet_Constructor (in=4, out=1, features [8,16,32,64,128], kernels (3,3,2)/(3,3,1)), one training step
= forward + pixel-weighted cross_entropy + backward + Adam, batch 4 of 4x256x256x32 patches per GPU, synthetic
seeded data, reference-default random-init weights.  A voxel is one input spatial site (B*X*Y*Z).

`value`   : device-timed (CUDA events, max over ranks) with the step's inputs already resident in HBM.
`e2e`     : the same step through the public API with HOST inputs -- pinned fp16 image / mask / pwl tensors like
            the reference dataloader yields (`transforms.py:133`), H2D copies (double-buffered on a copy stream)
            and the D2H read of the loss inside the timed region.
`roofline`: the kernel family with the largest share of the step, algorithmic bytes / CUDA-event time.
`cpu_baseline`: the oracle port (reference algorithm, torch CPU fp32) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line.  Native libraries write banners to file descriptor 1 (NCCL prints
# "NCCL version ..." at communicator creation), so fd 1 is pointed at stderr for the whole run and the result line is
# written to the saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "3d_unet_train_voxels_per_s"
UNIT = "voxel/s"
SHAPE = (4, 256, 256, 32)   # C, X, Y, Z of one patch
WORKLOAD = "README 3D Unet_Constructor [8,16,32,64,128] k(3,3,2)/(3,3,1): train step (fwd + pixel-weighted BCE + bwd + Adam)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HCUNET_BENCH_PRECISION", "mixed"), choices=["fp32", "mixed"])
    ap.add_argument("--batch", type=int, default=4, help="patches per GPU per step")
    ap.add_argument("--z", type=int, default=SHAPE[3])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (BASELINE configs 3 / 4 / 5) of the line")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (reference algorithm on torch CPU fp32)
# ----------------------------------------------------------------------------------------------------
def cpu_step_fn(z, batch):
    """One CPU training step of the workload: forward (train-mode BN) + pixel-weighted loss + backward + Adam.
    The UNMODIFIED reference modules (`hcat/unet.py`, `hcat/loss.py`; from /root/reference in the build container, from the
    staged copies under oracle/_ref on the GPU box) when they are there -> kind "reference"; else the oracle port."""
    import torch

    from oracle import ref_loader as R
    from oracle import unet_oracle as O

    x, mask, pwl = O.golden_inputs(O.README_3D, (batch, SHAPE[0], SHAPE[1], SHAPE[2], z), 0)
    image, mask, pwl = x.half(), mask.half(), pwl.half()      # what the reference dataloader yields (transforms.py:133)
    if R.reference_available():
        torch.manual_seed(0)
        model = R.build_reference_unet(**O.README_3D).train()
        loss_fn = R.load_reference_loss().cross_entropy
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)    # tests/r_unet_test.py:24

        def step():  # tests/r_unet_test.py:24-56 pattern
            opt.zero_grad()
            out = model(image.float())
            loss = loss_fn(out, mask, pwl, "pixel")
            loss.backward()
            opt.step()
            return float(loss)

        return step, batch * SHAPE[1] * SHAPE[2] * z, "reference", f"unmodified hcat/unet.py + hcat/loss.py from {R.REFERENCE_ROOT}"

    import hcunet_b200 as H

    torch.manual_seed(0)
    m = H.Unet_Constructor(**O.README_3D)  # parameter container only: identical init to the reference
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    opt = torch.optim.Adam(list(leaves.values()), lr=1e-3)

    def step():
        loss, _, grads, newbuf = O.train_step_grads(sd, O.README_3D, image.float(), mask, pwl)
        for k, p in leaves.items():
            p.grad = grads[k]
        opt.step()
        with torch.no_grad():
            for k, p in leaves.items():
                sd[k] = p.detach()
            sd.update(newbuf)
        return float(loss)

    return step, batch * SHAPE[1] * SHAPE[2] * z, "port", "oracle/unet_oracle.py (oracle/_ref not staged)"


def run_cpu(steps, warmup, z, batch=1):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, vox, kind, what = cpu_step_fn(z, batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    return {"value": vox / t, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} train step(s) (fwd + pixel loss + bwd + Adam) of batch {batch} x {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]}x{z}, "
                      f"{what}, torch {torch.__version__} CPU fp32, {cores} threads",
            "ms_per_step": t * 1e3, "batch": batch}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the GPU arm's config: batch 4 of the patch per step; steps bounded so the run ends within a few minutes
    k_steps, k_warm = max(1, min(args.steps, 60)), max(0, min(args.warmup, 5))
    cb = run_cpu(k_steps, k_warm, args.z, batch=args.batch)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": k_steps, "warmup": k_warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic (seeded), random-init weights",
            "config": {"workload": WORKLOAD, "patch": [SHAPE[0], SHAPE[1], SHAPE[2], args.z], "batch_per_gpu": args.batch,
                       "global_batch": args.batch, "precision": "fp32", "parallelism": "cpu", "optimizer": "Adam lr 1e-3",
                       "note": "the reference's own CPU path on this box's host cores; kind 'reference' = the unmodified "
                               "reference modules staged under oracle/_ref by build(), 'port' = the oracle restatement",
                       "timed_steps": k_steps},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def _stage(msg):
    """Progress marker on stderr (multi-GPU hangs are otherwise silent)."""
    if os.environ.get("HCUNET_BENCH_VERBOSE", "0") != "0":
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


# ---- extra legs: the other BASELINE.json configurations, reported inside the same JSON line ("extra") ---------------

def _timed_region(dist, world, dev, fn, steps):
    """`steps` calls of fn between barrier + synchronize, CUDA events, max over ranks -> seconds per call."""
    import torch

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / steps


def extra_cfg4(H, dist, world, rank, dev, model, opt, sync, precision, steps, flat=None):
    """BASELINE configs[3]: the same model and step on 256 x 256 x 64 patches, batch 4 per GPU, data-parallel (NCCL gradient
    all-reduce at world > 1), inputs resident, one CUDA graph per step."""
    import torch
    from hcunet_b200.graph import GraphedTrainStep

    B, C, X, Y, Z = 4, 4, 256, 256, 64
    loader = H.StackLoader(model)
    g = torch.Generator().manual_seed(4321 + rank)
    raw = torch.randint(0, 256, (B, Z, Y, X, C), generator=g, dtype=torch.uint8).to(dev)
    ext = loader.label_extent((B, Z, Y, X, C))
    msk = loader.labels((torch.rand((B, Z, Y, X), generator=g) > 0.7).half().pin_memory(), ext)
    pwl = loader.labels((torch.rand((B, Z, Y, X), generator=g) * 3).half().pin_memory(), ext)

    def eager():
        if flat is not None:
            flat.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        H.cross_entropy(model(loader.image(raw)), msk, pwl, "pixel").backward()
        if flat is not None:
            flat.sync_grad()
        sync.allreduce()
        opt.step()

    for _ in range(3):   # this shape's step cache
        eager()
    gstep = GraphedTrainStep(model, opt, lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel"), (raw, msk, pwl),
                             grad_sync=sync.allreduce if world > 1 else None, input_fn=loader.image, flat=flat)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step(i):
        flush.zero_()
        gstep.run()

    for i in range(3):
        step(i)
    t = _timed_region(dist, world, dev, step, steps)
    t_flush = _timed_region(dist, world, dev, lambda i: flush.zero_(), steps)
    t = max(1e-9, t - t_flush)
    del gstep
    return {"workload": "BASELINE configs[3]: README 3D model, train step on 256x256x64 patches, data-parallel",
            "patch": [C, X, Y, Z], "batch_per_gpu": B, "global_batch": B * world, "n_gpus": world, "precision": precision,
            "ms_per_step": t * 1e3, "value": B * X * Y * Z * world / t, "unit": UNIT, "steps": steps,
            "timing": "CUDA events, max over ranks, inputs resident, L2 flushed between steps (flush time subtracted)"}


def extra_cfg5(H, dist, world, rank, dev, precision):
    """BASELINE configs[4]: overlap-tile inference of a synthetic 4 x 2048 x 2048 x 128 stack (fp16, pinned host memory) with
    the README model in eval mode, the tile grid sharded over the ranks with NO communication (hcunet_b200.tiling): every
    rank copies its tiles host -> device and runs them; the time is the max over ranks."""
    import torch
    from hcunet_b200 import tiling

    kw = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
              kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
              upsample_stride=(2, 2, 1), dilation=1, groups=1)
    torch.manual_seed(0)
    model = H.Unet_Constructor(**kw)
    model.precision = precision
    model = model.to(dev)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        model.train()
        model(torch.randn((1, 4, 256, 256, 32), generator=g).half().to(dev))
        model.eval()
    X = Y = 2048
    Z = 128
    tile_out = 512 if world > 1 else 1024
    align, margin, mz = tiling.tile_geometry(model.model_specification)
    ox, oy = tiling.tiled_output_extent(model.model_specification, X), tiling.tiled_output_extent(model.model_specification, Y)
    all_tiles = tiling.tile_grid((ox, oy), tile_out, align)
    mine = tiling.shard_tiles(all_tiles, world, rank)
    # this rank's rows of the stack only (its tiles' input footprint), cheap synthetic content
    stack = torch.empty((1, 4, X, Y, Z), dtype=torch.float16).pin_memory()
    xs = sorted({t[0] for t in mine}) or [0]
    x_lo, x_hi = xs[0], min(X, max(t[1] for t in mine) + margin + align) if mine else 1
    for c in range(4):
        stack[0, c, x_lo:x_hi] = torch.randn((x_hi - x_lo, 1, 1), generator=g).half() * torch.randn((1, Y, Z), generator=g).half()
    out, tiles = tiling.predict_tiled(model, stack, tile_out=tile_out, world=world, rank=rank)     # warm-up (step cache)
    t = _timed_region(dist, world, dev, lambda i: tiling.predict_tiled(model, stack, tile_out=tile_out, world=world, rank=rank,
                                                                       out=out), 2)
    finite = bool(torch.isfinite(out).all())
    del out, stack
    return {"workload": "BASELINE configs[4]: README 3D model eval, overlap-tile inference of a 4x2048x2048x128 fp16 stack from "
                        "pinned host memory, tiles sharded over the ranks (no communication)",
            "n_gpus": world, "tile_out": tile_out, "tiles_total": len(all_tiles), "tiles_this_rank": len(tiles),
            "precision": precision, "seconds_whole_stack": t, "input_voxels_per_s": X * Y * Z / t,
            "output_voxels_per_s": ox * oy * (Z - mz) / t, "finite": finite,
            "timing": "CUDA events around every rank's share (tile H2D copies inside), max over ranks"}


def extra_cfg3(H, dev, precision, steps):
    """BASELINE configs[2]: classic 2D U-Net (the constructor's defaults: features [32..1024], 3x3, ConvTranspose 2x2 s2), train
    step on batch 16 of 572 x 572 x 3 tiles, one GPU, inputs resident."""
    import torch
    from hcunet_b200.graph import GraphedTrainStep

    B = 16
    torch.manual_seed(0)
    model = H.Unet_Constructor()
    model.precision = precision
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    g = torch.Generator().manual_seed(1)
    img = torch.randn((B, 3, 572, 572), generator=g).half().to(dev)
    msk = (torch.rand((B, 2, 572, 572), generator=g) > 0.7).half().to(dev)
    pwl = (torch.rand((B, 2, 572, 572), generator=g) * 3).half().to(dev)
    loss_fn = lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel")  # noqa: E731
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss_fn(model(img), msk, pwl).backward()
        opt.step()
    gstep = GraphedTrainStep(model, opt, loss_fn, (img, msk, pwl))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step(i):
        flush.zero_()
        gstep.run()

    for i in range(2):
        step(i)
    t = _timed_region(None, 1, dev, step, steps) - _timed_region(None, 1, dev, lambda i: flush.zero_(), steps)
    flops = 3 * 169471 * 572 * 572 * B
    del gstep
    return {"workload": "BASELINE configs[2]: classic 2D U-Net [32..1024] k(3,3), ConvTranspose2d s2, train step on 572x572x3 tiles",
            "batch": B, "n_gpus": 1, "precision": precision, "ms_per_step": t * 1e3, "pixels_per_s": B * 572 * 572 / t,
            "train_tflops": flops / t / 1e12, "of_bf16_sustained_peak": None, "steps": steps}


def main_ours(args):
    import torch
    import torch.distributed as dist

    import hcunet_b200 as H
    from hcunet_b200 import _lib
    from hcunet_b200.parallel import GradSync
    from oracle import unet_oracle as O  # cpu_baseline leg + workload kwargs only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _stage("process group up")
    B, (C, X, Y), Z = args.batch, SHAPE[:3], args.z

    torch.manual_seed(0)
    model = H.Unet_Constructor(**O.README_3D)
    model.precision = args.precision
    model = model.to(dev).train()
    sync = GradSync(model, world)          # broadcast params from rank 0; fp32 mean all-reduce of the gradients
    # HCUNET_AR_OVERLAP=1: the exchange is launched from inside backward (two buckets, communication stream) and captured with
    # the step's graph (GradSync.attach).  Measured on 8 B200 (profiles/r02_bench_n8_*.json): 3.01 ms per step against 2.97 ms
    # for the serial form (graph -> one all-reduce -> optimiser graph): the 2.9 MB collective costs less than the SMs its
    # kernels take from the backward, so the serial form is the default.
    overlap_ar = world > 1 and os.environ.get("HCUNET_AR_OVERLAP", "0") != "0"
    if overlap_ar:
        sync.attach()
    _stage("parameters broadcast")
    use_graph = os.environ.get("HCUNET_BENCH_GRAPH", "1") != "0"
    # the optimiser holds ONE flat parameter (every nn.Parameter is a view of it; its gradient is the engine's flat gradient
    # buffer): one elementwise Adam launch instead of a multi-tensor pass over 136 tensors + 136 step counters (0.10 -> 0.01 ms)
    flat = H.FlatParameters(model) if os.environ.get("HCUNET_FLAT_ADAM", "1") != "0" else None
    own_adam = flat is not None and os.environ.get("HCUNET_OWN_ADAM", "1") != "0"
    if own_adam:
        # the whole optimiser step (non-finite check of the fp16-storage path + Adam) is ONE launch of the library (hcu_adam_flat)
        opt = H.FlatAdam(flat, lr=1e-3)
    else:
        opt = torch.optim.Adam([flat.flat] if flat is not None else model.parameters(), lr=1e-3, fused=True, capturable=use_graph)
        if flat is not None and os.environ.get("HCUNET_GUARD", "1") != "0":
            flat.guard(opt)     # skip-step on non-finite gradients (fp16 storage), one pass over the flat gradient per step

    def zero_grads():
        if flat is not None:
            flat.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)

    def grads_ready():
        if flat is not None:
            flat.sync_grad()

    # Synthetic patches as the reference dataloader finds them on disk (dataloader.py:40-58): the RAW stack [Z, Y, X, C] uint8
    # as skimage.io.imread yields it, mask / pwl [Z, Y, X].  hcunet_b200.StackLoader is the input path (to_float -> reshape ->
    # normalize -> to_tensor on the device + the origin crop of mask / pwl the loss reads, loss.py:51-56).
    # NBUF distinct batches so consecutive steps never reuse a cached input.
    NBUF = 3
    loader = H.StackLoader(model)
    ext = loader.label_extent((B, Z, Y, X, C))
    g = torch.Generator().manual_seed(1234 + rank)
    host = []
    for _ in range(NBUF):
        img = torch.randint(0, 256, (B, Z, Y, X, C), generator=g, dtype=torch.uint8).pin_memory()
        msk = (torch.rand((B, Z, Y, X), generator=g) > 0.7).half().pin_memory()
        pwl = (torch.rand((B, Z, Y, X), generator=g) * 3).half().pin_memory()
        host.append((img, msk, pwl))
    # resident inputs: the raw stack in HBM + the label crops the loss reads
    resident = [(h[0].to(dev), loader.labels(h[1], ext), loader.labels(h[2], ext)) for h in host]
    crop_bytes = 2 * B * ext[0] * ext[1] * ext[2] * 2      # mask + pwl crops, fp16, read in place from pinned host memory
    h2d_bytes = host[0][0].numel() * host[0][0].element_size() + crop_bytes
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def eager_step(raw, msk, pwl):
        zero_grads()
        logits = model(loader.image(raw))
        loss = H.cross_entropy(logits, msk, pwl, "pixel")
        loss.backward()
        grads_ready()
        sync.allreduce()
        opt.step()
        return loss

    for _ in range(2):  # the first two steps record the step cache (per-layer weight packs / scatters, more launches)
        eager_step(*resident[0])
    _stage("eager recording steps done")
    l0 = _lib.launch_count()
    eager_step(*resident[0])
    torch.cuda.synchronize()
    _stage("eager steady-state step done")
    launches_per_step = _lib.launch_count() - l0  # library kernels of one steady-state step (graph replays re-launch the same set)
    if use_graph:
        from hcunet_b200.graph import GraphedTrainStep

        gstep = GraphedTrainStep(model, opt, lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel"), resident[0],
                                 grad_sync=sync.allreduce if world > 1 else None, input_fn=loader.image, flat=flat)
        # one CUDA-graph launch per step; the inputs are copied into the graph's static buffers (device->device here,
        # pinned host->device in the e2e leg) inside the timed region
        step = gstep
        _stage("graph captured")
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` steps between barrier+synchronize, CUDA events, max over ranks -> seconds."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- resident-input throughput (`value`) ----------------------------------------------------
    def res_step(i):
        flush.zero_()  # L2 flush between iterations (256 MB write); inside the region, ~40 us
        step(*resident[i % NBUF])

    for i in range(args.warmup):
        res_step(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _stage("warm-up done")
    t_res = timed(res_step, args.steps)
    _stage("timed region done")
    launches = launches_per_step * args.steps
    # the flush is not part of the workload: time it alone and subtract
    t_flush = timed(lambda i: flush.zero_(), args.steps)
    t_step = max(1e-9, (t_res - t_flush) / args.steps)
    vox_per_step = B * X * Y * Z * world

    # ---- end to end (`e2e`): pinned host inputs, H2D on a copy stream one step ahead, loss read back ----
    copy_stream = torch.cuda.Stream(device=dev)
    if use_graph:
        # Two captured graphs of the same step (same model, same optimiser), each with its own static input buffers: the
        # copy stream fills graph B's inputs from pinned host memory while graph A runs -- no device-to-device staging copy.
        gstep2 = GraphedTrainStep(model, opt, lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel"), resident[0],
                                  grad_sync=sync.allreduce if world > 1 else None, input_fn=loader.image, flat=flat)
        gsteps = [gstep, gstep2]
        slots = [tuple(g_.static_in) for g_ in gsteps]
        _stage("second graph captured (double-buffered inputs)")
    else:
        gsteps = None
        slots = [tuple(torch.empty_like(t) for t in resident[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            h = host[i % NBUF]
            slots[s][0].copy_(h[0], non_blocking=True)          # the raw uint8 stack: the only bulk H2D copy
            loader.labels(h[1], ext, out=slots[s][1])           # mask / pwl: the crop is gathered from pinned host memory
            loader.labels(h[2], ext, out=slots[s][2])
            ready[s].record(copy_stream)

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()     # D2H landing slots for the step's loss
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(steps):
        """Every step: inputs pinned host -> device (copy stream, one step ahead), the step, and a D2H read of its loss
        (4 bytes into pinned memory).  The host consumes the loss of step i after it has enqueued step i + 1, so the GPU
        never idles behind a host round trip (a `loss.item()` right after each step cost ~0.1 ms per step)."""
        for s in range(2):
            freed[s].record()
        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            loss = gsteps[s].run() if gsteps is not None else step(*slots[s])
            freed[s].record()
            loss_host[s:s + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_done[s].record()
            if i >= 1:
                loss_done[1 - s].synchronize()
                losses.append(float(loss_host[1 - s]))
        loss_done[(steps - 1) % 2].synchronize()
        losses.append(float(loss_host[(steps - 1) % 2]))

    e2e_run(max(2, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t) / args.steps
    _stage("e2e region done")
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event profile of the same step -> roofline of the dominant kernel ------------
    roof = None
    if rank == 0 and not args.no_profile:
        from hcunet_b200 import profiler

        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        prof = profiler.KernelProfile()
        hook_saved, model._engine.grad_ready_hook = model._engine.grad_ready_hook, None   # rank 0 alone: no collective here
        with prof:
            for i in range(min(args.steps, 5)):
                # Park the GPU behind a ~12 ms spin so the host enqueues the whole step before the first kernel runs:
                # the event pairs then bracket kernel time, not host launch latency (small kernels were inflated 2-10x).
                torch.cuda._sleep(int(12e-3 * 1.9e9))
                # eager: per-kernel events need individual launches.  No gradient all-reduce here: only rank 0 profiles
                # (a collective entered by one rank would dead-lock), and the collective is not one of this library's kernels
                zero_grads()
                H.cross_entropy(model(loader.image(resident[i % NBUF][0])), resident[i % NBUF][1], resident[i % NBUF][2],
                                "pixel").backward()
                grads_ready()
                opt.step()
        model._engine.grad_ready_hook = hook_saved
        roof = prof.roofline(peaks, t_step * min(args.steps, 5))
        # DRAM traffic of the dominant kernel: not measurable live (needs ncu); the committed capture of the same command
        # (tools/gpu_profile.sh -> profiles/*_traffic.json) is reported per launch, like `achieved`
        try:
            import glob
            for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
                tj = json.load(open(f))
                if roof is not None and tj.get("kernel") == roof.get("kernel"):
                    roof["traffic"] = tj["dram_bytes_per_launch"]
                    roof["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
                    roof["traffic_source"] = os.path.relpath(f, ROOT)
                    roof["algorithmic_bytes_per_launch"] = roof["achieved"] * 1e9 * roof["avg_launch_ms"] / 1e3 \
                        if roof.get("unit") == "GB/s" else None
        except Exception:
            pass
        if os.environ.get("HCUNET_PROFILE_OUT"):
            nst = min(args.steps, 5)
            rows = [dict(kernel=k[0], layer=k[1], ms=v["ms"] / nst, calls=v["calls"] / nst, bytes=v["bytes"] / nst,
                         flops=v["flops"] / nst) for k, v in prof.by_layer().items()]
            rows.sort(key=lambda r: -r["ms"])
            with open(os.environ["HCUNET_PROFILE_OUT"], "w") as f:
                for r in rows:
                    gbs = r["bytes"] / (r["ms"] * 1e6) if r["ms"] > 0 else 0
                    tfs = r["flops"] / (r["ms"] * 1e9) if r["ms"] > 0 else 0
                    f.write(f"{r['ms']:9.4f} ms  {r['calls']:5.1f}x  {gbs:8.1f} GB/s {tfs:8.2f} TF/s  {r['kernel']:28s} {r['layer']}\n")

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": vox_per_step / t_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp16 storage, fp32 accumulate" if args.precision == "mixed" else "fp32",
                "data": "synthetic (seeded), random-init weights",
                "config": {"workload": WORKLOAD, "patch": [C, X, Y, Z], "batch_per_gpu": B, "global_batch": B * world,
                           "precision": args.precision, "parallelism": f"dp{world}",
                           "optimizer": ("hcunet_b200.FlatAdam lr 1e-3 (one launch: skip-step check + Adam on the flat parameter buffer)" if own_adam else
                                         "Adam(fused) lr 1e-3" + (" on the flat parameter buffer (hcunet_b200.FlatParameters)" if flat is not None else "")),
                           "gradient_exchange": None if world == 1 else (
                               "NCCL all-reduce (mean, fp32, in place on the engine's flat gradient buffer) in two buckets launched "
                               "from inside backward on a communication stream, captured in the step's CUDA graph"
                               if overlap_ar else "one NCCL all-reduce between the backward graph and the optimiser graph"),
                           "input": "raw uint8 stack [B,Z,Y,X,C] -> hcunet_b200.StackLoader (hcu_load_stack inside the step)",
                           "launch": "one CUDA graph per step (hcunet_b200.graph.GraphedTrainStep)" if use_graph else "eager",
                           "l2": "256 MB buffer zeroed between iterations, its time measured alone and subtracted",
                           "output_voxels_per_step": B * world * 68 * 68 * (Z - 5)},
                "e2e": {"value": vox_per_step / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4, "ms_per_step": t_e2e * 1e3,
                        "note": "host inputs = the raw stack as the reference dataloader reads it (uint8 [B,Z,Y,X,C], pinned) + fp16 mask / pwl [B,Z,Y,X] (pinned); per step, on a copy stream one step ahead: the raw stack H2D straight into the static input of one of two alternating CUDA graphs, and the origin crop of mask / pwl that the loss reads (loss.py:51-56) gathered in place from pinned host memory (hcu_load_labels); the graph starts with hcu_load_stack (to_float / reshape / normalize / to_tensor on the device); every step's loss is copied D2H into pinned memory and read by the host one step later"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
                "loss_first_last": [losses[0], losses[-1]] if losses else None}
    # ---- the other BASELINE configurations, inside the same line ----
    extra = None
    if not args.no_extra:
        extra = {}
        k4 = max(3, min(args.steps, 10))
        for name, fn in (("cfg4", lambda: extra_cfg4(H, dist if world > 1 else None, world, rank, dev, model, opt, sync, args.precision, k4, flat)),
                         ("cfg5", lambda: extra_cfg5(H, dist if world > 1 else None, world, rank, dev, args.precision)),
                         ("cfg3", (lambda: extra_cfg3(H, dev, args.precision, 3)) if world == 1 else None)):
            if fn is None:
                continue
            try:
                extra[name] = fn()
            except Exception as e:   # an extra leg must never take the headline line down (all ranks see the same exception path)
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            _stage(f"extra {name} done")
            torch.cuda.empty_cache()
        if rank == 0 and extra.get("cfg3") and "train_tflops" in extra["cfg3"]:
            try:
                pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                extra["cfg3"]["of_bf16_sustained_peak"] = extra["cfg3"]["train_tflops"] / pk["bf16_tflops_sustained"]
            except Exception:
                pass
        if line is not None:
            line["extra"] = extra
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb = run_cpu(8, 2, Z, batch=1)   # bounded sample (batch 1 of the same patch): ~5 s of CPU work on the box's cores
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        # captured collectives: drop the graphs before the communicator, and never let a teardown problem hold the ranks
        import gc
        t_exit = threading.Timer(30.0, lambda: os._exit(0))
        t_exit.daemon = True
        t_exit.start()
        gstep = gstep2 = gsteps = step = None   # noqa: F841
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
