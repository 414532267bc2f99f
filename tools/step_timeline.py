#!/usr/bin/env python
"""Timeline of one steady-state training step of the bench model with the weight gradients on their side stream: start / end of
every C-ABI call relative to the first one, per stream, from CUDA events (the GPU is parked behind a spin while the host enqueues
the whole step, so the events see back-to-back kernels on both streams, as inside the CUDA graph).

    HCUNET_PROFILE_OVERLAP=1 python tools/step_timeline.py > gpurun_out/timeline.txt
"""
import os
import sys

import torch

os.environ.setdefault("HCUNET_PROFILE_OVERLAP", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hcunet_b200 as H  # noqa: E402
from hcunet_b200 import _lib, profiler  # noqa: E402

README_3D = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                 kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
                 upsample_stride=(2, 2, 1), dilation=1, groups=1)


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = H.Unet_Constructor(**README_3D)
    model.precision = "mixed"
    model = model.to(dev).train()
    B, X, Y, Z, C = 4, 256, 256, 32, 4
    loader = H.StackLoader(model)
    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 256, (B, Z, Y, X, C), generator=g, dtype=torch.uint8).to(dev)
    ext = loader.label_extent((B, Z, Y, X, C))
    msk = loader.labels((torch.rand((B, Z, Y, X), generator=g) > 0.7).half().pin_memory(), ext)
    pwl = loader.labels((torch.rand((B, Z, Y, X), generator=g) * 3).half().pin_memory(), ext)

    def step():
        model.zero_grad(set_to_none=True)
        H.cross_entropy(model(loader.image(raw)), msk, pwl, "pixel").backward()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    prof = profiler.KernelProfile()
    streams = {}
    orig_append = prof.records.append

    class Rec(list):
        def append(self, item):
            list.append(self, item + (torch.cuda.current_stream().cuda_stream,))

    prof.records = Rec()
    with prof:
        torch.cuda._sleep(int(30e-3 * 1.9e9))
        base = torch.cuda.Event(enable_timing=True)
        base.record()
        step()
        torch.cuda.synchronize()
    rows = []
    for name, note, e0, e1, st in prof.records:
        sid = streams.setdefault(st, len(streams))
        rows.append((base.elapsed_time(e0) * 1e3, base.elapsed_time(e1) * 1e3, sid, name.replace("hcu_", ""), note[0] if note else ""))
    rows.sort()
    t0 = rows[0][0]
    end = max(r[1] for r in rows) - t0
    print(f"# one training step (fwd + loss + bwd), {len(rows)} C-ABI calls, {end:.0f} us from the first kernel's start to the last one's end")
    print("# start_us   end_us  dur_us  stream  call  layer")
    busy = {}
    for a, b, sid, name, layer in rows:
        print(f"{a - t0:9.1f} {b - t0:8.1f} {b - a:7.1f}  s{sid}  {name:26s} {layer}")
        busy[sid] = busy.get(sid, 0.0) + (b - a)
    for sid, v in sorted(busy.items()):
        print(f"# stream s{sid}: {v:.0f} us inside calls")
    # idle gaps of the main stream while the side stream works, and the tail after the main stream's last call
    main_rows = [r for r in rows if r[2] == 0]
    side_rows = [r for r in rows if r[2] != 0]
    if side_rows:
        print(f"# main stream's last call ends at {max(r[1] for r in main_rows) - t0:.0f} us, side stream's at {max(r[1] for r in side_rows) - t0:.0f} us")


if __name__ == "__main__":
    main()
