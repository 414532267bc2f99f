#!/usr/bin/env python
"""Where the bench step's time goes: CUDA graphs of (a) the training forward + loss alone, (b) forward + loss + backward,
(c) the whole step with Adam -- each timed with CUDA events over L2-flushed replays -- with the weight gradients on the
side stream (default) and on the main stream (HCUNET_OVERLAP=0, second process).

    python tools/step_breakdown.py [--batch 4] [--z 32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hcunet_b200 as H  # noqa: E402

README_3D = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                 kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
                 upsample_stride=(2, 2, 1), dilation=1, groups=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--z", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = H.Unet_Constructor(**README_3D)
    model.precision = "mixed"
    model = model.to(dev).train()
    # the bench's optimiser: one launch on the flat parameter buffer (HCUNET_OWN_ADAM=0: torch's fused Adam over 136 tensors)
    flat = H.FlatParameters(model) if os.environ.get("HCUNET_OWN_ADAM", "1") != "0" else None
    opt = H.FlatAdam(flat, lr=1e-3) if flat is not None else torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    B, X, Y, Z, C = a.batch, 256, 256, a.z, 4
    loader = H.StackLoader(model)
    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 256, (B, Z, Y, X, C), generator=g, dtype=torch.uint8).to(dev)
    ext = loader.label_extent((B, Z, Y, X, C))
    msk = loader.labels((torch.rand((B, Z, Y, X), generator=g) > 0.7).half().pin_memory(), ext)
    pwl = loader.labels((torch.rand((B, Z, Y, X), generator=g) * 3).half().pin_memory(), ext)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def fwd():
        return H.cross_entropy(model(loader.image(raw)), msk, pwl, "pixel")

    def fwd_bwd():
        opt.zero_grad(set_to_none=True)
        loss = fwd()
        loss.backward()
        if flat is not None:
            flat.sync_grad()
        return loss

    def full():
        fwd_bwd()
        opt.step()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            full()
        with torch.no_grad():
            for _ in range(3):
                fwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = {}
    for name, fn, ng in (("forward+loss (no_grad, batch statistics)", fwd, True), ("forward+loss+backward", fwd_bwd, False),
                         ("whole step (+Adam)", full, False)):
        gr = torch.cuda.CUDAGraph()
        if ng:
            with torch.no_grad(), torch.cuda.graph(gr):
                fn()
        else:
            with torch.cuda.graph(gr):
                fn()
        graphs[name] = gr
    ts = {}
    for name, gr in graphs.items():
        for _ in range(3):
            gr.replay()
        t = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        t.sort()
        ts[name] = t[len(t) // 2]
        print(f"overlap={os.environ.get('HCUNET_OVERLAP', '1')}  {name:45s} {ts[name]:.3f} ms")
    k = list(ts)
    print(f"overlap={os.environ.get('HCUNET_OVERLAP', '1')}  => backward {ts[k[1]] - ts[k[0]]:.3f} ms, Adam {ts[k[2]] - ts[k[1]]:.3f} ms")


if __name__ == "__main__":
    main()
