#!/usr/bin/env python
"""Free-running mixed step vs the fp16-storage emulation, tensor by tensor (no teacher forcing): where does a deviation
enter and how does it grow?   python tools/mixed_trace.py g2d_small"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from oracle import mixed_oracle as M  # noqa: E402

import hcunet_b200 as H  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main(name):
    fx = load_golden(name)
    taps = {}
    loss_e, logits_e, grads_e, _ = M.train_step(fx["state_dict"], fx["kwargs"], fx["x"], fx["mask"], fx["pwl"], taps=taps)
    m = H.Unet_Constructor(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"])
    m.precision = "mixed"
    m = m.cuda().train()

    def tap(tag, t, c, sz):
        ref = taps[tag]
        if tag.endswith(".bn"):
            print(f"  {tag:34s} {max(rel(t[i], ref[i]) for i in range(4)):.2e}")
            return
        b, ch = ref.shape[:2]
        want = ref.reshape(b, ch, -1).permute(0, 2, 1)
        got = t[:, :, :c].float().cpu()
        ndiff = int((got != want).sum())
        print(f"  {tag:34s} {rel(got, want):.2e}   {ndiff} of {want.numel()} elements differ")

    m._engine.tap = tap
    logits = m(fx["x"].cuda())
    loss = H.cross_entropy(logits, fx["mask"].cuda(), fx["pwl"].cuda(), "pixel")
    loss.backward()
    for k, p in m.named_parameters():
        print(f"  grad {k:34s} {rel(p.grad, grads_e[k]):.2e}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "g2d_small")
