"""Data-parallel overlap on one GPU: the engine's `grad_ready_hook` (hcunet_b200.parallel.GradSync.attach) is called twice
per backward -- [offset of the first early level, n) when the up path and the two deepest levels are done, [0, that offset)
at the end -- with events that really cover the range (the hook reads it on another stream), and the gradients are the
same as without the hook.  Steady-state steps (batched scatters, weight gradients on the side stream)."""
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def test_backward_reports_two_final_gradient_ranges():
    import hcunet_b200 as H

    kw = dict(O.README_3D, feature_sizes=[8, 16, 32, 64])
    torch.manual_seed(0)
    m = H.Unet_Constructor(**kw)
    m.precision = "mixed"
    m = m.cuda().train()
    x, mask, pwl = O.golden_inputs(kw, (2, 4, 108, 108, 8), 5)
    x, mask, pwl = x.cuda(), mask.cuda(), pwl.cuda()

    def step():
        m.zero_grad(set_to_none=True)
        H.cross_entropy(m(x), mask, pwl, "pixel").backward()
        torch.cuda.synchronize()
        return torch.cat([p.grad.flatten() for p in m.parameters()]).clone()

    for _ in range(3):
        want = step()
    calls, snaps = [], []
    comm = torch.cuda.Stream()

    def hook(flat, lo, hi, events):
        for ev in events:
            comm.wait_event(ev)
        with torch.cuda.stream(comm):
            snaps.append(flat[lo:hi].clone())      # what a collective launched now would read
            done = torch.cuda.Event()
            done.record()
        calls.append((lo, hi))
        return done

    m2 = H.Unet_Constructor(**kw)
    m2.load_state_dict(m.state_dict())
    m2.precision = "mixed"
    m2 = m2.cuda().train()
    m2._engine.grad_ready_hook = hook
    m, m_ref = m2, m
    for _ in range(3):        # recording steps (bucket split recorded with the step cache), then a steady-state step
        calls.clear(); snaps.clear()
        got = step()
    n = want.numel()
    names = [k for k, _ in m.named_parameters()]
    first_early = sum(p.numel() for k, p in m.named_parameters() if k.startswith("out_conv") or k.startswith("down_steps.0.")
                      or k.startswith("down_steps.1."))
    assert calls == [(first_early, n), (0, first_early)], (calls, first_early, n, names[:3])
    # the ranges were final when their events fired
    assert torch.equal(snaps[0], got[first_early:]) and torch.equal(snaps[1], got[:first_early])
    # same gradients as the un-hooked model (weight-gradient atomics: ~1e-6)
    live = (want.abs() > 0)
    assert float((got - want).norm() / want.norm()) <= 1e-4 and bool(live.any())
