"""Per-kernel parity (GPU): each hand-written kernel on its own, through the C ABI, against the fp32 oracle
op (torch CPU) on the same inputs.  This is where the mixed path's 2e-3 gate lives: one layer, same inputs,
rel-L2 <= 2e-3 (fp16 operands, fp32 accumulate; measured ~3e-4)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_cl(x, cpitch=None, dtype=torch.float16):
    """[N,C,X,Y,Z] -> channels-last [N,X,Y,Z,cpitch] (zero padded channels)."""
    n, c = x.shape[:2]
    cp = cpitch or c
    out = torch.zeros((n,) + tuple(x.shape[2:]) + (cp,), dtype=dtype)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(dtype)
    return out.cuda().contiguous()


def from_cl(y):
    return y.permute(0, 4, 1, 2, 3).float().cpu()


def run_tc(x, w, *, bias=None, dil=(1, 1, 1), pad=(0, 0, 0), cpitch=None, in_affine=None, in_relu=False,
           out_affine=None, out_relu=False, want_stats=False, out_f32=False, use_simt=False, hint=0, tile=0):
    """x [N,Cin,X,Y,Z] fp32 (fp16-representable), w [Cout,Cin,kx,ky,kz] -> (y [N,Cout,...], stats or None)."""
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    n, cin = x.shape[:2]
    cout = w.shape[0]
    taps = tuple(w.shape[2:])
    isz = tuple(x.shape[2:])
    osz = tuple(isz[i] + 2 * pad[i] - (taps[i] - 1) * dil[i] for i in range(3))
    cp = cpitch or cin
    xin = to_cl(x, cp)
    odt = torch.float32 if out_f32 else torch.float16
    y = torch.full((n,) + osz + (cout,), float("nan"), dtype=odt, device="cuda")
    d = conv_desc(_lib.F16, _lib.F32 if out_f32 else _lib.F16, n, isz, cp, 0, cin, cin, osz, osz, cout, 0, cout, 1, taps,
                  dil, pad=pad, in_relu=int(in_relu), out_relu=int(out_relu))
    d.reserved[0] = hint   # 0 auto, 1 classic kernel, 2 K-streamed kernel
    d.reserved[1] = tile   # forced K-streamed tile: MB | Nc << 8 | PC << 20
    if hint:
        buf = C.create_string_buffer(256)
        lib.hcu_conv_tc_describe(C.byref(d), buf, 256)
        if tile and b"does not fit" in buf.value:
            pytest.skip("forced tile does not fit in shared memory")
        assert buf.value.decode().startswith("ks " if hint == 2 else "classic "), buf.value
    wg = w.permute(2, 3, 4, 1, 0).contiguous().float().cuda()  # [taps][cin][cout]
    stats = torch.zeros((_lib.STAT_BINS, 2, cout), dtype=torch.float64, device="cuda") if want_stats else None
    isc = ish = osc = osh = None
    if in_affine is not None:
        isc = torch.zeros(cp, device="cuda"); ish = torch.zeros(cp, device="cuda")
        isc[:cin], ish[:cin] = in_affine[0].cuda(), in_affine[1].cuda()
    if out_affine is not None:
        osc, osh = out_affine[0].cuda().contiguous(), out_affine[1].cuda().contiguous()
    b = bias.cuda() if bias is not None else None
    if use_simt:
        _lib.check(lib.hcu_conv_fwd(C.byref(d), P(xin), P(wg), P(b), P(isc), P(ish), P(osc), P(osh), P(y), P(stats),
                                    stream()), "conv_fwd")
    else:
        assert lib.hcu_conv_tc_supported(C.byref(d)) == 1
        nb = lib.hcu_conv_tc_packed_bytes(C.byref(d))
        packed = torch.empty(nb, dtype=torch.uint8, device="cuda")
        _lib.check(lib.hcu_conv_tc_pack(C.byref(d), P(wg), P(packed), stream()), "pack")
        _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), P(xin), P(packed), P(b), P(isc), P(ish), P(osc), P(osh), P(y),
                                       P(stats), stream()), "conv_tc_fwd")
    torch.cuda.synchronize()
    return from_cl(y), (stats.sum(0).cpu() if stats is not None else None)   # binned reduction buffers


def h16(t):
    return t.half().float()


CASES = [
    # (N, Cin, Cout, in size, kernel, dilation, cpitch)
    (1, 8, 8, (10, 12, 9), (3, 3, 2), (1, 1, 1), None),
    (2, 8, 8, (9, 11, 7), (3, 3, 1), (1, 1, 1), None),
    (1, 4, 8, (12, 10, 6), (3, 3, 2), (1, 1, 1), 8),       # first layer: 4 channels in a pitch of 8
    (1, 16, 16, (8, 20, 31), (3, 3, 1), (1, 1, 1), None),
    (2, 16, 32, (7, 9, 6), (3, 3, 2), (1, 1, 1), None),
    (1, 32, 32, (6, 30, 29), (3, 3, 1), (1, 1, 1), None),
    (1, 64, 64, (5, 12, 28), (3, 3, 2), (1, 1, 1), None),
    (1, 64, 128, (5, 12, 10), (3, 3, 2), (1, 1, 1), None),
    (1, 128, 128, (4, 10, 9), (3, 3, 1), (1, 1, 1), None),
    (1, 8, 16, (14, 13, 6), (3, 3, 2), (2, 2, 1), None),   # dilation (g3d_dil)
    (1, 16, 8, (40, 70, 31), (3, 3, 2), (1, 1, 1), None),  # several runs, several x segments
    (1, 8, 24, (6, 9, 8), (1, 1, 1), (1, 1, 1), None),     # 1x1x1, cout not a multiple of 16
]


@pytest.mark.parametrize("case", CASES)
def test_conv_tc_matches_fp32_conv(case):
    n, cin, cout, isz, k, dil, cp = case
    g = torch.Generator().manual_seed(hash(case) % 10000)
    x = h16(torch.randn((n, cin) + isz, generator=g))
    w = h16(torch.randn((cout, cin) + k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = F.conv3d(x, w, b, dilation=dil)
    y, stats = run_tc(x, w, bias=b, dil=dil, cpitch=cp, want_stats=True)
    assert not torch.isnan(y).any(), "some outputs were never written"
    assert rel_l2(y, ref) <= 1e-3, rel_l2(y, ref)   # fp16 output rounding only
    npix = ref.numel() / cout
    mean = stats[0] / npix
    var = stats[1] / npix - mean * mean
    assert torch.allclose(mean.float(), ref.mean(dim=(0, 2, 3, 4)), atol=2e-4, rtol=1e-4)
    assert torch.allclose(var.float(), ref.var(dim=(0, 2, 3, 4), unbiased=False), atol=2e-4, rtol=1e-3)
    # fp32 output: accumulation-order differences only
    y32, _ = run_tc(x, w, bias=b, dil=dil, cpitch=cp, out_f32=True)
    assert rel_l2(y32, ref) <= 2e-6, rel_l2(y32, ref)


def test_conv_tc_fused_input_bn_relu_and_output_affine():
    g = torch.Generator().manual_seed(3)
    x = h16(torch.randn((2, 16, 9, 10, 7), generator=g))
    w = h16(torch.randn((32, 16, 3, 3, 2), generator=g) / 17.0)
    sc, sh = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))  # the producer rounds A to fp16
    osc, osh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g)
    ref = F.relu(F.conv3d(a, w) * osc.view(1, -1, 1, 1, 1) + osh.view(1, -1, 1, 1, 1))
    y, _ = run_tc(x, w, in_affine=(sc, sh), in_relu=True, out_affine=(osc, osh), out_relu=True)
    assert rel_l2(y, ref) <= 1e-3, rel_l2(y, ref)


def test_conv_tc_zero_padding_is_dgrad():
    """pad = (k-1)*dil with a flipped kernel is the data gradient of a valid convolution."""
    g = torch.Generator().manual_seed(4)
    dy = h16(torch.randn((1, 16, 7, 9, 6), generator=g))
    w = h16(torch.randn((16, 8, 3, 3, 2), generator=g) / 12.0)  # forward weight [Cout=16, Cin=8]
    ref = torch.nn.grad.conv3d_input((1, 8, 9, 11, 7), w, dy)
    wd = w.flip(2, 3, 4).permute(1, 0, 2, 3, 4).contiguous()   # [Cin, Cout] flipped
    y, _ = run_tc(dy, wd, pad=(2, 2, 1))
    assert tuple(y.shape) == (1, 8, 9, 11, 7)
    assert rel_l2(y, ref) <= 1e-3, rel_l2(y, ref)


def test_conv_tc_agrees_with_simt_kernel():
    g = torch.Generator().manual_seed(5)
    x = h16(torch.randn((1, 32, 6, 14, 12), generator=g))
    w = h16(torch.randn((32, 32, 3, 3, 1), generator=g) / 17.0)
    a, sa = run_tc(x, w, want_stats=True)
    b, sb = run_tc(x, w, want_stats=True, use_simt=True)
    assert rel_l2(a, b) <= 6e-4
    assert torch.allclose(sa, sb, rtol=1e-4, atol=1e-3)


# ---- K-streamed tcgen05 kernel (conv_ks_kernel: >= 64 input channels, weights and activations streamed) ----------
KS_CASES = [
    # (N, Cin, Cout, in size, kernel, dilation, pad)
    (1, 64, 64, (5, 12, 28), (3, 3, 2), (1, 1, 1), (0, 0, 0)),      # 3D: x march = separate planes, (y, z) flat
    (2, 128, 64, (4, 10, 9), (3, 3, 1), (1, 1, 1), (0, 0, 0)),      # two images stacked in one flat run
    (3, 64, 128, (30, 30, 1), (3, 3, 1), (1, 1, 1), (0, 0, 0)),     # 2D: the whole image is the flat plane
    (16, 256, 64, (12, 12, 1), (3, 3, 1), (1, 1, 1), (0, 0, 0)),    # 2D bottom level: 16 tiny images in 5 runs
    (2, 512, 32, (14, 14, 1), (3, 3, 1), (1, 1, 1), (0, 0, 0)),     # 64 channel planes, 16 chunks
    (1, 64, 272, (9, 20, 1), (3, 3, 1), (1, 1, 1), (0, 0, 0)),      # cout not a multiple of the column chunk: N-split + tail
    (2, 64, 64, (40, 70, 1), (3, 3, 1), (1, 1, 1), (0, 0, 0)),      # several runs per image
    (1, 64, 32, (7, 9, 6), (3, 3, 2), (1, 1, 1), (2, 2, 1)),        # zero padding (the data-gradient form), 3D
    (2, 96, 64, (11, 13, 1), (3, 3, 1), (1, 1, 1), (2, 2, 0)),      # zero padding, 2D, 12 channel planes
    (1, 64, 64, (14, 13, 6), (3, 3, 2), (2, 2, 1), (0, 0, 0)),      # dilation
    (2, 128, 256, (6, 6, 1), (1, 1, 1), (1, 1, 1), (0, 0, 0)),      # 1x1 (the transposed convs' GEMM)
]


def ks_tile(mb, nc, pc):
    return mb | (nc << 8) | (pc << 20)


# the tile search picks small tiles for small problems: force the big ones the full-size layers use as well
KS_TILES = [0, ks_tile(4, 64, 4), ks_tile(2, 32, 8), ks_tile(1, 64, 4), ks_tile(4, 128, 8), ks_tile(2, 256, 4)]


@pytest.mark.parametrize("tile", KS_TILES)
@pytest.mark.parametrize("case", KS_CASES)
def test_conv_ks_matches_fp32_conv(case, tile):
    n, cin, cout, isz, k, dil, pad = case
    if tile and cout % ((tile >> 8) & 0xfff):
        pytest.skip("column chunk does not divide cout")
    g = torch.Generator().manual_seed(hash(case) % 10000)
    x = h16(torch.randn((n, cin) + isz, generator=g))
    w = h16(torch.randn((cout, cin) + k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = F.conv3d(x, w, b, dilation=dil, padding=pad)
    y, stats = run_tc(x, w, bias=b, dil=dil, pad=pad, want_stats=True, hint=2, tile=tile)
    assert not torch.isnan(y).any(), "some outputs were never written"
    assert rel_l2(y, ref) <= 1e-3, rel_l2(y, ref)   # fp16 output rounding only
    npix = ref.numel() / cout
    mean = stats[0] / npix
    var = stats[1] / npix - mean * mean
    assert torch.allclose(mean.float(), ref.mean(dim=(0, 2, 3, 4)), atol=2e-4, rtol=1e-4)
    assert torch.allclose(var.float(), ref.var(dim=(0, 2, 3, 4), unbiased=False), atol=2e-4, rtol=1e-3)


def test_conv_ks_fused_input_bn_relu_and_output_affine_and_classic_agreement():
    g = torch.Generator().manual_seed(13)
    x = h16(torch.randn((2, 64, 9, 10, 7), generator=g))
    w = h16(torch.randn((64, 64, 3, 3, 2), generator=g) / 34.0)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))  # the producer rounds A to fp16
    osc, osh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    ref = F.relu(F.conv3d(a, w) * osc.view(1, -1, 1, 1, 1) + osh.view(1, -1, 1, 1, 1))
    y, _ = run_tc(x, w, in_affine=(sc, sh), in_relu=True, out_affine=(osc, osh), out_relu=True, hint=2)
    assert rel_l2(y, ref) <= 1e-3, rel_l2(y, ref)
    # zero padding is applied AFTER the input transform (padding pixels stay zero), and both kernels agree
    yk, sk = run_tc(x, w, in_affine=(sc, sh), in_relu=True, pad=(2, 2, 1), want_stats=True, hint=2)
    yc, sc_ = run_tc(x, w, in_affine=(sc, sh), in_relu=True, pad=(2, 2, 1), want_stats=True, hint=1)
    refp = F.conv3d(a, w, padding=(2, 2, 1))
    assert rel_l2(yk, refp) <= 1e-3, rel_l2(yk, refp)
    assert rel_l2(yk, yc) <= 6e-4
    assert torch.allclose(sk, sc_, rtol=1e-4, atol=1e-3)


# ---- weight gradient on tensor cores (wgrad_mma.cu) ---------------------------------------------------------
def run_wgrad(x, dy, k, *, dil=(1, 1, 1), cpitch=None, in_affine=None, use_simt=False, use_tc5=False, use_ws=False,
              use_rows=False):
    """x [N,Cin,...] activations (pre-transform), dy [N,Cout,...] -> dW [Cout,Cin,kx,ky,kz] fp32."""
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    n, cin = x.shape[:2]
    cout = dy.shape[1]
    isz, osz = tuple(x.shape[2:]), tuple(dy.shape[2:])
    cp = cpitch or cin
    xin, dyin = to_cl(x, cp), to_cl(dy)
    d = conv_desc(_lib.F16, _lib.F16, n, isz, cp, 0, cin, cin, osz, osz, cout, 0, cout, 1, k, dil,
                  in_relu=int(in_affine is not None))
    T = k[0] * k[1] * k[2]
    isc = ish = None
    if in_affine is not None:
        isc = torch.zeros(cp, device="cuda"); ish = torch.zeros(cp, device="cuda")
        isc[:cin], ish[:cin] = in_affine[0].cuda(), in_affine[1].cuda()
    if use_simt:
        ns = 4
        part = torch.empty((ns, T * cin * cout), device="cuda")
        _lib.check(lib.hcu_conv_wgrad_partial(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(part), ns, stream()))
        wacc = part.sum(0)
    elif use_ws:
        assert lib.hcu_conv_wgrad_ws_supported(C.byref(d)) == 1
        wacc = torch.zeros((T * cin * cout,), device="cuda")
        _lib.check(lib.hcu_conv_wgrad_ws_acc(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(wacc), stream()), "wgrad_ws")
    elif use_rows:
        assert lib.hcu_conv_wgrad_rows_supported(C.byref(d)) == 1
        wacc = torch.zeros((T * cin * cout,), device="cuda")
        _lib.check(lib.hcu_conv_wgrad_rows_acc(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(wacc), stream()), "wgrad_rows")
    elif use_tc5:
        assert lib.hcu_conv_wgrad_tc5_supported(C.byref(d)) == 1
        wacc = torch.zeros((T * cin * cout,), device="cuda")   # the tcgen05 kernel accumulates into a zeroed buffer
        _lib.check(lib.hcu_conv_wgrad_tc5_acc(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(wacc), stream()), "wgrad_tc5")
    else:
        assert lib.hcu_conv_wgrad_tc_supported(C.byref(d)) == 1
        wacc = torch.full((T * cin * cout,), float("nan"), device="cuda")
        _lib.check(lib.hcu_conv_wgrad_tc(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(wacc), stream()), "wgrad_tc")
    torch.cuda.synchronize()
    return wacc.view(k[0], k[1], k[2], cin, cout).permute(4, 3, 0, 1, 2).cpu()


WG_CASES = [
    (1, 8, 8, (10, 12, 9), (3, 3, 2), (1, 1, 1), None),
    (2, 4, 8, (9, 11, 7), (3, 3, 2), (1, 1, 1), 8),
    (1, 8, 8, (9, 40, 31), (3, 3, 1), (1, 1, 1), None),
    (1, 8, 16, (8, 12, 10), (3, 3, 2), (1, 1, 1), None),
    (2, 16, 16, (7, 14, 9), (3, 3, 1), (1, 1, 1), None),
    (1, 16, 32, (7, 9, 8), (3, 3, 2), (1, 1, 1), None),
    (1, 32, 32, (6, 11, 9), (3, 3, 1), (1, 1, 1), None),
    (1, 32, 64, (5, 9, 8), (3, 3, 2), (1, 1, 1), None),
    (1, 64, 64, (5, 8, 7), (3, 3, 1), (1, 1, 1), None),
    (1, 64, 128, (4, 8, 6), (3, 3, 2), (1, 1, 1), None),
    (1, 128, 128, (4, 7, 6), (3, 3, 1), (1, 1, 1), None),
    (1, 8, 16, (12, 13, 6), (3, 3, 2), (2, 2, 1), None),
    (1, 16, 8, (5, 6, 7), (1, 1, 1), (1, 1, 1), None),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_wgrad_tc_matches_fp32(case):
    n, cin, cout, isz, k, dil, cp = case
    g = torch.Generator().manual_seed(hash(case) % 10000 + 1)
    osz = tuple(isz[i] - (k[i] - 1) * dil[i] for i in range(3))
    x = h16(torch.randn((n, cin) + isz, generator=g))
    dy = h16(torch.randn((n, cout) + osz, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, (cout, cin) + k, dy, dilation=dil)
    got = run_wgrad(x, dy, k, dil=dil, cpitch=cp)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)   # exact fp16 products, fp32 accumulate


# ---- warp-specialised weight gradient (wgrad_ws.cu, 8/16-channel levels) ------------------------------------
WGS_CASES = [c for c in WG_CASES if (c[6] or c[1]) <= 32 and c[2] <= 8] + [
    (2, 8, 8, (40, 60, 33), (3, 3, 2), (1, 1, 1), None),       # several runs, several x segments, ragged tail
    (1, 16, 8, (20, 45, 31), (3, 3, 1), (1, 1, 1), None),
    (1, 32, 8, (9, 30, 20), (3, 3, 2), (1, 1, 1), None),
]


@pytest.mark.parametrize("case", WGS_CASES)
def test_wgrad_ws_matches_fp32(case):
    n, cin, cout, isz, k, dil, cp = case
    g = torch.Generator().manual_seed(hash(case) % 10000 + 5)
    osz = tuple(isz[i] - (k[i] - 1) * dil[i] for i in range(3))
    x = h16(torch.randn((n, cin) + isz, generator=g))
    dy = h16(torch.randn((n, cout) + osz, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, (cout, cin) + k, dy, dilation=dil)
    got = run_wgrad(x, dy, k, dil=dil, cpitch=cp, use_ws=True)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)   # exact fp16 products, fp32 accumulate


def test_wgrad_ws_fused_input_bn_relu():
    g = torch.Generator().manual_seed(29)
    x = h16(torch.randn((2, 8, 12, 19, 17), generator=g))
    dy = h16(torch.randn((2, 8, 10, 17, 16), generator=g))
    sc, sh = torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))
    ref = torch.nn.grad.conv3d_weight(a, (8, 8, 3, 3, 2), dy)
    got = run_wgrad(x, dy, (3, 3, 2), in_affine=(sc, sh), use_ws=True)
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)


# ---- row-stacked tcgen05 weight gradient fed by TMA (wgrad_rows.cu, channel-poor levels) ---------------------
WGR_CASES = [
    (1, 8, 8, (10, 12, 9), (3, 3, 2), (1, 1, 1), None),        # one tile, one K group (8 z positions of 16)
    (2, 4, 8, (9, 11, 17), (3, 3, 2), (1, 1, 1), 8),           # 4 of 8 input channels live (the first layer)
    (1, 8, 8, (9, 40, 31), (3, 3, 1), (1, 1, 1), None),        # several row tiles, ragged last tile, z = 31 of 32
    (2, 8, 8, (40, 60, 33), (3, 3, 2), (1, 1, 1), None),       # two K groups, merged rows shorter than the z reach
    (2, 8, 8, (12, 37, 50), (3, 3, 2), (1, 1, 1), None),       # four K groups: 16-byte boxes
    (1, 8, 16, (8, 22, 18), (3, 3, 2), (1, 1, 1), None),       # two dy planes: one per CTA kind
    (2, 16, 16, (7, 24, 19), (3, 3, 1), (1, 1, 1), None),      # two input planes per CTA
    (1, 16, 8, (20, 45, 31), (3, 3, 1), (1, 1, 1), None),
    (1, 16, 32, (7, 19, 18), (3, 3, 2), (1, 1, 1), None),      # four dy planes
    (1, 32, 32, (6, 21, 19), (3, 3, 1), (1, 1, 1), None),      # four input planes
    (1, 8, 16, (12, 23, 16), (3, 3, 2), (2, 2, 1), None),      # dilation in x and y
    (1, 8, 8, (9, 20, 24), (3, 3, 2), (1, 1, 2), None),        # dilation in z
    (1, 16, 8, (5, 16, 17), (1, 1, 1), (1, 1, 1), None),       # 1x1: 16 rows per tile
    (4, 8, 8, (30, 33, 20), (3, 3, 1), (1, 1, 1), None),       # CTAs whose step range crosses tiles and images
    # interleaved channel planes (P > 1): one MMA covers all planes of a row on both sides
    (2, 16, 16, (9, 41, 30), (3, 3, 1), (1, 1, 1), None),      # 8 rows x 2 planes, ragged last row tile
    (2, 16, 32, (8, 30, 30), (3, 3, 2), (1, 1, 1), None),      # dy planes split over CTA kinds by TMEM capacity
    (2, 32, 32, (9, 31, 29), (3, 3, 1), (1, 1, 1), None),      # 4 rows x 4 planes
    (1, 32, 64, (7, 15, 29), (3, 3, 2), (1, 1, 1), None),      # 64 dy channels
    (1, 32, 16, (6, 13, 40), (3, 3, 2), (1, 1, 1), None),      # three K groups
    (1, 24, 16, (6, 13, 20), (3, 3, 2), (1, 1, 1), 32),        # 24 of 32 input channels live
]


@pytest.mark.parametrize("case", WGR_CASES)
def test_wgrad_rows_matches_fp32(case):
    n, cin, cout, isz, k, dil, cp = case
    g = torch.Generator().manual_seed(hash(case) % 10000 + 7)
    osz = tuple(isz[i] - (k[i] - 1) * dil[i] for i in range(3))
    x = h16(torch.randn((n, cin) + isz, generator=g))
    dy = h16(torch.randn((n, cout) + osz, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, (cout, cin) + k, dy, dilation=dil)
    got = run_wgrad(x, dy, k, dil=dil, cpitch=cp, use_rows=True)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)   # exact fp16 products, fp32 accumulate


def test_wgrad_rows_fused_input_bn_relu():
    g = torch.Generator().manual_seed(31)
    for cin, cout, isz, k in ((8, 8, (12, 19, 17), (3, 3, 2)), (16, 16, (9, 30, 20), (3, 3, 1)), (32, 32, (7, 22, 29), (3, 3, 1))):
        osz = tuple(isz[i] - k[i] + 1 for i in range(3))
        x = h16(torch.randn((2, cin) + isz, generator=g))
        dy = h16(torch.randn((2, cout) + osz, generator=g))
        sc, sh = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
        a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))
        ref = torch.nn.grad.conv3d_weight(a, (cout, cin) + k, dy)
        got = run_wgrad(x, dy, k, in_affine=(sc, sh), use_rows=True)
        assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)


@pytest.mark.parametrize("case", [(2, 4, 8, (9, 27, 20), (3, 3, 2), 8), (1, 8, 8, (8, 33, 31), (3, 3, 1), None),
                                  (1, 16, 16, (7, 20, 18), (3, 3, 2), None)])
def test_wgrad_rows_with_fused_bn_backward_apply_equals_the_two_launches(case):
    """dy = c1 * (bn(y) > 0 ? g : 0) + c2 * y + c3 computed on the staged tiles (hcu_conv_wgrad_rows_bnb_acc) against
    hcu_bn_bwd_apply followed by hcu_conv_wgrad_rows_acc: the same fp16 dy, the same MMAs -> the same bits."""
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    n, cin, cout, isz, k, cp = case
    gen = torch.Generator().manual_seed(41)
    osz = tuple(isz[i] - k[i] + 1 for i in range(3))
    x = to_cl(h16(torch.randn((n, cin) + isz, generator=gen)), cp or cin)
    g = to_cl(h16(torch.randn((n, cout) + osz, generator=gen)))
    y = to_cl(h16(torch.randn((n, cout) + osz, generator=gen)))
    sc = (torch.rand(cout, generator=gen) + 0.5).cuda(); sh = (torch.randn(cout, generator=gen) * 0.3).cuda()
    coef = torch.randn(3, cout, generator=gen).cuda().contiguous()
    npix = n * osz[0] * osz[1] * osz[2]
    d = conv_desc(_lib.F16, _lib.F16, n, isz, cp or cin, 0, cin, cin, osz, osz, cout, 0, cout, 1, k, (1, 1, 1))
    T = k[0] * k[1] * k[2]
    dy = torch.empty_like(g)
    _lib.check(lib.hcu_bn_bwd_apply(P(g), _lib.F16, P(y), _lib.F16, P(dy), _lib.F16, npix, cout, P(sc), P(sh), 1, P(coef), None, None,
                                    stream()), "bn_bwd_apply")
    two = torch.zeros(T * cin * cout, device="cuda")
    _lib.check(lib.hcu_conv_wgrad_rows_acc(C.byref(d), P(x), None, None, P(dy), P(two), stream()), "wgrad_rows")
    one = torch.zeros(T * cin * cout, device="cuda")
    _lib.check(lib.hcu_conv_wgrad_rows_bnb_acc(C.byref(d), P(x), None, None, P(g), P(y), P(sc), P(sh), P(coef), P(one), stream()),
               "wgrad_rows_bnb")
    torch.cuda.synchronize()
    assert torch.isfinite(one).all()
    assert rel_l2(one, two) <= 1e-6, rel_l2(one, two)   # same products; the fp32 atomics land in another order


def test_wgrad_rows_refuses_what_it_cannot_take():
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    ok = conv_desc(_lib.F16, _lib.F16, 1, (9, 12, 20), 8, 0, 8, 8, (7, 10, 19), (7, 10, 19), 8, 0, 8, 1, (3, 3, 2), (1, 1, 1))
    assert lib.hcu_conv_wgrad_rows_supported(C.byref(ok)) == 1
    flat = conv_desc(_lib.F16, _lib.F16, 1, (30, 30, 1), 8, 0, 8, 8, (28, 28, 1), (28, 28, 1), 8, 0, 8, 1, (3, 3, 1), (1, 1, 1))
    assert lib.hcu_conv_wgrad_rows_supported(C.byref(flat)) == 0      # 2D: no z rows to put on K
    wide = conv_desc(_lib.F16, _lib.F16, 1, (9, 12, 20), 64, 0, 64, 64, (7, 10, 19), (7, 10, 19), 64, 0, 64, 1, (3, 3, 2), (1, 1, 1))
    assert lib.hcu_conv_wgrad_rows_supported(C.byref(wide)) == 0      # channel-rich: wgrad_tc5.cu


# ---- weight gradient on tcgen05 (wgrad_tc5.cu, channel-rich levels) -----------------------------------------
WG5_CASES = [c for c in WG_CASES if c[1] >= 16 or c[2] >= 32] + [
    (2, 32, 32, (9, 20, 19), (3, 3, 1), (1, 1, 1), None),      # several runs per plane, several x segments
    (4, 64, 128, (12, 12, 28), (3, 3, 2), (1, 1, 1), None),    # d4.conv1 of the bench (tap groups of 4)
    (2, 128, 64, (6, 16, 9), (3, 3, 2), (1, 1, 1), None),      # Up.conv1 shape class
    (1, 32, 32, (11, 12, 10), (3, 3, 2), (2, 2, 1), None),     # dilation
    # channel blocks over CTAs (> 128 input / output channels) and the 2D "whole image = one flat plane" reading
    (3, 64, 64, (30, 28, 1), (3, 3, 1), (1, 1, 1), None),      # 2D, short rows: flat plane
    (2, 256, 128, (14, 14, 1), (3, 3, 1), (1, 1, 1), None),    # two input-channel blocks
    (2, 128, 256, (12, 13, 1), (3, 3, 1), (1, 1, 1), None),    # two output-channel blocks
    (1, 512, 384, (9, 10, 1), (3, 3, 1), (1, 1, 1), None),     # 4 x 3 blocks
    (1, 256, 256, (4, 9, 8), (3, 3, 2), (1, 1, 1), None),      # 3D with blocks
    (2, 256, 256, (6, 6, 1), (1, 1, 1), (1, 1, 1), None),      # 1x1
    (1, 32, 32, (9, 200, 1), (3, 3, 1), (1, 1, 1), None),      # 2D, long rows: row-by-row planes as before
]


@pytest.mark.parametrize("case", WG5_CASES)
def test_wgrad_tc5_matches_fp32(case):
    n, cin, cout, isz, k, dil, cp = case
    g = torch.Generator().manual_seed(hash(case) % 10000 + 3)
    osz = tuple(isz[i] - (k[i] - 1) * dil[i] for i in range(3))
    x = h16(torch.randn((n, cin) + isz, generator=g))
    dy = h16(torch.randn((n, cout) + osz, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, (cout, cin) + k, dy, dilation=dil)
    got = run_wgrad(x, dy, k, dil=dil, cpitch=cp, use_tc5=True)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)   # exact fp16 products, fp32 accumulate


def test_wgrad_tc5_fused_input_bn_relu():
    g = torch.Generator().manual_seed(19)
    x = h16(torch.randn((2, 32, 8, 9, 7), generator=g))
    dy = h16(torch.randn((2, 64, 6, 7, 6), generator=g))
    sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))
    ref = torch.nn.grad.conv3d_weight(a, (64, 32, 3, 3, 2), dy)
    got = run_wgrad(x, dy, (3, 3, 2), in_affine=(sc, sh), use_tc5=True)
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)
    # input-channel blocks take their own slice of the scale / shift vectors
    x = h16(torch.randn((2, 256, 9, 8, 1), generator=g))
    dy = h16(torch.randn((2, 64, 7, 6, 1), generator=g))
    sc, sh = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))
    ref = torch.nn.grad.conv3d_weight(a, (64, 256, 3, 3, 1), dy)
    got = run_wgrad(x, dy, (3, 3, 1), in_affine=(sc, sh), use_tc5=True)
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)


def test_wgrad_tc_fused_input_bn_relu():
    g = torch.Generator().manual_seed(9)
    x = h16(torch.randn((2, 16, 8, 9, 7), generator=g))
    dy = h16(torch.randn((2, 32, 6, 7, 6), generator=g))
    sc, sh = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g) * 0.3
    a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)))
    ref = torch.nn.grad.conv3d_weight(a, (32, 16, 3, 3, 2), dy)
    got = run_wgrad(x, dy, (3, 3, 2), in_affine=(sc, sh))
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)
    simt = run_wgrad(x, dy, (3, 3, 2), in_affine=(sc, sh), use_simt=True)
    assert rel_l2(simt, ref) <= 1e-3   # the FFMA kernel does not round the transformed activation to fp16


@pytest.mark.parametrize("shape", [(2, 8, 9, 7, 5), (1, 16, 12, 10, 6), (1, 32, 5, 8, 3)])
def test_fp16_vector_maxpool_matches_aten(shape):
    """The 16-byte fp16 pool kernels: fused BN+ReLU, first-maximum tie rule, remainder rows get zero gradient."""
    from hcunet_b200 import _lib

    lib = _lib.load()
    n, c, ix, iy, iz = shape
    g = torch.Generator().manual_seed(sum(shape))
    # values chosen so that x * sc + sh is exact in fp32 and fp16 (fma == mul + add): the comparison is bit-exact
    x = h16(torch.randint(-3, 4, shape, generator=g).float() * 0.5)
    sc = torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (c,), generator=g)]
    sh = torch.randint(-2, 3, (c,), generator=g).float() * 0.25
    a = F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    ref, _ = F.max_pool3d(a.requires_grad_(True), (2, 2, 1), return_indices=True)
    ox, oy, oz = ix // 2, iy // 2, iz
    cl = to_cl(x)
    pooled = torch.empty((n, ox, oy, oz, c), dtype=torch.float16, device="cuda")
    arg = torch.empty((n, ox, oy, oz, c), dtype=torch.uint8, device="cuda")
    scd, shd = sc.cuda(), sh.cuda()   # keep the device tensors alive while the kernel uses their pointers
    _lib.check(lib.hcu_bn_relu_maxpool(P(cl), _lib.F16, P(pooled), _lib.F16, P(arg), n, ix, iy, iz, c, 2, 2, 1,
                                       P(scd), P(shd), 1, stream()))
    torch.cuda.synchronize()
    assert torch.equal(from_cl(pooled), h16(ref.detach()))
    go = h16(torch.randn(ref.shape, generator=g))
    ref.backward(go)
    dfull = torch.full((n, ix, iy, iz, c), float("nan"), dtype=torch.float16, device="cuda")
    gcl = to_cl(go)
    _lib.check(lib.hcu_maxpool_bwd(P(gcl), _lib.F16, P(arg), P(dfull), _lib.F16, n, ix, iy, iz, c, 2, 2, 1,
                                   stream()))
    torch.cuda.synchronize()
    want = a.grad
    assert torch.equal(from_cl(dfull), want)


def test_input_layout_kernel_pads_channels():
    from hcunet_b200 import _lib

    lib = _lib.load()
    x = torch.randn(2, 4, 6, 5, 3)
    dst = torch.full((2, 6 * 5 * 3, 8), float("nan"), dtype=torch.float16, device="cuda")
    for src in (x.cuda(), x.half().cuda()):
        _lib.check(lib.hcu_nc_to_cl(P(src), _lib.F32 if src.dtype == torch.float32 else _lib.F16, P(dst), _lib.F16, 2, 4,
                                    90, 8, None, stream()))
        got = dst.view(2, 6, 5, 3, 8).cpu().float()
        assert torch.equal(got[..., :4], x.half().float().permute(0, 2, 3, 4, 1))
        assert torch.equal(got[..., 4:], torch.zeros(2, 6, 5, 3, 4))


@pytest.mark.parametrize("c,cpitch,c_off,npix", [(8, 8, 0, 20000), (16, 16, 0, 9001), (64, 64, 0, 4099), (32, 64, 32, 777),
                                                 (512, 512, 0, 300), (1, 8, 0, 5000), (24, 24, 0, 1234)])
def test_colsum_matches_fp64_sum(c, cpitch, c_off, npix):
    """Bias gradients of the layers without a BatchNorm behind them (`ConvTranspose`, `out_conv`): per-channel sum of a
    channels-last fp16 tensor, 16-byte vector kernel where the channel slice is 8-aligned, generic kernel otherwise."""
    from hcunet_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(c * 131 + npix)
    x = torch.randn((npix, cpitch), generator=g).half().cuda()
    scratch = torch.empty(4096, dtype=torch.float64, device="cuda")
    out = torch.empty(c, dtype=torch.float32, device="cuda")
    dscale = torch.tensor([0.25], dtype=torch.float32, device="cuda")
    _lib.check(lib.hcu_colsum(P(x), _lib.F16, npix, cpitch, c_off, c, 2.0, P(dscale), P(scratch), P(out), stream()), "colsum")
    want = x[:, c_off:c_off + c].double().sum(0) * 0.5
    assert float((out.double() - want).abs().max()) <= 1e-5 * float(x.double().abs().sum(0).max())


@pytest.mark.parametrize("case", [(2, 8, 8, (20, 22, 9), (3, 3, 1)), (1, 8, 8, (40, 36, 15), (3, 3, 2)),
                                  (2, 16, 16, (20, 22, 9), (3, 3, 1)), (2, 16, 8, (20, 22, 9), (3, 3, 2))])
def test_data_gradient_with_fused_bn_backward_statistics(case):
    """`hcu_conv_tc_fwd_bnbwd` == `hcu_conv_tc_fwd` (data gradient) followed by `hcu_bn_bwd_stats_fin`: the same g bit for
    bit, the same sums / coefficients / dgamma / dbeta to fp32 grouping."""
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    B, cdy, cout, osz, k = case
    isz = tuple(osz[i] + k[i] - 1 for i in range(3))
    g0 = torch.Generator(device="cuda").manual_seed(sum(osz) + cdy)
    dy = torch.randn((B,) + osz + (cdy,), device="cuda", generator=g0).half()
    d = conv_desc(_lib.F16, _lib.F16, B, osz, cdy, 0, cdy, cdy, isz, isz, cout, 0, cout, 1, k, pad=tuple(t - 1 for t in k))
    assert lib.hcu_conv_tc_bnbwd_supported(C.byref(d)) == 1
    T = k[0] * k[1] * k[2]
    w = torch.randn(T * cdy * cout, device="cuda", generator=g0) / (T * cdy) ** 0.5
    packed = torch.empty(lib.hcu_conv_tc_packed_bytes(C.byref(d)), dtype=torch.uint8, device="cuda")
    _lib.check(lib.hcu_conv_tc_pack(C.byref(d), P(w), P(packed), stream()))
    npix = B * isz[0] * isz[1] * isz[2]
    y = torch.randn((B,) + isz + (cout,), device="cuda", generator=g0).half()
    scale, shift = torch.rand(cout, device="cuda", generator=g0) + 0.5, torch.randn(cout, device="cuda", generator=g0) * 0.3
    mean, invstd = torch.randn(cout, device="cuda", generator=g0) * 0.1, torch.rand(cout, device="cuda", generator=g0) + 0.5
    gamma = torch.rand(cout, device="cuda", generator=g0) + 0.5
    outs = []
    for fused in (False, True):
        g = torch.full((B,) + isz + (cout,), float("nan"), dtype=torch.float16, device="cuda")
        ws = torch.zeros(2 * cout * _lib.STAT_BINS + 1, dtype=torch.float64, device="cuda")
        coef, dgamma, dbeta = torch.zeros((3, cout), device="cuda"), torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
        fin = _lib.HcuBnBwdFin(float(npix), gamma.data_ptr(), 1, 1.0, None, dgamma.data_ptr(), dbeta.data_ptr(), None,
                               coef.data_ptr(), ws.data_ptr() + 8 * 2 * cout * _lib.STAT_BINS)
        if fused:
            _lib.check(lib.hcu_conv_tc_fwd_bnbwd(C.byref(d), P(dy), P(packed), P(g), P(y), P(scale), P(shift), P(mean), P(invstd),
                                                 P(ws), C.byref(fin), stream()), "conv_tc_fwd_bnbwd")
        else:
            _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), P(dy), P(packed), None, None, None, None, None, P(g), None, stream()))
            _lib.check(lib.hcu_bn_bwd_stats_fin(P(g), _lib.F16, P(y), _lib.F16, npix, cout, P(scale), P(shift), P(mean), P(invstd),
                                                1, None, None, P(ws), C.byref(fin), stream()))
        torch.cuda.synchronize()
        outs.append((g, coef, dgamma, dbeta))
    a, b = outs
    assert not torch.isnan(b[0]).any() and torch.equal(a[0], b[0])
    for i in (1, 2, 3):
        assert float((a[i] - b[i]).abs().max()) <= 2e-6 * float(a[i].abs().max()), i


@pytest.mark.parametrize("shape,c,pool", [((2, 11, 12, 7), 8, (2, 2, 1)), ((1, 8, 9, 6), 16, (2, 2, 2)), ((1, 6, 6, 5), 64, (2, 2, 1)),
                                          ((1, 9, 7, 4), 32, (3, 2, 1))])
def test_pooled_bn_backward_statistics_match_the_scattered_gradient(shape, c, pool):
    """BatchNorm-backward statistics of a pooled layer computed at POOLED resolution (`bn_bwd_stats_pool_h8_kernel`: pooled
    gradient + y gathered at the argmax positions) == the sums over the full-resolution gradient that max-pool backward would
    scatter (torch, float64), including extents that the pooling floors."""
    from hcunet_b200 import _lib

    lib = _lib.load()
    B, X, Y, Z = shape
    px, py, pz = pool
    g0 = torch.Generator().manual_seed(X * 100 + c)
    y = (torch.randn((B, X, Y, Z, c), generator=g0) * 1.5).half().cuda()
    scale = (torch.rand(c, generator=g0) + 0.5).cuda()
    shift = (torch.randn(c, generator=g0) * 0.3).cuda()
    mean = (torch.randn(c, generator=g0) * 0.1).cuda()
    invstd = (torch.rand(c, generator=g0) + 0.5).cuda()
    ox, oy, oz = X // px, Y // py, Z // pz
    pooled = torch.empty((B, ox, oy, oz, c), dtype=torch.float16, device="cuda")
    argmax = torch.empty((B, ox, oy, oz, c), dtype=torch.uint8, device="cuda")
    _lib.check(lib.hcu_bn_relu_maxpool(P(y), _lib.F16, P(pooled), _lib.F16, P(argmax), B, X, Y, Z, c, px, py, pz, P(scale), P(shift), 1,
                                       stream()), "bn_relu_maxpool")
    dpool = torch.randn((B, ox, oy, oz, c), generator=g0).half().cuda()
    npix = B * X * Y * Z
    sums = torch.zeros((_lib.STAT_BINS, 2, c), dtype=torch.float64, device="cuda")
    geom = _lib.HcuPoolGeom(B, X, Y, Z, px, py, pz)
    _lib.check(lib.hcu_bn_bwd_stats(P(dpool), _lib.F16, P(y), _lib.F16, npix, c, P(scale), P(shift), P(mean), P(invstd), 1, P(argmax),
                                    C.byref(geom), P(sums), stream()), "bn_bwd_stats")
    torch.cuda.synchronize()
    got = sums.sum(0).cpu()
    # reference: scatter the pooled gradient to the argmax voxel of every window, then the plain sums
    yc, dp, am = y.double().cpu(), dpool.double().cpu(), argmax.cpu().long()
    full = torch.zeros((B, X, Y, Z, c), dtype=torch.float64)
    wz, wq = am % pz, am // pz
    wy, wx = wq % py, wq // py
    bi, qx, qy, qz, ci = torch.meshgrid(torch.arange(B), torch.arange(ox), torch.arange(oy), torch.arange(oz), torch.arange(c), indexing="ij")
    full[bi, qx * px + wx, qy * py + wy, qz * pz + wz, ci] = dp
    act = yc * scale.double().cpu() + shift.double().cpu()
    g = torch.where(act > 0, full, torch.zeros_like(full))
    want1 = g.sum(dim=(0, 1, 2, 3))
    want2 = (g * (yc - mean.double().cpu()) * invstd.double().cpu()).sum(dim=(0, 1, 2, 3))
    den = max(float(want1.abs().max()), float(want2.abs().max()), 1.0)
    assert float((got[0] - want1).abs().max()) <= 1e-4 * den and float((got[1] - want2).abs().max()) <= 1e-4 * den


# ---- Adam on the flat parameter buffer (optim.cu) ------------------------------------------------------------------
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_flat_adam_matches_torch_adam_and_skips_non_finite_steps(wd):
    """hcu_adam_flat against torch.optim.Adam on the same parameters and gradients (5 steps, odd length: scalar tail), then
    one step with an inf gradient: parameters, moments and the step counter must not move."""
    import hcunet_b200 as H

    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11)).cuda()      # 2 607 parameters, not a multiple of 4
    ref = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11)).cuda()
    ref.load_state_dict(net.state_dict())
    fp = H.FlatParameters(net)
    opt = H.FlatAdam(fp, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    topt = torch.optim.Adam(ref.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(5):
        g = torch.randn(fp.numel, device="cuda", generator=gen)
        fp.flat.grad = g
        off = 0
        for p in ref.parameters():
            p.grad = g[off:off + p.numel()].view_as(p).clone()
            off += p.numel()
        opt.step()
        topt.step()
    assert not opt.skipped() and int(opt.step_count.item()) == 5
    want = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    assert rel_l2(fp.flat.detach(), want) <= 1e-6, rel_l2(fp.flat.detach(), want)
    for p, q in zip(net.parameters(), ref.parameters()):      # the model's parameters are views of the flat buffer
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)
    before = (fp.flat.detach().clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone())
    g = torch.randn(fp.numel, device="cuda", generator=gen)
    g[1234] = float("inf")
    fp.flat.grad = g
    opt.step()
    assert opt.skipped() and int(opt.step_count.item()) == 5
    assert torch.equal(fp.flat.detach(), before[0]) and torch.equal(opt.exp_avg, before[1]) and torch.equal(opt.exp_avg_sq, before[2])
    g[1234] = 0.0
    opt.step()
    assert not opt.skipped() and int(opt.step_count.item()) == 6


def test_wgrad_rows_randomised_geometry_sweep():
    """Seeded sweep over the geometry corners of wgrad_rows.cu: ragged last row tile / K group, step ranges that cross tiles and
    images, merged and 16-byte TMA boxes, one to eight dy planes, plain and interleaved input planes, with and without the fused
    BatchNorm + ReLU of the input.  (compute-sanitizer is closed on this pool: wide coverage against the fp32 op stands in.)"""
    import random

    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    rng = random.Random(20261019)
    g = torch.Generator().manual_seed(77)
    done = 0
    for _ in range(60):
        cin = rng.choice([3, 4, 8, 16, 24, 32])
        cout = rng.choice([1, 8, 16, 32, 64] if cin > 8 else [1, 8, 16, 32])
        k = rng.choice([(3, 3, 2), (3, 3, 1), (1, 1, 1), (2, 2, 2), (3, 1, 2)])
        n = rng.choice([1, 2, 3])
        isz = (rng.randint(k[0], 9), rng.randint(k[1] + 1, 45), rng.randint(k[2] + 7, 70))
        osz = tuple(isz[i] - k[i] + 1 for i in range(3))
        cp = (cin + 7) // 8 * 8
        cop = (cout + 7) // 8 * 8
        d = conv_desc(_lib.F16, _lib.F16, n, isz, cp, 0, cin, cin, osz, osz, cop, 0, cout, 1, k, (1, 1, 1), in_relu=1)
        if not lib.hcu_conv_wgrad_rows_supported(C.byref(d)):
            continue
        x = h16(torch.randn((n, cin) + isz, generator=g))
        dy = h16(torch.randn((n, cout) + osz, generator=g))
        affine = rng.random() < 0.5
        sc, sh = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
        a = h16(F.relu(x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))) if affine else x
        ref = torch.nn.grad.conv3d_weight(a, (cout, cin) + k, dy)
        xin, dyin = to_cl(x, cp), to_cl(dy, cop)
        isc = ish = None
        if affine:
            isc = torch.zeros(cp, device="cuda"); ish = torch.zeros(cp, device="cuda")
            isc[:cin], ish[:cin] = sc.cuda(), sh.cuda()
        T = k[0] * k[1] * k[2]
        wacc = torch.zeros((T * cin * cout,), device="cuda")
        _lib.check(lib.hcu_conv_wgrad_rows_acc(C.byref(d), P(xin), P(isc), P(ish), P(dyin), P(wacc), stream()), "wgrad_rows")
        torch.cuda.synchronize()
        got = wacc.view(k[0], k[1], k[2], cin, cout).permute(4, 3, 0, 1, 2).cpu()
        assert torch.isfinite(got).all(), (cin, cout, k, n, isz)
        assert rel_l2(got, ref) <= 1e-5, (rel_l2(got, ref), cin, cout, k, n, isz, affine)
        done += 1
    assert done >= 30, done


@pytest.mark.parametrize("case", [(2, 16, 8, (9, 21, 17), (2, 2, 2), (2, 2, 1), True),      # up_steps.3.up_conv shape class
                                  (1, 32, 16, (7, 13, 19), (2, 2, 2), (2, 2, 1), True),     # 64 virtual dy channels, two CTA kinds
                                  (2, 16, 8, (6, 14, 12), (2, 2, 2), (2, 2, 2), False),     # eight phases, no taps left
                                  (1, 32, 8, (5, 11, 16), (4, 4, 2), (2, 2, 1), False)])    # (2,2,2) taps per phase, padding in x and y
def test_wgrad_rows_transposed_conv_phases_agree_with_the_mma_kernel(case):
    """Weight gradient of a ConvTranspose with its stride phases folded into the dy channels (HcuConvDesc.ophase) and the low-side
    padding its taps need: the TMA-fed row-stacked kernel (one tensor map per phase, padded positions zeroed after the fused
    BatchNorm + ReLU) against wgrad_mma.cu on the same descriptor."""
    from hcunet_b200 import _lib
    from hcunet_b200.engine import conv_desc

    lib = _lib.load()
    n, cin, cout, isz, k, s, affine = case
    gen = torch.Generator().manual_seed(91)
    nph = s[0] * s[1] * s[2]
    J = tuple(k[i] // s[i] for i in range(3))
    osz = tuple((isz[i] - 1) * s[i] + k[i] for i in range(3))
    Q = tuple(osz[i] // s[i] for i in range(3))
    x = to_cl(h16(torch.randn((n, cin) + isz, generator=gen)))
    dy = to_cl(h16(torch.randn((n, cout) + osz, generator=gen)))
    d = conv_desc(_lib.F16, _lib.F16, n, isz, cin, 0, cin, cin, Q, osz, nph * cout, 0, nph * cout, 1, J,
                  pad=tuple(j - 1 for j in J), ostep=s, in_relu=int(affine))
    d.ophase = s[0] | (s[1] << 8) | (s[2] << 16)
    isc = ish = None
    if affine:
        isc = (torch.rand(cin, generator=gen) + 0.5).cuda(); ish = (torch.randn(cin, generator=gen) * 0.3).cuda()
    assert lib.hcu_conv_wgrad_rows_supported(C.byref(d)) == 1 and lib.hcu_conv_wgrad_tc_supported(C.byref(d)) == 1
    T = J[0] * J[1] * J[2]
    want = torch.zeros(T * cin * nph * cout, device="cuda")
    _lib.check(lib.hcu_conv_wgrad_tc_acc(C.byref(d), P(x), P(isc), P(ish), P(dy), P(want), stream()), "wgrad_tc")
    got = torch.zeros_like(want)
    _lib.check(lib.hcu_conv_wgrad_rows_acc(C.byref(d), P(x), P(isc), P(ish), P(dy), P(got), stream()), "wgrad_rows")
    torch.cuda.synchronize()
    assert torch.isfinite(got).all() and float(want.abs().max()) > 0
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)
