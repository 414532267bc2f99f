#!/usr/bin/env python
"""Times single kernels at the README model's layer shapes (CUDA events, L2 flushed between launches).

    python tools/kernel_bench.py conv  [layer ...]     # tcgen05 forward conv
    python tools/kernel_bench.py wgrad [layer ...]     # tensor-core weight gradient
    python tools/kernel_bench.py conv d0.conv2 --once  # one launch (for ncu)
    python tools/kernel_bench.py conv c3.d4.conv2 --hint=2            # force the K-streamed kernel (1: classic)
    python tools/kernel_bench.py conv c3.d4.conv2 --hint=2 --tile=4,128,8   # ... and its tile (MB, Nc, PC)
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcunet_b200 import _lib  # noqa: E402
from hcunet_b200.engine import conv_desc  # noqa: E402

B = 4
# name: (cin, cout, in size, kernel)   -- bench shape: batch 4 of 4x256x256x32
LAYERS = {
    "d0.conv1": (8, 8, (256, 256, 32), (3, 3, 2)),
    "d0.conv2": (8, 8, (254, 254, 31), (3, 3, 1)),
    "d1.conv1": (8, 16, (126, 126, 31), (3, 3, 2)),
    "d1.conv2": (16, 16, (124, 124, 30), (3, 3, 1)),
    "d2.conv1": (16, 32, (61, 61, 30), (3, 3, 2)),
    "d2.conv2": (32, 32, (59, 59, 29), (3, 3, 1)),
    "d3.conv1": (32, 64, (28, 28, 29), (3, 3, 2)),
    "d3.conv2": (64, 64, (26, 26, 28), (3, 3, 1)),
    "d4.conv1": (64, 128, (12, 12, 28), (3, 3, 2)),
    "d4.conv2": (128, 128, (10, 10, 27), (3, 3, 1)),
    "u0.conv1": (64, 64, (16, 16, 28), (3, 3, 2)),
    "u1.conv1": (32, 32, (24, 24, 28), (3, 3, 2)),
    "u3.conv1": (8, 8, (72, 72, 28), (3, 3, 2)),
    "x.n48": (8, 48, (256, 256, 32), (3, 3, 2)),
    "x.n128": (8, 128, (256, 256, 32), (3, 3, 2)),
    # classic 2D U-Net, batch 16 of 3 x 572 x 572 (BASELINE config 3): 5th entry = batch
    "c3.d0.conv2": (32, 32, (570, 570, 1), (3, 3, 1), 16),
    "c3.d1.conv1": (32, 64, (284, 284, 1), (3, 3, 1), 16),
    "c3.d1.conv2": (64, 64, (282, 282, 1), (3, 3, 1), 16),
    "c3.d2.conv1": (64, 128, (140, 140, 1), (3, 3, 1), 16),
    "c3.d2.conv2": (128, 128, (138, 138, 1), (3, 3, 1), 16),
    "c3.d3.conv1": (128, 256, (68, 68, 1), (3, 3, 1), 16),
    "c3.d3.conv2": (256, 256, (66, 66, 1), (3, 3, 1), 16),
    "c3.d4.conv1": (256, 512, (32, 32, 1), (3, 3, 1), 16),
    "c3.d4.conv2": (512, 512, (30, 30, 1), (3, 3, 1), 16),
    "c3.d5.conv1": (512, 1024, (14, 14, 1), (3, 3, 1), 16),
    "c3.d5.conv2": (1024, 1024, (12, 12, 1), (3, 3, 1), 16),
    "c3.u0.conv1": (512, 512, (20, 20, 1), (3, 3, 1), 16),
    "c3.u1.conv1": (256, 256, (32, 32, 1), (3, 3, 1), 16),
    "c3.u2.conv1": (128, 128, (56, 56, 1), (3, 3, 1), 16),
    "c3.u3.conv1": (64, 64, (104, 104, 1), (3, 3, 1), 16),
}


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    notransform = "--raw" in sys.argv
    hint, tile = 0, 0
    for a in sys.argv[1:]:
        if a.startswith("--hint="):
            hint = int(a[7:])
        if a.startswith("--tile="):
            mb, nc, pc = (int(v) for v in a[7:].split(","))
            tile = mb | (nc << 8) | (pc << 20)
    kind = args[0] if args else "conv"
    names = args[1:] or list(LAYERS)
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for name in names:
        cin, cout, isz, k = LAYERS[name][:4]
        B = LAYERS[name][4] if len(LAYERS[name]) > 4 else 4
        osz = tuple(isz[i] - k[i] + 1 for i in range(3))
        x = torch.randn((B,) + isz + (cin,), device="cuda").half()
        sc = torch.rand(cin, device="cuda") + 0.5
        sh = torch.randn(cin, device="cuda") * 0.1
        if notransform:
            sc = sh = None
        T = k[0] * k[1] * k[2]
        nin, nout = x.numel(), B * osz[0] * osz[1] * osz[2] * cout
        if kind in ("conv", "dgrad", "eval"):
            if kind == "dgrad":   # data gradient of this layer: dy [osz, cout] -> dx [isz, cin], zero padding k - 1, nothing fused
                pad = tuple(t - 1 for t in k)
                x = torch.randn((B,) + osz + (cout,), device="cuda").half()
                d = conv_desc(_lib.F16, _lib.F16, B, osz, cout, 0, cout, cout, isz, isz, cin, 0, cin, 1, k, pad=pad)
                cin, cout, osz, sc, sh = cout, cin, isz, None, None
                nin, nout = x.numel(), B * osz[0] * osz[1] * osz[2] * cout
            elif kind == "eval":
                sc = sh = None
                d = conv_desc(_lib.F16, _lib.F16, B, isz, cin, 0, cin, cin, osz, osz, cout, 0, cout, 1, k, out_relu=1)
            else:
                d = conv_desc(_lib.F16, _lib.F16, B, isz, cin, 0, cin, cin, osz, osz, cout, 0, cout, 1, k, in_relu=1)
            d.reserved[0], d.reserved[1] = hint, tile
            buf = C.create_string_buffer(256)
            lib.hcu_conv_tc_describe(C.byref(d), buf, 256)
            if buf.value.startswith(b"unsupported"):
                print(f"{kind:5s} {name:12s} {buf.value.decode()}")
                continue
            w = torch.randn(T * cin * cout, device="cuda") / (T * cin) ** 0.5
            packed = torch.empty(lib.hcu_conv_tc_packed_bytes(C.byref(d)), dtype=torch.uint8, device="cuda")
            _lib.check(lib.hcu_conv_tc_pack(C.byref(d), P(w), P(packed), st))
            y = torch.empty((B,) + osz + (cout,), dtype=torch.float16, device="cuda")
            stats = torch.zeros((_lib.STAT_BINS, 2, cout), dtype=torch.float64, device="cuda")
            bias = torch.randn(cout, device="cuda")
            osc, osh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
            if kind == "dgrad":
                fn = lambda: _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), P(x), P(packed), None, None, None, None, None, P(y),
                                                            None, st))
            elif kind == "eval":
                fn = lambda: _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), P(x), P(packed), None, None, None, P(osc), P(osh), P(y),
                                                            None, st))
            else:   # training forward: conv bias + BatchNorm statistics (+ the previous layer's BN + ReLU on load unless --raw)
                fn = lambda: _lib.check(lib.hcu_conv_tc_fwd(C.byref(d), P(x), P(packed), P(bias), P(sc), P(sh), None, None, P(y),
                                                            P(stats), st))
        else:
            dy = torch.randn((B,) + osz + (cout,), device="cuda").half()
            d = conv_desc(_lib.F16, _lib.F16, B, isz, cin, 0, cin, cin, osz, osz, cout, 0, cout, 1, k, in_relu=1)
            wacc = torch.empty(T * cin * cout, device="cuda")
            if kind == "wgradrows":
                wacc.zero_()
                if not lib.hcu_conv_wgrad_rows_supported(C.byref(d)):
                    print(f"{kind:5s} {name:12s} unsupported")
                    continue
                fn = lambda: _lib.check(lib.hcu_conv_wgrad_rows_acc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
            elif kind == "wgradrowsbnb":   # BatchNorm-backward apply fused into the dy staging (first layer: untransformed input)
                wacc.zero_()
                yy = torch.randn_like(dy)
                bsc, bsh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.3
                coef = torch.randn(3, cout, device="cuda")
                fn = lambda: _lib.check(lib.hcu_conv_wgrad_rows_bnb_acc(C.byref(d), P(x), None, None, P(dy), P(yy), P(bsc), P(bsh), P(coef),
                                                                        P(wacc), st))
            elif kind == "wgrad5":
                wacc.zero_()
                fn = lambda: _lib.check(lib.hcu_conv_wgrad_tc5_acc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
            elif kind == "wgrad_auto":   # what the engine picks: warp-specialised kernel when it takes the shape
                wacc.zero_()
                if lib.hcu_conv_wgrad_ws_supported(C.byref(d)):
                    fn = lambda: _lib.check(lib.hcu_conv_wgrad_ws_acc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
                else:
                    fn = lambda: _lib.check(lib.hcu_conv_wgrad_tc_acc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
            elif kind == "wgradws":
                wacc.zero_()
                fn = lambda: _lib.check(lib.hcu_conv_wgrad_ws_acc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
            else:
                fn = lambda: _lib.check(lib.hcu_conv_wgrad_tc(C.byref(d), P(x), P(sc), P(sh), P(dy), P(wacc), st))
        if once:
            fn()
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        byt = (nin + nout) * 2
        fl = 2 * (nout // cout) * T * cin * cout
        cfg = ("  " + buf.value.decode()) if kind in ("conv", "dgrad", "eval") else ""
        print(f"{kind:5s} {name:12s} {ms*1e3:8.1f} us  {byt/ms/1e6:7.1f} GB/s  {fl/ms/1e9:7.2f} TF/s  (min {min(ts)*1e3:.1f} us){cfg}")


if __name__ == "__main__":
    main()
