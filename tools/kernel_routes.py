"""Which kernel and tile configuration every conv of a model gets (host only: no GPU needed).

    python tools/kernel_routes.py cfg3            # 2D classic U-Net, 16 x 3 x 572 x 572
    python tools/kernel_routes.py cfg2            # README 3D model, 4 x 4 x 256 x 256 x 32
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcunet_b200 import _lib  # noqa: E402
from hcunet_b200.engine import ConvGeom, conv_desc, plan_unet  # noqa: E402
from hcunet_b200.unet import Unet_Constructor  # noqa: E402

README_3D = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                 kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
                 upsample_stride=(2, 2, 1), dilation=1, groups=1)


def describe(lib, d):
    buf = C.create_string_buffer(256)
    lib.hcu_conv_tc_describe(C.byref(d), buf, 256)
    return buf.value.decode()


def main(which="cfg3", hint=0):
    lib = _lib.load()
    if which == "cfg3":
        spec, shape = Unet_Constructor().model_specification, (16, 3, 572, 572)
    else:
        spec, shape = Unet_Constructor(**README_3D).model_specification, (4, 4, 256, 256, 32)
    plan = plan_unet(spec, shape)
    F16 = _lib.F16
    for g in plan.steps:
        if not isinstance(g, ConvGeom):
            print(g.name, g.cin, "->", g.cout, g.in_sz, "->", g.out_sz)
            continue
        cpi, cpo = max(8, g.cin_t), max(8, g.cout_t)
        d = conv_desc(F16, F16, plan.batch, g.in_sz, cpi, 0, g.cin_g, min(cpi, max(g.cin_g, 8)) if g.cin_g < 8 else g.cin_g,
                      g.out_sz, g.out_sz, cpo, 0, g.cout_g, g.groups, g.taps, g.dil)
        d.reserved[0] = hint
        pad = tuple((g.taps[i] - 1) * g.dil[i] for i in range(3))
        dd = conv_desc(F16, F16, plan.batch, g.out_sz, cpo, 0, g.cout_g, g.cout_g, g.in_sz, g.in_sz, cpi, 0, g.cin_g, g.groups,
                       g.taps, g.dil, pad=pad)
        dd.reserved[0] = hint
        m = plan.batch * g.out_sz[0] * g.out_sz[1] * g.out_sz[2]
        gf = 2 * m * g.cout_t * g.cin_g * g.taps[0] * g.taps[1] * g.taps[2] / 1e9
        w5 = lib.hcu_conv_wgrad_tc5_supported(C.byref(d))
        ws = lib.hcu_conv_wgrad_ws_supported(C.byref(d))
        wt = lib.hcu_conv_wgrad_tc_supported(C.byref(d))
        wr = int(bool(lib.hcu_conv_wgrad_rows_supported(C.byref(d))) and cpi <= 32 and cpo <= 64)   # the engine's routing limits
        print(f"{g.name:20s} {g.cin_t:4d}->{g.cout_t:4d} {str(g.in_sz):16s} {gf:6.1f} GF | fwd: {describe(lib, d)}")
        print(f"{'':55s} | dgrad: {describe(lib, dd)}")
        print(f"{'':55s} | wgrad: rows={wr} tc5={w5} ws={ws} mma={wt}  -> {'rows' if wr else 'tc5' if (w5 and cpi >= 32) else 'ws' if ws else 'mma' if wt else 'simt'}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cfg3", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
