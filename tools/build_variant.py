#!/usr/bin/env python
"""Build a VARIANT of libhcunet_b200.so for A/B measurements (extra nvcc flags, e.g. -DHCU_MBAR_NOHINT):

    python tools/build_variant.py nohint -DHCU_MBAR_NOHINT     ->  hcunet_b200/libhcunet_b200.nohint.so
    HCUNET_LIB=hcunet_b200/libhcunet_b200.nohint.so python bench.py ...

Objects go to hcunet_b200/csrc/_obj.<tag>/ so the main build's incremental state is untouched."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcunet_b200 import build as B  # noqa: E402


def main():
    tag, flags = sys.argv[1], sys.argv[2:]
    obj = B.OBJ + "." + tag
    os.makedirs(obj, exist_ok=True)
    nvcc = B._nvcc()
    srcs = B.sources()
    objs = [os.path.join(obj, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def one(pair):
        r = subprocess.run([nvcc] + B.NVCC_FLAGS + flags + ["-c", pair[0], "-o", pair[1]], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(one, zip(srcs, objs)))
    lib = os.path.join(B.HERE, f"libhcunet_b200.{tag}.so")
    subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-lcudart", "-lcuda"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
