#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the
built library (no GPU needed):  UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTMALDG / UTMASTG =
cp.async.bulk.tensor (TMA tensor maps), SYNCS = mbarrier, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hcunet_b200", "libhcunet_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "LDGSTS", "FFMA", "total"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for k in KEYS:
                if op.startswith(k) and k != "total":
                    cur[k] += 1
            if op.startswith("UTC") and op.endswith("MMA") and not op.startswith("UTCHMMA"):
                cur["UTCHMMA"] += 1
    names = demangle(list(counts))
    short = {}
    for k, v in names.items():
        v = re.sub(r"\(.*$", "", v)
        v = v.replace("hcu::", "").replace("(anonymous namespace)::", "")
        short[k] = v[:78]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a) -- instruction counts per kernel")
    print("# " + " ".join(f"{k:>8s}" for k in KEYS) + "  kernel")
    tot = collections.Counter()
    rows = []
    for k, c in counts.items():
        rows.append((short[k], c))
        tot.update(c)
    # group the conv_tc_kernel variants
    for name, c in sorted(rows, key=lambda r: r[0]):
        print("  " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS) + "  " + name)
    print("  " + " ".join(f"{tot.get(k, 0):8d}" for k in KEYS) + "  ALL KERNELS")


if __name__ == "__main__":
    main()
