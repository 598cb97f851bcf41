#!/usr/bin/env python
"""Static SASS evidence of the shipped library: per kernel, how often the mnemonics that identify the Blackwell paths occur
(`cuobjdump -sass recommender_b200/lib/librecsys_b200.so`).  tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
cp.async.bulk.tensor -> UTMALDG/UTMASTG, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR, mma.sync -> HMMA, cp.async -> LDGSTS.

    python scripts/sass_static.py > profiles/rN_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "recommender_b200", "lib", "librecsys_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "LDG", "STG", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.[A-Z0-9_.]+)?", line)
        if cur and m:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dem):
        demangle[n] = d
    print("# Static SASS histogram of librecsys_b200.so (sm_100a)\n")
    print("`python scripts/sass_static.py`; counts are SASS lines per kernel (all template instances summed per base name).\n")
    agg = collections.OrderedDict()
    for n, c in kernels.items():
        base = re.sub(r"<.*", "", demangle.get(n, n)).replace("void ", "").strip()
        base = re.sub(r"\(.*", "", base)
        agg.setdefault(base, collections.Counter()).update(c)
    total = collections.Counter()
    for c in agg.values():
        total.update(c)
    cols = [w for w in WATCH if total[w]]
    print("| kernel | SASS lines | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for base, c in sorted(agg.items(), key=lambda kv: -kv[1]["_total"]):
        if not any(c[w] for w in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS")):
            continue
        print(f"| `{base}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    print(f"| **whole library** | {total['_total']} | " + " | ".join(str(total[w]) for w in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
