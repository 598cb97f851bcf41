#!/usr/bin/env python
"""Opcode histogram (executed warp instructions and stall samples) from `ncu --page source --csv`."""
import collections
import csv
import sys

path, per = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
i_src, i_ex, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot, nlines = collections.Counter(), collections.Counter(), 0, 0
for r in rows:
    if len(r) <= i_samp or not r[i_ex].isdigit():
        continue
    t = r[i_src].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    ops[op] += int(r[i_ex])
    samp[op] += int(r[i_samp])
    tot += int(r[i_ex])
    nlines += 1
print(f"total warp instructions {tot} ({tot / per:.1f} per unit), SASS lines {nlines}, stall samples {sum(samp.values())}")
for op, c in ops.most_common(28):
    print(f"{op:12s} {c / per:9.1f} per unit   {100 * c / tot:5.1f}% of instr   {100 * samp[op] / max(1, sum(samp.values())):5.1f}% of samples")
