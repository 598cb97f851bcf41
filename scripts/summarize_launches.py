#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

    python scripts/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = row["Kernel Name"][:110]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| us total | launches | share | us/step ({steps} steps) | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | {v[1] / steps:.1f} | `{k}` |")
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")


if __name__ == "__main__":
    main()
