#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path):
    hdr, agg, unit = None, collections.OrderedDict(), "ns"
    for r in csv.reader(open(path, errors="replace")):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = d.get("Metric Unit", unit)
        k = d["Kernel Name"].replace("<unnamed>::", "").split("(")[0]
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches)")
    print()
    print("| kernel | launches | total us | avg us | min us | max us | share |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1]*scale:.1f} | {a[1]/a[0]*scale:.2f} | {a[2]*scale:.2f} | {a[3]*scale:.2f} | {a[1]/total:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
