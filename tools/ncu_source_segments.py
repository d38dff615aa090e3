#!/usr/bin/env python
"""Summarise the source page of an ncu report into runs of SASS instructions with equal execution counts.

    python tools/ncu_source_segments.py report.ncu-rep [kernel-regex] [--sass]
Shows where the dynamic instructions and the stall samples of a kernel go (needs -lineinfo / --import-source).
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    regex = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "."
    show_sass = "--sass" in sys.argv
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
            continue
        if len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    seen = set()
    for b in blocks:
        hdr, data = b["hdr"], b["data"]
        ie, te, ss, src = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
        tot = sum(int(r[ie]) for r in data)
        key = (b["name"], tot)
        if key in seen:
            continue
        seen.add(key)
        tsamples = sum(int(r[ss]) for r in data) or 1
        print(f"## {b['name'][:140]}\n   dynamic warp instructions {tot}, static {len(data)}, stall samples {tsamples}")
        seg = []
        for i, r in enumerate(data):
            n, t = int(r[ie]), int(r[te])
            if seg and abs(seg[-1][2] - n) < 0.02 * max(n, 1) + 1:
                seg[-1][1] = i
                seg[-1][3] += n
                seg[-1][4] += t
                seg[-1][5] += int(r[ss])
            else:
                seg.append([i, i, n, n, t, int(r[ss])])
        for s in seg:
            if s[3] > tot * 0.004 or s[5] > tsamples * 0.01:
                print(f"   sass {s[0]:4d}-{s[1]:4d} ({s[1]-s[0]+1:3d} ins) x{s[2]:8d}  = {100*s[3]/tot:5.1f}% of instr, lanes {s[4]/max(s[3],1):5.1f}/32, {100*s[5]/tsamples:5.1f}% of samples")
                if show_sass:
                    for r in data[s[0]:s[1] + 1]:
                        print("        ", r[src].strip()[:90], "  samples", r[ss])


if __name__ == "__main__":
    main()
