#!/usr/bin/env python
"""Per-source-line and per-phase profile of one kernel from an ncu report captured with --import-source on.

    python tools/ncu_line_profile.py report.ncu-rep <kernel-regex> [--phases file.cu:line=name,...] [--top N]

The source page of the report lists, per SASS instruction, the executed warp instructions, thread instructions and
stall samples; `nvdisasm -g` of the same build (extracted from parallel-rdp_b200/librdp_b200.so) gives every SASS
instruction's source line.  Both are in program order, so they are joined by position.  Output: totals per source line,
and per phase (the tile kernel's phases are told apart by line ranges of backend_cuda.cu; an instruction inlined from a
header counts for the phase of the closest preceding backend_cuda.cu line)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_source_rows(path, regex):
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
    return blocks


def disasm_lines(kernel_regex):
    lib = os.path.join(ROOT, "parallel-rdp_b200", "librdp_b200.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
        cubins = sorted((os.path.getsize(os.path.join(tmp, f)), f) for f in os.listdir(tmp) if f.endswith(".cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubins[-1][1])], capture_output=True, text=True).stdout
    out, cur, active = {}, None, None
    for l in text.splitlines():
        if l.startswith(".text."):
            active = l[6:].rstrip(":")
            out[active] = []
            cur = None
            continue
        if active is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out[active].append((cur, m.group(2).strip()))
    return out


def main():
    path, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    phases = []
    if "--phases" in sys.argv:
        for item in sys.argv[sys.argv.index("--phases") + 1].split(","):
            loc, name = item.split("=")
            f, ln = loc.split(":")
            phases.append((f, int(ln), name))
    blocks = ncu_source_rows(path, regex)
    listings = disasm_lines(regex)
    seen = set()
    for b in blocks:
        hdr, data = b["hdr"], b["data"]
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        te = hdr.index("Thread Instructions Executed")
        tot = sum(int(r[ie]) for r in data)
        if (b["name"], tot) in seen or not tot:
            continue
        seen.add((b["name"], tot))
        # the disassembly section with the same number of instructions whose mangled name matches
        cands = [k for k, v in listings.items() if len(v) == len(data)]
        if not cands:
            print(f"## {b['name'][:120]}: no disassembly section with {len(data)} instructions (different build?)")
            continue
        lst = listings[cands[0]]
        samples = sum(int(r[ss]) for r in data)
        print(f"## {b['name'][:140]}\n   {tot} warp instructions, {len(data)} SASS instructions, {samples} stall samples")
        per_line, per_phase = {}, {}
        phase = "prologue"
        for (loc, _), r in zip(lst, data):
            n, t, s = int(r[ie]), int(r[te]), int(r[ss])
            if loc and phases and any(loc[0] == f for f, _, _ in phases):
                for f, ln, name in phases:
                    if loc[0] == f and loc[1] >= ln:
                        phase = name
            a = per_line.setdefault(loc, [0, 0, 0])
            a[0] += n; a[1] += t; a[2] += s
            p = per_phase.setdefault(phase, [0, 0, 0])
            p[0] += n; p[1] += t; p[2] += s
        if phases:
            print("   phase                      warp-instr   share  lanes  samples share")
            for name, (n, t, s) in per_phase.items():
                print(f"   {name:<24} {n:>12} {100.0 * n / tot:6.1f}% {t / max(n, 1):5.1f} {s:>8} {100.0 * s / max(samples, 1):5.1f}%")
        print(f"   top {top} source lines by warp instructions")
        for loc, (n, t, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
            name = f"{loc[0]}:{loc[1]}" if loc else "?"
            print(f"   {name:<28} {n:>12} {100.0 * n / tot:6.2f}% lanes {t / max(n, 1):5.1f} samples {s:>6} {100.0 * s / max(samples, 1):5.2f}%")


if __name__ == "__main__":
    main()
