#!/usr/bin/env python
"""Writes source-annotated SASS listings of the hot kernels in parallel-rdp_b200/librdp_b200.so (cuobjdump -xelf +
nvdisasm -g) under profiles/: one line per instruction, `index file:line  instruction`, plus an opcode histogram.

    python tools/sass_listing.py <tag>      ->  profiles/<tag>_sass_<kernel>.txt
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {  # output name -> substrings the mangled name must contain
    "k_tiles_deferred_zb_opa_tex_modulate_2c": ["k_tiles", "ELb1ELm17456926977ELj2096953ELj987152EEELb1ELb0"],
    "k_front": ["k_frontEN4rdpb"],
    "k_span_bin_native": ["k_span_binILb0"],
    "k_vi_filter": ["k_vi_filter"],
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "sass"
    lib = os.path.join(ROOT, "parallel-rdp_b200", "librdp_b200.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
        cubin = max((os.path.getsize(os.path.join(tmp, f)), f) for f in os.listdir(tmp) if f.endswith(".cubin"))[1]
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    sections, active, cur = {}, None, None
    for l in text.splitlines():
        if l.startswith(".text."):
            active = l[6:].rstrip(":")
            sections[active] = []
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
            sections[active].append((cur, m.group(2).strip()))
    for name, needles in KERNELS.items():
        match = [k for k in sections if all(n in k for n in needles)]
        if not match:
            print("no section for", name)
            continue
        rows = sections[match[0]]
        hist = collections.Counter(re.split(r"[ .]", ins.lstrip("@!P0123456789T ").strip())[0] for _, ins in rows)
        out = os.path.join(ROOT, "profiles", f"{tag}_sass_{name}.txt")
        with open(out, "w") as f:
            f.write(f"# {match[0]}\n# {len(rows)} SASS instructions (sm_100a, nvdisasm -g of librdp_b200.so); opcode histogram:\n")
            f.write("# " + ", ".join(f"{op} {n}" for op, n in hist.most_common(24)) + "\n")
            for i, (loc, ins) in enumerate(rows):
                f.write(f"{i:5d} {(loc[0] + ':' + str(loc[1])) if loc else '?':<28} {ins}\n")
        print(out, len(rows))


if __name__ == "__main__":
    main()
