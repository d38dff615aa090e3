#!/usr/bin/env python
"""Builds an experimental variant of the product library next to it: parallel-rdp_b200/variants/librdp_b200_<name>.so
with extra -D flags.  RDP_B200_LIBRARY=<that path> makes bench.py / the binding load it (kernel experiments only).

    python tools/build_variant.py <name> [-DFOO=1 ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(g.PKG, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"librdp_b200_{name}.so")
    cmd = ["nvcc"] + g.NVCC_FLAGS + defs + ["-shared", "-o", out, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(g.PKG, "csrc"),
                                              "-I", os.path.join(g.PKG, "host")] + g._sources() + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(out_dir, f"ptxas_{name}.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        print(log[-3000:])
        raise SystemExit(1)
    for line in log.splitlines():
        if "17456926977ELj2096953ELj987152EEELb1ELb0" in line and "Compiling" in line:
            idx = log.splitlines().index(line)
            print(name, "|", " ".join(l.strip() for l in log.splitlines()[idx + 1: idx + 4]))
    print(out)


if __name__ == "__main__":
    main()
