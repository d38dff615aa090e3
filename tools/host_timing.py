#!/usr/bin/env python
"""Times the host front-end alone (command ring + command processor + renderer pass recording) on the C2 stream,
with the test-only simulator's launches turned into no-ops.  No GPU needed.

    RDP_HOSTSIM_NULL=1 python tools/host_timing.py
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("RDP_HOSTSIM_NULL", "1")
from harness import workloads as wl  # noqa: E402
from sidebyside import HOSTSIM_LIB, build_hostsim  # noqa: E402

rdp = importlib.import_module("parallel-rdp_b200")
build_hostsim()
rng = np.random.RandomState(2)
frames = [wl.c2_frame(rng, 5000) for _ in range(8)]
rdram = np.zeros(wl.RDRAM_SIZE, dtype=np.uint8)
p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, library=HOSTSIM_LIB)
opts = p.default_scanout_options()
for mode in ("stream", "stream+scanout"):
    for it in range(2):
        h0 = p.host_timing()
        t0 = time.perf_counter()
        n = 200
        for i in range(n):
            p.enqueue_command_stream_borrowed(frames[i % 8]) if os.environ.get("BORROWED") else p.enqueue_command_stream(frames[i % 8])
            if mode != "stream":
                p.scanout(opts)
        p.idle()
        dt = time.perf_counter() - t0
    h1 = p.host_timing()
    print(f"{mode}: {n / dt:.0f} frames/s host-only, {1e6 * dt / n:.1f} us/frame, {1e9 * dt / n / 5001:.1f} ns/primitive; command thread us/frame: "
          + ", ".join(f"{k} {1e6 * (h1[k] - h0[k]) / n:.1f}" for k in h1))
p.close()
