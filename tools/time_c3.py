#!/usr/bin/env python
"""BASELINE C3 (fill-rate stress: 64 large alpha-blended CI8+TLUT texrects per 640x480 RGBA16 frame) through the public
API on cuda:0: frames/s and shaded Mpix/s.  Diagnostic (the bench line is C2); run on a GPU box."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness import workloads as wl  # noqa: E402

rdp = importlib.import_module("parallel-rdp_b200")


def main(streams=4, frames=200):
    rng = np.random.RandomState(3)
    procs = []
    for s in range(streams):
        rdram = np.zeros(wl.RDRAM_SIZE, np.uint8)
        tex, tlut = wl.c3_assets()
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex.view(np.uint8)
        t8 = np.frombuffer(tlut.tobytes(), np.uint8)
        rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + t8.size] = t8
        p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE)
        p.end_write_rdram()
        procs.append((p, rdram, [wl.c3_frame(rng, 64) for _ in range(8)]))
    opts = procs[0][0].default_scanout_options()
    opts.blend_previous_frame = 1
    opts.upscale_deinterlacing = 0
    vi = [np.zeros(14, np.uint32) for _ in (0, 1)]
    for f in (0, 1):
        for k, v in wl.c3_vi_registers(f).items():
            vi[f][rdp.VI_REG[k]] = v & 0xffffffff
    outs = [np.empty(640 * 576, np.uint32) for _ in procs]

    def run(n):
        for i in range(n):
            for (p, _, fr), out in zip(procs, outs):
                p.enqueue_command_stream(fr[i % len(fr)])
                p.set_vi_registers(vi[i & 1])
                p.scanout_sync(opts, out) if i == n - 1 else p.scanout(opts)
        for p, _, _ in procs:
            p.idle()

    run(5)
    t0 = time.perf_counter()
    run(frames)
    dt = time.perf_counter() - t0
    fps = frames * streams / dt
    print(f"C3: {streams} stream(s), {frames} frames each: {fps:.0f} frames/s, variant {procs[0][0].last_tile_variant()}")
    for p, _, _ in procs:
        p.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
