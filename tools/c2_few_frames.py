#!/usr/bin/env python
"""C2 frames and C3 frames through the public API on cuda:0, every scan-out image and the final RDRAM / hidden RDRAM checked
against the oracle: a small driver for running the hot kernels under compute-sanitizer (tools/gpu_sanitizer.sh) and for
the 100-frame parity run (tools/gpu_r2_final.sh).  usage: c2_few_frames.py [triangles] [c2 frames] [c3 frames]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import workloads as wl  # noqa: E402
from oracle import angrylion  # noqa: E402

rdp = importlib.import_module("parallel-rdp_b200")
VI_REG = rdp.VI_REG


def run(kind, frames_fn, regs_fn, image, n):
    rdram = image.copy()
    p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE)
    o = angrylion.Oracle(wl.RDRAM_SIZE)
    o.rdram[:] = image
    o.hidden[:] = 0
    p.begin_read_hidden_rdram()[:] = 0
    p.end_write_hidden_rdram()
    p.end_write_rdram()
    opts = p.default_scanout_options()
    opts.blend_previous_frame = 1
    opts.upscale_deinterlacing = 0
    ok = True
    for i in range(n):
        words = frames_fn()
        p.enqueue_command_stream(words)
        o.command_stream(words)
        for k, v in regs_fn(i & 1).items():
            p.set_vi_register(VI_REG[k], v)
            o.set_vi_register(VI_REG[k], v)
        pix, w, h = p.scanout_sync(opts)
        ref = o.update_screen()
        ok = ok and ref is not None and ref.shape == pix.shape and np.array_equal(ref & 0xFFFFFF, pix & 0xFFFFFF)
    p.idle()
    ok = ok and np.array_equal(p.begin_read_rdram(), o.rdram) and np.array_equal(p.begin_read_hidden_rdram()[:wl.HIDDEN_SIZE] & 3, o.hidden[:wl.HIDDEN_SIZE] & 3)
    print(f"{kind}: {n} frames, variant {p.last_tile_variant()}, {'bit-exact vs oracle' if ok else 'MISMATCH'}")
    p.close()
    o.close()
    return ok


def main():
    tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rng = np.random.RandomState(2)
    img = np.zeros(wl.RDRAM_SIZE, np.uint8)
    img[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
    ok = run("C2", lambda: wl.c2_frame(rng, tris), wl.c2_vi_registers, img, n)
    img = np.zeros(wl.RDRAM_SIZE, np.uint8)
    tex, tlut = wl.c3_assets()
    img[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
    img[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
    n3 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ok = run("C3", lambda: wl.c3_frame(rng, 64 if n3 > 1 else 6), wl.c3_vi_registers, img, n3) and ok
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
