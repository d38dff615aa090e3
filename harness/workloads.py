"""harness/workloads.py -- the synthetic command streams BASELINE.json's configs describe.

C2 "triangles": 640x480 RGBA32, ~5,000 textured, Gouraud-shaded, Z-buffered triangles per frame, 2-cycle combiner,
   bilinear TMEM sampling, 8 LoadTile'd 32x32 RGBA16 textures, colour + Z clears through fill rectangles, SyncFull
   and a 480i weave scan-out per frame.
C3 "fill-rate": 640x480 RGBA16, 64 large alpha-blended texture rectangles per frame, CI8 + TLUT textures, coverage AA,
   dither, depth writes from primitive depth.
All randomness is numpy RandomState(seed) (MT19937), so streams are reproducible.
"""
from __future__ import annotations

import numpy as np

from harness.rdp_builder import *  # noqa: F401,F403
from harness import rdp_builder as rb

RDRAM_SIZE = 8 << 20
HIDDEN_SIZE = 4 << 20

COLOR_ADDR = 0x100000
Z_ADDR = 0x300000
TEX_ADDR = 0x400000
LIST_ADDR = 0x600000  # where the CPU baseline parks the display list (the oracle reads it from RDRAM)
WIDTH, HEIGHT = 640, 480


def c2_textures(seed: int = 2) -> np.ndarray:
    """8 textures of 32x32 RGBA16 (2 KiB each), random bytes."""
    return np.random.RandomState(seed ^ 0x7E57).randint(0, 256, size=8 * 32 * 32 * 2, dtype=np.uint8)


def c2_frame(rng: np.random.RandomState, triangles: int = 5000) -> np.ndarray:
    """One frame of the C2 stream as packed command words."""
    b = Builder()
    om = b.om

    # Z clear: bind the depth buffer as a 16-bit colour image and fill it (the usual N64 idiom).
    b.set_color_image(FMT_RGBA, SIZE_16, Z_ADDR, WIDTH)
    b.set_scissor(0, 0, WIDTH, HEIGHT)
    om.__init__()
    om.cycle_type = CYCLE_FILL
    b.touch_other_modes()
    b.set_fill_color(0xFFFCFFFC)
    b.fill_rectangle(0, 0, WIDTH, HEIGHT)

    # Colour clear.
    b.set_color_image(FMT_RGBA, SIZE_32, COLOR_ADDR, WIDTH)
    b.set_depth_image(Z_ADDR)
    b.set_fill_color(0)
    b.fill_rectangle(0, 0, WIDTH, HEIGHT)

    # Scene state.
    om.cycle_type = CYCLE_2
    om.perspective = True
    om.sample_quad = True
    om.bilerp0 = om.bilerp1 = True
    om.aa = True
    om.z_compare = om.z_update = True
    om.z_mode = Z_OPAQUE
    om.image_read = True
    om.cvg_mode = CVG_CLAMP
    om.rgb_dither = DITHER_OFF
    om.alpha_dither = ADITHER_OFF
    om.blend = [[BL_1A_PIXEL, BL_1B_PIXEL_A, BL_2A_MEMORY, BL_2B_INV_PIXEL_A]] * 2
    om.force_blend = False
    b.touch_other_modes()
    b.set_combine(((CC.A_TEXEL0, CC.B_ZERO, CC.C_SHADE, CC.D_ZERO), (CC.AA_TEXEL0, CC.AA_ZERO, CC.AC_SHADE, CC.AA_ZERO)),
                  ((CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_COMBINED), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_COMBINED)))

    per_tex = (triangles + 7) // 8
    for k in range(8):
        n = min(per_tex, triangles - k * per_tex)
        if n <= 0:
            break
        b.set_texture_image(TEX_ADDR + k * 2048, FMT_RGBA, SIZE_16, 32)
        b.set_tile(0, FMT_RGBA, SIZE_16, 64, 0, mask_s=5, mask_t=5)
        b.load_tile(0, 0, 0, 32, 32)
        b.set_tile_size(0, 0, 0, 32, 32)

        c = rng.uniform(-1.0, 1.0, (n, 1, 2))
        v = c + rng.uniform(-0.1, 0.1, (n, 3, 2))
        xy = np.stack([(0.5 * v[:, :, 0] + 0.5) * WIDTH, (0.5 * v[:, :, 1] + 0.5) * HEIGHT], axis=2)
        z = rng.uniform(0.0, 0.95, (n, 3))
        w = rng.uniform(1.0, 5.0, (n, 3))
        uv = rng.uniform(-1000.0, 1000.0, (n, 3, 2))
        rgba = rng.uniform(0.0, 1.0, (n, 3, 4))
        b.triangles(setup_triangles(xy, z, 1.0 / w, uv, rgba, tile=0, max_level=0, perspective=True))

    b.sync_full()
    return b.words()


def c2_vi_registers(field: int):
    """480i weave scan-out of the 640x480 RGBA32 frame (SURVEY 8d, C4(i) register set with a 32-bit framebuffer)."""
    control = (rb.VI_CONTROL_TYPE_RGBA8888 | rb.VI_CONTROL_AA_RESAMP_EXTRA | rb.VI_CONTROL_DIVOT | rb.VI_CONTROL_DITHER_FILTER |
               rb.VI_CONTROL_GAMMA | rb.VI_CONTROL_GAMMA_DITHER | rb.VI_CONTROL_SERRATE)
    return {
        "Control": control,
        "Origin": COLOR_ADDR,
        "Width": WIDTH,
        "VSync": 525,
        "HStart": rb.vi_start_register(108, 108 + 640),
        "VStart": rb.vi_start_register(34, 34 + 480),
        "XScale": rb.vi_scale_register(0x400, 0),
        "YScale": rb.vi_scale_register(0x800, 0x400 if field else 0),
        "VCurrentLine": field & 1,
    }


def c3_assets(seed: int = 3):
    rng = np.random.RandomState(seed ^ 0xC1)
    tex = rng.randint(0, 256, size=64 * 32, dtype=np.uint8)        # CI8 64x32
    tlut = rng.randint(0, 256, size=256 * 2, dtype=np.uint8)       # 256 RGBA16 entries
    return tex, tlut


def c3_frame(rng: np.random.RandomState, rects: int = 64) -> np.ndarray:
    """Fill-rate stress: large overlapping alpha-blended texture rectangles, RGBA16 target."""
    b = Builder()
    om = b.om
    om.__init__()
    b.set_color_image(FMT_RGBA, SIZE_16, COLOR_ADDR, WIDTH)
    b.set_depth_image(Z_ADDR)
    b.set_scissor(0, 0, WIDTH, HEIGHT)

    om.cycle_type = CYCLE_1
    om.tlut = True
    om.sample_quad = True
    om.bilerp0 = om.bilerp1 = True
    om.aa = True
    om.cvg_mode = CVG_WRAP
    om.image_read = True
    om.force_blend = True
    om.blend = [[BL_1A_PIXEL, BL_1B_PIXEL_A, BL_2A_MEMORY, BL_2B_INV_PIXEL_A]] * 2
    om.rgb_dither = DITHER_MAGIC
    om.alpha_dither = ADITHER_PATTERN
    om.z_update = True
    om.prim_depth = True
    b.touch_other_modes()
    b.set_combine(((CC.A_TEXEL0, CC.B_ZERO, CC.C_ENV, CC.D_ZERO), (CC.AA_TEXEL0, CC.AA_ZERO, CC.AC_ENV, CC.AA_ZERO)))
    b.set_env_color(255, 255, 255, 160)

    # TLUT into upper TMEM, CI8 texture into lower TMEM.
    b.set_texture_image(TEX_ADDR + 0x1000, FMT_RGBA, SIZE_16, 256)
    b.set_tile(7, FMT_RGBA, SIZE_4, 0, 0x800)
    b.load_tlut(7, 0, 0, 256, 1)
    b.set_texture_image(TEX_ADDR, FMT_CI, SIZE_8, 64)
    b.set_tile(0, FMT_CI, SIZE_8, 64, 0, mask_s=6, mask_t=5)
    b.load_tile(0, 0, 0, 64, 32)
    b.set_tile_size(0, 0, 0, 64, 32)

    for _ in range(rects):
        x, y = int(rng.randint(0, 256)), int(rng.randint(0, 256))
        w, h = int(rng.randint(256, 641)), int(rng.randint(256, 481))
        b.set_prim_depth(int(rng.randint(0, 0x8000)), int(rng.randint(0, 0x10000)))
        b.tex_rect(0, x << 2, y << 2, w << 2, h << 2, int(rng.randint(0, 2048)) & 0xFFFF, int(rng.randint(0, 1024)) & 0xFFFF,
                   1 << 9, 1 << 9)
    b.sync_full()
    return b.words()


def c3_vi_registers(field: int):
    regs = c2_vi_registers(field)
    regs["Control"] = (regs["Control"] & ~3) | rb.VI_CONTROL_TYPE_RGBA5551
    return regs


def frame_stats(words: np.ndarray):
    """(number of primitive commands, command bytes) of a packed stream."""
    prims = 0
    for c in rb.split_commands(words):
        op = (int(c[0]) >> 24) & 63
        if 0x08 <= op <= 0x0F or op in (0x24, 0x25, 0x36):
            prims += 1
    return prims, int(words.size) * 4
