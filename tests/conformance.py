"""Randomised differential conformance driver, modelled on the reference's rdp_conformance.cpp
(run_conformance_rasterization :283-607, run_conformance_load_tile :609-971, suite table :1019-1668).

Every suite feeds identical command words and identical initial RDRAM / hidden RDRAM / TMEM into the Angrylion
oracle and into RDP::CommandProcessor, then requires RDRAM + hidden RDRAM (or TMEM) to be byte-identical after
every iteration.  RNG is numpy's RandomState(1337) (MT19937), so streams are reproducible.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from harness.rdp_builder import *  # noqa: F401,F403
from harness import rdp_builder as rb


@dataclass
class Variant:
    interlace: bool = False
    aa: bool = False
    color: bool = False
    depth: bool = False
    color_depth_alias: bool = False
    depth_compare: bool = False
    texture: bool = False
    pipelined_texel1: bool = False
    tlut: bool = False
    tlut_type: bool = False
    mid_texel: bool = False
    convert_one: bool = False
    bilerp0: bool = True
    bilerp1: bool = True
    dither: int = rb.DITHER_OFF
    alpha_dither: int = rb.ADITHER_OFF
    texture_format: int = rb.FMT_RGBA
    texture_size: int = rb.SIZE_4
    cycle_type: int = rb.CYCLE_1
    z_mode: int = rb.Z_OPAQUE
    coverage_mode: int = rb.CVG_CLAMP
    fb_fmt: int = rb.FMT_RGBA
    fb_size: int = rb.SIZE_16
    primitive_count: int = 1
    sample_quad: bool = True
    perspective: bool = False
    alpha_test: bool = False
    alpha_test_dither: bool = False
    cvg_times_alpha: bool = False
    alpha_cvg_select: bool = False
    combiner_inputs: bool = False
    blending: bool = False
    lod_frac: bool = False
    lod_sharpen: bool = False
    lod_detail: bool = False
    randomize_rdram: bool = False
    image_read_enable: bool = False
    color_on_coverage: bool = False
    force_flip: bool = False
    fill_rect: bool = False
    tex_rect: bool = False
    prim_depth: bool = False
    ym_out_of_range: bool = False


class Rng:
    def __init__(self, seed=1337):
        self.r = np.random.RandomState(seed)

    def u32(self):
        return int(self.r.randint(0, 1 << 32, dtype=np.uint64))

    def boolean(self):
        return bool(self.u32() & 1)

    def uniform(self, lo, hi, size=None):
        return self.r.uniform(lo, hi, size)


def _pick(rng, base, keep_if, options):
    if base == keep_if:
        return base
    k = rng.u32() % (len(options) + 1)
    return options[k] if k < len(options) and options[k] is not None else base


def _random_combiner_cycle(rng, base_rgb, base_alpha):
    a, b, c, d = base_rgb
    aa, ab, ac, ad = base_alpha
    a = _pick(rng, a, CC.A_COMBINED, [CC.A_ENV, CC.A_ONE, CC.A_PRIM, CC.A_NOISE])
    b = _pick(rng, b, -1, [CC.B_ENV, CC.B_K4, CC.B_KEY_CENTER, CC.B_PRIM])
    c = _pick(rng, c, CC.C_COMBINED_A, [CC.C_PRIM, CC.C_PRIM_A, CC.C_ENV, CC.C_ENV_A, CC.C_K5, CC.C_KEY_SCALE])
    d = _pick(rng, d, CC.D_COMBINED, [None, CC.D_ENV, CC.D_ONE])
    aa = _pick(rng, aa, CC.AA_COMBINED, [None, CC.AA_ENV, CC.AA_ONE, CC.AA_PRIM])
    ab = _pick(rng, ab, CC.AA_COMBINED, [None, CC.AA_ENV, CC.AA_ONE, CC.AA_PRIM])
    ac = _pick(rng, ac, -1, [None, CC.AC_ENV, CC.AC_PRIM])
    ad = _pick(rng, ad, CC.AA_COMBINED, [None, CC.AA_ENV, CC.AA_ONE, CC.AA_PRIM])
    return (a, b, c, d), (aa, ab, ac, ad)


def random_primitives(rng: Rng, count: int, v: Variant):
    """count random triangles in a 320x240 viewport -> (xy, z, inv_w, uv, rgba)."""
    if v.force_flip:
        x = np.stack([rng.uniform(-1.0, -0.5, count), rng.uniform(0.5, 1.0, count), rng.uniform(-1.0, -0.5, count)], 1)
        y = np.stack([rng.uniform(-1.0, -0.5, count), rng.uniform(-0.3, 0.3, count), rng.uniform(0.5, 1.0, count)], 1)
    else:
        x = rng.uniform(-1.0, 1.0, (count, 3))
        y = rng.uniform(-1.0, 1.0, (count, 3))
    z = rng.uniform(0.0, 0.95, (count, 3))
    w = rng.uniform(1.0, 5.0, (count, 3))
    uv = rng.uniform(-1000.0, 1000.0, (count, 3, 2))
    if v.color:
        rgba = rng.uniform(-5.0, 5.0, (count, 3, 4))
    elif v.depth:
        rgba = np.zeros((count, 3, 4))
    else:
        rgba = np.ones((count, 3, 4))
    if not v.depth:
        z = np.zeros_like(z)
    xy = np.stack([(0.5 * x + 0.5) * 320.0, (0.5 * y + 0.5) * 240.0], axis=2)
    return xy, z, 1.0 / w, uv, rgba


def apply_ym_offsets(tri_words: np.ndarray, deltas: np.ndarray):
    """Edge-case generator: push YM above the top vertex or YL below the bottom (reference: draw_triangle_ym_out_of_range)."""
    tri = tri_words.copy()
    for i in range(tri.shape[0]):
        d = int(deltas[i])
        w0, w1 = int(tri[i, 0]), int(tri[i, 1])
        yl = w0 & 0x3FFF
        yh = w1 & 0x3FFF
        if d < 0:
            ym = (yh + d) & 0x3FFF
            w1 = (w1 & ~(0x3FFF << 16)) | (ym << 16)
        elif d > 0:
            yl = (yl + d) & 0x3FFF
            w0 = (w0 & ~0x3FFF) | yl
        tri[i, 0], tri[i, 1] = w0, w1
    return tri


def run_rasterization(sbs, v: Variant, iterations: int, seed: int = 1337, first: int = 0, digests=None):
    """Returns a list of error strings (empty = pass).  With `digests` (a list) the SHA-256 of RDRAM + hidden RDRAM
    after every iteration is appended (taken from the oracle when present, else from the GPU side)."""
    rng = Rng(seed)
    b = Builder()
    om = b.om

    b.set_color_image(v.fb_fmt, v.fb_size, 0, 320)
    b.set_depth_image(0 if v.color_depth_alias else (1 << 20))
    om.sample_quad = v.sample_quad

    if v.texture:
        sbs.write_tmem(rng.r.randint(0, 256, 4096, dtype=np.uint8))
        om.perspective = v.perspective
        if v.lod_frac:
            b.set_combine(((CC.A_TEXEL1, CC.B_TEXEL0, CC.C_LOD_FRAC, CC.D_TEXEL0), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_TEXEL0)),
                          ((CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_COMBINED), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_COMBINED)))
        else:
            d = CC.D_TEXEL1 if v.pipelined_texel1 else CC.D_TEXEL0
            b.set_combine(((CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, d), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_TEXEL0)))
        om.tex_lod = v.lod_frac
    else:
        b.set_combine(((CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_SHADE), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_SHADE)))

    om.alpha_test = v.alpha_test
    om.alpha_test_dither = v.alpha_test_dither
    b.set_blend_color(0, 0, 0, 130)
    om.aa = v.aa
    om.rgb_dither = v.dither
    om.alpha_dither = v.alpha_dither
    om.cvg_times_alpha = v.cvg_times_alpha
    om.alpha_cvg_select = v.alpha_cvg_select
    om.cycle_type = v.cycle_type
    om.sharpen = v.lod_sharpen
    om.detail = v.lod_detail
    om.z_mode = v.z_mode
    om.z_compare = v.depth_compare
    om.cvg_mode = v.coverage_mode
    om.image_read = v.image_read_enable
    om.color_on_cvg = v.color_on_coverage
    om.prim_depth = v.prim_depth
    om.mid_texel = v.mid_texel
    om.convert_one = v.convert_one
    om.bilerp0 = v.bilerp0
    om.bilerp1 = v.bilerp1

    errors = []
    for index in range(iterations):
        active = index >= first
        if active:
            sbs.clear_rdram()

        if index & 2:
            b.set_scissor_subpixels(19, 14, 1162, 801, v.interlace, v.interlace and bool(index & 1))
        else:
            b.set_scissor_subpixels(0, 0, 1023 + ((index >> 3) & 7), 800, v.interlace, v.interlace and bool(index & 1))

        b.set_fill_color(rng.u32())
        om.z_update = v.depth and (rng.u32() & 1) != 0
        if v.prim_depth:
            b.set_prim_depth(rng.u32() & 0xFFFF, rng.u32() & 0xFFFF)

        two = v.cycle_type == CYCLE_2
        if v.combiner_inputs:
            c0 = _random_combiner_cycle(rng, (CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_ZERO), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_TEXEL0))
            if two:
                c1 = _random_combiner_cycle(rng, (CC.A_COMBINED, CC.B_COMBINED, CC.C_COMBINED_A, CC.D_COMBINED),
                                            (CC.AA_COMBINED, CC.AA_COMBINED, CC.AC_SHADE, CC.AA_COMBINED))
            else:
                c1 = _random_combiner_cycle(rng, (CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_ZERO), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_TEXEL0))
            b.set_combine(c0, c1)
            b.set_env_color(*[rng.u32() & 0xFF for _ in range(4)])
            b.set_key_r(rng.u32() & 0xFFF, rng.u32() & 0xFF, rng.u32() & 0xFF)
            b.set_key_gb(rng.u32() & 0xFFF, rng.u32() & 0xFF, rng.u32() & 0xFF, rng.u32() & 0xFFF, rng.u32() & 0xFF, rng.u32() & 0xFF)

        if v.combiner_inputs or v.convert_one or not v.bilerp0 or (two and not v.bilerp1):
            b.set_convert(*[rng.u32() & 0xFFFF for _ in range(6)])

        b.set_prim_color(16, 0xAA, *[rng.u32() & 0xFF for _ in range(4)])

        if v.blending:
            om.blend[0] = [BL_1A_PIXEL, BL_1B_SHADE_A, BL_2A_FOG, BL_2B_INV_PIXEL_A]
            if two:
                om.blend[1] = [BL_1A_PIXEL, BL_1B_PIXEL_A, BL_2A_MEMORY, BL_2B_INV_PIXEL_A]
            om.force_blend = True
            b.set_fog_color(*[rng.u32() & 0xFF for _ in range(4)])
        elif v.color_on_coverage:
            om.blend[0] = [BL_1A_PIXEL, BL_1B_SHADE_A, BL_2A_MEMORY, BL_2B_INV_PIXEL_A]
            om.force_blend = False
        elif v.z_mode == Z_INTERPENETRATING:
            om.blend[0] = [BL_1A_PIXEL, BL_1B_SHADE_A, BL_2A_MEMORY, BL_2B_MEMORY_A]
            om.force_blend = False
        else:
            om.blend[0] = [BL_1A_PIXEL, BL_1B_PIXEL_A, BL_2A_PIXEL, BL_2B_INV_PIXEL_A]
            om.force_blend = False

        if v.texture:
            mask_s, mask_t, shift_s, shift_t, palette = [rng.u32() & 0xF for _ in range(5)]
            flags = [rng.boolean() for _ in range(4)]
            om.tlut, om.tlut_ia = v.tlut, v.tlut_type
            slo, tlo = rng.u32() & 0xF, rng.u32() & 0xF
            width, height = 4 + (rng.u32() & 0xFF), 4 + (rng.u32() & 0xFF)
            offset = 8
            for i in range(8):
                b.set_tile(i, v.texture_format, v.texture_size, 24, offset, palette, mask_s, shift_s, mask_t, shift_t,
                           clamp_s=flags[0], clamp_t=flags[1], mirror_s=flags[2], mirror_t=flags[3])
                b.set_tile_size_subpixels(i, slo, tlo, width, height)
                offset += 16

        b.touch_other_modes()
        b.flush_other_modes()

        if v.randomize_rdram and active:
            sbs.randomize_rdram(rng.r)
        elif v.randomize_rdram:
            rng.r.randint(0, 256, size=sbs.rdram_size, dtype=np.uint8)
            rng.r.randint(0, 4, size=sbs.hidden_size, dtype=np.uint8)

        n = v.primitive_count
        if v.fill_rect:
            for _ in range(n):
                b.fill_rectangle_subpixels(rng.u32() & 63, rng.u32() & 63, rng.u32() & 2047, rng.u32() & 2047)
        elif v.tex_rect:
            for _ in range(n):
                x, y, w, h = rng.u32() & 63, rng.u32() & 63, rng.u32() & 2047, rng.u32() & 2047
                s, t, dsdx, dtdy = [rng.u32() & 0xFFFF for _ in range(4)]
                if rng.u32() & 1:
                    b.tex_rect(3, x, y, w, h, s, t, dsdx, dtdy, flip=False)
                else:
                    b.tex_rect(2, x, y, w, h, s, t, dsdx, dtdy, flip=True)
        else:
            xy, z, iw, uv, rgba = random_primitives(rng, n, v)
            tri = setup_triangles(xy, z, iw, uv, rgba, tile=0, max_level=6, perspective=True)
            if v.ym_out_of_range and tri.shape[0]:
                tri = apply_ym_offsets(tri, rng.r.randint(-8, 8, tri.shape[0]))
            b.triangles(tri)

        b.sync_full()
        words = b.take()
        if not active:
            # keep state in sync on both sides without paying for the compare
            sbs.run(words)
            continue
        sbs.run(words)
        if digests is not None:
            digests.append(sbs.digest())
        errs = sbs.compare_rdram() if sbs.oracle else []
        if errs:
            errors.append(f"iteration {index}: " + "; ".join(errs))
            break
    return errors


# ---------------------------------------------------------------------------------------------------------------------
# Suite table (names follow the reference's ctest names, CMakeLists.txt:114-352)
# ---------------------------------------------------------------------------------------------------------------------
def _suites():
    S = {}
    S["fill-ym-out-of-range"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_16, ym_out_of_range=True)
    S["fill-8"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_8)
    S["fill-16"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_16)
    S["fill-16-interlace"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_16, interlace=True)
    S["fill-16-ia"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_16, fb_fmt=FMT_IA)
    S["fill-32"] = Variant(cycle_type=CYCLE_FILL, fb_size=SIZE_32)

    def copy(tmem, fbs, atest=False, lut=False, persp=False):
        return Variant(cycle_type=CYCLE_COPY, texture=True, force_flip=fbs != SIZE_8, texture_format=FMT_RGBA,
                       texture_size=tmem, randomize_rdram=fbs == SIZE_4, fb_size=fbs, alpha_test=atest, tlut=lut,
                       perspective=persp)

    for name, tmem, fbs in (("32bpp-fb8", SIZE_32, SIZE_8), ("32bpp-fb16", SIZE_32, SIZE_16), ("4bpp-fb8", SIZE_4, SIZE_8),
                            ("4bpp-fb16", SIZE_4, SIZE_16), ("4bpp-fb4", SIZE_4, SIZE_4), ("8bpp-fb4", SIZE_8, SIZE_4),
                            ("16bpp-fb4", SIZE_16, SIZE_4), ("32bpp-fb4", SIZE_32, SIZE_4), ("8bpp-fb8", SIZE_8, SIZE_8),
                            ("8bpp-fb16", SIZE_8, SIZE_16), ("16bpp-fb8", SIZE_16, SIZE_8), ("16bpp-fb16", SIZE_16, SIZE_16)):
        S["copy-" + name] = copy(tmem, fbs)
    S["copy-4bpp-fb16-tlut"] = copy(SIZE_4, SIZE_16, lut=True)
    S["copy-8bpp-fb16-tlut"] = copy(SIZE_8, SIZE_16, lut=True)
    S["copy-16bpp-fb16-tlut"] = copy(SIZE_16, SIZE_16, lut=True)
    S["copy-16bpp-fb16-alpha-test"] = copy(SIZE_16, SIZE_16, atest=True)
    S["copy-8bpp-fb8-perspective"] = copy(SIZE_8, SIZE_8, persp=True)
    S["copy-8bpp-fb16-perspective"] = copy(SIZE_8, SIZE_16, persp=True)
    S["copy-16bpp-fb8-perspective"] = copy(SIZE_16, SIZE_8, persp=True)
    S["copy-16bpp-fb16-perspective"] = copy(SIZE_16, SIZE_16, persp=True)

    S["fill-rect"] = Variant(fill_rect=True, cycle_type=CYCLE_FILL)
    S["tex-rect"] = Variant(tex_rect=True, texture=True, color=True, cycle_type=CYCLE_COPY, texture_size=SIZE_16, fb_size=SIZE_16)
    S["tex-rect-i8"] = Variant(tex_rect=True, texture=True, color=True, cycle_type=CYCLE_1, texture_size=SIZE_16, fb_size=SIZE_8)
    S["rasterization-noaa"] = Variant()
    S["rasterization-aa"] = Variant(aa=True)
    S["rasterization-interlace-aa"] = Variant(aa=True, interlace=True)
    S["interpolation-color"] = Variant(color=True)
    S["interpolation-depth"] = Variant(depth=True)
    S["interpolation-color-depth"] = Variant(color=True, depth=True)
    S["combiner-1cycle"] = Variant(color=True, depth=True, combiner_inputs=True)
    S["combiner-2cycle"] = Variant(color=True, depth=True, combiner_inputs=True, cycle_type=CYCLE_2)
    S["combiner-2cycle-alpha-test-color"] = Variant(color=True, depth=True, cycle_type=CYCLE_2, alpha_test=True, alpha_dither=ADITHER_PATTERN)
    S["combiner-2cycle-alpha-test-texture"] = Variant(texture=True, color=True, depth=True, cycle_type=CYCLE_2, alpha_test=True, alpha_dither=ADITHER_PATTERN)

    fbs = (("i4", FMT_RGBA, SIZE_4), ("i8", FMT_RGBA, SIZE_8), ("rgba5551", FMT_RGBA, SIZE_16), ("ia88", FMT_IA, SIZE_16), ("rgba8888", FMT_RGBA, SIZE_32))
    for name, fmt, size in fbs:
        S["blender-fog-color-1cycle-" + name] = Variant(color=True, depth=True, blending=True, dither=DITHER_MAGIC,
                                                         alpha_dither=ADITHER_PATTERN, fb_fmt=fmt, fb_size=size)
    for name, fmt, size in fbs:
        S["blender-fog-color-2cycle-" + name] = Variant(color=True, depth=True, blending=True, dither=DITHER_MAGIC,
                                                         alpha_dither=ADITHER_PATTERN, cycle_type=CYCLE_2, randomize_rdram=True,
                                                         fb_fmt=fmt, fb_size=size)

    S["interpolation-color-depth-cvg-times-alpha"] = Variant(color=True, depth=True, cvg_times_alpha=True, alpha_test=True)
    S["interpolation-color-depth-aa-cvg-times-alpha"] = Variant(color=True, depth=True, aa=True, cvg_times_alpha=True, alpha_test=True)
    S["interpolation-color-depth-aa-cvg-times-alpha-alpha-cvg-select"] = Variant(color=True, depth=True, aa=True, cvg_times_alpha=True, alpha_test=True, alpha_cvg_select=True)
    S["interpolation-color-depth-aa-alpha-cvg-select"] = Variant(color=True, depth=True, aa=True, alpha_test=True, alpha_cvg_select=True)

    for name, mode in (("clamp", CVG_CLAMP), ("wrap", CVG_WRAP), ("zap", CVG_ZAP), ("save", CVG_SAVE)):
        S["coverage-" + name] = Variant(color=True, aa=True, coverage_mode=mode, randomize_rdram=True)
        S["coverage-" + name + "-image-read"] = Variant(color=True, aa=True, coverage_mode=mode, image_read_enable=True, randomize_rdram=True)
    S["color-on-coverage"] = Variant(color=True, aa=True, coverage_mode=CVG_WRAP, image_read_enable=True, randomize_rdram=True, color_on_coverage=True)

    for name, mode in (("opaque", Z_OPAQUE), ("interpenetrating", Z_INTERPENETRATING), ("transparent", Z_TRANSPARENT), ("decal", Z_DECAL)):
        for suffix, pd in (("", False), ("-prim-depth", True)):
            S["depth-compare-" + name + suffix] = Variant(color=True, depth=True, aa=True, depth_compare=True, randomize_rdram=True,
                                                          image_read_enable=True, z_mode=mode, prim_depth=pd)

    dithers = (("off", DITHER_OFF, ADITHER_OFF, False), ("test", DITHER_OFF, ADITHER_OFF, True), ("all", DITHER_NOISE, ADITHER_NOISE, True),
               ("magic-pattern", DITHER_MAGIC, ADITHER_PATTERN, False), ("magic-flip", DITHER_MAGIC, ADITHER_INV_PATTERN, False),
               ("magic-noise", DITHER_MAGIC, ADITHER_NOISE, False), ("bayer-pattern", DITHER_BAYER, ADITHER_PATTERN, False),
               ("bayer-flip", DITHER_BAYER, ADITHER_INV_PATTERN, False), ("bayer-noise", DITHER_BAYER, ADITHER_NOISE, False),
               ("off-pattern", DITHER_OFF, ADITHER_PATTERN, False), ("off-flip", DITHER_OFF, ADITHER_INV_PATTERN, False),
               ("off-noise", DITHER_OFF, ADITHER_NOISE, False), ("noise-off", DITHER_NOISE, ADITHER_OFF, False),
               ("noise-pattern", DITHER_NOISE, ADITHER_PATTERN, False), ("noise-flip", DITHER_NOISE, ADITHER_INV_PATTERN, False),
               ("noise-noise", DITHER_NOISE, ADITHER_NOISE, False))
    for name, rgb, alpha, td in dithers:
        S["interpolation-color-depth-alpha-test-dither-" + name] = Variant(color=True, depth=True, dither=rgb, alpha_dither=alpha,
                                                                          alpha_test=True, alpha_test_dither=td)

    lod = dict(color=True, texture=True, texture_format=FMT_RGBA, texture_size=SIZE_16, lod_frac=True, cycle_type=CYCLE_2)
    S["interpolation-color-texture-2cycle-lod-frac"] = Variant(**lod)
    S["interpolation-color-texture-perspective-2cycle-lod-frac"] = Variant(perspective=True, **lod)
    S["interpolation-color-texture-perspective-2cycle-lod-frac-sharpen"] = Variant(perspective=True, lod_sharpen=True, **lod)
    S["interpolation-color-texture-perspective-2cycle-lod-frac-detail"] = Variant(perspective=True, lod_detail=True, **lod)
    S["interpolation-color-texture-perspective-2cycle-lod-frac-sharpen-detail"] = Variant(perspective=True, lod_detail=True, lod_sharpen=True, **lod)

    S["rasterization-many-primitives"] = Variant(color=True, primitive_count=5 * 1024)
    S["rasterization-many-primitives-alias"] = Variant(color=True, depth=True, primitive_count=1024, color_depth_alias=True,
                                                       depth_compare=True, randomize_rdram=True)

    t16 = dict(color=True, texture=True, texture_size=SIZE_16, texture_format=FMT_RGBA)
    S["interpolation-color-texture-pipelined-texel1"] = Variant(pipelined_texel1=True, **t16)
    S["interpolation-color-texture-pipelined-texel1-perspective"] = Variant(pipelined_texel1=True, perspective=True, **t16)
    S["interpolation-color-texture-2cycle-convert-bilerp"] = Variant(cycle_type=CYCLE_2, convert_one=True, **t16)
    S["interpolation-color-texture-2cycle-convert-factors"] = Variant(cycle_type=CYCLE_2, convert_one=True, bilerp1=False, **t16)
    S["interpolation-color-texture-2cycle-implicit-convert-factors"] = Variant(cycle_type=CYCLE_2, sample_quad=False, bilerp1=False, **t16)
    S["interpolation-color-texture-2cycle-implicit-convert-factors-bilerp"] = Variant(cycle_type=CYCLE_2, sample_quad=False, **t16)
    S["interpolation-color-texture-yuv16-nearest"] = Variant(color=True, texture=True, texture_size=SIZE_16, texture_format=FMT_YUV,
                                                             bilerp0=False, bilerp1=False, sample_quad=False)

    def tex(fmt, size, tlut=False, ia=False, quad=True, mid=False):
        return Variant(color=True, texture=True, texture_format=fmt, texture_size=size, tlut=tlut, tlut_type=ia, sample_quad=quad, mid_texel=mid)

    sizes = ((4, SIZE_4), (8, SIZE_8), (16, SIZE_16), (32, SIZE_32))
    for bits, size in sizes:
        S[f"interpolation-color-texture-rgba{bits}"] = tex(FMT_RGBA, size)
        S[f"interpolation-color-texture-rgba{bits}-nearest"] = tex(FMT_RGBA, size, quad=False)
        S[f"interpolation-color-texture-ci{bits}"] = tex(FMT_CI, size)
        S[f"interpolation-color-texture-ia{bits}"] = tex(FMT_IA, size)
        S[f"interpolation-color-texture-i{bits}"] = tex(FMT_I, size)
        S[f"interpolation-color-texture-ci{bits}-mid-texel"] = tex(FMT_CI, size, mid=True)
        S[f"interpolation-color-texture-ci{bits}-tlut-mid-texel"] = tex(FMT_CI, size, tlut=True, mid=True)
        S[f"interpolation-color-texture-ci{bits}-tlut-mid-texel-no-quad"] = tex(FMT_CI, size, tlut=True, quad=False, mid=True)
        S[f"interpolation-color-texture-ci{bits}-tlut"] = tex(FMT_CI, size, tlut=True)
        S[f"interpolation-color-texture-ci{bits}-tlut-no-quad"] = tex(FMT_CI, size, tlut=True, quad=False)
        S[f"interpolation-color-texture-ia{bits}-tlut"] = tex(FMT_IA, size, tlut=True)
        S[f"interpolation-color-texture-i{bits}-tlut"] = tex(FMT_I, size, tlut=True)
        S[f"interpolation-color-texture-rgba{bits}-tlut"] = tex(FMT_RGBA, size, tlut=True)
        S[f"interpolation-color-texture-ci{bits}-tlut-ia16"] = tex(FMT_CI, size, tlut=True, ia=True)
    S["interpolation-color-texture-yuv16"] = tex(FMT_YUV, SIZE_16)
    S["interpolation-color-texture-rgba16-mid-texel"] = tex(FMT_RGBA, SIZE_16, mid=True)
    S["interpolation-color-texture-perspective"] = Variant(color=True, texture=True, perspective=True, texture_size=SIZE_16)

    for cname, quad, mid, bil in (("plain", False, False, False), ("quad", True, False, False), ("mid", False, True, False),
                                  ("quad-mid", True, True, False), ("bilerp", False, False, True), ("quad-bilerp", True, False, True),
                                  ("bilerp-mid", False, True, True), ("quad-mid-bilerp", True, True, True)):
        for cyc_name, cyc in (("1cycle", CYCLE_1), ("2cycle", CYCLE_2)):
            for fmt_name, fmt in (("YUV", FMT_YUV), ("RGBA", FMT_RGBA)):
                for conv_name, conv in (("", False), ("-conv", True)):
                    S[f"texture-convert-{cyc_name}-{fmt_name}{conv_name}{cname}"] = Variant(
                        color=True, texture=True, perspective=True, texture_size=SIZE_16, texture_format=fmt, cycle_type=cyc,
                        sample_quad=quad, mid_texel=mid, bilerp0=bil, bilerp1=bil, convert_one=conv)
    return S


SUITES = _suites()


# ---------------------------------------------------------------------------------------------------------------------
# Texture load tests (reference: run_conformance_load_tile, rdp_conformance.cpp:609-971): known byte ramp in RDRAM,
# one LoadTile / LoadBlock / LoadTLUT, then all 2,048 TMEM words must match the oracle.
# ---------------------------------------------------------------------------------------------------------------------
OP_TILE, OP_BLOCK, OP_TLUT = "tile", "block", "tlut"


def _load_case(sbs, width, height, tmem_offset, tmem_stride, rdram_offset, op, vram_size, tile_size, dxt, yuv):
    sbs.write_tmem(np.zeros(4096, dtype=np.uint8))
    if op == OP_TLUT:
        height = 1
    # byte ramp 0,1,2,... in big-endian memory order == this uint32 pattern in RDRAM storage order
    i = np.arange(64 * 1024, dtype=np.uint64)
    ramp = ((((4 * i + 0) & 0xFF) << 24) | (((4 * i + 1) & 0xFF) << 16) | (((4 * i + 2) & 0xFF) << 8) | ((4 * i + 3) & 0xFF)).astype(np.uint32)
    sbs.write_rdram(2 * 1024 * 1024, ramp)

    b = Builder()
    b.set_texture_image(2 * 1024 * 1024 + rdram_offset, FMT_RGBA, vram_size, width)
    b.set_tile(0, FMT_YUV if yuv else FMT_RGBA, tile_size, tmem_stride, tmem_offset)
    if op == OP_TLUT:
        b.load_tlut(0, 1, 1, width, height)
    elif op == OP_TILE:
        b.load_tile_subpixels(0, (width & 3) + 8, (height & 3) + 8, (width << 2) | (height & 3), (height << 2) | (width & 3))
    else:
        b.load_block(0, 1, 3, width, dxt if dxt else ((1 << 10) >> (height & 3)))
    sbs.run(b.words())
    errs = sbs.compare_tmem()
    if errs:
        return [f"load {op} w={width} h={height} off={tmem_offset} stride={tmem_stride} rdram+{rdram_offset} "
                f"vram_size={vram_size} tile_size={tile_size} dxt={dxt} yuv={yuv}: {errs[0]}"]
    return []


def run_load_cases(sbs, op, vram_size, tile_size, yuv=False, thorough=True):
    errs = []

    def case(w, h, off, stride, roff, dxt=0):
        if not errs:
            errs.extend(_load_case(sbs, w, h, off, stride, roff, op, vram_size, tile_size, dxt, yuv))

    case(32, 7, 0, 128, 0)
    if tile_size == SIZE_16:
        case(64, 32, 0, 128, 0)
    if op == OP_TILE:
        case(128, 64, 0, 128 + 8, 0)
        case(4, 1024, 0, 8, 0)
        case(4, 1023, 0, 8, 0)
    if op == OP_BLOCK:
        if tile_size == SIZE_16 and not yuv:
            case(1600, 1, 0, 0, 0, 103)
            case(1600, 1, 0, 8, 0, 103)
        if (vram_size == SIZE_32 and tile_size == SIZE_32 and not yuv) or (vram_size == SIZE_16 and tile_size == SIZE_16 and yuv):
            case(799, 1, 0, 0, 0, 872)
            case(100, 1, 0, 32, 0, 872)
    case(32, 7, 0x800, 128, 0)
    case(32, 7, 128, 128, 0)
    case(32, 9, 0, 128, 1)
    step = 1 if thorough else 5
    for width in range(9, 32, step):
        for height in range(1, 4):
            case(width, height, 0, 0, height & 3)
    for width in range(8, 32, step):
        case(width, 3, 0, 256, width & 3)
    for stride in (8, 16, 24):
        for width in range(8, 64, step):
            case(width, 3, 0, stride, 0)
    return errs


LOAD_SUITES = {
    "texture-load-tile-16-yuv": [(OP_TILE, SIZE_16, SIZE_16, True)],
    "texture-load-block-16-yuv": [(OP_BLOCK, SIZE_16, SIZE_16, True)],
    "texture-load-tile-4": [(OP_TILE, SIZE_8, SIZE_4, False)],
    "texture-load-tile-8": [(OP_TILE, SIZE_8, SIZE_8, False), (OP_TILE, SIZE_16, SIZE_8, False)],
    "texture-load-tile-16": [(OP_TILE, SIZE_16, SIZE_16, False), (OP_TILE, SIZE_8, SIZE_16, False), (OP_TILE, SIZE_32, SIZE_16, False)],
    "texture-load-tile-32": [(OP_TILE, SIZE_32, SIZE_32, False), (OP_TILE, SIZE_16, SIZE_32, False), (OP_TILE, SIZE_8, SIZE_32, False)],
    "texture-load-block-8": [(OP_BLOCK, SIZE_8, SIZE_8, False)],
    "texture-load-block-16": [(OP_BLOCK, SIZE_16, SIZE_16, False), (OP_BLOCK, SIZE_8, SIZE_16, False)],
    "texture-load-block-32": [(OP_BLOCK, SIZE_32, SIZE_32, False), (OP_BLOCK, SIZE_16, SIZE_32, False), (OP_BLOCK, SIZE_8, SIZE_32, False)],
    "texture-load-tlut-4": [(OP_TLUT, SIZE_16, SIZE_4, False), (OP_TLUT, SIZE_8, SIZE_4, False), (OP_TLUT, SIZE_32, SIZE_4, False)],
    "texture-load-tlut-8": [(OP_TLUT, SIZE_16, SIZE_8, False), (OP_TLUT, SIZE_8, SIZE_8, False), (OP_TLUT, SIZE_32, SIZE_8, False)],
    "texture-load-tlut-16": [(OP_TLUT, SIZE_16, SIZE_16, False), (OP_TLUT, SIZE_8, SIZE_16, False), (OP_TLUT, SIZE_32, SIZE_16, False)],
}


def run_load_suite(sbs, name, thorough=True):
    for op, vram_size, tile_size, yuv in LOAD_SUITES[name]:
        errs = run_load_cases(sbs, op, vram_size, tile_size, yuv, thorough)
        if errs:
            return errs
    return []
