"""The reference's analytic VI test (vi_conformance.cpp:163-267, `per-scanline-xh`): HStart / XScale latched per VI line,
checked against a closed-form model of the horizontal scaler instead of the oracle (Angrylion has no raster effects) --
with and without a crop rectangle, and, as in the reference, also through 2x internal upscaling."""
import numpy as np
import pytest

import vi_conformance as vc
from harness import rdp_builder as rb
from sidebyside import SideBySide, rdp

REG = rdp.VI_REG


def _expected(scale: int, crop: bool):
    w, h = 640 * scale, 240 * scale
    ref = np.zeros((h, w), np.uint32)

    def region(y0, y1, h_start, h_end, left_clamp, right_clamp, x_add, x_start):
        x_base = h_start
        if not left_clamp:
            h_start += 8
        if not right_clamp:
            h_end -= 7
        x_begin, x_end = max(h_start, 0) * scale, min(h_end, 640) * scale
        x = np.arange(x_begin, x_end)
        sample_x = (x - x_base * scale) * x_add + x_start * scale
        frac = (sample_x >> 5) & 31
        lo = (sample_x >> 10) // scale
        hi = ((sample_x >> 10) + 1) // scale
        ref[y0 * scale:y1 * scale, x_begin:x_end] = ((lo * (32 - frac) + hi * frac + 16) >> 5).astype(np.uint32)  # red channel

    region(20, 50, 0, 320, False, False, 256, 0)  # the top 20 lines are blanked
    region(50, 100, 0, 640, False, False, 240, 200)
    region(100, 150, -8, 648, True, True, 220, 400)
    region(150, 200, 8, 640, False, True, 210, 600)
    if crop:
        ref = ref[11 * scale:h - 12 * scale, 9 * scale:w - 10 * scale]
    return ref


def _run(backend: str, upscale: bool, crop: bool):
    scale = 2 if upscale else 1
    sbs = SideBySide(backend, with_oracle=False, flags=(1 << 2) if upscale else 0)
    try:
        vc._set_defaults(sbs, vc.ViVariant(fmt=rb.VI_CONTROL_TYPE_RGBA8888, aa=rb.VI_CONTROL_AA_RESAMP_ONLY))
        sbs.gpu.idle()
        fb = sbs.gpu.begin_read_rdram()[4096:4096 + 200 * 240 * 4].view(np.uint32).reshape(240, 200)
        fb[:, :] = (np.arange(200, dtype=np.uint32) << 24)[None, :]
        sbs.gpu.end_write_rdram()
        sbs.set_vi_register(REG["Origin"], 4096)
        sbs.set_vi_register(REG["Width"], 200)
        v, hoff = vc.V_OFFSET_NTSC, vc.H_OFFSET_NTSC
        sbs.set_vi_register(REG["VStart"], rb.vi_start_register(v + 20 * 2, v + 200 * 2))
        g = sbs.gpu
        g.begin_vi_register_per_scanline(g.PER_SCANLINE_HSTART | g.PER_SCANLINE_XSCALE)
        for line, (h0, h1), (x_add, x_start) in ((v, (hoff, hoff + 320), (256, 0)),
                                                 (v + 50 * 2, (hoff, hoff + 640), (240, 200)),
                                                 (v + 100 * 2, (hoff - 8, hoff + 648), (220, 400)),
                                                 (v + 150 * 2, (hoff + 8, hoff + 648), (210, 600))):
            g.set_vi_register_for_scanline(g.PER_SCANLINE_HSTART, rb.vi_start_register(h0, h1))
            g.set_vi_register_for_scanline(g.PER_SCANLINE_XSCALE, rb.vi_scale_register(x_add, x_start))
            g.latch_vi_register_for_scanline(line)
        g.end_vi_register_per_scanline()
        opts = g.default_scanout_options()
        opts.blend_previous_frame = 1
        opts.upscale_deinterlacing = 0
        if crop:
            opts.crop_enable, opts.crop_left, opts.crop_right, opts.crop_top, opts.crop_bottom = 1, 9, 10, 11, 12
        pix, w, h = g.scanout_sync(opts)
        pix = pix.copy()
    finally:
        sbs.close()
    ref = _expected(scale, crop)
    assert (h, w) == ref.shape
    assert np.array_equal(pix & 0xFFFFFF, ref)


@pytest.mark.parametrize("upscale", [False, True])
@pytest.mark.parametrize("crop", [False, True])
def test_hostsim_per_scanline_xh(upscale, crop):
    _run("hostsim", upscale, crop)


@pytest.mark.gpu
@pytest.mark.parametrize("upscale", [False, True])
@pytest.mark.parametrize("crop", [False, True])
def test_gpu_per_scanline_xh(upscale, crop):
    _run("cuda", upscale, crop)
