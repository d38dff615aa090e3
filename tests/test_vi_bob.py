"""Bob deinterlacing of an upscaled scan-out (ScanoutOptions::upscale_deinterlacing, video_interface.cpp:1066-1123,
1355-1389; shaders/vi_deinterlace.vert/.frag).

The reference reaches its deinterlacer only for a serrated frame that is scanned out upscaled as a single field
(upscale_deinterlacing) and then halved (downscale_steps > 0): the halved field is resampled through a clamped linear
sampler at a vertical offset of -+ scale'/4 rows, scale' being what is left of the upscaling factor, into an image
of twice the rows when scale' is 1.  Every sample position is a whole or a half source row, so the filter is a row copy
or the mean of two rows -- restated here in numpy over the un-bobbed field image (the same scan-out with the serrate
override off, which takes the single-field path without the deinterlacer)."""
import numpy as np
import pytest

from harness import workloads as wl
from sidebyside import SideBySide, rdp

FLAGS = {2: 1 << 2, 4: 1 << 3, 8: 1 << 4}


def _scanouts(backend, scale, steps, field, triangles=150):
    sbs = SideBySide(backend, rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, with_oracle=False, flags=FLAGS[scale])
    try:
        sbs.clear_rdram()
        sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
        sbs.run(wl.c2_frame(np.random.RandomState(2), triangles))
        sbs.gpu.idle()
        for k, v in wl.c2_vi_registers(field).items():
            sbs.set_vi_register(rdp.VI_REG[k], v)
        outs = []
        for serrate in (0, 1):
            opts = sbs.gpu.default_scanout_options()
            opts.blend_previous_frame = 0
            opts.upscale_deinterlacing = 1
            opts.downscale_steps = steps
            opts.vi_serrate = serrate
            opts.vi_gamma_dither = 0  # seeded by the frame counter
            pix, w, h = sbs.gpu.scanout_sync(opts)
            assert pix.shape == (h, w)
            outs.append(pix.copy())
        return outs
    finally:
        sbs.close()


def _bob(src, remaining_scale, field):
    h = src.shape[0]
    out_h = 2 * h if remaining_scale == 1 else h
    y = np.arange(out_h)
    if remaining_scale == 1:
        ty2 = y - 1 if field else y
    else:
        ty2 = 2 * y - remaining_scale // 2 if field else 2 * y + remaining_scale // 2
    r0 = np.clip(ty2 >> 1, 0, h - 1)
    r1 = np.clip((ty2 >> 1) + (ty2 & 1), 0, h - 1)
    out = np.zeros((out_h, src.shape[1]), np.uint32)
    for sh in (0, 8, 16, 24):
        a = (src[r0] >> sh) & 0xff
        b = (src[r1] >> sh) & 0xff
        out |= ((a + b + 1) >> 1) << sh
    return out


def _check(backend, scale, steps, field):
    plain, bobbed = _scanouts(backend, scale, steps, field)
    remaining = max(1, scale >> steps)
    assert plain.shape == (240 * scale >> steps, 640 * scale >> steps)  # one field, halved
    assert bobbed.shape == (plain.shape[0] * (2 if remaining == 1 else 1), plain.shape[1])
    assert plain.any()
    # field == 0 is the scan-out with VCurrentLine == 0: sampled a quarter (scaled) row up
    assert np.array_equal(bobbed, _bob(plain, remaining, field == 0))
    return bobbed


@pytest.mark.parametrize("scale,steps,field", [(2, 1, 0), (2, 1, 1), (4, 1, 0), (4, 1, 1), (4, 2, 0), (8, 1, 1)])
def test_hostsim_bob_is_the_half_row_resample_of_the_field(scale, steps, field):
    _check("hostsim", scale, steps, field)


def test_hostsim_no_bob_without_downscale_steps():
    """upscale_deinterlacing alone only selects the single-field scan-out (video_interface.cpp:1355)."""
    a, b = _scanouts("hostsim", 2, 0, 0)
    assert a.shape == b.shape == (480, 1280)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("scale,steps,field", [(2, 1, 0), (2, 1, 1), (4, 1, 1), (4, 2, 0)])
def test_gpu_bob_matches_scalar_statement(scale, steps, field):
    assert np.array_equal(_check("cuda", scale, steps, field), _check("hostsim", scale, steps, field))
