"""Internal upscaling (BASELINE C4-ii): eligible passes are rendered a second time into an upscaled domain.

What can be pinned: (1) native RDRAM + hidden RDRAM stay byte-exact against the oracle with upscaling on (the reference
renders the 1x pass and the upscaled pass on every flush, rdp_renderer.cpp:2955-2980); (2) fill-mode rectangles come out
replicated exactly in every slice; (3) the CUDA kernels and the scalar host-simulator statement of the same code agree
byte for byte on the upscaled domain.  The upscaled image itself has no CPU oracle (Angrylion is 1x only), so beyond that
only self-consistency is checked: box-filtering the upscaled colour image back down lands close to the native image."""
import hashlib

import numpy as np
import pytest

import conformance as cf
from harness import workloads as wl
from harness import rdp_builder as rb
from harness.rdp_builder import Builder
from oracle import angrylion
from sidebyside import SideBySide, rdp

FLAGS = {2: 1 << 2, 4: 1 << 3, 8: 1 << 4}
W, H, COLOR, DEPTH = 640, 480, 0x100000, 0x300000


def _upscaled_image(gpu, scale, addr, width, height, bytes_per_pixel):
    up, hidden = gpu.read_upscaled_rdram()
    dtype = np.uint32 if bytes_per_pixel == 4 else np.uint16
    img = np.zeros((height * scale, width * scale), dtype)
    for sy in range(scale):
        for sx in range(scale):
            sl = up[sy * scale + sx][addr:addr + width * height * bytes_per_pixel].view(dtype).reshape(height, width)
            if bytes_per_pixel == 2:
                sl = sl.reshape(height, width // 2, 2)[:, :, ::-1].reshape(height, width)  # 16-bit index i lives at i ^ 1
            img[sy::scale, sx::scale] = sl
    return img, up, hidden


def _c2(sbs, triangles, frames=1):
    sbs.clear_rdram()
    sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
    rng = np.random.RandomState(2)
    for _ in range(frames):
        sbs.run(wl.c2_frame(rng, triangles))
    sbs.gpu.idle()


def _check_c2(backend, scale, triangles):
    sbs = SideBySide(backend, rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, flags=FLAGS[scale])
    try:
        assert sbs.gpu.upscaling_factor() == scale
        _c2(sbs, triangles)
        errs = sbs.compare_rdram()
        assert not errs, errs  # native result untouched by the second pass
        assert sbs.gpu.upscaled_pass_count() == 2  # Z clear (RGBA16 640 wide) and the main pass
        native = sbs.gpu.begin_read_rdram()[COLOR:COLOR + W * H * 4].view(np.uint32).reshape(H, W).copy()
        img, up, hidden = _upscaled_image(sbs.gpu, scale, COLOR, W, H, 4)
        digest = hashlib.sha256(up.tobytes() + (hidden & 3).tobytes()).hexdigest()
    finally:
        sbs.close()
    for shift in (24, 16, 8):
        a = ((native >> shift) & 0xff).astype(np.float64)
        b = ((img >> shift) & 0xff).astype(np.float64).reshape(H, scale, W, scale).mean(axis=(1, 3))
        # same picture (the textures are random noise, so the four sub-samples of a pixel do not average to the native
        # sample exactly: ~10 of 255 per covered pixel, against ~85 for unrelated images) ...
        assert np.abs(a - b)[native != 0].mean() < 16.0
    assert sum(np.array_equal(img[sy::scale, sx::scale], native) for sy in range(scale) for sx in range(scale)) == 0  # ... but not a replica
    covered_native = (native != 0).sum()
    covered_up = (img != 0).sum() / (scale * scale)
    assert 0.85 * covered_native < covered_up < 1.05 * covered_native
    return digest


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scale,triangles", [(2, 200), (4, 200), (8, 40)])
def test_hostsim_upscaled_c2(scale, triangles):
    _check_c2("hostsim", scale, triangles)


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_fill_rect_is_replicated_and_cpu_writes_propagate():
    scale = 2
    sbs = SideBySide("hostsim", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, flags=FLAGS[scale])
    try:
        sbs.clear_rdram()
        # the "CPU" paints a stripe into the framebuffer before anything is rendered (inside the rows the pass will
        # touch: the domain synchronisation covers the pass's deduced framebuffer height, as in the reference) ...
        stripe = np.full(320 * 2 * 8, 0x5a, np.uint8)
        sbs.write_rdram(COLOR + 320 * 2 * 30, stripe)
        b = Builder()
        b.set_color_image(0, 2, COLOR, 320)  # RGBA16, 320 wide
        b.set_scissor(0, 0, 320, 240)
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.flush_other_modes()
        b.set_fill_color(0xf801f801)
        b.fill_rectangle(16, 8, 200, 60)
        b.sync_full()
        sbs.run(b.take())
        sbs.gpu.idle()
        assert not sbs.compare_rdram()
        assert sbs.gpu.upscaled_pass_count() == 1
        native = sbs.gpu.begin_read_rdram().copy()
        native_hidden = sbs.gpu.begin_read_hidden_rdram().copy()
        up, hidden = sbs.gpu.read_upscaled_rdram()
        lo, hi = COLOR, COLOR + 320 * 68 * 2
        n16 = native[lo:hi].view(np.uint16).reshape(68, 160, 2)[:, :, ::-1].reshape(68, 320)
        nh = native_hidden[lo // 2:hi // 2].reshape(68, 320)
        assert (n16[8:68, 16:216] == 0xf801).all() and (n16[30:38, 216:] == 0x5a5a).all()
        for s in range(scale * scale):
            u16 = up[s][lo:hi].view(np.uint16).reshape(68, 160, 2)[:, :, ::-1].reshape(68, 320)
            uh = hidden[s][lo // 2:hi // 2].reshape(68, 320)
            # ... and both are in every slice: the stripe by the PRE broadcast, the fill by rendering -- fill rectangles
            # are never upscaled (rdp_device.cpp:730), every sub-pixel repeats its native pixel, edges included.
            keep = np.ones((68, 320), bool)
            assert np.array_equal(u16[keep], n16[keep])
            # hidden bits travel with a colour word that changed or was rendered; untouched words keep the domain's
            # initial value (3, rdp_renderer.cpp:593)
            written = keep & (n16 != 0)
            assert np.array_equal(uh[written] & 3, nh[written] & 3)
    finally:
        sbs.close()


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["interpolation-color-depth", "depth-compare-opaque", "interpolation-color-texture-rgba16",
                                  "interpolation-color-texture-pipelined-texel1-perspective", "tex-rect", "fill-16-interlace",
                                  "copy-16bpp-fb8-perspective", "interpolation-color-texture-perspective-2cycle-lod-frac-sharpen-detail"])
def test_hostsim_conformance_native_result_unchanged_by_upscaling(name):
    sbs = SideBySide("hostsim", flags=FLAGS[2])
    try:
        errs = cf.run_rasterization(sbs, cf.SUITES[name], 3)
        passes = sbs.gpu.upscaled_pass_count()
    finally:
        sbs.close()
    assert not errs, errs
    assert passes > 0 or "fb8" in name


# --- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scale,triangles", [(2, 1500), (4, 1500), (8, 300)])
def test_gpu_upscaled_c2_matches_scalar_statement(scale, triangles):
    """C2 frame: native exact vs oracle, self-consistency, and the CUDA upscaled domain == host simulator's."""
    gpu = _check_c2("cuda", scale, triangles)
    cpu = _check_c2("hostsim", scale, triangles)
    assert gpu == cpu


@pytest.mark.gpu
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["interpolation-color-depth", "depth-compare-opaque", "interpolation-color-texture-rgba16",
                                  "interpolation-color-texture-pipelined-texel1-perspective", "tex-rect", "fill-rect",
                                  "combiner-2cycle", "interpolation-color-texture-perspective-2cycle-lod-frac-sharpen-detail"])
def test_gpu_conformance_with_upscaling(name):
    """Native result still exact with 4x upscaling on, and the upscaled domain equals the host simulator's."""
    digests = {}
    for backend in ("cuda", "hostsim"):
        sbs = SideBySide(backend, flags=FLAGS[4])
        try:
            errs = cf.run_rasterization(sbs, cf.SUITES[name], 2)
            assert not errs, errs
            assert sbs.gpu.upscaled_pass_count() > 0
            up, hidden = sbs.gpu.read_upscaled_rdram()
            digests[backend] = hashlib.sha256(up.tobytes() + (hidden & 3).tobytes()).hexdigest()
        finally:
            sbs.close()
    assert digests["cuda"] == digests["hostsim"]


# --- scan-out of the upscaled domain (video_interface.cpp:1215-1426 with scaling_factor > 1) -------------------------
def _scanout_c2(backend, scale, triangles, steps):
    """One C2 frame, scanned out with `downscale_steps` = each of `steps`: list of (pixels, w, h), plus the oracle image."""
    sbs = SideBySide(backend, rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, flags=FLAGS.get(scale, 0))
    try:
        _c2(sbs, triangles)
        for k, v in wl.c2_vi_registers(0).items():
            sbs.set_vi_register(rdp.VI_REG[k], v)
        ref = sbs.oracle.update_screen()
        outs = []
        for n in steps:
            opts = sbs.gpu.default_scanout_options()
            opts.blend_previous_frame = 0
            opts.upscale_deinterlacing = 0
            opts.downscale_steps = n
            opts.vi_gamma_dither = 0  # the dither noise is seeded by the frame counter: keep repeated scan-outs comparable
            pix, w, h = sbs.gpu.scanout_sync(opts)
            outs.append((pix.copy(), w, h))
        return outs, ref
    finally:
        sbs.close()


def _rgb_planes(img):
    return [((img >> s) & 0xff).astype(np.float64) for s in (0, 8, 16)]


def _check_scanout(backend, scale, triangles):
    log2 = scale.bit_length() - 1
    outs, ref = _scanout_c2(backend, scale, triangles, list(range(log2 + 2)))
    (native_like, w0, h0) = outs[log2]  # halved all the way down: the native output size
    assert (w0, h0) == (ref.shape[1], ref.shape[0])
    for n, (pix, w, h) in enumerate(outs):
        eff = min(n, log2)
        assert (w, h) == (ref.shape[1] * scale >> eff, ref.shape[0] * scale >> eff)
        assert pix.any()
    # more steps than there is scale to undo changes nothing (video_interface.cpp:1009)
    assert np.array_equal(outs[log2][0], outs[log2 + 1][0])
    # each step is exactly the rounded 2x2 mean of the previous image
    for n in range(log2):
        big, small = outs[n][0], outs[n + 1][0]
        for a, b in zip(_rgb_planes(big), _rgb_planes(small)):
            box = np.floor((a.reshape(b.shape[0], 2, b.shape[1], 2).sum(axis=(1, 3)) + 2) / 4)
            assert np.array_equal(box, b)
    # and the fully halved image is the native scan-out up to what supersampling changes (edges, noise-texture detail)
    # and the oracle's gamma dither, which this comparison leaves on: ~10-20 of 255 per lit pixel, against ~85 for
    # unrelated images
    for a, b in zip(_rgb_planes(native_like), _rgb_planes(ref)):
        assert np.abs(a - b)[ref != 0].mean() < 32.0
    return hashlib.sha256(b"".join(o[0].tobytes() for o in outs)).hexdigest()


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_upscaled_scanout():
    _check_scanout("hostsim", 2, 150)


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_downscale_steps_are_ignored_without_upscaling():
    outs, ref = _scanout_c2("hostsim", 1, 150, [0, 2])
    assert outs[0][1:] == outs[1][1:] == (ref.shape[1], ref.shape[0])
    assert np.array_equal(outs[0][0], outs[1][0])


@pytest.mark.gpu
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scale", [2, 4])
def test_gpu_upscaled_scanout_matches_scalar_statement(scale):
    assert _check_scanout("cuda", scale, 800) == _check_scanout("hostsim", scale, 800)


# --- quirks (rdp_device.hpp:76-109): which primitives an upscaled pass really upscales ---------------------------------
def _tex_rect_slices(native_tex_rect: bool):
    """A scaled texture rectangle over a cleared 320x240 RGBA16 target at 2x: the four slices of the colour image."""
    scale = 2
    sbs = SideBySide("hostsim", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, flags=FLAGS[scale])
    try:
        sbs.clear_rdram()
        sbs.gpu.set_quirks(native_tex_rect, False)
        sbs.write_tmem(np.random.RandomState(5).randint(0, 256, 4096, dtype=np.uint8))
        b = Builder()
        b.set_color_image(0, 2, COLOR, 320)
        b.set_scissor(0, 0, 320, 240)
        b.om.cycle_type = rb.CYCLE_1
        b.om.sample_quad = True
        b.touch_other_modes()
        b.flush_other_modes()
        CC = rb.CC
        b.set_combine(((CC.A_ZERO, CC.B_ZERO, CC.C_ZERO, CC.D_TEXEL0), (CC.AA_ZERO, CC.AA_ZERO, CC.AC_ZERO, CC.AA_TEXEL0)))
        b.set_tile(0, rb.FMT_RGBA, rb.SIZE_16, 64, 0, mask_s=5, mask_t=5)
        b.set_tile_size(0, 0, 0, 32, 32)
        b.tex_rect(0, 20, 12, 150, 90, 0, 0, 150, 170)  # s/t steps that are not whole texels: the filter has work to do
        b.sync_full()
        sbs.run(b.take())
        sbs.gpu.idle()
        assert not sbs.compare_rdram()
        assert sbs.gpu.upscaled_pass_count() == 1
        up, _ = sbs.gpu.read_upscaled_rdram()
        lo, hi = COLOR, COLOR + 320 * 110 * 2
        return [up[s][lo:hi].view(np.uint16).reshape(110, 320).copy() for s in range(scale * scale)]
    finally:
        sbs.close()


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_native_resolution_tex_rect_quirk():
    on = _tex_rect_slices(True)    # the default: texture rectangles are not upscaled -> every sub-pixel repeats its pixel
    assert on[0].any()
    assert all(np.array_equal(on[0], s) for s in on[1:])
    off = _tex_rect_slices(False)  # upscaled like a triangle: the sub-pixels sample the texture at their own positions
    assert off[0].any()
    assert not all(np.array_equal(off[0], s) for s in off[1:])
