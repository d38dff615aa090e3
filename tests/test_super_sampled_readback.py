"""Super-sampled read-back (COMMAND_PROCESSOR_FLAG_SUPER_SAMPLED_READ_BACK_BIT / _DITHER_BIT, rdp_device.hpp:50-51).

With this flag the reference renders every pass in the upscaled domain only and resolves it into native RDRAM
(rdp_renderer.cpp:2296-2334, 2955-2980; shaders/update_upscaled_domain_resolve.comp): a native pixel one of whose
S*S samples was stored to takes the rounded average of its samples' channels (16 / 32-bit colour), its first sample
(8-bit colour, depth).  Native RDRAM is then no longer Angrylion's image, so what is pinned here is (1) the resolve
arithmetic, restated in numpy over the upscaled domain the pass left behind, (2) exactness where averaging cannot
change anything (fill rectangles, which are never upscaled), (3) the refusal without upscaling, and (4) on the GPU,
byte equality of native RDRAM and the upscaled domain with the scalar statement in the host simulator."""
import hashlib

import numpy as np
import pytest

import conformance as cf
from harness import workloads as wl
from harness import rdp_builder as rb
from harness.rdp_builder import Builder
from oracle import angrylion
from sidebyside import SideBySide, build_hostsim, rdp

UPSCALE = {2: 1 << 2, 4: 1 << 3, 8: 1 << 4}
SSAA, SSAA_DITHER = 1 << 5, 1 << 6
W, H, COLOR, DEPTH = 640, 480, 0x100000, 0x300000

BAYER = np.array([[0, 4, 1, 5], [4, 0, 5, 1], [3, 7, 2, 6], [7, 3, 6, 2]], np.uint32)


def _run_c2(backend, scale, triangles, flags):
    sbs = SideBySide(backend, rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, with_oracle=False, flags=UPSCALE[scale] | flags)
    try:
        sbs.clear_rdram()
        sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
        sbs.run(wl.c2_frame(np.random.RandomState(2), triangles))
        sbs.gpu.idle()
        passes = sbs.gpu.upscaled_pass_count()
        native = sbs.gpu.begin_read_rdram().copy()
        up, hidden = sbs.gpu.read_upscaled_rdram()
        return native, up.copy(), hidden.copy(), passes
    finally:
        sbs.close()


def _resolve32(up, addr, count):
    n = up.shape[0]
    words = np.stack([up[s][addr:addr + count * 4].view(np.uint32) for s in range(n)]).astype(np.uint64)
    out = np.zeros(count, np.uint64)
    for shift in (24, 16, 8, 0):
        out |= ((((words >> shift) & 0xff).sum(axis=0) + (n >> 1)) // n) << shift
    return out.astype(np.uint32)


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_hostsim_c2_native_image_is_the_average_of_its_samples(scale):
    native, up, hidden, passes = _run_c2("hostsim", scale, 200 if scale < 8 else 40, SSAA)
    assert passes == 2  # Z clear and main pass, both upscaled only
    got = native[COLOR:COLOR + W * H * 4].view(np.uint32)
    want = _resolve32(up, COLOR, W * H)
    assert np.array_equal(got, want)
    assert (got != 0).sum() > 4000  # something was rendered
    # depth cannot be filtered: the first sample
    assert np.array_equal(native[DEPTH:DEPTH + W * H * 2], up[0][DEPTH:DEPTH + W * H * 2])
    # an anti-aliased picture, not a replica of one sample
    img = got.reshape(H, W)
    assert any(not np.array_equal(img, up[s][COLOR:COLOR + W * H * 4].view(np.uint32).reshape(H, W)) for s in range(scale * scale))


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_c2_resolved_image_is_close_to_the_native_render():
    """Same picture as the oracle's 1x render, up to what averaging four samples of a noise texture does."""
    native, _, _, _ = _run_c2("hostsim", 2, 200, SSAA)
    sbs = SideBySide("hostsim", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE)
    try:
        sbs.clear_rdram()
        sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
        sbs.run(wl.c2_frame(np.random.RandomState(2), 200))
        assert not sbs.compare_rdram()
        ref = sbs.oracle.rdram[COLOR:COLOR + W * H * 4].view(np.uint32).copy()
    finally:
        sbs.close()
    got = native[COLOR:COLOR + W * H * 4].view(np.uint32)
    covered = ref != 0
    assert 0.85 * covered.sum() < (got != 0).sum() < 1.15 * covered.sum()
    for shift in (24, 16, 8):
        a = ((ref >> shift) & 0xff).astype(np.float64)
        b = ((got >> shift) & 0xff).astype(np.float64)
        assert np.abs(a - b)[covered].mean() < 16.0
    assert not np.array_equal(got, ref)


def _fill_rect_16(backend, flags, scale=2):
    """RGBA16 320-wide: a fill rectangle (never upscaled)."""
    sbs = SideBySide(backend, with_oracle=angrylion.available(), flags=UPSCALE[scale] | flags)
    try:
        sbs.clear_rdram()
        b = Builder()
        b.set_color_image(0, 2, COLOR, 320)
        b.set_scissor(0, 0, 320, 240)
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.flush_other_modes()
        b.set_fill_color(0x39cf39cf)
        b.fill_rectangle(16, 8, 200, 60)
        b.sync_full()
        sbs.run(b.take())
        sbs.gpu.idle()
        native = sbs.gpu.begin_read_rdram().copy()
        oracle = sbs.oracle.rdram.copy() if sbs.oracle else None
        up, _ = sbs.gpu.read_upscaled_rdram()
        return native, oracle, up.copy()
    finally:
        sbs.close()


@pytest.mark.parametrize("flags", [SSAA, SSAA | SSAA_DITHER])
def test_hostsim_fill_rectangle_resolves_exactly(flags):
    native, oracle, up = _fill_rect_16("hostsim", flags)
    lo, hi = COLOR, COLOR + 320 * 68 * 2
    n16 = native[lo:hi].view(np.uint16).reshape(68, 160, 2)[:, :, ::-1].reshape(68, 320)
    # every sample holds the fill colour, so neither the rounding nor the dither (< 1 LSB of 5 bits) can move it
    assert (n16[8:68, 16:216] == 0x39cf).all()
    assert (n16[:8] == 0).all() and (n16[:, :16] == 0).all()
    if oracle is not None:
        assert np.array_equal(native, oracle)  # RDRAM only: hidden RDRAM is not resolved
    for s in range(4):
        assert np.array_equal(up[s][lo:hi], native[lo:hi])


def test_hostsim_cpu_writes_survive_the_next_resolve():
    """What the write mask is for: a pixel the pass did not store to is not resolved, so a value the CPU put into native
    RDRAM between two passes stays (and reaches every sample through the PRE synchronisation)."""
    sbs = SideBySide("hostsim", with_oracle=False, flags=UPSCALE[2] | SSAA)
    try:
        sbs.clear_rdram()

        def fill(x, y, w, h, color):
            b = Builder()
            b.set_color_image(0, 2, COLOR, 320)
            b.set_scissor(0, 0, 320, 240)
            b.om.cycle_type = rb.CYCLE_FILL
            b.touch_other_modes()
            b.flush_other_modes()
            b.set_fill_color(color)
            b.fill_rectangle(x, y, w, h)
            b.sync_full()
            sbs.run(b.take())
            sbs.gpu.idle()

        fill(0, 0, 320, 100, 0x08430843)
        stripe = np.full(320 * 2 * 4, 0x5a, np.uint8)  # rows 40..43, inside the first rectangle
        sbs.write_rdram(COLOR + 320 * 2 * 40, stripe)
        fill(16, 60, 100, 30, 0xf801f801)  # same framebuffer, rows 60..89: rows 0..89 are synchronised and resolved
        native = sbs.gpu.begin_read_rdram().copy()
        up, _ = sbs.gpu.read_upscaled_rdram()
    finally:
        sbs.close()
    n16 = native[COLOR:COLOR + 320 * 100 * 2].view(np.uint16).reshape(100, 160, 2)[:, :, ::-1].reshape(100, 320)
    assert (n16[40:44] == 0x5a5a).all()
    assert (n16[:40] == 0x0843).all() and (n16[44:60] == 0x0843).all()
    assert (n16[60:90, 16:116] == 0xf801).all() and (n16[60:90, 116:] == 0x0843).all() and (n16[90:] == 0x0843).all()
    for s in range(4):
        u16 = up[s][COLOR:COLOR + 320 * 90 * 2].view(np.uint16).reshape(90, 160, 2)[:, :, ::-1].reshape(90, 320)
        assert np.array_equal(u16, n16[:90])


def _resolve16(up, addr, count, width, dither):
    n = up.shape[0]
    # 16-bit element i lives at i ^ 1; undo it per pair to get pixel order
    words = np.stack([up[s][addr:addr + count * 2].view(np.uint16).reshape(-1, 2)[:, ::-1].reshape(-1) for s in range(n)]).astype(np.uint32)
    x = np.arange(count, dtype=np.uint32) % width
    y = np.arange(count, dtype=np.uint32) // width
    d = BAYER[y & 3, x & 3] * n
    out = np.zeros(count, np.uint32)
    for shift, bits in ((11, 31), (6, 31), (1, 31), (0, 1)):
        c = ((words >> shift) & bits).sum(axis=0)
        c = (8 * c + d) // (8 * n) if dither else (c + (n >> 1)) // n
        out |= c << shift
    return out.astype(np.uint16)


@pytest.mark.parametrize("name,dither", [("interpolation-color-depth", False), ("interpolation-color-depth", True),
                                         ("interpolation-color-texture-rgba16", True)])
def test_hostsim_16bit_resolve_arithmetic(name, dither):
    """RGBA16 320-wide conformance scene: native colour == numpy restatement of the resolve over the samples."""
    sbs = SideBySide("hostsim", with_oracle=False, flags=UPSCALE[2] | SSAA | (SSAA_DITHER if dither else 0))
    try:
        assert not cf.run_rasterization(sbs, cf.SUITES[name], 1)
        sbs.gpu.idle()
        native = sbs.gpu.begin_read_rdram().copy()
        up, _ = sbs.gpu.read_upscaled_rdram()
    finally:
        sbs.close()
    rows = 200  # scissor 0..800/4
    got = native[:320 * rows * 2].view(np.uint16).reshape(-1, 2)[:, ::-1].reshape(-1)
    want = _resolve16(up, 0, 320 * rows, 320, dither)
    written = np.zeros(320 * rows, bool)
    for s in range(4):
        written |= up[s][:320 * rows * 2].view(np.uint16).reshape(-1, 2)[:, ::-1].reshape(-1) != 0
    assert written.sum() > 1000
    # (unwritten pixels hold zeros in every sample, which resolve to zero with or without being touched)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_pass_too_wide_to_upscale_is_rendered_natively():
    """Own limit (span X edges are 16-bit x.3 values): a 1024-wide target at 8x would be 8192 upscaled pixels wide, so the
    pass takes the ordinary native route and stays exact against the oracle, hidden RDRAM included."""
    sbs = SideBySide("hostsim", flags=UPSCALE[8] | SSAA)
    try:
        sbs.clear_rdram()
        b = Builder()
        b.set_color_image(0, 2, COLOR, 1024)
        b.set_scissor(0, 0, 1024, 64)
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.flush_other_modes()
        b.set_fill_color(0x39cf39cf)
        b.fill_rectangle(100, 4, 800, 40)
        b.sync_full()
        sbs.run(b.take())
        assert not sbs.compare_rdram()
        assert sbs.gpu.upscaled_pass_count() == 0
    finally:
        sbs.close()


def test_super_sampled_readback_needs_upscaling():
    rdram = np.zeros(1 << 22, np.uint8)
    with pytest.raises(RuntimeError):
        rdp.CommandProcessor(rdram, 1 << 21, flags=SSAA, library=build_hostsim())


# --- GPU ------------------------------------------------------------------------------------------------------------
def _digest(native, up, hidden):
    return hashlib.sha256(native.tobytes() + up.tobytes() + (hidden & 3).tobytes()).hexdigest()


@pytest.mark.gpu
@pytest.mark.parametrize("scale,triangles,flags", [(2, 1500, SSAA), (4, 600, SSAA), (2, 600, SSAA | SSAA_DITHER)])
def test_gpu_c2_super_sampled_matches_scalar_statement(scale, triangles, flags):
    gpu = _run_c2("cuda", scale, triangles, flags)
    cpu = _run_c2("hostsim", scale, triangles, flags)
    assert gpu[3] == cpu[3] == 2
    assert np.array_equal(gpu[0], cpu[0])  # native RDRAM
    assert _digest(*gpu[:3]) == _digest(*cpu[:3])
    got = gpu[0][COLOR:COLOR + W * H * 4].view(np.uint32)
    assert np.array_equal(got, _resolve32(gpu[1], COLOR, W * H))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["interpolation-color-depth", "depth-compare-opaque", "interpolation-color-texture-rgba16",
                                  "tex-rect", "fill-rect", "fill-8", "combiner-2cycle", "copy-16bpp-fb8-perspective",
                                  "rasterization-many-primitives-alias"])
def test_gpu_conformance_super_sampled_matches_scalar_statement(name):
    if name not in cf.SUITES:
        pytest.skip("no such suite")
    results = {}
    for backend in ("cuda", "hostsim"):
        sbs = SideBySide(backend, with_oracle=False, flags=UPSCALE[2] | SSAA | SSAA_DITHER)
        try:
            digests = []
            assert not cf.run_rasterization(sbs, cf.SUITES[name], 3, digests=digests)
            assert sbs.gpu.upscaled_pass_count() > 0
            up, hidden = sbs.gpu.read_upscaled_rdram()
            results[backend] = (digests, hashlib.sha256(up.tobytes() + (hidden & 3).tobytes()).hexdigest())
        finally:
            sbs.close()
    assert results["cuda"] == results["hostsim"]


@pytest.mark.gpu
def test_gpu_fill_rectangle_resolves_exactly():
    native, oracle, up = _fill_rect_16("cuda", SSAA | SSAA_DITHER)
    if oracle is not None:
        assert np.array_equal(native, oracle)
