"""Host logic added in round 2, on the test-only host simulator (no GPU):

  * RDRAM mirror: page-granular host -> device uploads (1 KiB pages + exact partial pages), exact-range device -> host
    read-back, CPU data lying between the colour and the depth image survives a read-back;
  * command ring: idle workers cost no CPU time (the 8-GPU cliff of SCALE_r01), the idle kick still flushes;
  * meta operations are a private ring record kind: opcodes 1..7 in a guest stream are two-word no-ops, whatever
    their payload (opcode 5 used to be dereferenced as a pointer);
  * scan-out fences survive an interleaved idle() (one timeline queue for two streams).
"""
import os
import time

import numpy as np
import pytest

from harness import rdp_builder as rb
from harness import workloads as wl
from oracle import angrylion
from sidebyside import HOSTSIM_LIB, build_hostsim, rdp

VI_REG = rdp.VI_REG
COLOR, DEPTH, BETWEEN = 0x100000, 0x140000, 0x130000  # 320x240 RGBA16 = 0x25800 bytes each; BETWEEN lies in the gap


# Every test runs on the host simulator (CPU suite) and, marked gpu, through librdp_b200.so on the device.
BACKENDS = [pytest.param(HOSTSIM_LIB, id="hostsim"), pytest.param(None, id="cuda", marks=pytest.mark.gpu)]


def _proc(library=HOSTSIM_LIB, explicit=True):
    """explicit: RDP_B200_FLAG_EXPLICIT_RDRAM_SYNC -- CPU writes are declared, nothing else crosses the bus (what the byte
    counts below are about); False: the default, the reference's imported-host-memory behaviour."""
    if library is not None:
        build_hostsim()
    rdram = np.zeros(wl.RDRAM_SIZE, dtype=np.uint8)
    p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=rdp.FLAG_EXPLICIT_RDRAM_SYNC if explicit else 0, library=library)
    p.begin_read_hidden_rdram()[:] = 0
    p.end_write_hidden_rdram()
    p.end_write_rdram()
    p.idle()
    return p


def _fill_pass(color, x0, y0, x1, y1, addr=COLOR, width=320):
    b = rb.Builder()
    b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, addr, width)
    b.set_scissor(0, 0, width, 240)
    b.om.__init__()
    b.om.cycle_type = rb.CYCLE_FILL
    b.touch_other_modes()
    b.set_fill_color(color)
    b.fill_rectangle(x0, y0, x1 - x0, y1 - y0)
    return b.words()


@pytest.mark.parametrize("library", BACKENDS)
def test_one_poked_pixel_uploads_at_most_one_page(library):
    p = _proc(library)
    try:
        p.enqueue_command_stream(_fill_pass(0x11111111, 0, 0, 320, 240))
        p.idle()
        base = p.rdram_mirror_stats()
        # The CPU pokes one 16-bit pixel and says so; the next pass must not cost more than one 1 KiB page of upload.
        p.rdram[COLOR + 2000:COLOR + 2002] = (0xAB, 0xCD)
        p.mark_rdram_written(COLOR + 2000, 2)
        p.enqueue_command_stream(_fill_pass(0x22222222, 100, 100, 110, 110))
        p.idle()
        st = p.rdram_mirror_stats()
        assert 0 < st["h2d_bytes"] - base["h2d_bytes"] <= 1024
        assert st["h2d_copies"] - base["h2d_copies"] == 1
        # ... and the poke survived the round trip (device RDRAM took it, the read-back returned it).
        assert tuple(p.rdram[COLOR + 2000:COLOR + 2002]) == (0xAB, 0xCD)
        # Read-back is the rendered ranges only: here one colour image, no depth writes.
        assert st["d2h_bytes"] - base["d2h_bytes"] <= 320 * 240 * 2
    finally:
        p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_page_bitmap_and_exact_edges(library):
    p = _proc(library)
    try:
        base = p.rdram_mirror_stats()
        # 5000 bytes starting 100 bytes into a page: a 924-byte head, 3 whole pages, a 1004-byte tail -> one merged run.
        lo = BETWEEN + 100
        p.rdram[lo:lo + 5000] = 0x5A
        p.mark_rdram_written(lo, 5000)
        # A second, disjoint 3-byte write in another page.
        p.rdram[BETWEEN + 0x4000 + 7:BETWEEN + 0x4000 + 10] = 0x77
        p.mark_rdram_written(BETWEEN + 0x4000 + 7, 3)
        p.enqueue_command_stream(_fill_pass(0x33333333, 0, 0, 8, 8))
        p.idle()
        st = p.rdram_mirror_stats()
        assert st["h2d_bytes"] - base["h2d_bytes"] == 5003
        assert st["h2d_copies"] - base["h2d_copies"] == 2
    finally:
        p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_cpu_data_between_colour_and_depth_survives_readback_mid_list(library):
    """the old mirror read back one bounding box from the colour image to the depth image and uploaded all
    8 MiB afterwards, so CPU data written between the two (textures, display lists) was overwritten with stale device
    bytes.  Here the CPU writes a texture there mid-list, without idle(), while both images have pending GPU writes."""
    p = _proc(library)
    o = angrylion.Oracle(wl.RDRAM_SIZE)
    try:
        o.rdram[:] = 0
        o.hidden[:] = 0
        b = rb.Builder()
        b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, COLOR, 320)
        b.set_depth_image(DEPTH)
        b.set_scissor(0, 0, 320, 240)
        b.om.__init__()
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.set_fill_color(0x10011001)
        b.fill_rectangle(0, 0, 320, 240)
        # Z "clear" through the colour image path, the usual idiom: both images now have pending GPU writes.
        b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, DEPTH, 320)
        b.set_fill_color(0xfffcfffc)
        b.fill_rectangle(0, 0, 320, 240)
        first = b.words()
        p.enqueue_command_stream(first)
        o.command_stream(first)

        # CPU writes a 32x32 RGBA16 texture between the images -- no idle(), no end_write_rdram().
        tex = np.random.RandomState(5).randint(0, 256, 2048).astype(np.uint8)
        p.rdram[BETWEEN:BETWEEN + 2048] = tex
        p.mark_rdram_written(BETWEEN, 2048)
        o.rdram[BETWEEN:BETWEEN + 2048] = tex

        b = rb.Builder()
        b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, COLOR, 320)
        b.set_texture_image(BETWEEN, rb.FMT_RGBA, rb.SIZE_16, 32)
        b.set_tile(0, rb.FMT_RGBA, rb.SIZE_16, 64, 0, mask_s=5, mask_t=5)
        b.load_tile(0, 0, 0, 32, 32)
        b.set_tile_size(0, 0, 0, 32, 32)
        b.om.__init__()
        b.om.cycle_type = rb.CYCLE_COPY
        b.touch_other_modes()
        b.tex_rect(0, 16 << 2, 16 << 2, 32 << 2, 32 << 2, 0, 0, 4 << 10, 1 << 10)
        second = b.words()
        p.enqueue_command_stream(second)
        o.command_stream(second)
        p.idle()

        assert np.array_equal(p.rdram[BETWEEN:BETWEEN + 2048], tex), "CPU data between the images was clobbered"
        assert np.array_equal(p.begin_read_rdram(), o.rdram), "RDRAM differs from the oracle"
        st = p.rdram_mirror_stats()
        # Read-back never exceeds the two images (plus nothing in between).
        assert st["d2h_bytes"] <= 3 * 2 * 320 * 240 * 2
    finally:
        p.close()
        o.close()


def test_idle_processors_cost_no_cpu():
    """SCALE_r01: 16 idle processors per rank used to yield-spin (4 000 yields every 500 us each).  Idle command
    threads must sleep: process CPU time over half a second of idleness stays far below one core."""
    procs = [_proc() for _ in range(8)]
    try:
        time.sleep(0.05)  # past the one idle kick
        t0, c0 = time.perf_counter(), time.process_time()
        time.sleep(0.5)
        wall, cpu = time.perf_counter() - t0, time.process_time() - c0
        assert cpu < 0.05 * wall, f"idle command threads burned {cpu:.3f} s of CPU in {wall:.3f} s"
    finally:
        for p in procs:
            p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_idle_kick_flushes_batched_work(library):
    """The reference's MetaIdle (command_ring.cpp:89-134): 500 us after the last command the renderer submits what it
    has batched, without any flush from the caller."""
    p = _proc(library)
    try:
        before = p.pass_count()
        words = _fill_pass(0x44444444, 0, 0, 16, 16)
        p.enqueue_command_stream(words)  # no SyncFull, no flush
        deadline = time.perf_counter() + 2.0
        while p.pass_count() == before and time.perf_counter() < deadline:
            time.sleep(0.005)
        assert p.pass_count() == before + 1
    finally:
        p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_guest_opcodes_1_to_7_are_two_word_nops(library):
    p = _proc(library)
    try:
        # Opcode 5 with a wild "pointer" payload used to be executed as a scan-out request; every opcode below 8 must be
        # skipped as a two-word no-op, keeping the stream in sync with the commands that follow.
        junk = []
        for op in range(1, 8):
            junk += [(op << 24) | 0x00dead, 0xdeadbeef]
        words = np.concatenate([np.array(junk, dtype=np.uint32), _fill_pass(0x55555555, 0, 0, 4, 4),
                                np.array(junk, dtype=np.uint32), _fill_pass(0x66666666, 4, 0, 8, 4)])
        p.enqueue_command_stream(words)
        for c in rb.split_commands(words):
            p.enqueue_command(c)
        p.idle()
        assert p.primitive_count() == 4
        row = p.begin_read_rdram()[COLOR:COLOR + 16]
        assert row[:8].tolist() == [0x55] * 8 and row[8:16].tolist() == [0x66] * 8
    finally:
        p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_scanout_fence_survives_interleaved_idle(library):
    p = _proc(library)
    try:
        rdram = p.rdram
        tex, tlut = wl.c3_assets()
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
        rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
        p.end_write_rdram()
        opts = p.default_scanout_options()
        rng = np.random.RandomState(3)
        p.enqueue_command_stream(wl.c3_frame(rng, 2))
        for k, v in wl.c3_vi_registers(0).items():
            p.set_vi_register(VI_REG[k], v)
        buf = np.zeros(640 * 576, dtype=np.uint32)
        ticket = p.scanout_async(opts, buf)
        p.idle()  # a timeline point on the compute stream, waited for before the scan-out fence
        w, h = p.scanout_wait(ticket)
        assert w == 640 and h in (240, 480) and buf[: w * h].any()
    finally:
        p.close()


@pytest.mark.parametrize("library", BACKENDS)
def test_default_mode_sees_undeclared_cpu_writes(library):
    """The reference's contract (integration_example.cpp:490, rdp_conformance.cpp:609-665): the emulator writes RDRAM through
    the pointer it handed over and calls nothing.  By default a texture load's source, a pass's colour / depth image and
    the scan-out's framebuffer are uploaded before the GPU reads them, and rendered memory comes back by itself."""
    p = _proc(library, explicit=False)
    o = angrylion.Oracle(wl.RDRAM_SIZE)
    try:
        o.rdram[:] = 0
        o.hidden[:] = 0
        rng = np.random.RandomState(9)
        for it in range(3):
            # undeclared: a texture, and a CPU-drawn background in the colour image
            tex = rng.randint(0, 256, 2048).astype(np.uint8)
            back = rng.randint(0, 256, 320 * 8 * 2).astype(np.uint8)
            for mem in (p.rdram, o.rdram):
                mem[BETWEEN:BETWEEN + 2048] = tex
                mem[COLOR + 320 * 2 * 100:COLOR + 320 * 2 * 108] = back
            b = rb.Builder()
            b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, COLOR, 320)
            b.set_depth_image(DEPTH)
            b.set_scissor(0, 0, 320, 240)
            b.set_texture_image(BETWEEN, rb.FMT_RGBA, rb.SIZE_16, 32)
            b.set_tile(0, rb.FMT_RGBA, rb.SIZE_16, 64, 0, mask_s=5, mask_t=5)
            b.load_tile(0, 0, 0, 32, 32)
            b.set_tile_size(0, 0, 0, 32, 32)
            b.om.__init__()
            b.om.cycle_type = rb.CYCLE_COPY
            b.touch_other_modes()
            b.tex_rect(0, (16 + 40 * it) << 2, 16 << 2, 32 << 2, 32 << 2, 0, 0, 4 << 10, 1 << 10)
            b.sync_full()
            words = b.words()
            p.enqueue_command_stream(words)
            o.command_stream(words)
            # no idle(): wait for the timeline the way an emulator does after SyncFull
            p.wait_for_timeline(p.signal_timeline())
            assert np.array_equal(p.rdram, o.rdram), f"iteration {it}: RDRAM differs from the oracle"
        st = p.rdram_mirror_stats()
        assert st["h2d_bytes"] > 0 and st["d2h_bytes"] > 0
    finally:
        p.close()
        o.close()
