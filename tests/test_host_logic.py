"""Host-side logic that needs no GPU: the command ring (single commands, bulk stream records, wrap-around),
asynchronous scan-out fences, and pass summaries.  Runs on the test-only host simulator."""
import numpy as np
import pytest

from harness import rdp_builder as rb
from harness import workloads as wl
from oracle import angrylion
from sidebyside import HOSTSIM_LIB, build_hostsim, rdp

VI_REG = rdp.VI_REG


def _proc(single_threaded=False, monkeypatch=None):
    build_hostsim()
    if monkeypatch is not None:
        if single_threaded:
            monkeypatch.setenv("PARALLEL_RDP_SINGLE_THREADED_COMMAND", "1")
        else:
            monkeypatch.delenv("PARALLEL_RDP_SINGLE_THREADED_COMMAND", raising=False)
    rdram = np.zeros(wl.RDRAM_SIZE, dtype=np.uint8)
    tex, tlut = wl.c3_assets()
    rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
    rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
    p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, library=HOSTSIM_LIB)
    p.begin_read_hidden_rdram()[:] = 0
    p.end_write_hidden_rdram()
    p.end_write_rdram()
    return p


def _render(feed, monkeypatch, single_threaded=False, frames=3):
    p = _proc(single_threaded, monkeypatch)
    try:
        rng = np.random.RandomState(7)
        for _ in range(frames):
            feed(p, wl.c3_frame(rng, 2))
        p.idle()
        return p.begin_read_rdram().copy(), (p.begin_read_hidden_rdram() & 3).copy(), p.primitive_count()
    finally:
        p.close()


def _feed_stream(p, words):
    p.enqueue_command_stream(words)


_KEEP_ALIVE = []


def _feed_borrowed(p, words):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    _KEEP_ALIVE.append(w)  # borrowed: must outlive the processing
    p.enqueue_command_stream_borrowed(w)


def _feed_mixed(p, words):
    # Alternate copied single commands and borrowed runs inside one pass: staged and borrowed word segments interleave.
    cmds = rb.split_commands(words)
    i = 0
    while i < len(cmds):
        run = np.ascontiguousarray(np.concatenate(cmds[i:i + 3]), dtype=np.uint32)
        _KEEP_ALIVE.append(run)
        p.enqueue_command_stream_borrowed(run)
        i += 3
        if i < len(cmds):
            p.enqueue_command(cmds[i])
            i += 1


def _feed_per_command(p, words):
    for c in rb.split_commands(words):
        p.enqueue_command(c)


def test_ring_bulk_stream_equals_per_command_and_direct(monkeypatch):
    a = _render(_feed_stream, monkeypatch)
    b = _render(_feed_per_command, monkeypatch)
    c = _render(_feed_stream, monkeypatch, single_threaded=True)
    d = _render(_feed_borrowed, monkeypatch)
    e = _render(_feed_mixed, monkeypatch)
    f = _render(_feed_borrowed, monkeypatch, single_threaded=True)
    for other in (b, c, d, e, f):
        assert a[2] == other[2]
        assert np.array_equal(a[0], other[0]) and np.array_equal(a[1], other[1])


def test_ring_wraps_and_splits_large_streams(monkeypatch):
    """More command words than the ring holds (1 Mi words), in streams larger than one record (32 Ki words):
    exercises record splitting at command boundaries, the PAD/wrap path and producer back-pressure."""
    p = _proc(False, monkeypatch)
    try:
        b = rb.Builder()
        b.set_color_image(rb.FMT_RGBA, rb.SIZE_16, wl.COLOR_ADDR, 64)
        b.set_scissor(0, 0, 64, 64)
        b.om.__init__()
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.set_fill_color(0x12341234)
        head = b.words()
        # 100k fill rectangles + interleaved NOPs of odd total length so records never align with the ring size.
        rect = rb.Builder()
        rect.fill_rectangle(0, 0, 8, 8)
        one = rect.words()
        nop = np.zeros(2, dtype=np.uint32)
        body = np.concatenate([one, nop, one] * 35000)
        total = 0
        for _ in range(8):
            p.enqueue_command_stream(np.concatenate([head, body]))
            total += 70000
        p.idle()
        assert p.primitive_count() == total
    finally:
        p.close()


def _scanout_frames(mode, monkeypatch):
    p = _proc(False, monkeypatch)
    try:
        rng = np.random.RandomState(11)
        opts = p.default_scanout_options()
        opts.blend_previous_frame = 1
        opts.upscale_deinterlacing = 0
        frames, fences = [], []
        bufs = [np.zeros(640 * 576, dtype=np.uint32) for _ in range(3)]
        for i in range(3):
            p.enqueue_command_stream(wl.c3_frame(rng, 2))
            for k, v in wl.c3_vi_registers(i & 1).items():
                p.set_vi_register(VI_REG[k], v)
            if mode == "sync":
                pix, w, h = p.scanout_sync(opts)
                frames.append(pix.copy())
            else:
                fences.append(p.scanout_async(opts, bufs[i]))
        for i, ticket in enumerate(fences):
            w, h = p.scanout_wait(ticket)
            assert (w, h) == (640, 480)
            frames.append(bufs[i][: w * h].reshape(h, w).copy())
        return frames
    finally:
        p.close()


def test_scanout_async_matches_sync(monkeypatch):
    """Three frames in flight with deferred fence waits give the same images as the blocking call."""
    a = _scanout_frames("sync", monkeypatch)
    b = _scanout_frames("async", monkeypatch)
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.any() and np.array_equal(x, y)


def test_bulk_vi_register_write_equals_fourteen_single_writes():
    """rdp_b200_set_vi_registers (extension) against rdp_b200_set_vi_register: same scan-out."""
    import numpy as np
    from harness import workloads as wl
    from sidebyside import SideBySide, rdp

    images = []
    for bulk in (False, True):
        sbs = SideBySide("hostsim", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, with_oracle=False)
        try:
            sbs.randomize_rdram(np.random.RandomState(7))
            regs = wl.c2_vi_registers(1)
            if bulk:
                block = np.zeros(14, np.uint32)
                for k, v in regs.items():
                    block[rdp.VI_REG[k]] = v & 0xffffffff
                sbs.gpu.set_vi_registers(block)
            else:
                for k, v in regs.items():
                    sbs.gpu.set_vi_register(k, v)
            opts = sbs.gpu.default_scanout_options()
            opts.upscale_deinterlacing = 0
            pix, w, h = sbs.gpu.scanout_sync(opts)
            images.append(pix.copy())
        finally:
            sbs.close()
    assert images[0].size and np.array_equal(images[0], images[1])


def test_validation_callback_reports_hardware_lockups():
    """rdp_b200_set_validation_callback (ValidationInterface, rdp_renderer.hpp:75-95; report sites
    rdp_renderer.cpp:1425-1462, 3101-3107): Fill mode with depth compare / image read, Copy mode on a 32-bit target and a
    multi-line TLUT load are reported, a legal clear is not."""
    from harness.rdp_builder import Builder
    p = _proc()
    seen = []
    try:
        p.set_validation_callback(lambda err, msg: seen.append((err, msg)))
        b = Builder()
        b.set_color_image(0, 2, 0x100000, 320)
        b.set_scissor(0, 0, 320, 240)
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.flush_other_modes()
        b.set_fill_color(0x00010001)
        b.fill_rectangle(0, 0, 64, 16)
        b.sync_full()
        p.enqueue_command_stream(b.take())
        p.idle()
        assert seen == []

        b.om.z_compare = True
        b.om.image_read = True
        b.touch_other_modes()
        b.flush_other_modes()
        b.fill_rectangle(0, 0, 64, 16)
        b.sync_full()
        p.enqueue_command_stream(b.take())
        p.idle()
        assert {e for e, _ in seen} == {1, 3}  # FILL_DEPTH_TEST, FILL_IMAGE_READ_ENABLE
        assert all("Fill mode" in m for _, m in seen)

        del seen[:]
        b.om.z_compare = False
        b.om.image_read = False
        b.om.cycle_type = rb.CYCLE_COPY
        b.touch_other_modes()
        b.flush_other_modes()
        b.set_color_image(0, 3, 0x100000, 320)
        b.tex_rect(0, 0, 0, 32, 8, 0, 0, 4 << 10, 1 << 10)
        b.sync_full()
        p.enqueue_command_stream(b.take())
        p.idle()
        assert {e for e, _ in seen} == {4}  # COPY_32BPP

        del seen[:]
        b.set_texture_image(wl.TEX_ADDR + 0x1000, 0, 2, 16)
        b.set_tile(7, 0, 2, 8, 2048)
        b.load_tlut(7, 0, 0, 16, 2)  # two lines
        b.sync_full()
        p.enqueue_command_stream(b.take())
        p.idle()
        assert {e for e, _ in seen} == {6}  # INVALID_MULTILINE_LOAD_TLUT

        del seen[:]
        p.set_validation_callback(None)
        b.load_tlut(7, 0, 0, 16, 2)
        b.sync_full()
        p.enqueue_command_stream(b.take())
        p.idle()
        assert seen == []
    finally:
        p.close()


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_state_snapshots_are_interned_per_pass():
    """A stream that toggles between two states for every primitive stays one pass: snapshots are interned (the reference's
    StateCache, rdp_data_structures.hpp:256-362), not appended until the 1024-entry table forces a flush."""
    from harness.rdp_builder import Builder
    from sidebyside import SideBySide
    sbs = SideBySide("hostsim")
    try:
        sbs.clear_rdram()
        b = Builder()
        b.set_color_image(0, 2, 0x100000, 320)
        b.set_scissor(0, 0, 320, 240)
        b.om.cycle_type = rb.CYCLE_FILL
        b.touch_other_modes()
        b.flush_other_modes()
        rng = np.random.RandomState(5)
        for i in range(3000):
            b.set_fill_color(0xf801f801 if i & 1 else 0x07c107c1)
            x, y = int(rng.randint(0, 300)), int(rng.randint(0, 220))
            b.fill_rectangle(x, y, int(rng.randint(1, 20)), int(rng.randint(1, 20)))
        b.sync_full()
        before = sbs.gpu.pass_count()
        sbs.run(b.take())
        assert not sbs.compare_rdram()
        assert sbs.gpu.primitive_count() >= 3000
        assert sbs.gpu.pass_count() - before == 1
    finally:
        sbs.close()


def test_tile_row_reciprocal_is_exact_over_every_frame_shape():
    """FbInfo::tile_row_recip (rdp_types.h; filled in by the CUDA backend at launch): the tile kernels split a tile index
    into (row, column) as umulhi(tile, ceil(2^32 / tiles_x)).  Exact for every tile index below 2^32 / tiles_x -- checked
    here for every tiles_x a native or upscaled pass can have (upscaled width < 8192 pixels -> at most 1024 tile columns)
    at the indices where a rounded-up reciprocal fails first: the last tile of the largest image and the row boundaries."""
    for tiles_x in range(2, 1025):
        recip = 0xFFFFFFFF // tiles_x + 1
        rows = 8192 // 4  # tile rows of the tallest upscaled image (1024 lines at 8x, 4 lines per tile)
        r = np.arange(rows, dtype=np.uint64)
        for tile in (r * tiles_x, r * tiles_x + (tiles_x - 1)):  # first and last tile of every row
            got = (tile * np.uint64(recip)) >> np.uint64(32)
            assert np.array_equal(got, tile // np.uint64(tiles_x)), tiles_x
