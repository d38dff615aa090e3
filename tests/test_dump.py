"""RDPDUMP2 reader / writer: file-format interchange with the reference's own reader and writer
(oracle/_ref/libref_dump.so, built from /root/reference/rdp_dump.cpp and parallel-rdp/rdp_dump_write.cpp), malformed
input handling, and replay of a dump through RDP::CommandProcessor against the Angrylion oracle -- the equivalent of the
reference's rdp-validate-dump (rdp_validate_dump.cpp:84-198)."""
import os

import numpy as np
import pytest

import conformance as cf
from harness import workloads as wl
from harness.rdp_builder import split_commands
from oracle import angrylion, ref_dump
from sidebyside import HOSTSIM_LIB, SideBySide, build_hostsim, gpu_available, rdp

RDRAM, HIDDEN = 8 << 20, 4 << 20


def _script(rng):
    """A few frames' worth of dump records touching every record type: (method, args...)."""
    rdram = np.zeros(RDRAM, np.uint8)
    hidden = np.zeros(HIDDEN, np.uint8)
    ops = []
    for frame in range(3):
        # dirty a few scattered regions: inside one 4 KiB block, across a block boundary, a long run, the last block
        for off, n in ((100 + frame, 50), (4096 * 7 - 10, 20), (4096 * 100, 4096 * 5 + 17), (RDRAM - 33, 33)):
            rdram[off:off + n] = rng.randint(0, 256, n)
        hidden[4096 * 3 + frame: 4096 * 3 + frame + 9000] = rng.randint(0, 4, 9000)
        ops.append(("flush_dram", rdram.copy()))
        ops.append(("flush_hidden_dram", hidden.copy()))
        for _ in range(5):
            n = int(rng.choice([2, 4, 8, 44]))
            ops.append(("emit_command", int(rng.randint(0, 64)), rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)))
        ops.append(("emit_command", 0x29, np.zeros(0, np.uint32)))  # zero-length payload
        ops.append(("set_vi_register", int(rng.randint(0, 14)), int(rng.randint(0, 2 ** 32, dtype=np.uint64))))
        ops.append(("signal_complete",))
        ops.append(("end_frame",))
    return ops


def _expected_events(ops):
    ev = []
    for op in ops:
        if op[0] == "flush_dram":
            ev.append((7, 0, 0, op[1].tobytes()))
        elif op[0] == "flush_hidden_dram":
            ev.append((9, 0, 0, op[1].tobytes()))
        elif op[0] == "emit_command":
            ev.append((2, op[1], len(op[2]), op[2].tobytes()))
        elif op[0] == "set_vi_register":
            ev.append((3, op[1], op[2], b""))
        elif op[0] == "signal_complete":
            ev.append((5, 0, 0, b""))
        elif op[0] == "end_frame":
            ev.append((4, 0, 0, b""))
    ev.append((6, 0, 0, b""))
    return ev


def _write(writer, ops):
    for op in ops:
        getattr(writer, op[0])(*op[1:])
    writer.close()


def _read_ours(path, lib):
    player = rdp.DumpPlayer(path, library=lib)
    try:
        assert player.get_rdram_size() == RDRAM and player.get_hidden_rdram_size() == HIDDEN
        out = []
        while True:
            e = player.next_event()
            if e is None:
                break
            kind, a, b, payload = e
            if kind == rdp.DUMP_RDP_COMMAND:
                out.append((kind, a, len(payload), payload.tobytes()))
            else:
                out.append((kind, a if kind == rdp.DUMP_SET_VI_REGISTER else 0, b if kind == rdp.DUMP_SET_VI_REGISTER else 0,
                            payload.tobytes() if payload is not None else b""))
        return out
    finally:
        player.close()


@pytest.fixture(scope="module")
def lib():
    return build_hostsim()  # the dump code is host C++: identical in the product library and the host simulator


def test_round_trip_and_rewind(tmp_path, lib):
    ops = _script(np.random.RandomState(1))
    path = str(tmp_path / "ours.rdp")
    _write(rdp.DumpWriter(path, RDRAM, HIDDEN, library=lib), ops)
    assert open(path, "rb").read(8) == b"RDPDUMP2"
    assert _read_ours(path, lib) == _expected_events(ops)
    player = rdp.DumpPlayer(path, library=lib)
    first = player.next_event()
    for _ in range(7):
        player.next_event()
    assert player.rewind()
    again = player.next_event()
    assert first[0] == again[0] and np.array_equal(first[3], again[3])
    player.close()


@pytest.mark.skipif(not ref_dump.available(), reason="oracle/_ref/libref_dump.so not built")
def test_reference_reader_reads_our_files(tmp_path, lib):
    ops = _script(np.random.RandomState(2))
    path = str(tmp_path / "ours.rdp")
    _write(rdp.DumpWriter(path, RDRAM, HIDDEN, library=lib), ops)
    assert ref_dump.play(path) == _expected_events(ops)


@pytest.mark.skipif(not ref_dump.available(), reason="oracle/_ref/libref_dump.so not built")
def test_our_reader_reads_reference_files(tmp_path, lib):
    ops = _script(np.random.RandomState(3))
    path = str(tmp_path / "ref.rdp")
    _write(ref_dump.Writer(path, RDRAM, HIDDEN), ops)
    assert _read_ours(path, lib) == _expected_events(ops)
    # and the two writers agree on everything but how dirty blocks are grouped: ours is never larger
    ours = str(tmp_path / "ours.rdp")
    _write(rdp.DumpWriter(ours, RDRAM, HIDDEN, library=lib), ops)
    assert os.path.getsize(ours) <= os.path.getsize(path)


def test_malformed_files(tmp_path, lib):
    bad = tmp_path / "bad.rdp"
    bad.write_bytes(b"RDPDUMP1" + bytes(8))
    with pytest.raises(ValueError):
        rdp.DumpPlayer(str(bad), library=lib)
    bad.write_bytes(b"RDPDUMP2" + np.array([3 << 20, 4 << 20], np.uint32).tobytes())  # unsupported RDRAM size
    with pytest.raises(ValueError):
        rdp.DumpPlayer(str(bad), library=lib)
    with pytest.raises(ValueError):
        rdp.DumpPlayer(str(tmp_path / "missing.rdp"), library=lib)
    hdr = b"RDPDUMP2" + np.array([RDRAM, HIDDEN], np.uint32).tobytes()
    cases = {
        "truncated command": np.array([2, 0x36, 4, 1, 2], np.uint32).tobytes(),
        "patch beyond the image": np.array([1, RDRAM - 4, 8], np.uint32).tobytes() + bytes(8),
        "unknown record": np.array([77], np.uint32).tobytes(),
        "truncated register": np.array([3, 1], np.uint32).tobytes(),
    }
    for name, body in cases.items():
        bad.write_bytes(hdr + np.array([4], np.uint32).tobytes() + body)
        p = rdp.DumpPlayer(str(bad), library=lib)
        assert p.next_event()[0] == rdp.DUMP_END_FRAME, name
        assert p.next_event() is None, name
        p.close()


# --- replay: a dump of rendering work through the processor vs the oracle (rdp-validate-dump) ------------------------
def _make_render_dump(path, lib, frames=2):
    """C2-style frames (textured, shaded, Z-buffered triangles + clears + texture loads) as an RDPDUMP2 file."""
    w = rdp.DumpWriter(path, RDRAM, HIDDEN, library=lib)
    rdram = np.zeros(RDRAM, np.uint8)
    hidden = np.zeros(HIDDEN, np.uint8)
    rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
    rng = np.random.RandomState(2)
    for i in range(frames):
        w.flush_dram(rdram)
        w.flush_hidden_dram(hidden)
        for cmd in split_commands(wl.c2_frame(rng, 150)):
            w.emit_command((int(cmd[0]) >> 24) & 63, cmd)
        w.signal_complete()
        for k, v in wl.c2_vi_registers(i & 1).items():
            w.set_vi_register(rdp.VI_REG[k], v)
        w.end_frame()
        # the "CPU" scribbles over a texture between frames: the next flush carries only the dirty blocks
        rdram[wl.TEX_ADDR + 64: wl.TEX_ADDR + 64 + 512] = rng.randint(0, 256, 512)
    w.close()


from validate_dump import validate_dump as _validate_dump  # tools-style entry point, shared with tests/validate_dump.py


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_hostsim_replays_dump_like_the_oracle(tmp_path, lib):
    path = str(tmp_path / "render.rdp")
    _make_render_dump(path, lib)
    errs, checks = _validate_dump("hostsim", path, lib)
    assert not errs, errs
    assert checks == 4


@pytest.mark.gpu
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_gpu_replays_dump_like_the_oracle(tmp_path):
    path = str(tmp_path / "render.rdp")
    _make_render_dump(path, None, frames=3)
    errs, checks = _validate_dump("cuda", path, None)
    assert not errs, errs
    assert checks == 6


# --- PARALLEL_RDP_DUMP_PATH: the processor records what it is fed (ref rdp_device.cpp:60-76, 874-897, 970-975, 1087-1093) ---
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_dump_path_env_records_a_replayable_dump(tmp_path, lib, monkeypatch):
    path = str(tmp_path / "live.rdp")
    monkeypatch.setenv("PARALLEL_RDP_DUMP_PATH", path)
    rdram = np.zeros(RDRAM, np.uint8)
    rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
    p = rdp.CommandProcessor(rdram, HIDDEN, library=lib)
    monkeypatch.delenv("PARALLEL_RDP_DUMP_PATH")  # the replaying processors below must not dump
    try:
        p.begin_read_hidden_rdram()[:] = 0
        p.end_write_hidden_rdram()
        p.end_write_rdram()
        rng = np.random.RandomState(2)
        opts = p.default_scanout_options()
        for i, feed in enumerate(("single", "stream", "borrowed")):
            frame = wl.c2_frame(rng, 120)
            if feed == "single":
                for cmd in split_commands(frame):
                    p.enqueue_command(cmd)
            elif feed == "stream":
                p.enqueue_command_stream(frame)
            else:
                p.enqueue_command_stream_borrowed(frame)
            for k, v in wl.c2_vi_registers(i & 1).items():
                p.set_vi_register(k, v)
            p.scanout_sync(opts)
            # the "CPU" scribbles over a texture between frames
            p.begin_read_rdram()[wl.TEX_ADDR + 64: wl.TEX_ADDR + 64 + 512] = rng.randint(0, 256, 512)
            p.end_write_rdram()
        p.idle()
        live_rdram = p.begin_read_rdram().copy()
        live_hidden = (p.begin_read_hidden_rdram() & 3).copy()
    finally:
        p.close()  # closes the dump
    assert open(path, "rb").read(8) == b"RDPDUMP2"

    # the file replays against the oracle like any other dump: 3 x (SIGNAL_COMPLETE + END_FRAME)
    errs, checks = _validate_dump("hostsim", path, lib)
    assert not errs, errs
    assert checks == 6

    # and a replay ends where the live run ended
    player = rdp.DumpPlayer(path, library=lib)
    q = rdp.CommandProcessor(np.zeros(RDRAM, np.uint8), HIDDEN, library=lib)
    try:
        while True:
            n, stop = player.replay(q)
            if stop == rdp.DUMP_EOF:
                break
            if stop == rdp.DUMP_END_FRAME:
                q.scanout_sync(q.default_scanout_options())
        q.idle()
        # (the last scribble happened after the last recorded state, so compare outside it)
        a, b = q.begin_read_rdram().copy(), live_rdram.copy()
        a[wl.TEX_ADDR + 64: wl.TEX_ADDR + 64 + 512] = 0
        b[wl.TEX_ADDR + 64: wl.TEX_ADDR + 64 + 512] = 0
        assert np.array_equal(a, b)
        assert np.array_equal(q.begin_read_hidden_rdram() & 3, live_hidden)
    finally:
        player.close()
        q.close()
    if ref_dump.available():
        events = ref_dump.play(path)
        assert sum(1 for e in events if e[0] == rdp.DUMP_SIGNAL_COMPLETE) == 3
        assert sum(1 for e in events if e[0] == rdp.DUMP_END_FRAME) == 3
