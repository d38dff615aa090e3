"""ctypes binding over the C-ABI of librdp_b200.so (include/rdp_b200.h).

This is plumbing for tests and bench.py; the product is the shared library.  The method names mirror
RDP::CommandProcessor of the reference (parallel-rdp/rdp_device.hpp:111-197) so that tests read like the
reference's own harness (replayer_driver_parallel.cpp).

There is no CPU fallback: if the CUDA library has not been built, or no CUDA device is usable, construction raises.
Tests that exercise the host logic without a GPU pass the test-only host simulator's path explicitly
(tests/hostsim); nothing in this package ever selects it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RDP_B200_LIBRARY: kernel experiments only (tools/build_variant.py builds alternative .so files next to the product).
PRODUCT_LIBRARY = os.environ.get("RDP_B200_LIBRARY") or os.path.join(_HERE, "librdp_b200.so")

FLAG_EXPLICIT_RDRAM_SYNC = 1 << 16  # RDP_B200_FLAG_EXPLICIT_RDRAM_SYNC (include/rdp_b200.h, "RDRAM coherency")

VI_REG = {name: i for i, name in enumerate(
    ["Control", "Origin", "Width", "Intr", "VCurrentLine", "Timing", "VSync", "HSync", "Leap", "HStart", "VStart",
     "VBurst", "XScale", "YScale"])}


class ScanoutOptions(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "crop_overscan_pixels", "crop_left", "crop_right", "crop_top", "crop_bottom", "crop_enable",
        "downscale_steps", "persist_frame_on_invalid_input", "blend_previous_frame", "upscale_deinterlacing",
        "vi_aa", "vi_scale", "vi_serrate", "vi_dither_filter", "vi_divot_filter", "vi_gamma_dither")]


def _bind(lib):
    P, U32, U64, SZ = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
    sig = {
        "rdp_b200_create": (P, [C.c_int, P, SZ, SZ, SZ, U32]),
        "rdp_b200_destroy": (None, [P]),
        "rdp_b200_device_is_supported": (C.c_int, [P]),
        "rdp_b200_enqueue_command": (None, [P, C.c_uint, P]),
        "rdp_b200_enqueue_command_direct": (None, [P, C.c_uint, P]),
        "rdp_b200_enqueue_command_stream": (None, [P, SZ, P]),
        "rdp_b200_enqueue_command_stream_borrowed": (None, [P, SZ, P]),
        "rdp_b200_flush": (None, [P]),
        "rdp_b200_signal_timeline": (U64, [P]),
        "rdp_b200_wait_for_timeline": (None, [P, U64]),
        "rdp_b200_idle": (None, [P]),
        "rdp_b200_begin_frame_context": (None, [P]),
        "rdp_b200_set_quirks": (None, [P, C.c_int, C.c_int]),
        "rdp_b200_set_validation_callback": (None, [P, P, P]),
        "rdp_b200_begin_read_rdram": (P, [P]),
        "rdp_b200_end_write_rdram": (None, [P]),
        "rdp_b200_mark_rdram_written": (None, [P, SZ, SZ]),
        "rdp_b200_rdram_mirror_stats": (None, [P, P]),
        "rdp_b200_set_command_thread_affinity": (C.c_int, [P, P, C.c_uint]),
        "rdp_b200_begin_read_hidden_rdram": (P, [P]),
        "rdp_b200_end_write_hidden_rdram": (None, [P]),
        "rdp_b200_get_rdram_size": (SZ, [P]),
        "rdp_b200_get_hidden_rdram_size": (SZ, [P]),
        "rdp_b200_get_tmem": (P, [P]),
        "rdp_b200_set_vi_register": (None, [P, U32, U32]),
        "rdp_b200_set_vi_registers": (None, [P, P]),
        "rdp_b200_begin_vi_register_per_scanline": (None, [P, U32]),
        "rdp_b200_set_vi_register_for_scanline": (None, [P, U32, U32]),
        "rdp_b200_latch_vi_register_for_scanline": (None, [P, C.c_uint]),
        "rdp_b200_end_vi_register_per_scanline": (None, [P]),
        "rdp_b200_scanout": (P, [P, P, P, P]),
        "rdp_b200_scanout_sync": (None, [P, P, P, SZ, P, P]),
        "rdp_b200_default_scanout_options": (None, [P]),
        "rdp_b200_scanout_async": (U64, [P, P, P, SZ, P, P]),
        "rdp_b200_scanout_wait": (None, [P, U64]),
        "rdp_b200_backend_name": (C.c_char_p, []),
        "rdp_b200_launch_count": (U64, [P]),
        "rdp_b200_primitive_count": (U64, [P]),
        "rdp_b200_pass_count": (U64, [P]),
        "rdp_b200_cuda_stream": (P, [P]),
        "rdp_b200_last_tile_variant": (C.c_char_p, [P]),
        "rdp_b200_host_timing": (None, [P, P]),
        "rdp_b200_upscaling_factor": (C.c_uint, [P]),
        "rdp_b200_upscaled_pass_count": (U64, [P]),
        "rdp_b200_read_upscaled_rdram": (C.c_int, [P, P, P]),
        "rdp_b200_dump_open": (P, [C.c_char_p]),
        "rdp_b200_dump_close": (None, [P]),
        "rdp_b200_dump_rdram_size": (SZ, [P]),
        "rdp_b200_dump_hidden_rdram_size": (SZ, [P]),
        "rdp_b200_dump_rewind": (C.c_int, [P]),
        "rdp_b200_dump_next": (C.c_int, [P, P]),
        "rdp_b200_dump_replay": (C.c_long, [P, P, P]),
        "rdp_b200_dump_writer_open": (P, [C.c_char_p, U32, U32]),
        "rdp_b200_dump_writer_flush_dram": (None, [P, P, U32]),
        "rdp_b200_dump_writer_flush_hidden_dram": (None, [P, P, U32]),
        "rdp_b200_dump_writer_command": (None, [P, U32, P, U32]),
        "rdp_b200_dump_writer_set_vi_register": (None, [P, U32, U32]),
        "rdp_b200_dump_writer_end_frame": (None, [P]),
        "rdp_b200_dump_writer_signal_complete": (None, [P]),
        "rdp_b200_dump_writer_close": (None, [P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


EXPORTED_SYMBOLS = None  # filled lazily from the table above for the symbol-presence test


def exported_symbol_names():
    import re
    header = os.path.join(os.path.dirname(_HERE), "include", "rdp_b200.h")
    text = open(header).read()
    return sorted(set(re.findall(r"\b(rdp_b200_[a-z_0-9]+)\s*\(", text)))


def load_library(path: str | None = None):
    path = path or PRODUCT_LIBRARY
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    return _bind(C.CDLL(path))


class CommandProcessor:
    """Mirror of RDP::CommandProcessor over the C-ABI.  `rdram` is a caller-owned numpy uint8 array (borrowed)."""

    def __init__(self, rdram: np.ndarray, hidden_rdram_size: int, flags: int = 0, cuda_device: int = 0,
                 library: str | None = None):
        assert rdram.dtype == np.uint8 and rdram.flags["C_CONTIGUOUS"]
        self.lib = load_library(library)
        self.rdram = rdram
        self._h = self.lib.rdp_b200_create(cuda_device, rdram.ctypes.data, 0, rdram.size, hidden_rdram_size, flags)
        if not self._h or not self.lib.rdp_b200_device_is_supported(self._h):
            raise RuntimeError("rdp_b200: device is not supported (no usable CUDA device / backend init failed)")
        self.hidden_size = hidden_rdram_size

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rdp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- commands -------------------------------------------------------------------------------------------
    def enqueue_command(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        self.lib.rdp_b200_enqueue_command(self._h, w.size, w.ctypes.data)

    def enqueue_command_stream(self, words: np.ndarray):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        self.lib.rdp_b200_enqueue_command_stream(self._h, w.size, w.ctypes.data)

    def enqueue_command_stream_borrowed(self, words: np.ndarray):
        """Zero-copy: `words` (uint32, contiguous, ideally pinned) must stay alive and unmodified until a later idle(),
        wait_for_timeline() or scan-out fence wait returns."""
        assert words.dtype == np.uint32 and words.flags["C_CONTIGUOUS"]
        self.lib.rdp_b200_enqueue_command_stream_borrowed(self._h, words.size, words.ctypes.data)

    def enqueue_command_stream_ptr(self, ptr: int, num_words: int):
        self.lib.rdp_b200_enqueue_command_stream(self._h, num_words, ptr)

    def flush(self):
        self.lib.rdp_b200_flush(self._h)

    def signal_timeline(self) -> int:
        return self.lib.rdp_b200_signal_timeline(self._h)

    def wait_for_timeline(self, value: int):
        self.lib.rdp_b200_wait_for_timeline(self._h, value)

    def idle(self):
        self.lib.rdp_b200_idle(self._h)

    def begin_frame_context(self):
        self.lib.rdp_b200_begin_frame_context(self._h)

    # --- memory ---------------------------------------------------------------------------------------------
    def begin_read_rdram(self) -> np.ndarray:
        self.lib.rdp_b200_begin_read_rdram(self._h)
        return self.rdram

    def end_write_rdram(self):
        self.lib.rdp_b200_end_write_rdram(self._h)

    def mark_rdram_written(self, offset: int, size: int):
        """The CPU wrote self.rdram[offset:offset + size] (extension; see include/rdp_b200.h)."""
        self.lib.rdp_b200_mark_rdram_written(self._h, offset, size)

    def rdram_mirror_stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        self.lib.rdp_b200_rdram_mirror_stats(self._h, out)
        return {"h2d_bytes": int(out[0]), "h2d_copies": int(out[1]), "d2h_bytes": int(out[2]), "d2h_copies": int(out[3])}

    def set_command_thread_affinity(self, cpus):
        arr = (C.c_int * len(cpus))(*[int(c) for c in cpus])
        return self.lib.rdp_b200_set_command_thread_affinity(self._h, arr, len(cpus))

    def begin_read_hidden_rdram(self) -> np.ndarray:
        p = self.lib.rdp_b200_begin_read_hidden_rdram(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.hidden_size,))

    def end_write_hidden_rdram(self):
        self.lib.rdp_b200_end_write_hidden_rdram(self._h)

    def get_tmem(self) -> np.ndarray:
        p = self.lib.rdp_b200_get_tmem(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(4096,))

    # --- VI -------------------------------------------------------------------------------------------------
    def set_vi_register(self, reg, value: int):
        self.lib.rdp_b200_set_vi_register(self._h, VI_REG[reg] if isinstance(reg, str) else int(reg), value & 0xffffffff)

    # Raster effects (ref rdp_device.hpp:186-196): HStart / XScale latched per VI line between begin and end.
    PER_SCANLINE_HSTART, PER_SCANLINE_XSCALE = 1, 2

    def begin_vi_register_per_scanline(self, flags: int = 3):
        self.lib.rdp_b200_begin_vi_register_per_scanline(self._h, flags)

    def set_vi_register_for_scanline(self, reg_bit: int, value: int):
        self.lib.rdp_b200_set_vi_register_for_scanline(self._h, reg_bit, value & 0xffffffff)

    def latch_vi_register_for_scanline(self, vi_line: int):
        self.lib.rdp_b200_latch_vi_register_for_scanline(self._h, vi_line)

    def end_vi_register_per_scanline(self):
        self.lib.rdp_b200_end_vi_register_per_scanline(self._h)

    def set_validation_callback(self, fn):
        """fn(error: int, message: str) for every condition that locks up real hardware (ValidationInterface,
        rdp_renderer.hpp:75-95); None removes it.  Called from the command thread."""
        if fn is None:
            self._validation_cb = None
            self.lib.rdp_b200_set_validation_callback(self._h, None, None)
            return
        proto = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p)
        self._validation_cb = proto(lambda user, err, msg: fn(int(err), msg.decode()))  # kept alive with the processor
        self.lib.rdp_b200_set_validation_callback(self._h, C.cast(self._validation_cb, C.c_void_p), None)

    def set_quirks(self, native_resolution_tex_rect: bool = True, native_texture_lod: bool = False):
        self.lib.rdp_b200_set_quirks(self._h, int(native_resolution_tex_rect), int(native_texture_lod))

    def set_vi_registers(self, values: np.ndarray):
        """All 14 registers at once (uint32[14], register-number order)."""
        assert values.dtype == np.uint32 and values.size == 14
        self.lib.rdp_b200_set_vi_registers(self._h, values.ctypes.data)

    def default_scanout_options(self) -> ScanoutOptions:
        o = ScanoutOptions()
        self.lib.rdp_b200_default_scanout_options(C.byref(o))
        return o

    def scanout_sync(self, opts: ScanoutOptions | None = None, out: np.ndarray | None = None):
        """Returns (pixels[h, w] uint32 RGBA8 little-endian bytes r,g,b,a, w, h)."""
        if out is None:
            out = np.empty(640 * 576 * self.upscaling_factor() ** 2, dtype=np.uint32)
        w, h = C.c_uint(0), C.c_uint(0)
        self.lib.rdp_b200_scanout_sync(self._h, C.byref(opts) if opts is not None else None, out.ctypes.data, out.size,
                                       C.byref(w), C.byref(h))
        n = w.value * h.value
        return out[:n].reshape(h.value, w.value) if n else out[:0].reshape(0, 0), w.value, h.value

    def scanout_async(self, opts: ScanoutOptions | None, out: np.ndarray):
        """Queues scan-out + asynchronous copy into `out` (uint32, ideally pinned) behind the commands enqueued so far.
        Returns a ticket for scanout_wait(), which returns (w, h)."""
        w, h = C.c_uint(0), C.c_uint(0)
        fence = self.lib.rdp_b200_scanout_async(self._h, C.byref(opts) if opts is not None else None, out.ctypes.data, out.size,
                                                C.byref(w), C.byref(h))
        return (fence, w, h, out, opts)  # the ticket keeps everything the library writes to alive

    def scanout_wait(self, ticket):
        fence, w, h = ticket[0], ticket[1], ticket[2]
        self.lib.rdp_b200_scanout_wait(self._h, fence)
        return w.value, h.value

    def scanout(self, opts: ScanoutOptions | None = None):
        """Device-pointer scan-out (no host copy).  Returns (device_ptr, w, h)."""
        w, h = C.c_uint(0), C.c_uint(0)
        p = self.lib.rdp_b200_scanout(self._h, C.byref(opts) if opts is not None else None, C.byref(w), C.byref(h))
        return p, w.value, h.value

    # --- introspection ----------------------------------------------------------------------------------------
    def backend_name(self) -> str:
        return self.lib.rdp_b200_backend_name().decode()

    def launch_count(self) -> int:
        return self.lib.rdp_b200_launch_count(self._h)

    def primitive_count(self) -> int:
        return self.lib.rdp_b200_primitive_count(self._h)

    def pass_count(self) -> int:
        return self.lib.rdp_b200_pass_count(self._h)

    def host_timing(self):
        out = (C.c_double * 4)()
        self.lib.rdp_b200_host_timing(self._h, out)
        return {"ring_busy_s": out[0], "submit_s": out[1], "scanout_s": out[2], "backpressure_s": out[3],
                "parse_s": out[0] - out[1] - out[2] - out[3]}

    def upscaling_factor(self) -> int:
        return self.lib.rdp_b200_upscaling_factor(self._h)

    def upscaled_pass_count(self) -> int:
        return self.lib.rdp_b200_upscaled_pass_count(self._h)

    def read_upscaled_rdram(self):
        """(rdram, hidden): uint8 arrays of shape (S*S, rdram_size) and (S*S, rdram_size // 2) -- one row per slice."""
        s = self.upscaling_factor()
        if s == 1:
            raise RuntimeError("internal upscaling is off")
        n = self.lib.rdp_b200_get_rdram_size(self._h)
        rdram = np.empty((s * s, n), np.uint8)
        hidden = np.empty((s * s, n // 2), np.uint8)
        if self.lib.rdp_b200_read_upscaled_rdram(self._h, rdram.ctypes.data, hidden.ctypes.data) != 0:
            raise RuntimeError("rdp_b200_read_upscaled_rdram failed")
        return rdram, hidden

    def last_tile_variant(self) -> str:
        return self.lib.rdp_b200_last_tile_variant(self._h).decode()

    def cuda_stream(self) -> int:
        return self.lib.rdp_b200_cuda_stream(self._h) or 0


# ---- RDPDUMP2 dumps (include/rdp_b200.h; reference: rdp_dump.cpp, parallel-rdp/rdp_dump_write.cpp) -------------------
DUMP_RDP_COMMAND, DUMP_SET_VI_REGISTER, DUMP_END_FRAME, DUMP_SIGNAL_COMPLETE, DUMP_EOF = 2, 3, 4, 5, 6
DUMP_UPDATE_RDRAM, DUMP_UPDATE_HIDDEN_RDRAM = 7, 9


class DumpEvent(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("arg0", C.c_uint32), ("arg1", C.c_uint32), ("num_words", C.c_uint32),
                ("words", C.POINTER(C.c_uint32)), ("data", C.POINTER(C.c_uint8)), ("size", C.c_size_t)]


class DumpPlayer:
    """Mirror of RDP::DumpPlayer: iterate over the records of an RDPDUMP2 file."""

    def __init__(self, path: str, library: str | None = None):
        self.lib = load_library(library)
        self._h = self.lib.rdp_b200_dump_open(os.fsencode(path))
        if not self._h:
            raise ValueError(f"{path}: not a readable RDPDUMP2 file")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rdp_b200_dump_close(self._h)
            self._h = None

    def get_rdram_size(self) -> int:
        return self.lib.rdp_b200_dump_rdram_size(self._h)

    def get_hidden_rdram_size(self) -> int:
        return self.lib.rdp_b200_dump_hidden_rdram_size(self._h)

    def rewind(self) -> bool:
        return bool(self.lib.rdp_b200_dump_rewind(self._h))

    def next_event(self):
        """(kind, arg0, arg1, payload) or None at the end; payload is a numpy copy of the words / image."""
        ev = DumpEvent()
        if not self.lib.rdp_b200_dump_next(self._h, C.byref(ev)):
            return None
        payload = None
        if ev.kind == DUMP_RDP_COMMAND:
            payload = np.ctypeslib.as_array(ev.words, shape=(ev.num_words,)).copy() if ev.num_words else np.zeros(0, np.uint32)
        elif ev.kind in (DUMP_UPDATE_RDRAM, DUMP_UPDATE_HIDDEN_RDRAM):
            payload = np.ctypeslib.as_array(ev.data, shape=(ev.size,)).copy()
        return ev.kind, ev.arg0, ev.arg1, payload

    def replay(self, processor: "CommandProcessor"):
        """Feed records into `processor` up to the next END_FRAME / SIGNAL_COMPLETE / EOF: (commands, stop_kind)."""
        ev = DumpEvent()
        n = self.lib.rdp_b200_dump_replay(self._h, processor._h, C.byref(ev))
        return n, (ev.kind if n >= 0 else None)


class DumpWriter:
    """Mirror of RDP::RDPDumpWriter."""

    def __init__(self, path: str, rdram_size: int, hidden_rdram_size: int, library: str | None = None):
        self.lib = load_library(library)
        self._h = self.lib.rdp_b200_dump_writer_open(os.fsencode(path), rdram_size, hidden_rdram_size)
        if not self._h:
            raise OSError(f"cannot create {path}")

    def flush_dram(self, rdram: np.ndarray):
        self.lib.rdp_b200_dump_writer_flush_dram(self._h, rdram.ctypes.data, rdram.size)

    def flush_hidden_dram(self, hidden: np.ndarray):
        self.lib.rdp_b200_dump_writer_flush_hidden_dram(self._h, hidden.ctypes.data, hidden.size)

    def emit_command(self, op: int, words: np.ndarray):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        self.lib.rdp_b200_dump_writer_command(self._h, op, words.ctypes.data, words.size)

    def set_vi_register(self, reg: int, value: int):
        self.lib.rdp_b200_dump_writer_set_vi_register(self._h, reg, value)

    def end_frame(self):
        self.lib.rdp_b200_dump_writer_end_frame(self._h)

    def signal_complete(self):
        self.lib.rdp_b200_dump_writer_signal_complete(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rdp_b200_dump_writer_close(self._h)
            self._h = None
