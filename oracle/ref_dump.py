"""ctypes driver for oracle/_ref/libref_dump.so: the reference's own RDPDUMP2 reader (rdp_dump.cpp) and writer
(parallel-rdp/rdp_dump_write.cpp), compiled from /root/reference by oracle/Makefile behind ref_dump_shim.cpp.
TEST INFRASTRUCTURE ONLY (format interchange checks of the product's reader / writer)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_dump.so")
_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t)


def available() -> bool:
    return os.path.exists(_LIB)


def _lib():
    lib = C.CDLL(_LIB)
    lib.refdump_play.restype = C.c_long
    lib.refdump_play.argtypes = [C.c_char_p, _CB, C.c_void_p]
    lib.refdump_writer_open.restype = C.c_void_p
    lib.refdump_writer_open.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
    for name, args in {"flush_dram": [C.c_void_p, C.c_void_p, C.c_uint32], "flush_hidden_dram": [C.c_void_p, C.c_void_p, C.c_uint32],
                       "command": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32], "set_vi_register": [C.c_void_p, C.c_uint32, C.c_uint32],
                       "end_frame": [C.c_void_p], "signal_complete": [C.c_void_p], "close": [C.c_void_p]}.items():
        fn = getattr(lib, "refdump_writer_" + name)
        fn.restype = None
        fn.argtypes = args
    return lib


def play(path: str):
    """All events the reference DumpPlayer delivers for `path`: list of (kind, a, b, payload bytes), or None if it
    refuses the file.  Kinds use the file's command ids (2 command, 3 VI register, 4 end frame, 5 signal complete,
    6 eof, 7 RDRAM image, 9 hidden RDRAM image)."""
    events = []

    def cb(_user, kind, a, b, data, size):
        payload = C.string_at(data, size) if size else b""
        events.append((int(kind), int(a), int(b), payload))

    n = _lib().refdump_play(os.fsencode(path), _CB(cb), None)
    return None if n < 0 else events


class Writer:
    def __init__(self, path: str, rdram_size: int, hidden_size: int):
        self.lib = _lib()
        self.h = self.lib.refdump_writer_open(os.fsencode(path), rdram_size, hidden_size)
        if not self.h:
            raise OSError(path)

    def flush_dram(self, a: np.ndarray):
        self.lib.refdump_writer_flush_dram(self.h, a.ctypes.data, a.size)

    def flush_hidden_dram(self, a: np.ndarray):
        self.lib.refdump_writer_flush_hidden_dram(self.h, a.ctypes.data, a.size)

    def emit_command(self, op: int, words: np.ndarray):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        self.lib.refdump_writer_command(self.h, op, w.ctypes.data, w.size)

    def set_vi_register(self, reg: int, value: int):
        self.lib.refdump_writer_set_vi_register(self.h, reg, value)

    def end_frame(self):
        self.lib.refdump_writer_end_frame(self.h)

    def signal_complete(self):
        self.lib.refdump_writer_signal_complete(self.h)

    def close(self):
        if self.h:
            self.lib.refdump_writer_close(self.h)
            self.h = None
