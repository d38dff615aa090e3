"""oracle/angrylion.py -- TEST INFRASTRUCTURE ONLY: ctypes driver for oracle/_ref/libangrylion_ref.so.

The shared object is the reference's own CPU renderer (Angrylion-Plus as bundled by paraLLEl-RDP) compiled
unmodified from /root/reference by oracle/Makefile plus our shim (oracle/ref_shim.c).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this module.

Angrylion keeps all state in process globals (angrylion rdp.c:345, rdram.c:34), so every Oracle instance loads a
private copy of the library to stay independent of other instances in the same process.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIBRARY = os.path.join(_HERE, "_ref", "libangrylion_ref.so")


def available() -> bool:
    return os.path.exists(REF_LIBRARY)


class Oracle:
    def __init__(self, rdram_size: int = 8 << 20, parallel: bool = False, num_workers: int = 0):
        if not available():
            raise RuntimeError(f"{REF_LIBRARY} missing: run `make -C oracle` where /root/reference exists")
        # Private copy => private globals.
        self._tmpdir = tempfile.mkdtemp(prefix="angrylion_ref_")
        private = os.path.join(self._tmpdir, "libangrylion_ref.so")
        shutil.copy(REF_LIBRARY, private)
        lib = C.CDLL(private)
        lib.ref_init.argtypes = [C.c_uint32, C.c_int, C.c_uint32]
        lib.ref_init.restype = C.c_int
        lib.ref_rdram.restype = C.c_void_p
        lib.ref_hidden_rdram.restype = C.c_void_p
        lib.ref_hidden_rdram_size.restype = C.c_uint32
        lib.ref_tmem.restype = C.c_void_p
        lib.ref_command.argtypes = [C.c_void_p]
        lib.ref_command_stream.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_command_stream.restype = C.c_uint32
        lib.ref_set_vi_register.argtypes = [C.c_uint32, C.c_uint32]
        lib.ref_update_screen.restype = C.c_int
        lib.ref_frame.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_frame.restype = C.c_void_p
        lib.ref_set_verbose.argtypes = [C.c_int]
        self.lib = lib
        if lib.ref_init(rdram_size, int(parallel), num_workers) != 0:
            raise RuntimeError("ref_init failed")
        self.rdram_size = rdram_size
        self.rdram = np.ctypeslib.as_array(C.cast(lib.ref_rdram(), C.POINTER(C.c_uint8)), shape=(rdram_size,))
        hs = lib.ref_hidden_rdram_size()
        self.hidden = np.ctypeslib.as_array(C.cast(lib.ref_hidden_rdram(), C.POINTER(C.c_uint8)), shape=(hs,))
        self.tmem = np.ctypeslib.as_array(C.cast(lib.ref_tmem(), C.POINTER(C.c_uint8)), shape=(4096,))

    def close(self):
        if getattr(self, "lib", None) is not None:
            self.lib.ref_close()
            self.lib = None
            shutil.rmtree(self._tmpdir, ignore_errors=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def command(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        pad = np.zeros(44, dtype=np.uint32)
        pad[: w.size] = w
        self.lib.ref_command(pad.ctypes.data)

    def command_stream(self, words: np.ndarray) -> int:
        w = np.ascontiguousarray(words, dtype=np.uint32)
        return self.lib.ref_command_stream(w.ctypes.data, w.size)

    def process_list(self, words: np.ndarray, list_addr: int) -> None:
        """Run a packed stream as an in-RDRAM display list (engages the oracle's worker pool when parallel=True)."""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        self.lib.ref_process_list.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        self.lib.ref_process_list.restype = C.c_int
        if self.lib.ref_process_list(w.ctypes.data, w.size, list_addr) != 0:
            raise RuntimeError("ref_process_list: display list does not fit in RDRAM")

    def set_vi_register(self, reg: int, value: int):
        self.lib.ref_set_vi_register(reg, value & 0xffffffff)

    def update_screen(self):
        """Runs the VI.  Returns pixels[h, w] uint32 (bytes r,g,b,a) or None when nothing was presented."""
        if not self.lib.ref_update_screen():
            return None
        w, h = C.c_uint32(0), C.c_uint32(0)
        p = self.lib.ref_frame(C.byref(w), C.byref(h))
        if not p or w.value == 0 or h.value == 0:
            return None
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(h.value, w.value))
        return arr.copy()
