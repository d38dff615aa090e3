"""Differential test harness: same command words + same initial memory into the Angrylion oracle and into
RDP::CommandProcessor (through the C-ABI), then byte-compare RDRAM, hidden RDRAM, TMEM and scan-out.

Mirrors the reference's SideBySideDriver (replayer_driver.cpp:138-144) and compare_rdram / compare_image
(conformance_utils.hpp:282-334).
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

rdp = importlib.import_module("parallel-rdp_b200")
from oracle import angrylion  # noqa: E402

HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, "_build", "librdp_hostsim.so")

RDRAM_SIZE = 4 << 20
HIDDEN_SIZE = 2 << 20


def build_hostsim():
    subprocess.run(["make", "-C", HOSTSIM_DIR, "-s"], check=True)
    return HOSTSIM_LIB


def gpu_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class SideBySide:
    """backend: 'hostsim' (CPU build of the device code, tests only) or 'cuda' (the product library)."""

    def __init__(self, backend: str, rdram_size: int = RDRAM_SIZE, hidden_size: int = HIDDEN_SIZE, with_oracle=True, flags: int = 0):
        self.rdram_size = rdram_size
        self.hidden_size = hidden_size
        self.oracle = angrylion.Oracle(rdram_size) if with_oracle else None
        self.rdram = np.zeros(rdram_size, dtype=np.uint8)
        lib = build_hostsim() if backend == "hostsim" else None
        self.gpu = rdp.CommandProcessor(self.rdram, hidden_size, flags=flags, library=lib)
        # The oracle initialises hidden RDRAM to 3 like the processor does; tests overwrite both anyway.
        self.frames_oracle = []
        self.frames_gpu = []

    def close(self):
        self.gpu.close()
        if self.oracle:
            self.oracle.close()

    # --- memory ---------------------------------------------------------------------------------------------
    def clear_rdram(self):
        self.gpu.idle()
        self.gpu.begin_read_rdram()[:] = 0
        self.gpu.end_write_rdram()
        self.gpu.begin_read_hidden_rdram()[:] = 0
        self.gpu.end_write_hidden_rdram()
        if self.oracle:
            self.oracle.rdram[:] = 0
            self.oracle.hidden[:] = 0

    def randomize_rdram(self, rng: np.random.RandomState):
        self.gpu.idle()
        data = rng.randint(0, 256, size=self.rdram_size, dtype=np.uint8)
        hid = rng.randint(0, 4, size=self.hidden_size, dtype=np.uint8)
        self.gpu.begin_read_rdram()[:] = data
        self.gpu.end_write_rdram()
        self.gpu.begin_read_hidden_rdram()[:] = hid
        self.gpu.end_write_hidden_rdram()
        if self.oracle:
            self.oracle.rdram[:] = data
            self.oracle.hidden[: self.hidden_size] = hid

    def write_rdram(self, offset: int, data: np.ndarray):
        """data: bytes in RDRAM storage order (host-endian 32-bit words of big-endian memory)."""
        self.gpu.idle()
        d = np.frombuffer(np.ascontiguousarray(data).tobytes(), dtype=np.uint8)
        self.gpu.begin_read_rdram()[offset:offset + d.size] = d
        self.gpu.end_write_rdram()
        if self.oracle:
            self.oracle.rdram[offset:offset + d.size] = d

    def write_tmem(self, data: np.ndarray):
        self.gpu.idle()
        d = np.frombuffer(np.ascontiguousarray(data).tobytes(), dtype=np.uint8)
        self.gpu.get_tmem()[:] = d
        if self.oracle:
            self.oracle.tmem[:] = d

    # --- commands -------------------------------------------------------------------------------------------
    def run(self, words: np.ndarray):
        if self.oracle:
            self.oracle.command_stream(words)
        self.gpu.enqueue_command_stream(words)

    def set_vi_register(self, reg: int, value: int):
        if self.oracle:
            self.oracle.set_vi_register(reg, value)
        self.gpu.set_vi_register(reg, value)

    def end_frame(self, opts=None):
        """Scan out on both sides; returns (oracle_pixels | None, gpu_pixels, w, h)."""
        ref = self.oracle.update_screen() if self.oracle else None
        if opts is None:
            opts = self.gpu.default_scanout_options()
            opts.blend_previous_frame = 1
            opts.upscale_deinterlacing = 0
        pix, w, h = self.gpu.scanout_sync(opts)
        return ref, pix.copy(), w, h

    # --- compare --------------------------------------------------------------------------------------------
    def digest(self) -> str:
        """SHA-256 over RDRAM and the two meaningful bits of hidden RDRAM (oracle side when present)."""
        import hashlib
        h = hashlib.sha256()
        if self.oracle:
            h.update(self.oracle.rdram.tobytes())
            h.update((self.oracle.hidden[: self.hidden_size] & 3).tobytes())
        else:
            self.gpu.idle()
            h.update(self.gpu.begin_read_rdram().tobytes())
            h.update((self.gpu.begin_read_hidden_rdram() & 3).tobytes())
        return h.hexdigest()

    def compare_rdram(self) -> list[str]:
        self.gpu.idle()
        errs = []
        a = self.gpu.begin_read_rdram()
        b = self.oracle.rdram
        if not np.array_equal(a, b):
            idx = np.nonzero(a != b)[0]
            i = int(idx[0])
            w = i & ~3
            errs.append(f"RDRAM differs at {len(idx)} bytes; first byte 0x{i:x}: "
                        f"gpu word 0x{int(a[w:w+4].view(np.uint32)[0]):08x} ref word 0x{int(b[w:w+4].view(np.uint32)[0]):08x}")
        ha = self.gpu.begin_read_hidden_rdram()
        hb = self.oracle.hidden[: self.hidden_size]
        if not np.array_equal(ha & 3, hb & 3):
            idx = np.nonzero((ha & 3) != (hb & 3))[0]
            i = int(idx[0])
            errs.append(f"hidden RDRAM differs at {len(idx)} bytes; first 0x{i:x}: gpu {int(ha[i])} ref {int(hb[i])}")
        return errs

    def compare_tmem(self) -> list[str]:
        self.gpu.idle()
        a = self.gpu.get_tmem().view(np.uint16)
        b = self.oracle.tmem.view(np.uint16)
        if np.array_equal(a, b):
            return []
        idx = np.nonzero(a != b)[0]
        i = int(idx[0])
        return [f"TMEM differs at {len(idx)} words; first slot 0x{i:x}: gpu 0x{int(a[i]):04x} ref 0x{int(b[i]):04x}"]
