"""ctypes front of harness/farm_driver.cpp: the native driver loops bench.py's end-to-end legs and the C5 stream farm
run in (only public C-ABI calls of include/rdp_b200.h; see the .cpp).  The driver library carries no link-time
dependency: farm_bind() resolves the entry points on the handle of whichever rdp library it is given (the product, or the
host simulator in the CPU tests)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "harness", "farm_driver.cpp")
LIB = os.path.join(ROOT, "harness", "_build", "libfarm_driver.so")

FARM_READ_BACK_RDRAM = 1
FARM_NO_SCANOUT_COPY = 2


class ScanoutOptions(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "crop_overscan_pixels", "crop_left", "crop_right", "crop_top", "crop_bottom", "crop_enable",
        "downscale_steps", "persist_frame_on_invalid_input", "blend_previous_frame", "upscale_deinterlacing",
        "vi_aa", "vi_scale", "vi_serrate", "vi_dither_filter", "vi_divot_filter", "vi_gamma_dither")]


class FarmStream(C.Structure):
    _fields_ = [("proc", C.c_void_p), ("frames", C.POINTER(C.c_void_p)), ("frame_words", C.POINTER(C.c_uint64)),
                ("num_frames", C.c_uint32), ("depth", C.c_uint32), ("vi_fields", C.c_void_p), ("outs", C.POINTER(C.c_void_p)),
                ("out_capacity", C.c_uint64), ("opts", ScanoutOptions), ("expect_w", C.c_uint32), ("expect_h", C.c_uint32)]


class FarmJob(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("rdram_image", C.c_void_p), ("frames", C.POINTER(C.c_void_p)),
                ("frame_words", C.POINTER(C.c_uint64)), ("num_frames", C.c_uint32), ("frames_to_run", C.c_uint32),
                ("vi_fields", C.c_void_p), ("expect_w", C.c_uint32), ("expect_h", C.c_uint32), ("downscale_steps", C.c_uint32)]


class FarmSlot(C.Structure):
    _fields_ = [("proc", C.c_void_p), ("kind", C.c_uint32), ("depth", C.c_uint32), ("outs", C.POINTER(C.c_void_p)),
                ("out_capacity", C.c_uint64), ("rdram", C.c_void_p), ("rdram_bytes", C.c_uint64)]


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(ROOT, "include", "rdp_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", LIB, SRC, "-I", os.path.join(ROOT, "include"),
                    "-lpthread", "-ldl"], check=True)
    return LIB


def load(rdp_library_path: str):
    """Loads the driver and binds it (dlsym on the library's own handle) to the rdp library at `rdp_library_path`."""
    lib = C.CDLL(build())
    lib.farm_bind.restype = C.c_int
    lib.farm_bind.argtypes = [C.c_char_p]
    if lib.farm_bind(rdp_library_path.encode()) != 0:
        raise RuntimeError("farm driver: cannot bind to " + rdp_library_path)
    lib.farm_run.restype = C.c_int
    lib.farm_run.argtypes = [C.POINTER(FarmStream), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.farm_run_pool.restype = C.c_int
    lib.farm_run_pool.argtypes = [C.POINTER(FarmSlot), C.c_uint32, C.POINTER(FarmJob), C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(C.c_uint64)]
    return lib


def pointer_array(arrays):
    """(void*[n], keep-alive) over numpy arrays."""
    arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return arr


def words_array(arrays):
    return (C.c_uint64 * len(arrays))(*[int(a.size) for a in arrays])


def vi_fields(reg_fn, vi_reg_index) -> np.ndarray:
    """uint32[2][14]: the registers of field 0 and field 1 (registers the workload does not program stay 0)."""
    out = np.zeros((2, 14), np.uint32)
    for f in (0, 1):
        for name, v in reg_fn(f).items():
            out[f, vi_reg_index[name]] = v & 0xffffffff
    return out


def copy_options(dst: ScanoutOptions, src) -> None:
    for name, _ in ScanoutOptions._fields_:
        setattr(dst, name, getattr(src, name))
