"""Parity of the BASELINE.json benchmark workloads (C2 triangles, C3 fill-rate) -- RDRAM, hidden RDRAM and scan-out
against the Angrylion oracle, and specialised tile kernels against the generic uber-kernel."""
import os

import numpy as np
import pytest

from harness import workloads as wl
from oracle import angrylion
from sidebyside import HOSTSIM_LIB, SideBySide, build_hostsim, rdp

VI_REG = rdp.VI_REG


def _setup(sbs, which):
    sbs.clear_rdram()
    if which == "c2":
        sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
    else:
        tex, tlut = wl.c3_assets()
        sbs.write_rdram(wl.TEX_ADDR, tex)
        sbs.write_rdram(wl.TEX_ADDR + 0x1000, tlut)


def _frames(which, count, size):
    rng = np.random.RandomState(2 if which == "c2" else 3)
    return [wl.c2_frame(rng, size) if which == "c2" else wl.c3_frame(rng, size) for _ in range(count)]


def _run_side_by_side(backend, which, count, size):
    sbs = SideBySide(backend, rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE)
    errs = []
    variants = set()
    try:
        _setup(sbs, which)
        regs = wl.c2_vi_registers if which == "c2" else wl.c3_vi_registers
        for i, words in enumerate(_frames(which, count, size)):
            sbs.run(words)
            sbs.gpu.idle()
            variants.add(sbs.gpu.last_tile_variant())
            for k, v in regs(i & 1).items():
                sbs.set_vi_register(VI_REG[k], v)
            ref, pix, w, h = sbs.end_frame()
            errs += sbs.compare_rdram()
            if ref is None or ref.shape != pix.shape or not np.array_equal(ref & 0xFFFFFF, pix & 0xFFFFFF):
                errs.append(f"{which} frame {i}: scan-out differs")
            if errs:
                break
    finally:
        sbs.close()
    return errs, variants


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,size", [("c2", 120), ("c3", 3)])
def test_hostsim_workload_matches_oracle(which, size):
    build_hostsim()
    errs, _ = _run_side_by_side("hostsim", which, 2, size)
    assert not errs, errs


@pytest.mark.gpu
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,size,variant", [("c2", 5000, "zb-opa-tex-modulate-2c/rgba32+deferred"), ("c3", 64, "xlu-sprite-ci8-tlut-1c/rgba16")])
def test_gpu_workload_matches_oracle_full_size(which, size, variant):
    """BASELINE.json's full frame sizes, 3 frames, byte-compared against the oracle; also pins which instantiation ran."""
    errs, variants = _run_side_by_side("cuda", which, 3, size)
    assert not errs, errs
    assert variants == {variant}


def _render_digest(which, count, size, force_generic, env=None):
    import hashlib
    old = os.environ.pop("RDP_B200_FORCE_GENERIC", None)
    if force_generic:
        os.environ["RDP_B200_FORCE_GENERIC"] = "1"
    for k in env or ():
        os.environ[k] = "1"
    try:
        sbs = SideBySide("cuda", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE, with_oracle=False)
        try:
            _setup(sbs, which)
            h = hashlib.sha256()
            names = set()
            for words in _frames(which, count, size):
                sbs.gpu.enqueue_command_stream(words)
                sbs.gpu.idle()
                names.add(sbs.gpu.last_tile_variant())
                h.update(sbs.gpu.begin_read_rdram().tobytes())
                h.update((sbs.gpu.begin_read_hidden_rdram() & 3).tobytes())
            return h.hexdigest(), names
        finally:
            sbs.close()
    finally:
        os.environ.pop("RDP_B200_FORCE_GENERIC", None)
        for k in env or ():
            os.environ.pop(k, None)
        if old is not None:
            os.environ["RDP_B200_FORCE_GENERIC"] = old


@pytest.mark.gpu
@pytest.mark.parametrize("which,size", [("c2", 5000), ("c3", 64)])
def test_gpu_specialised_equals_generic(which, size):
    """Size-independent property: the specialised instantiation and the uber-kernel produce identical memory."""
    a, names_a = _render_digest(which, 4, size, force_generic=False)
    b, names_b = _render_digest(which, 4, size, force_generic=True)
    assert names_b == {"generic"}
    assert "generic" not in names_a
    assert a == b


@pytest.mark.gpu
def test_gpu_deferred_equals_in_order():
    """The deferred tile pipeline (coverage/Z -> ordered depth -> compacted shading -> ordered blend) against the
    in-order kernels on the full C2 frame: specialised deferred, specialised in-order, dynamic deferred."""
    a, names_a = _render_digest("c2", 4, 5000, force_generic=False)
    b, names_b = _render_digest("c2", 4, 5000, force_generic=False, env=["RDP_B200_NO_DEFERRED"])
    c, names_c = _render_digest("c2", 4, 5000, force_generic=False, env=["RDP_B200_FORCE_DEFERRED"])
    assert names_a == {"zb-opa-tex-modulate-2c/rgba32+deferred"}
    assert names_b == {"zb-opa-tex-modulate-2c/rgba32"}
    assert names_c == {"generic+deferred"}
    assert a == b == c
