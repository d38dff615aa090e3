"""Conservativeness of the per-(tile, primitive) depth rejection bound derived in DESIGN.md 7b -- groundwork for the next
tile-kernel optimisation, checked on the CPU before any kernel uses it.

The host simulator (RDP_HOSTSIM_HIZ=1, tests/hostsim/backend_cpu.cpp hiz_check_tile) walks every tile of a deferrable
pass primitive-major through the same defer_candidate / defer_depth code the kernels run, evaluates the bound for
every entry with the information a warp would have (exact tile maxima at the start of each 32-primitive batch, widened
inside the batch by what earlier entries may have written), and counts entries that the bound rejects although one of
their fragments passes the depth test.  That count has to be zero; the rendered result is still compared with the
oracle, so the walk itself is known to be the real one."""
import ctypes as C

import numpy as np
import pytest

import conformance as cf
from harness import workloads as wl
from oracle import angrylion
from sidebyside import SideBySide


def _stats(sbs):
    sbs.gpu.idle()
    lib = sbs.gpu.lib
    lib.rdp_b200_backend_device.restype = C.c_void_p
    out = (C.c_ulonglong * 8)()
    lib.rdpb_hostsim_hiz_stats(C.c_void_p(lib.rdp_b200_backend_device(sbs.gpu._h)), out)
    return dict(zip(("entries", "rejected", "hidden", "false_rejections", "rejected_with_exact_zmin", "unsound_z_bounds",
                     "rejected_per_row", "false_rejections_per_row"), out))


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_bound_never_rejects_a_visible_entry_on_c2(monkeypatch):
    monkeypatch.setenv("RDP_HOSTSIM_HIZ", "1")
    sbs = SideBySide("hostsim", rdram_size=wl.RDRAM_SIZE, hidden_size=wl.HIDDEN_SIZE)
    try:
        sbs.clear_rdram()
        sbs.write_rdram(wl.TEX_ADDR, wl.c2_textures())
        rng = np.random.RandomState(2)
        for _ in range(5):
            sbs.run(wl.c2_frame(rng, 5000))
        s = _stats(sbs)
        assert not sbs.compare_rdram()
    finally:
        sbs.close()
    print(s)
    assert s["false_rejections"] == 0 and s["unsound_z_bounds"] == 0 and s["false_rejections_per_row"] == 0
    assert s["entries"] > 400000
    # about a third of the entries have no visible fragment; the conservative bound finds a good part of those
    assert 0.25 * s["entries"] < s["hidden"] < 0.40 * s["entries"]
    assert s["rejected"] > 0.10 * s["entries"]
    assert s["rejected"] <= s["rejected_with_exact_zmin"] <= s["hidden"]


@pytest.mark.parametrize("name", ["depth-compare-opaque", "depth-compare-interpenetrating", "depth-compare-decal",
                                  "depth-compare-transparent", "depth-compare-opaque-prim-depth", "interpolation-color-depth"])
def test_bound_never_rejects_a_visible_entry_in_depth_suites(name, monkeypatch):
    if name not in cf.SUITES:
        pytest.skip("no such suite")
    monkeypatch.setenv("RDP_HOSTSIM_HIZ", "1")
    sbs = SideBySide("hostsim", with_oracle=angrylion.available())
    try:
        assert not cf.run_rasterization(sbs, cf.SUITES[name], 24)
        s = _stats(sbs)
    finally:
        sbs.close()
    assert s["entries"] > 0
    assert s["false_rejections"] == 0 and s["unsound_z_bounds"] == 0 and s["false_rejections_per_row"] == 0
