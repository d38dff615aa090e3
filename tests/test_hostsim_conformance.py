"""CPU-only conformance: the product's host C++ plus the device headers compiled with g++ (tests/hostsim) against the
Angrylion oracle and the golden digests.  This covers the host logic and the integer pixel maths without a GPU; the
`-m gpu` tests run the same suites through the CUDA library."""
import json
import os

import pytest

import conformance as cf
from oracle import angrylion
from sidebyside import SideBySide

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "rdp_digests.json")))

# A spread over every stage of the path; the full table runs on the GPU.
CPU_SUITES = [
    "fill-8", "fill-16-interlace", "fill-32", "fill-rect", "fill-ym-out-of-range", "tex-rect", "tex-rect-i8",
    "copy-4bpp-fb16-tlut", "copy-16bpp-fb8-perspective", "copy-32bpp-fb4",
    "rasterization-aa", "rasterization-interlace-aa", "interpolation-color-depth", "combiner-2cycle",
    "blender-fog-color-1cycle-i8", "blender-fog-color-2cycle-ia88", "coverage-wrap-image-read", "color-on-coverage",
    "depth-compare-interpenetrating", "depth-compare-decal-prim-depth",
    "interpolation-color-depth-alpha-test-dither-all", "interpolation-color-depth-aa-cvg-times-alpha-alpha-cvg-select",
    "interpolation-color-texture-perspective-2cycle-lod-frac-sharpen-detail",
    "interpolation-color-texture-pipelined-texel1-perspective", "interpolation-color-texture-rgba32",
    "interpolation-color-texture-yuv16", "interpolation-color-texture-ci4-tlut-mid-texel-no-quad",
    "interpolation-color-texture-ia16-tlut", "texture-convert-2cycle-YUV-convquad-mid-bilerp",
    "texture-convert-1cycle-RGBA-convbilerp",
]


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", CPU_SUITES)
def test_hostsim_matches_oracle(name):
    sbs = SideBySide("hostsim")
    try:
        digests = []
        errs = cf.run_rasterization(sbs, cf.SUITES[name], GOLDEN["iterations"], digests=digests)
    finally:
        sbs.close()
    assert not errs, errs
    assert digests == GOLDEN["digests"][name]


# The deferred tile pipeline's phase split (rdp_blend.cuh defer_*), in its scalar form, on every suite it accepts.
DEFERRED_SUITES = [n for n in CPU_SUITES if not any(k in n for k in ("copy-", "alpha-test", "cvg-times-alpha"))] + [
    "depth-compare-opaque", "depth-compare-transparent", "interpolation-color-texture-rgba16", "fill-rect",
]


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(set(DEFERRED_SUITES)))
def test_hostsim_deferred_matches_oracle(name, monkeypatch):
    monkeypatch.setenv("RDP_HOSTSIM_DEFERRED", "1")
    sbs = SideBySide("hostsim")
    try:
        digests = []
        errs = cf.run_rasterization(sbs, cf.SUITES[name], GOLDEN["iterations"], digests=digests)
        variant = sbs.gpu.last_tile_variant()
    finally:
        sbs.close()
    assert not errs, errs
    assert digests == GOLDEN["digests"][name]
    if name in ("depth-compare-opaque", "interpolation-color-texture-rgba16", "combiner-2cycle"):
        assert variant == "generic+deferred"


@pytest.mark.parametrize("name", ["fill-16", "interpolation-color-texture-rgba16", "depth-compare-opaque"])
def test_hostsim_matches_golden_without_oracle(name):
    sbs = SideBySide("hostsim", with_oracle=False)
    try:
        digests = []
        cf.run_rasterization(sbs, cf.SUITES[name], GOLDEN["iterations"], digests=digests)
    finally:
        sbs.close()
    assert digests == GOLDEN["digests"][name]


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_many_primitives_single_pass():
    """More primitives than the reference's 256-per-pass cap, all in one pass, must still blend in order."""
    v = cf.Variant(color=True, depth=True, primitive_count=700, depth_compare=True, aa=True, randomize_rdram=True, image_read_enable=True)
    sbs = SideBySide("hostsim")
    try:
        errs = cf.run_rasterization(sbs, v, 2)
        passes = sbs.gpu.pass_count()
    finally:
        sbs.close()
    assert not errs, errs
    assert passes <= 4


# --- texture loads -----------------------------------------------------------------------------------------------
@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(cf.LOAD_SUITES))
def test_hostsim_texture_loads(name):
    sbs = SideBySide("hostsim")
    try:
        errs = cf.run_load_suite(sbs, name, thorough=False)
    finally:
        sbs.close()
    assert not errs, errs


# --- VI ----------------------------------------------------------------------------------------------------------
import vi_conformance as vc  # noqa: E402

VI_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "vi_digests.json")))


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(vc.VI_SUITES))
def test_hostsim_vi_matches_oracle(name):
    sbs = SideBySide("hostsim")
    try:
        digests = []
        errs = vc.run_vi(sbs, vc.VI_SUITES[name], VI_GOLDEN["iterations"], digests=digests)
    finally:
        sbs.close()
    assert not errs, errs
    assert digests == VI_GOLDEN["digests"][name]
