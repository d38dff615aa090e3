"""GPU parity tests proper: every conformance suite through librdp_b200.so (CUDA, via the C-ABI) against the
Angrylion oracle (when oracle/_ref travelled to the box) and against the committed golden digests."""
import json
import os

import pytest

import conformance as cf
from oracle import angrylion
from sidebyside import SideBySide

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "rdp_digests.json")))

pytestmark = pytest.mark.gpu

ALL = sorted(cf.SUITES)


@pytest.mark.parametrize("name", ALL)
def test_gpu_matches_oracle(name):
    v = cf.SUITES[name]
    iters = 2 if v.primitive_count > 1 else 12
    sbs = SideBySide("cuda", with_oracle=angrylion.available())
    try:
        assert sbs.gpu.backend_name() == "cuda-sm_100a"
        digests = []
        errs = cf.run_rasterization(sbs, v, iters, digests=digests)
        launches = sbs.gpu.launch_count()
    finally:
        sbs.close()
    assert not errs, errs
    assert launches > 0
    if name in GOLDEN["digests"]:
        n = GOLDEN["iterations"]
        assert digests[:n] == GOLDEN["digests"][name]


@pytest.mark.parametrize("name", ALL)
def test_gpu_deferred_matches_oracle(name, monkeypatch):
    """Same suites with every deferrable pass forced through the deferred tile kernel (dynamic instantiation)."""
    monkeypatch.setenv("RDP_B200_FORCE_DEFERRED", "1")
    v = cf.SUITES[name]
    iters = 2 if v.primitive_count > 1 else 6
    sbs = SideBySide("cuda", with_oracle=angrylion.available())
    try:
        digests = []
        errs = cf.run_rasterization(sbs, v, iters, digests=digests)
    finally:
        sbs.close()
    assert not errs, errs
    if name in GOLDEN["digests"]:
        n = min(GOLDEN["iterations"], iters)
        assert digests[:n] == GOLDEN["digests"][name][:n]


@pytest.mark.parametrize("name", sorted(cf.LOAD_SUITES))
def test_gpu_texture_loads(name):
    if not angrylion.available():
        pytest.skip("texture load suites compare TMEM against the oracle")
    sbs = SideBySide("cuda")
    try:
        errs = cf.run_load_suite(sbs, name, thorough=True)
    finally:
        sbs.close()
    assert not errs, errs


import vi_conformance as vc  # noqa: E402

VI_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "vi_digests.json")))


@pytest.mark.parametrize("name", sorted(vc.VI_SUITES))
def test_gpu_vi_matches_oracle(name):
    sbs = SideBySide("cuda", with_oracle=angrylion.available())
    try:
        digests = []
        errs = vc.run_vi(sbs, vc.VI_SUITES[name], 16, digests=digests)
    finally:
        sbs.close()
    assert not errs, errs
    n = VI_GOLDEN["iterations"]
    assert digests[:n] == VI_GOLDEN["digests"][name]


# --- the tile kernels' overflow paths --------------------------------------------------------------------------------
# Random screen-sized triangles pile hundreds of primitives onto every tile: more than one 128-entry list per tile
# (several walker runs, no mask staging), more than 256 fragments per 32-entry chunk (several batches in the deferred
# walker), and -- above 8192 primitives -- the coarse-mask gather instead of the single-round one.
STRESS = {
    "deferred-3000": (cf.Variant(color=True, depth=True, primitive_count=3000, depth_compare=True, aa=True, randomize_rdram=True,
                                 image_read_enable=True), "generic"),
    "deferred-texture-2cycle-3000": (cf.Variant(color=True, depth=True, texture=True, texture_size=2,
                                                cycle_type=1, perspective=True, primitive_count=3000, depth_compare=True, aa=True,
                                                randomize_rdram=True, image_read_enable=True), "generic"),
    "in-order-alpha-test-3000": (cf.Variant(color=True, depth=True, primitive_count=3000, depth_compare=True, aa=True, alpha_test=True,
                                            randomize_rdram=True, image_read_enable=True), "generic"),
    "coarse-gather-9000": (cf.Variant(color=True, depth=True, primitive_count=9000, depth_compare=True, aa=True, randomize_rdram=True,
                                      image_read_enable=True), "generic"),
}


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(STRESS))
@pytest.mark.parametrize("force_deferred", [False, True])
def test_gpu_many_overlapping_primitives(name, force_deferred, monkeypatch):
    if force_deferred:
        monkeypatch.setenv("RDP_B200_FORCE_DEFERRED", "1")
    v, _ = STRESS[name]
    sbs = SideBySide("cuda")
    try:
        errs = cf.run_rasterization(sbs, v, 1)
        passes = sbs.gpu.pass_count()
        variant = sbs.gpu.last_tile_variant()
    finally:
        sbs.close()
    assert not errs, errs
    assert passes <= 4
    if force_deferred and "alpha-test" not in name:
        assert variant == "generic+deferred"
