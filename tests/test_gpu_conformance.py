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
