"""The reference's own conformance programs against this library.

oracle/Makefile compiles rdp_conformance.cpp / vi_conformance.cpp (and the command builder, triangle converter, Angrylion
replayer and side-by-side comparison they drive) UNMODIFIED from /root/reference, with
oracle/ref_harness/replayer_driver_b200.cpp -- the reference's ReplayerDriver interface on top of the C-ABI of
include/rdp_b200.h, i.e. INTEGRATION.md's binding compiled for real -- in the GPU seat.  So these are the reference's
std::mt19937{1337} command streams, the reference's comparisons, the reference's Angrylion glue; the 255 (suite, range)
pairs are its ctest registrations (tests/golden/reference_ctest_list.txt).

CPU suite: every registered test, a few iterations each, against the test-only host simulator (same device code, same
host code).  GPU suite: every registered test through librdp_b200.so, five iterations each; the full depth
(--range 0 1000) is tools/run_reference_conformance.py --scale 1, run once per round with the log under profiles/."""
import os
import subprocess
import sys

import pytest

from sidebyside import HOSTSIM_LIB, build_hostsim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tools", "run_reference_conformance.py")
BINARIES = [os.path.join(ROOT, "oracle", "_ref", n) for n in ("rdp-conformance-b200", "vi-conformance-b200")]
needs_harness = pytest.mark.skipif(not all(os.path.exists(b) for b in BINARIES), reason="oracle/_ref conformance programs not built")


def _run(*args, timeout=1500):
    return subprocess.run([sys.executable, RUNNER, *args], capture_output=True, text=True, timeout=timeout)


def test_registered_list_matches_the_reference():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_reference_conformance as rrc
    tests = rrc.registered_tests()
    assert len(tests) == 255 and sum(1 for t in tests if t[0] == "vi") == 24
    assert ("rdp", "fill-16", 1000) in tests and ("vi", "per-scanline-xh", 1000) in tests
    assert all(n in (100, 1000) for _, _, n in tests)
    if os.path.exists("/root/reference/CMakeLists.txt"):
        import re
        txt = open("/root/reference/CMakeLists.txt").read()
        names = [m.group(2).strip() for m in re.finditer(r"^\s*add_(rdp_test_reduced|rdp_test|vi_test)\(([^)]+)\)", txt, re.M)]
        assert names == [t[1] for t in tests]


@needs_harness
@pytest.mark.timeout(1800)
def test_every_reference_ctest_passes_on_the_host_simulator():
    build_hostsim()
    r = _run("--library", HOSTSIM_LIB, "--scale", "0.002", "--jobs", str(max(2, (os.cpu_count() or 4) - 1)))
    assert r.returncode == 0 and "255 / 255" in r.stdout, r.stdout[-3000:]


@needs_harness
@pytest.mark.gpu
@pytest.mark.timeout(1800)
def test_every_reference_ctest_passes_on_the_gpu():
    # One process per program: contexts sharing a GPU are time-sliced, and 255 start-ups cost more than the suites.
    r = _run("--single-process", "5")
    assert r.returncode == 0 and "255 / 255" in r.stdout, r.stdout[-3000:]
