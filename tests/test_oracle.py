"""Pins the oracle: the reference's CPU renderer (oracle/_ref) must reproduce the committed golden digests, which
were produced by the same library from the conformance streams (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import angrylion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "rdp_digests.json")))

pytestmark = pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_fill_rect_known_answer():
    """100x50 fill rectangle writes exactly 5,000 16-bit pixels (SURVEY 8c probe)."""
    from harness.rdp_builder import Builder, FMT_RGBA, SIZE_16, CYCLE_FILL
    o = angrylion.Oracle(4 << 20)
    b = Builder()
    b.set_color_image(FMT_RGBA, SIZE_16, 0, 320)
    b.set_scissor(0, 0, 320, 240)
    b.om.cycle_type = CYCLE_FILL
    b.set_fill_color(0xFFFFFFFF)
    b.fill_rectangle(10, 20, 100, 50)
    b.sync_full()
    o.command_stream(b.words())
    px = o.rdram.view(np.uint16)
    assert np.count_nonzero(px) == 5000
    assert np.all(px[np.nonzero(px)] == 0xFFFF)
    o.close()


@pytest.mark.parametrize("name", ["fill-16", "interpolation-color-depth", "depth-compare-interpenetrating",
                                  "interpolation-color-texture-ci8-tlut", "texture-convert-2cycle-YUV-convquad-mid-bilerp",
                                  "copy-8bpp-fb16-tlut", "blender-fog-color-2cycle-rgba8888"])
def test_oracle_reproduces_golden(name):
    import conformance as cf
    from golden.make_golden import OracleOnly
    sbs = OracleOnly()
    digests = []
    cf.run_rasterization(sbs, cf.SUITES[name], GOLDEN["iterations"], digests=digests)
    sbs.close()
    assert digests == GOLDEN["digests"][name]
