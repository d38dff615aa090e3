"""Generates tests/golden/rdp_digests.json by running the conformance streams through the reference's CPU renderer
(oracle/_ref, built from /root/reference).  Run here, where the reference exists; the JSON travels with the repo so
GPU-side tests can be checked even where the oracle library is unavailable.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import conformance as cf  # noqa: E402
from sidebyside import SideBySide  # noqa: E402

GOLDEN_ITERATIONS = 4


class OracleOnly(SideBySide):
    """Drives only the oracle (no GPU object at all)."""

    def __init__(self):
        from oracle import angrylion
        import numpy as np
        self.rdram_size, self.hidden_size = 4 << 20, 2 << 20
        self.oracle = angrylion.Oracle(self.rdram_size)
        self.gpu = None

    def close(self):
        self.oracle.close()

    def clear_rdram(self):
        self.oracle.rdram[:] = 0
        self.oracle.hidden[:] = 0

    def randomize_rdram(self, rng):
        import numpy as np
        self.oracle.rdram[:] = rng.randint(0, 256, size=self.rdram_size, dtype=np.uint8)
        self.oracle.hidden[: self.hidden_size] = rng.randint(0, 4, size=self.hidden_size, dtype=np.uint8)

    def write_tmem(self, data):
        import numpy as np
        self.oracle.tmem[:] = np.frombuffer(np.ascontiguousarray(data).tobytes(), dtype=np.uint8)

    def run(self, words):
        self.oracle.command_stream(words)

    def compare_rdram(self):
        return []


def main():
    out = {}
    for name, variant in cf.SUITES.items():
        if variant.primitive_count > 1:
            continue
        sbs = OracleOnly()
        digests = []
        cf.run_rasterization(sbs, variant, GOLDEN_ITERATIONS, digests=digests)
        sbs.close()
        out[name] = digests
    path = os.path.join(ROOT, "tests", "golden", "rdp_digests.json")
    with open(path, "w") as f:
        json.dump({"iterations": GOLDEN_ITERATIONS, "seed": 1337, "digests": out}, f, indent=0, sort_keys=True)
    print("wrote", path, len(out), "suites")


if __name__ == "__main__":
    main()


def make_vi_golden():
    import vi_conformance as vc

    class OracleOnlyVI(OracleOnly):
        def set_vi_register(self, reg, value):
            self.oracle.set_vi_register(reg, value)

        def end_frame(self, opts=None):
            return self.oracle.update_screen(), None, 0, 0

    out = {}
    for name, variant in vc.VI_SUITES.items():
        sbs = OracleOnlyVI()
        digests = []
        vc.run_vi(sbs, variant, GOLDEN_ITERATIONS, digests=digests)
        sbs.close()
        out[name] = digests
    path = os.path.join(ROOT, "tests", "golden", "vi_digests.json")
    with open(path, "w") as f:
        json.dump({"iterations": GOLDEN_ITERATIONS, "seed": 1337, "digests": out}, f, indent=0, sort_keys=True)
    print("wrote", path, len(out), "suites")


if __name__ == "__main__":
    make_vi_golden()
