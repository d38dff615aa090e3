"""bench.py's reference arm runs on host cores alone, so its JSON line can be checked without a GPU: the keys the driver
reads, the tier's cpu_baseline / e2e shapes, and that a missing CUDA device makes the product arm fail loudly instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from oracle import angrylion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=600)


@pytest.mark.skipif(not angrylion.available(), reason="oracle/_ref not built")
def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "vi_frames_per_s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a CUDA device is present")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--small", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]
