"""GPU: randomised parity run (tools/fuzz_extract.py): random image sizes, constructor arguments, image statistics and staging
paths through the single-frame and batch entry points against the CPU oracle — every accepted case must be bit-identical."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_extractor_parity():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_extract.py"), "30", "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "30 cases, 0 mismatches" in r.stdout


def test_randomised_matcher_parity():
    """tools/fuzz_match.py: random query subsets, radii, occupancy, ratios and vocabulary groupings through all five matcher entry
    points (the parallel and the sequential resolve) against the oracle's sequential restatements."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_match.py"), "25", "21"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "25 cases, 0 mismatches" in r.stdout
