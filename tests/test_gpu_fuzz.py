"""Randomised parity sweep (tools/fuzz_parity.py): random frame sizes, feature counts, scale factors, level counts, FAST thresholds and frame
contents; the CUDA path with both FAST formulations against the C oracle and the reference's compiled ORBextractor.cpp, then the three matcher
engines on random problem sizes.  The tool runs
hundreds of cases on demand (620 cases over two seeds at the end of round 2: 0 mismatches); the test keeps a short sweep in the suite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_geometries_and_parameters(built, oracle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "16", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "16 cases, 0 mismatches" in r.stdout
