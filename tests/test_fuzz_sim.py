"""A short fixed-seed slice of the randomised sim-vs-oracle sweep (tests/fuzz_sim.py) in the CPU suite."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fuzz_slice():
    p = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_sim.py"), "5", "25"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "bad 0" in p.stdout


@pytest.mark.gpu
def test_fuzz_slice_gpu():
    """The same sweep through the CUDA library (another seed, 40 s)."""
    p = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_sim.py"), "6", "40", "engine", "0", "0"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "bad 0" in p.stdout
