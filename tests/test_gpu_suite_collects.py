"""CPU guard for the GPU suite: `pytest -m gpu --setup-plan` resolves every collected item's fixtures without running
anything, so a helper that pytest mistakes for a test (round 1: `testfortran_system(lib, N)` -> "fixture 'lib' not found",
which stopped the driver's `-x` run) fails here, on the CPU, instead of on the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_gpu_item_resolves_its_fixtures():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--setup-plan", "-q",
                        "-p", "no:cacheprovider"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert " error" not in out.splitlines()[-1].lower(), out[-3000:]
    assert "fixture '" not in out, out[-3000:]
