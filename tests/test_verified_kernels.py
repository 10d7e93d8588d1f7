"""Guard for CPU-only editing sessions: every kernel that was verified on the B200 (profiles/r1_verified_sass.json,
written after the last green GPU run) must still compile to byte-identical SASS. A deliberate kernel change has
to go back to the GPU, after which the fingerprint file is re-saved (tools/sass_fingerprint.py --save)."""
import os
import shutil
import subprocess
import sys

import pytest

from common import ROOT, api


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_gpu_verified_kernels_are_byte_identical():
    if not os.path.exists(api.PRODUCT_LIB):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_fingerprint.py"), "--check",
                        os.path.join(ROOT, "profiles", "r1_verified_sass.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
