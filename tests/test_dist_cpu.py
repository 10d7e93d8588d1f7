"""Host-side logic of the multi-GPU path, covered on CPU with a world_size-2 gloo group."""
import os
import subprocess
import sys

import common as cm


def test_world_size_2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(cm.ROOT, "tests", "dist_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "[dist-cpu] ALL OK" in r.stdout
