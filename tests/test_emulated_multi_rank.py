"""The multi-GPU parity script (tests/mgpu_check.py, the one `-m gpu` runs under torchrun on 2 and 4 B200s), run
on the CPU: one process per rank, kernels through the execution-model emulator (tests/cusim/cusim.h),
collectives through a shared-memory stand-in for NCCL (tests/cusim/fake_nccl.cpp) that the engine loads via
EMDEE_NCCL_LIB. Covers what hangs or corrupts real multi-GPU runs: slab ownership, per-step halo exchange,
neighbor-only migration at rebuilds, the two-phase distributed rebuild criterion, collective downloads -- with
3 ranks as the first case where the up and down neighbors are different ranks. A collective that not every rank
enters is reported by the stand-in after FAKE_NCCL_TIMEOUT seconds instead of hanging.
"""
import os
import subprocess
import sys

import pytest

import common as cm

WORLDS = [2, 3, 8]   # 4 ranks run on real GPUs (tests/test_multi_gpu.py); 8 covers every multi-neighbor pattern here


@pytest.fixture(scope="module")
def libs():
    sys.path.insert(0, os.path.join(cm.ROOT, "tests", "cusim"))
    import build as cusim_build
    cusim_build.build()
    return cusim_build.build_fake_nccl()


@pytest.mark.parametrize("world", WORLDS)
def test_slab_decomposition_on_emulator(world, libs):
    if world > 2 * (os.cpu_count() or 1):
        pytest.skip(f"{world} spinning ranks on {os.cpu_count()} cores")
    env = dict(os.environ, EMDEE_MGPU_EMULATED="1", EMDEE_NCCL_LIB=libs, FAKE_NCCL_TIMEOUT="120", EMDEE_QUIET="1",
               OMP_NUM_THREADS="1")
    if world >= 6:
        env["EMDEE_MGPU_NCELL"] = "21"
    if world >= 4:
        env["EMDEE_MGPU_SKIP_VERLET"] = "1"    # verlet_step with migration is covered at 2 and 3 ranks
        env["EMDEE_MGPU_SKIP_CHARGED"] = "1"   # the charged LJ system repeats the neutral one with a Coulomb model
    if world >= 3:   # keep the CPU suite short: the Ewald crystal grows with the rank count (three cell layers per rank)
        env["EMDEE_MGPU_SKIP_EWALD"] = "1"
    env["EMDEE_MGPU_BODY_STEPS"] = "3" if world <= 3 else "0"   # rigid-body dynamics: 2 and 3 ranks cover both neighbor patterns
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world),
           os.path.join(cm.ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "[mgpu] ALL OK" in r.stdout


def test_bond_longer_than_the_halo_is_rejected_on_every_rank(libs):
    """several GPUs: a bonded partner outside the halo would be read with stale coordinates; the library stops instead
    (Engine::add_bonded, k_bonded `toolong`), on all ranks (the flag is all-reduced), with its usual message format"""
    env = dict(os.environ, EMDEE_MGPU_EMULATED="1", EMDEE_NCCL_LIB=libs, FAKE_NCCL_TIMEOUT="60", EMDEE_QUIET="1",
               OMP_NUM_THREADS="1", EMDEE_MGPU_LONG_BOND="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29711", os.path.join(cm.ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode != 0
    assert "NOT rejected" not in r.stdout
    assert r.stderr.count("Error in bonded force computation: a bond or angle arm is longer than Rc + skin") >= 1, r.stderr[-3000:]
