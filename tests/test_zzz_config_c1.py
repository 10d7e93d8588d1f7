"""BASELINE.json configs[0] at full size: the reference's test/testfortran.f90 driven by examples/data.inp -- 10 000
atoms on a simple-cubic lattice at rho* = 0.88, Rc = 2.5, skin = 0.4, two atom types with the same LJ model set on
all three type pairs, every pair |i - j| < 10 excluded (EmDee_ignore_pair, 89 955 calls), momenta uploaded from the
host, NVE with dt = 0.002 and Options%Compute only every Nprop-th step. The product against the oracle through the
C ABI. (Written after the last GPU session of round 1; the emulator run uses the same recipe with 1000 atoms.)"""
import os

import numpy as np
import pytest

import common as cm
from test_gpu_parity import both

pytestmark = pytest.mark.gpu


def _make_testfortran_system(lib, N, seed=86245, rho=0.88, Temp=2.0):
    L = (N / rho) ** (1.0 / 3.0)
    Nd = int(np.ceil(N ** (1.0 / 3.0) - 1e-9))
    m = np.arange(N)
    k, j, i = m // (Nd * Nd), (m % (Nd * Nd)) // Nd, m % Nd
    R = (L / Nd) * (np.stack([i, j, k], axis=1) + 0.5)                    # create_configuration, testfortran.f90:146-160
    rng = np.random.default_rng(seed)
    V = rng.normal(size=(N, 3))
    V -= V.mean(axis=0)
    V *= np.sqrt(Temp * (3 * N - 3) / (V * V).sum())
    types = np.where(np.arange(N) < N // 2, 1, 2).astype(np.int32)
    s = lib.system(2, 1, 2.5, 0.4, N, types, None, None)
    pair = lib.EmDee_pair_lj_cut(1.0, 1.0)
    s.set_pair_model(1, 1, pair, 1.0)
    s.set_pair_model(2, 2, pair, 1.0)
    s.set_pair_model(1, 2, pair, 1.0)
    for a in range(1, N):                                                  # testfortran.f90:76-80
        for b in range(a + 1, min(a + 10, N + 1)):
            s.ignore_pair(a, b)
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    s.upload("momenta", V)
    return s


def test_testfortran_data_inp():
    N = int(os.environ.get("EMDEE_TEST_C1_ATOMS", "10000"))
    nsteps, nprop = (100, 50) if N == 10000 else (20, 10)
    sp, so = both(lambda lib: _make_testfortran_system(lib, N))
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-12
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-13
    dt = 0.002
    for step in range(1, nsteps + 1):
        for s in (sp, so):
            s.md.Options.Compute = (step % nprop == 0)
            s.boost(1.0, 0.0, 0.5 * dt)
            s.displace(1.0, 0.0, dt)
            s.boost(1.0, 0.0, 0.5 * dt)
    assert sp.md.Builds == so.md.Builds and sp.md.Builds >= 2
    assert np.array_equal(sp.pairs(), so.pairs())
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
    for a, b in ((sp.md.Energy.Potential, so.md.Energy.Potential), (sp.md.Virial.Total, so.md.Virial.Total),
                 (sp.md.Kinetic.Total, so.md.Kinetic.Total)):
        assert cm.rel(a, b) < 1e-9
    sp.finalize(), so.finalize()
