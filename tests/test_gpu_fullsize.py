"""Parity at BASELINE.json's full size (1M-atom LJ box, configs[3]): direct comparison with the CPU oracle
(it finishes in seconds on a multicore host) plus size-independent properties of the domain."""
import os

import numpy as np
import pytest

import bench
import common as cm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    R, P, L = bench.make_workload(bench.NCELL_DEFAULT)
    sp = bench.build_system(cm.product(), R, P, L, 1)
    yield sp, R, P, L
    sp.finalize()


def test_full_size_matches_oracle(big):
    sp, R, P, L = big
    threads = min(os.cpu_count() or 1, 32)
    so = bench.build_system(cm.oracle(), R, P, L, threads)
    assert np.array_equal(sp.pairs(), so.pairs()), "neighbor pair sets differ at 1M atoms"
    err = cm.rel_force_error(sp.download("forces"), so.download("forces"))
    assert err <= 1e-10, err
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) <= 1e-12
    assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) <= 1e-12
    # a few resident MD steps, then compare again (rebuild decisions included)
    for _ in range(12):
        bench.md_step(sp)
        bench.md_step(so)
    assert sp.md.Builds == so.md.Builds
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) <= 1e-9
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) <= 1e-11
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) <= 1e-11
    so.finalize()


def test_full_size_properties(big):
    sp, R, P, L = big
    F = sp.download("forces")
    # Newton's third law: the net force vanishes (to rounding of a 1M-term sum)
    assert np.abs(F.sum(axis=0)).max() < 1e-7 * np.abs(F).sum() / F.shape[0] * np.sqrt(F.shape[0])
    # every pair within Rc+skin appears exactly once, indices valid, no self pairs
    pairs = sp.pairs()
    assert pairs.min() >= 0 and pairs.max() < F.shape[0] and np.all(pairs[:, 0] < pairs[:, 1])
    key = pairs[:, 0].astype(np.int64) * F.shape[0] + pairs[:, 1]
    assert np.unique(key).size == key.size
    # idempotence: recomputing on the same configuration changes nothing, bit for bit
    sp.upload("coordinates", sp.download("coordinates"))
    U0 = sp.md.Energy.Potential
    sp.compute_forces()
    assert sp.md.Energy.Potential == U0 or cm.rel(sp.md.Energy.Potential, U0) < 1e-15
    assert np.array_equal(sp.download("forces"), F) or sp.md.Builds > 1


def test_energy_conservation_short_nve(big):
    sp, R, P, L = big
    e0 = sp.md.Energy.Potential + sp.md.Kinetic.Total
    for _ in range(40):
        bench.md_step(sp)
    e1 = sp.md.Energy.Potential + sp.md.Kinetic.Total
    # plain truncated LJ (energy jumps by E(Rc) whenever a pair crosses Rc) on a melting lattice: loose bound
    assert abs(e1 - e0) < 2e-2 * abs(sp.md.Kinetic.Total)
