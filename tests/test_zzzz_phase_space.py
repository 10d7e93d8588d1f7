"""EmDee_memory_address and EmDee_share_phase_space on device-resident state (reference src/EmDeeCode.f90:212-269,
test/test_phase_space_sharing.f90) against the oracle. The product hands out pointers into CUDA managed memory
(Engine::expose) and aliases device buffers between systems (Engine::share_phase_space). Written after the last GPU
session of round 1: logic verified through the emulator (tests/test_emulated_kernels.py)."""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import both

pytestmark = pytest.mark.gpu


def _lj(lib, Rc=None):
    s, c = cm.lj_sample_system(lib, lambda l, e, sg: l.EmDee_shifted_force(l.EmDee_pair_lj_cut(e, sg)), Rc=Rc)
    s.info = c
    return s


def test_memory_address_views_are_the_live_arrays():
    sp, so = both(_lj)
    for s in (sp, so):
        s.random_momenta(0.8, True, 4242)
    views = {}
    for tag, s in (("p", sp), ("o", so)):
        views[tag] = {k: s.memory_address(k) for k in ("coordinates", "momenta", "forces")}
        for k in ("coordinates", "momenta", "forces"):
            assert np.array_equal(views[tag][k], s.download(k)), k
    # the client moves atoms and changes momenta THROUGH the pointers; the next force call must see it
    for tag, s in (("p", sp), ("o", so)):
        v = views[tag]
        v["coordinates"][10] += np.array([0.05, -0.03, 0.02])
        v["coordinates"][500] += np.array([1.3, 0.0, 0.0])      # far enough to trigger a list rebuild
        v["momenta"][:] *= 1.01
        s.compute_forces()
    assert sp.md.Builds == so.md.Builds == 2
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(np.array(views["p"]["forces"]), np.array(views["o"]["forces"])) < 1e-10
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-12
    # dynamics keeps writing into the same arrays
    for _ in range(5):
        for s in (sp, so):
            s.boost(1.0, 0.0, 0.001)
            s.displace(1.0, 0.0, 0.002)
            s.boost(1.0, 0.0, 0.001)
    assert np.abs(np.array(views["p"]["coordinates"]) - np.array(views["o"]["coordinates"])).max() < 1e-10
    assert np.abs(np.array(views["p"]["momenta"]) - np.array(views["o"]["momenta"])).max() < 1e-9
    assert np.array_equal(views["p"]["coordinates"], sp.download("coordinates"))
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-11
    sp.finalize(), so.finalize()


def _shared_pair(lib, bodies):
    if bodies:
        f = lambda l: l.EmDee_coul_sf()
        a, c = cm.spce_sample_system(lib, f)
        b, _ = cm.spce_sample_system(lib, f, Rc=0.5 * c["Rc"])
        a.random_momenta(c["kB"] * c["Temp"], True, 31)
        dt = 1.0
    else:
        a, b = _lj(lib), _lj(lib, Rc=1.5)
        c = a.info
        a.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        dt = c["Dt"]
    return a, b, dt


@pytest.mark.parametrize("bodies", [False, True])
def test_share_phase_space(bodies):
    """reference test/test_phase_space_sharing.f90:38-90: a second system with a shorter cutoff takes over the phase
    space of the first and evaluates its own forces and energies on it while the first one drives the dynamics."""
    (ap, bp, dt), (ao, bo, _) = _shared_pair(cm.product(), bodies), _shared_pair(cm.oracle(), bodies)

    def nve(s, n):
        for _ in range(n):
            s.boost(1.0, 0.0, 0.5 * dt)
            s.displace(1.0, 0.0, dt)
            s.boost(1.0, 0.0, 0.5 * dt)

    for a, b in ((ap, bp), (ao, bo)):
        nve(a, 4)
        a.share_phase_space(b)
        assert b.md.Kinetic.Total == a.md.Kinetic.Total
        b.md.Options.Compute = True
        b.compute_forces()
    assert np.abs(bp.download("coordinates") - bo.download("coordinates")).max() < 1e-9
    assert cm.rel(bp.md.Energy.Potential, bo.md.Energy.Potential) < 1e-9
    assert np.array_equal(bp.pairs(), bo.pairs())
    for a, b in ((ap, bp), (ao, bo)):
        nve(a, 6)                      # the first system moves the shared atoms ...
        b.compute_forces()             # ... and the second one evaluates its model on the new configuration
    assert np.array_equal(bp.download("coordinates"), ap.download("coordinates"))
    assert np.abs(bp.download("coordinates") - bo.download("coordinates")).max() < 1e-8
    assert cm.rel(bp.md.Energy.Potential, bo.md.Energy.Potential) < 1e-8
    assert cm.rel(bp.md.Virial.Total, bo.md.Virial.Total) < 1e-7
    assert bp.md.Builds == bo.md.Builds and ap.md.Builds == ao.md.Builds
    assert cm.rel_force_error(bp.download("forces"), bo.download("forces")) < 1e-7
    # and the other way round: the taker can drive the dynamics too
    for b in (bp, bo):
        b.random_momenta(0.5 * (ap.info["kB"] * ap.info["Temp"] if not bodies else 0.3), True, 5)
        nve(b, 3)
    assert np.abs(ap.download("coordinates") - ao.download("coordinates")).max() < 1e-8
    for s in (bp, bo, ap, ao):
        s.finalize()
