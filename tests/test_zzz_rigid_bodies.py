"""Device-resident rigid-body dynamics (engine_bodies.cuh: k_body_frame / k_body_boost / k_body_move / ... and the
EmDee_verlet_step bookkeeping) against the CPU oracle's restatement of reference src/ArBee.f90,
src/EmDeeData.f90:157-189,823-922 and src/EmDeeCode.f90:659-801,950-1211, through the C ABI: the same calls the
reference's own rigid-body programs make (test/test_rigid_body_exact.f90, test_rigid_body_miller.f90,
test_rigid_body_setup.f90, test_verlet.f90, test_coul_*.f90). Written after the last GPU session of round 1: the
kernels' logic is verified on the CPU through the emulator (tests/test_emulated_kernels.py); the file name sorts late.

Tolerances: single evaluations 1e-9 relative to the largest entry of the array (the device uses a quaternion-product
formulation and fused multiply-adds where the oracle uses the reference's 4x3 matrices, and the analytic 3x3
eigen-solver loses a few digits to cancellation); trajectories of a few steps 1e-8 absolute on coordinates
(Angstrom), 1e-8 relative on energies. On the emulator (no FMA contraction) the same comparisons hold at 1e-11.
"""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import both

pytestmark = pytest.mark.gpu

BODY_ITEMS = {"quaternions": 4, "quatmom": 4, "quattau": 4, "angmom": 3, "bodycoord": 3, "bodymom": 3,
              "bodyforces": 3, "torques": 3, "inertia": 3}


def _spce(lib, mode=0, seed=4321, momenta=True):
    s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0))
    s.md.Options.RotationMode = mode
    if momenta:
        s.random_momenta(c["kB"] * c["Temp"], True, seed)
    s.info = c
    return s


def _close(a, b, rtol, what=""):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err < rtol, f"{what}: {err:.3e}"


def _compare_bodies(sp, so, nb, rtol, skip=()):
    for item, w in BODY_ITEMS.items():
        if item in skip:
            continue
        _close(sp.download(item, (nb, w)), so.download(item, (nb, w)), rtol, item)


def _compare_scalars(sp, so, rtol):
    for grp, names in (("Energy", ("Potential", "Dispersion", "Coulomb")), ("Virial", ("Total", "Body")),
                       ("Kinetic", ("Total", "Rotational"))):
        for n in names:
            a, b = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
            assert cm.rel(a, b) < rtol, f"{grp}.{n}: {a!r} vs {b!r}"


def test_body_frames_and_random_momenta():
    sp, so = both(lambda lib: _spce(lib, momenta=False))
    nb = 750
    _compare_bodies(sp, so, nb, 1e-9, skip=("quatmom", "angmom", "bodymom", "bodyforces", "torques", "quattau"))
    _close(sp.download("centersOfMass", (nb, 3)), so.download("centersOfMass", (nb, 3)), 1e-13, "centersOfMass")
    kT = sp.info["kB"] * sp.info["Temp"]
    for s in (sp, so):
        s.random_momenta(kT, True, 97531)
    _compare_bodies(sp, so, nb, 1e-9, skip=("bodyforces", "torques", "quattau"))
    _close(sp.download("momenta"), so.download("momenta"), 1e-9, "momenta")
    _compare_scalars(sp, so, 1e-8)
    assert sp.md.DoF == so.md.DoF and sp.md.RotDoF == so.md.RotDoF
    sp.finalize(), so.finalize()


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_spce_nve_trajectory(mode):
    """boost / displace / boost with rigid bodies: exact free rotor (mode 0) and NO_SQUISH splitting (mode n)."""
    sp, so = both(lambda lib: _spce(lib, mode))
    dt = 1.0
    for step in range(1, 9):
        for s in (sp, so):
            s.md.Options.Compute = (step % 4 == 0)
            s.boost(1.0, 0.0, 0.5 * dt)
            s.displace(1.0, 0.0, dt)
            s.boost(1.0, 0.0, 0.5 * dt)
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
    _compare_bodies(sp, so, 750, 1e-8)
    _close(sp.download("momenta"), so.download("momenta"), 1e-8, "momenta")
    _compare_scalars(sp, so, 1e-8)
    assert sp.md.Builds == so.md.Builds
    assert np.array_equal(sp.pairs(), so.pairs())
    sp.finalize(), so.finalize()


def test_translate_and_rotate_switches():
    """Options%Translate / Options%Rotate gate the two halves of boost and move (src/EmDeeData.f90:836-857, 879-893)."""
    for translate, rotate in ((True, False), (False, True)):
        sp, so = both(lambda lib: _spce(lib, 0, seed=11))
        for s in (sp, so):
            s.md.Options.Translate = translate
            s.md.Options.Rotate = rotate
            for _ in range(3):
                s.boost(1.0, 0.0, 0.5)
                s.displace(1.0, 0.0, 1.0)
                s.boost(1.0, 0.0, 0.5)
        assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
        _compare_bodies(sp, so, 750, 1e-8)
        _compare_scalars(sp, so, 1e-8)
        sp.finalize(), so.finalize()


def test_momenta_upload_with_bodies():
    sp, so = both(lambda lib: _spce(lib, 0, seed=5))
    rng = np.random.default_rng(3)
    P = rng.normal(size=(2250, 3)) * 0.02
    for s in (sp, so):
        s.upload("momenta", P)
    _compare_bodies(sp, so, 750, 1e-9, skip=("bodyforces", "torques", "quattau"))
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-12
    assert cm.rel(sp.md.Kinetic.Rotational, so.md.Kinetic.Rotational) < 1e-12
    _close(sp.download("momenta"), so.download("momenta"), 1e-9, "momenta")
    sp.finalize(), so.finalize()


@pytest.mark.parametrize("mode", [0, 2])
def test_verlet_step_with_shadow_terms(mode):
    sp, so = both(lambda lib: _spce(lib, mode, seed=2468))
    for step in range(1, 7):
        for s in (sp, so):
            s.md.Options.Compute = (step % 2 == 0)
            s.verlet_step(1.0)
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
    _compare_scalars(sp, so, 1e-8)
    for grp, n in (("Energy", "ShadowPotential"), ("Kinetic", "ShadowKinetic"), ("Kinetic", "ShadowRotational")):
        a, b = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
        assert cm.rel(a, b) < 1e-8, f"{grp}.{n}: {a!r} vs {b!r}"
    sp.finalize(), so.finalize()


def test_verlet_step_free_atoms_only():
    """reference test/test_verlet.f90: the LJ sample (no bodies) driven by EmDee_verlet_step."""
    def make(lib):
        s, c = cm.lj_sample_system(lib, lambda l, e, sg: l.EmDee_pair_lj_cut(e, sg))
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        s.info = c
        return s
    sp, so = both(make)
    for step in range(1, 11):
        for s in (sp, so):
            s.md.Options.Compute = (step % 5 == 0)
            s.verlet_step(sp.info["Dt"])
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-10
    for grp, n in (("Energy", "Potential"), ("Kinetic", "Total"), ("Energy", "ShadowPotential"), ("Kinetic", "ShadowKinetic")):
        a, b = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
        assert cm.rel(a, b) < 1e-9, f"{grp}.{n}: {a!r} vs {b!r}"
    assert sp.md.Kinetic.ShadowRotational == 0.0 and sp.md.Builds == so.md.Builds
    sp.finalize(), so.finalize()


def test_mixed_bodies_and_free_atoms():
    """reference test/test_rigid_body_setup.f90:36: every third molecule dissolved into free atoms."""
    def make(lib):
        c = cm.load_fixture("NIST_spce_sample")
        mol = c["molecule"].copy()
        mol[mol % 3 == 0] = 0
        s = lib.system(2, 1, c["Rc"], c["Rs"], c["N"], c["atomType"], c["mass"], mol)
        eps = c["epsilon"] / c["mvv2e"]
        for i in range(2):
            model = lib.EmDee_pair_none() if eps[i] == 0 else lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[i], c["sigma"][i]))
            s.set_pair_model(i + 1, i + 1, model, c["kCoul"])
        s.set_coul_model(lib.EmDee_coul_sf())
        s.upload("charges", c["Q"])
        s.upload("coordinates", c["R"])
        s.upload("box", np.array([c["L"]]))
        s.random_momenta(c["kB"] * c["Temp"], True, 99)
        return s
    sp, so = both(make)
    assert sp.md.DoF == so.md.DoF == 3 * 750 + 6 * 500 - 3
    for _ in range(5):
        for s in (sp, so):
            s.boost(1.0, 0.0, 0.25)
            s.displace(1.0, 0.0, 0.5)
            s.boost(1.0, 0.0, 0.25)
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
    _close(sp.download("centersOfMass", (1250, 3)), so.download("centersOfMass", (1250, 3)), 1e-9, "centersOfMass")
    _close(sp.download("momenta"), so.download("momenta"), 1e-8, "momenta")
    _compare_scalars(sp, so, 1e-8)
    sp.finalize(), so.finalize()


def test_replicated_spce_rigid_nve_properties():
    """Size-independent properties on a replicated SPC/E box (4^3 x 2250 = 144 000 atoms on the GPU; the emulator run
    keeps one replica): n^3-replica identity of the energies, then rigid-body NVE on the device -- total energy and
    linear momentum conserved, every molecule still rigid, centres of mass consistent with the member coordinates."""
    import os
    n = int(os.environ.get("EMDEE_TEST_REPLICAS", "4"))
    f = lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0)
    s1, c1 = cm.spce_sample_system(cm.product(), f)
    U1, W1 = s1.md.Energy.Potential, s1.md.Virial.Total
    s1.finalize()
    s, c = cm.spce_sample_system(cm.product(), f, replicas=n)
    assert cm.rel(s.md.Energy.Potential, n ** 3 * U1) < 1e-10 and cm.rel(s.md.Virial.Total, n ** 3 * W1) < 1e-9
    N, nb = c["N"], c["N"] // 3
    s.random_momenta(c["kB"] * c["Temp"], True, 2024)
    R0 = s.download("coordinates")
    d0 = np.linalg.norm(R0[0::3] - R0[1::3], axis=1)
    E0 = s.md.Energy.Potential + s.md.Kinetic.Total
    for _ in range(10):
        s.boost(1.0, 0.0, 0.5)
        s.displace(1.0, 0.0, 1.0)
        s.boost(1.0, 0.0, 0.5)
    E1 = s.md.Energy.Potential + s.md.Kinetic.Total
    assert abs(E1 - E0) < 5e-4 * s.md.Kinetic.Total
    R, P = s.download("coordinates"), s.download("momenta")
    assert np.abs(np.linalg.norm(R[0::3] - R[1::3], axis=1) - d0).max() < 1e-10          # O-H bonds still rigid
    assert np.abs(P.sum(axis=0)).max() < 1e-9 * np.abs(P).sum()                          # no net momentum
    m = c["mass"][c["atomType"] - 1]
    com = (m[:, None] * R).reshape(nb, 3, 3).sum(axis=1) / m.reshape(nb, 3).sum(axis=1)[:, None]
    assert np.abs(s.download("bodycoord", (nb, 3)) - com).max() < 1e-9
    K = 0.5 * (P ** 2 / m[:, None]).sum()
    assert cm.rel(K, s.md.Kinetic.Total) < 1e-10                                         # atom momenta carry the same kinetic energy
    q = s.download("quaternions", (nb, 4))
    assert np.abs((q ** 2).sum(axis=1) - 1.0).max() < 1e-12
    s.finalize()


def test_free_rotor_long_exact_rotation():
    """One isolated body, many periods of torque-free motion in single calls: Jacobi / Carlson code paths with period
    jumps (src/ArBee.f90:262-268) on the device against the oracle."""
    from test_oracle_rigid import free_rotor
    for seed, t in ((7, 37.0), (8, 120.0), (9, 3.0)):
        sp, _ = free_rotor(cm.product(), mode=0, seed=seed)
        so, _ = free_rotor(cm.oracle(), mode=0, seed=seed)
        for s in (sp, so):
            s.displace(1.0, 0.0, t)
            s.displace(1.0, 0.0, 0.5 * t)
        assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
        _close(sp.download("quaternions", (1, 4)), so.download("quaternions", (1, 4)), 1e-9, "q")
        _close(sp.download("angmom", (1, 3)), so.download("angmom", (1, 3)), 1e-9, "omega")
        sp.finalize(), so.finalize()
