"""Reciprocal-space Ewald solver of the CPU oracle (restated from reference src/kspace_ewald.f90, src/modelClass_kspace.f90,
src/coul_long.f90, src/EmDeeData.f90:689-700). The reference's own Ewald test is disabled (test/test_kspace_ewald.bak)
and test_coul_long.f90 carries no expected values: "parity unpinned". Pinned here by the Madelung constant of rock salt,
by F = -dU/dR on a distorted crystal, and by the equivalence of the two routes that take the smooth part of excluded
intramolecular pairs back out (rigid bodies at initialization vs bond/angle structures every step)."""
import numpy as np
import pytest

import common as cm

MADELUNG_NACL = 1.7475645946331822


def rock_salt(lib, ncell=4, Rc=2.8, skin=0.3, accuracy=1e-6, R=None, kCoul=1.0):
    g = np.arange(2 * ncell)
    grid = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    sign = 1 - 2 * (grid.sum(axis=1) % 2)
    L = 2.0 * ncell
    R0 = grid.astype(float) + 0.25
    if R is None:
        R = R0
    N = len(R0)
    types = np.where(sign > 0, 1, 2).astype(np.int32)
    s = lib.system(2, 1, Rc, skin, N, types, np.array([1.0, 1.0]), None)
    s.set_pair_model(1, 1, lib.EmDee_pair_none(), kCoul)
    s.set_pair_model(2, 2, lib.EmDee_pair_none(), kCoul)
    s.set_coul_model(lib.EmDee_coul_long())
    s.set_kspace_model(lib.EmDee_kspace_ewald(accuracy))
    s.upload("charges", sign.astype(float))
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    return s, R0, L


def test_madelung_constant_of_rock_salt():
    s, R0, L = rock_salt(cm.oracle())
    N = len(R0)
    assert cm.rel(s.md.Energy.Coulomb, -MADELUNG_NACL * N / 2) < 5e-6      # requested accuracy: 1e-6
    assert cm.rel(s.md.Virial.Total, s.md.Energy.Coulomb) < 1e-14       # W(long) closes the 1/r virial theorem (1251)
    assert np.abs(s.download("forces")).max() < 1e-6                      # perfect crystal: no net force
    s.finalize()
    s2, _, _ = rock_salt(cm.oracle(), kCoul=2.5)                          # the Coulomb constant scales everything
    assert cm.rel(s2.md.Energy.Coulomb, -2.5 * MADELUNG_NACL * N / 2) < 5e-6
    s2.finalize()
    s3, _, _ = rock_salt(cm.oracle(), accuracy=1e-9)                      # floor: the reference's erfc is Abramowitz-Stegun 7.1.26 (1.5e-7 abs)
    assert cm.rel(s3.md.Energy.Coulomb, -MADELUNG_NACL * N / 2) < 2e-6
    s3.finalize()


def test_ewald_forces_are_the_gradient_of_the_total_energy():
    lib = cm.oracle()
    rng = np.random.default_rng(5)
    s0, R0, L = rock_salt(lib)
    s0.finalize()
    R = R0 + rng.normal(scale=0.08, size=R0.shape)
    s, _, _ = rock_salt(lib, R=R)
    F = s.download("forces")
    s.finalize()
    h = 1e-5

    def U(Rx):
        t, _, _ = rock_salt(lib, R=Rx)
        u = t.md.Energy.Potential
        t.finalize()
        return u

    for a, x in ((0, 0), (17, 1), (200, 2)):
        Rp, Rm = R.copy(), R.copy()
        Rp[a, x] += h
        Rm[a, x] -= h
        fd = -(U(Rp) - U(Rm)) / (2 * h)
        # the real-space term is truncated at Rc with accuracy 1e-6: its tiny jump limits the agreement
        assert abs(fd - F[a, x]) < 2e-4 * max(1.0, abs(F[a, x])), (a, x, fd, F[a, x])


def _water(lib, mode):
    """The NIST SPC/E sample (750 waters) with coul_long + kspace_ewald(1e-4): waters as rigid bodies ("rigid"), or as free
    atoms tied by bond / angle structures with `none` models ("structures")."""
    c = cm.load_fixture("NIST_spce_sample")
    N = c["N"]
    mol = c["molecule"]
    s = lib.system(2, 1, c["Rc"], c["Rs"], N, c["atomType"], c["mass"], mol if mode == "rigid" else None)
    eps = c["epsilon"] / c["mvv2e"]
    for i in range(2):
        model = lib.EmDee_pair_none() if eps[i] == 0 else lib.EmDee_pair_lj_cut(eps[i], c["sigma"][i])
        s.set_pair_model(i + 1, i + 1, model, c["kCoul"])
    s.set_coul_model(lib.EmDee_coul_long())
    s.set_kspace_model(lib.EmDee_kspace_ewald(1e-4))
    if mode == "structures":
        for m in range(N // 3):
            o = 3 * m + 1
            lib.EmDee_add_bond(s.md, o, o + 1, lib.EmDee_bond_none())
            lib.EmDee_add_bond(s.md, o, o + 2, lib.EmDee_bond_none())
            lib.EmDee_add_angle(s.md, o + 1, o, o + 2, lib.EmDee_angle_none())
    s.upload("charges", c["Q"])
    s.upload("coordinates", c["R"])
    s.upload("box", np.array([c["L"]]))
    return s, c


def test_rigid_and_structure_discounts_agree():
    """reference test/test_coul_long.f90:36-50 (rigid SPC/E, coul_long + kspace_ewald(1e-4)) against the same water with
    bond/angle structures instead of bodies: both remove the erf part of the three intramolecular pairs and the self term."""
    lib = cm.oracle()
    a, c = _water(lib, "rigid")
    b, _ = _water(lib, "structures")
    assert cm.rel(a.md.Energy.Coulomb, b.md.Energy.Coulomb) < 1e-11
    assert cm.rel(a.md.Energy.Dispersion, b.md.Energy.Dispersion) < 1e-13
    # intermolecular forces agree; the structures route also carries the (internal, net-zero) discount forces
    Fa, Fb = a.download("forces"), b.download("forces")
    net = lambda F: F.reshape(-1, 3, 3).sum(axis=1)
    assert np.abs(net(Fa) - net(Fb)).max() < 1e-10 * np.abs(net(Fa)).max()
    # Ewald total vs. the damped, smoothed real-space sum (alpha = 0.2/A, Rc = 10 A): the same physics to a few percent
    d, _ = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0),
                                 pair_factory=lambda l, i, e, sg: l.EmDee_pair_none() if e == 0 else l.EmDee_pair_lj_cut(e, sg))
    assert abs(a.md.Energy.Coulomb - d.md.Energy.Coulomb) < 0.05 * abs(a.md.Energy.Coulomb)
    for s in (a, b, d):
        s.finalize()
