"""Bonded terms of the CPU oracle (restated from reference src/EmDeeData.f90:443-550, src/bond_harmonic.f90,
src/angle_harmonic.f90, src/EmDeeCode.f90:574-655). The reference has no test with expected values for them
("parity unpinned"): pinned here by F = -dU/dR and W = -dU/dln(lambda) finite differences on a flexible water box."""
import numpy as np
import pytest

import common as cm


def flexible_water(lib, R=None, L=None, nmol=40, bonded=True, seed=3):
    """SPC/E geometry made flexible: O-H harmonic bonds, H-O-H harmonic angle, LJ on O, coul_sf; no rigid bodies."""
    c = cm.load_fixture("NIST_spce_sample")
    n = 3 * nmol
    L0 = 15.0   # >= 2.5*(Rc + skin): the 5x5x5 cell stencil covers the cutoff sphere
    if R is None:
        rng = np.random.default_rng(seed)
        g = np.array([[i, j, k] for i in range(4) for j in range(4) for k in range(4)], dtype=float)[:nmol] * 3.0 + 0.7
        mol = np.array([[0.0, 0.0, 0.0], [0.8, 0.58, 0.0], [-0.8, 0.58, 0.0]])
        R = (g[:, None, :] + mol[None, :, :]).reshape(-1, 3) + rng.normal(scale=0.05, size=(n, 3))
        L = L0
    types = np.tile(np.array([1, 2, 2], dtype=np.int32), nmol)
    s = lib.system(2, 1, 5.0, 0.8, n, types, c["mass"], None)
    eps = c["epsilon"] / c["mvv2e"]
    s.set_pair_model(1, 1, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[0], c["sigma"][0])), c["kCoul"])
    s.set_pair_model(2, 2, lib.EmDee_pair_none(), c["kCoul"])
    s.set_coul_model(lib.EmDee_shifted_force(lib.EmDee_coul_cut()))
    if bonded:
        bond = lib.EmDee_bond_harmonic(0.9, 1.0)
        angle = lib.EmDee_angle_harmonic(0.15, np.deg2rad(109.47))
        for m in range(nmol):
            o = 3 * m + 1
            s.lib.EmDee_add_bond(s.md, o, o + 1, bond)
            s.lib.EmDee_add_bond(s.md, o, o + 2, bond)
            s.lib.EmDee_add_angle(s.md, o + 1, o, o + 2, angle)
    s.upload("charges", np.tile(np.array([-0.8476, 0.4238, 0.4238]), nmol))
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    return s, R, L


def test_bonded_forces_are_the_gradient_and_virial_the_scaling_derivative():
    lib = cm.oracle()
    s, R, L = flexible_water(lib)
    F = s.download("forces")
    U0, W0 = s.md.Energy.Potential, s.md.Virial.Total
    assert s.md.Energy.Bond > 0 and s.md.Energy.Angle > 0
    assert cm.rel(s.md.Energy.Potential, s.md.Energy.Dispersion + s.md.Energy.Coulomb + s.md.Energy.Bond + s.md.Energy.Angle) < 1e-14
    s.finalize()
    h = 1e-5

    def U(Rx, Lx):
        t, _, _ = flexible_water(lib, Rx, Lx)
        u = t.md.Energy.Potential
        t.finalize()
        return u

    for a, x in ((0, 0), (1, 2), (2, 1), (61, 0)):
        Rp, Rm = R.copy(), R.copy()
        Rp[a, x] += h
        Rm[a, x] -= h
        fd = -(U(Rp, L) - U(Rm, L)) / (2 * h)
        assert abs(fd - F[a, x]) < 1e-6 * max(1.0, abs(F[a, x])), (a, x, fd, F[a, x])
    dU = (U(R * (1 + h), L * (1 + h)) - U(R * (1 - h), L * (1 - h))) / (2 * h)
    assert abs(dU + W0) < 1e-6 * max(1.0, abs(W0))
    # bonded pairs left the neighbor list (EmDee_add_bond / add_angle call EmDee_ignore_pair)
    t, _, _ = flexible_water(lib, R, L, bonded=False)
    s2, _, _ = flexible_water(lib, R, L)
    assert lib.EmDeeX_pair_count(t.md) - lib.EmDeeX_pair_count(s2.md) == 3 * 40
    t.finalize(), s2.finalize()


def test_bonded_layer_switch_and_model_errors():
    lib = cm.oracle()
    c = cm.load_fixture("NIST_spce_sample")
    # layer 2 has bonded terms switched off (EmDee_layer_based_parameters, Bonded = [1, 0])
    s = lib.system(1, 2, 5.0, 0.8, 6, np.array([1, 2, 2, 1, 2, 2], dtype=np.int32), c["mass"], None)
    lj = lib.EmDee_pair_lj_cut(0.1, 3.1)
    s.set_pair_model(1, 1, lj, 0.0)
    s.set_pair_model(2, 2, lib.EmDee_pair_none(), 0.0)
    s.layer_based_parameters(5.0, [0, 0], [1, 0])
    lib.EmDee_add_bond(s.md, 1, 2, lib.EmDee_bond_harmonic(1.0, 0.9))
    lib.EmDee_add_angle(s.md, 2, 1, 3, lib.EmDee_angle_none())
    s.upload("box", np.array([30.0]))
    s.upload("coordinates", np.array([[1, 1, 1], [2.1, 1, 1], [1, 2.0, 1], [8, 8, 8], [9, 8, 8], [8, 9, 8]], dtype=float))
    assert s.md.Energy.Bond == pytest.approx(0.5 * 1.0 * (1.1 - 0.9) ** 2, rel=1e-12) and s.md.Energy.Angle == 0.0
    s.switch_model_layer(2)
    assert s.md.Energy.Bond == 0.0
    s.finalize()
