"""Harmonic bonds and angles on the device (k_bonded, engine_bonded.cuh) against the oracle's restatement of reference
src/EmDeeData.f90:443-550: flexible water with LJ + Coulomb, with and without rigid bodies next to it, NVE steps, layers
with bonded terms switched off. Written after the last GPU session of round 1 (logic verified through the emulator)."""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import both
from test_oracle_bonded import flexible_water

pytestmark = pytest.mark.gpu


def _scalars(sp, so, rtol=1e-11):
    for grp, names in (("Energy", ("Potential", "Dispersion", "Coulomb", "Bond", "Angle")), ("Virial", ("Total",))):
        for n in names:
            a, b = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
            assert abs(a - b) <= rtol * max(abs(b), abs(so.md.Energy.Potential)), f"{grp}.{n}: {a!r} vs {b!r}"


def test_flexible_water_forces_and_dynamics():
    sp, so = both(lambda lib: flexible_water(lib)[0])
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    _scalars(sp, so)
    for s in (sp, so):
        s.random_momenta(0.0006, True, 13)
    for step in range(1, 13):
        for s in (sp, so):
            s.md.Options.Compute = (step % 3 == 0)
            s.boost(1.0, 0.0, 0.25)
            s.displace(1.0, 0.0, 0.5)
            s.boost(1.0, 0.0, 0.25)
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-8
    _scalars(sp, so, 1e-9)
    assert sp.md.Builds == so.md.Builds
    sp.finalize(), so.finalize()


def test_bonded_terms_next_to_rigid_bodies_and_layers():
    """Half of the waters are rigid bodies, the other half flexible (bonds + angle); two layers, bonded terms only in
    the first (EmDee_layer_based_parameters). The bonded forces on body members enter the rigid-body virial."""
    def make(lib):
        c = cm.load_fixture("NIST_spce_sample")
        nmol = 60
        R = c["R"][:3 * nmol].copy()
        first = R[0::3].repeat(3, axis=0)
        R = R - c["L"] * np.round((R - first) / c["L"])            # whole molecules
        types = np.tile(np.array([1, 2, 2], dtype=np.int32), nmol)
        bodies = np.repeat(np.arange(1, nmol + 1), 3).astype(np.int32)
        bodies[3 * (nmol // 2):] = 0
        s = lib.system(2, 2, 9.0, 1.0, 3 * nmol, types, c["mass"], bodies)
        eps = c["epsilon"] / c["mvv2e"]
        s.set_pair_model(1, 1, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[0], c["sigma"][0])), c["kCoul"])
        s.set_pair_model(2, 2, lib.EmDee_pair_none(), c["kCoul"])
        s.set_coul_model(lib.EmDee_shifted_force(lib.EmDee_coul_cut()))
        s.layer_based_parameters(9.0, [0, 0], [1, 0])
        bond, angle = lib.EmDee_bond_harmonic(0.9, 1.0), lib.EmDee_angle_harmonic(0.15, np.deg2rad(109.47))
        for m in range(nmol // 2, nmol):
            o = 3 * m + 1
            lib.EmDee_add_bond(s.md, o, o + 1, bond)
            lib.EmDee_add_bond(s.md, o, o + 2, bond)
            lib.EmDee_add_angle(s.md, o + 1, o, o + 2, angle)
        # one bond between a body member and a free atom: its force feeds the body virial
        lib.EmDee_add_bond(s.md, 2, 3 * (nmol // 2) + 1, lib.EmDee_bond_harmonic(0.01, 4.0))
        lib.EmDee_add_dihedral(s.md, 1, 4, 7, 10, lib.EmDee_dihedral_none())
        s.upload("charges", np.tile(np.array([-0.8476, 0.4238, 0.4238]), nmol))
        s.upload("box", np.array([c["L"]]))
        s.upload("coordinates", R)
        return s
    sp, so = both(make)
    assert np.array_equal(sp.pairs(), so.pairs())
    for layer in (1, 2):
        for s in (sp, so):
            s.switch_model_layer(layer)
        assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
        _scalars(sp, so)
        assert abs(sp.md.Virial.Body - so.md.Virial.Body) < 1e-11 * abs(so.md.Energy.Potential)
        assert (sp.md.Energy.Bond > 0) == (layer == 1)
    for s in (sp, so):
        s.switch_model_layer(1)
        s.random_momenta(0.0006, True, 77)
        for _ in range(4):
            s.boost(1.0, 0.0, 0.25)
            s.displace(1.0, 0.0, 0.5)
            s.boost(1.0, 0.0, 0.25)
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
    _scalars(sp, so, 1e-9)
    sp.finalize(), so.finalize()
