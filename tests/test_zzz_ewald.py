"""coul_long + kspace_ewald on the device (k_ewald_structure / k_ewald_forces, engine_ewald.cuh, and the Ewald discount
of bonded pairs in k_bonded) against the oracle's restatement of reference src/kspace_ewald.f90 and
src/modelClass_kspace.f90, through the C ABI: the calls of reference test/test_coul_long.f90. Written after the last
GPU session of round 1 (logic verified through the emulator, tests/test_emulated_kernels.py)."""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import both
from test_oracle_ewald import MADELUNG_NACL, _water, rock_salt

pytestmark = pytest.mark.gpu


def _scalars(sp, so, rtol):
    ref = abs(so.md.Energy.Potential)
    for grp, names in (("Energy", ("Potential", "Dispersion", "Coulomb", "Bond", "Angle")), ("Virial", ("Total", "Body"))):
        for n in names:
            a, b = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
            assert abs(a - b) <= rtol * max(abs(b), ref), f"{grp}.{n}: {a!r} vs {b!r}"


def test_rock_salt_madelung_and_distorted_crystal():
    sp, R0, L = rock_salt(cm.product())
    assert cm.rel(sp.md.Energy.Coulomb, -MADELUNG_NACL * len(R0) / 2) < 5e-6
    sp.finalize()
    R = R0 + np.random.default_rng(5).normal(scale=0.08, size=R0.shape)
    sp, so = both(lambda lib: rock_salt(lib, R=R)[0])
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    _scalars(sp, so, 1e-11)
    sp.finalize(), so.finalize()


@pytest.mark.parametrize("mode", ["rigid", "structures"])
def test_spce_coul_long(mode):
    """reference test/test_coul_long.f90:36-50 (rigid SPC/E) and its flexible twin (bond / angle structures)."""
    sp, so = both(lambda lib: _water(lib, mode)[0])
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    _scalars(sp, so, 1e-11)
    if mode == "rigid":
        c = cm.load_fixture("NIST_spce_sample")
        for s in (sp, so):
            s.random_momenta(c["kB"] * c["Temp"], True, 17)
            for _ in range(3):
                s.boost(1.0, 0.0, 0.5)
                s.displace(1.0, 0.0, 1.0)
                s.boost(1.0, 0.0, 0.5)
        assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
        _scalars(sp, so, 1e-9)
    sp.finalize(), so.finalize()
