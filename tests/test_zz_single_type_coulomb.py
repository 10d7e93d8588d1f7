"""One atom type, charged atoms, plain pair_lj_cut + coul_sf: the combination BASELINE.json configs[4] names
("synthetic LJ + coul_sf"). In the product it selects a compile-time specialised force kernel
(k_pair_forces<LJ, NONE, COUL_SF, NONE, single type>), distinct from the generic multi-type kernel the other
Coulomb tests exercise. Also run through EmDee_layer_based_parameters, which restores coul_sf's shifts (Q1).
File name sorts late on purpose (added after the last GPU session of round 1).
"""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import assert_state_parity, both

pytestmark = pytest.mark.gpu


def _system(lib, restore_shifts):
    R, L = cm.fcc_lj_box(7, rho=0.8442, jitter=0.06, seed=21)   # 1372 atoms
    N = R.shape[0]
    Q = np.where(np.arange(N) % 2 == 0, 0.5, -0.5)
    Q[::7] = 0.0                                                # some neutral atoms
    s = lib.system(2, 1, 2.5, 0.3, N, None, None, None)
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 1.0)
    s.set_coul_model(lib.EmDee_coul_sf())
    if restore_shifts:
        s.layer_based_parameters(2.5, [0], [1])   # after the models: re-runs cutoff_setup on them
    s.upload("charges", Q)
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    return s


@pytest.mark.parametrize("restore_shifts", [False, True])
def test_single_type_lj_coul_sf(restore_shifts):
    sp, so = both(lambda lib: _system(lib, restore_shifts))
    assert so.md.Energy.Coulomb != 0.0
    assert_state_parity(sp, so)
    for s in (sp, so):                       # virial-only mode goes through the COMPUTE=false instantiation
        s.md.Options.Compute = False
        s.compute_forces()
    assert_state_parity(sp, so, energies=False)
    for s in (sp, so):
        s.md.Options.Compute = True
        s.random_momenta(1.0, True, 99)
    for _ in range(20):
        for s in (sp, so):
            s.boost(1.0, 0.0, 0.002)
            s.displace(1.0, 0.0, 0.004)
            s.boost(1.0, 0.0, 0.002)
    assert sp.md.Builds == so.md.Builds
    assert abs(sp.md.Energy.Potential - so.md.Energy.Potential) <= 1e-9 * abs(so.md.Energy.Potential)
    assert abs(sp.md.Energy.Coulomb - so.md.Energy.Coulomb) <= 1e-9 * max(abs(so.md.Energy.Coulomb), 1.0)
    sp.finalize(), so.finalize()
