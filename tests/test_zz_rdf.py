"""EmDee_rdf on the device (k_rdf over the resident neighbor list) against the oracle. The kernel was written after
the last GPU session of round 1 (logic verified through the emulator, tests/test_emulated_kernels.py); file name
sorts late on purpose."""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import COUL_VARIANTS, _two_type_system, both

pytestmark = pytest.mark.gpu


def test_rdf_matches_oracle():
    """EmDee_rdf from the resident list (k_rdf) against the oracle's restatement of reference
    src/EmDeeCode.f90:1281-1395: integer pair counts, so g must agree to rounding of the normalisation."""
    sp, so = both(lambda lib: _two_type_system(lib, COUL_VARIANTS["coul_sf"]))
    for s in (sp, so):
        s.random_momenta(0.9, True, 777)
    for _ in range(12):                      # a few steps: the list goes stale (no rebuild) and is then rebuilt
        for s in (sp, so):
            s.boost(1.0, 0.0, 0.002)
            s.displace(1.0, 0.0, 0.004)
            s.boost(1.0, 0.0, 0.002)
        gp = sp.rdf(50, 2.5, [1, 1, 2], [1, 2, 2])
        go = so.rdf(50, 2.5, [1, 1, 2], [1, 2, 2])
        assert gp.shape == go.shape and np.allclose(gp, go, rtol=1e-12, atol=0.0)
    gp, go = sp.rdf(2000, 2.85, [2], [1]), so.rdf(2000, 2.85, [2], [1])      # many bins, beyond Rc into the skin
    assert np.allclose(gp, go, rtol=1e-12, atol=0.0) and gp.sum() > 0
    gp, go = sp.rdf(20000, 2.0, [1, 2], [1, 2]), so.rdf(20000, 2.0, [1, 2], [1, 2])   # histogram too big for smem
    assert np.allclose(gp, go, rtol=1e-12, atol=0.0)
    sp.finalize(), so.finalize()
