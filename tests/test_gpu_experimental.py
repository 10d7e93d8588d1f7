"""Opt-in kernel paths that have been written but NOT yet confirmed on a GPU. Gated twice: the `gpu` marker and
EMDEE_TEST_EXPERIMENTAL=1, so the default `pytest -m gpu` run never depends on unconfirmed code. Once a path is
confirmed its test moves to test_gpu_parity.py (as the brick and duo paths did).

    EMDEE_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import os

import numpy as np
import pytest

import common as cm
from test_gpu_parity import (COUL_VARIANTS, KTOL, PAIR_VARIANTS, _lj, _two_type_system, assert_state_parity, both)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("EMDEE_TEST_EXPERIMENTAL") != "1",
                                 reason="unconfirmed kernel paths: set EMDEE_TEST_EXPERIMENTAL=1")]


def _all_model_families():
    for variant in ("lj_cut", "lj_shifted_force", "softcore_0.7"):
        sp, so = both(lambda lib: cm.lj_sample_system(lib, PAIR_VARIANTS[variant])[0])
        assert_state_parity(sp, so)
        sp.finalize(), so.finalize()
    sp, so = both(lambda lib: cm.spce_sample_system(lib, COUL_VARIANTS["coul_damped_smoothed"])[0])
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()
    sp, so = both(lambda lib: _two_type_system(lib, COUL_VARIANTS["coul_sf"]))
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()
    # dynamics with rebuilds: the reference's 100-step known-answer run (on the emulator, where a shuffle costs
    # a fiber barrier per lane, a 20-step prefix of it)
    steps = int(os.environ.get("EMDEE_TEST_REPLAY_STEPS", "100"))
    outs, builds = [], []
    for lib in (cm.product(), cm.oracle()):
        s, c = cm.lj_sample_system(lib, _lj)
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        outs.append(cm.run_nve(s, c, steps))
        builds.append(s.md.Builds)
        s.finalize()
    if steps == 100:
        assert np.abs(outs[0] - cm.kats()["lj_cut"]).max() < KTOL
    assert np.abs(outs[0] - outs[1]).max() < KTOL and builds[0] == builds[1]


@pytest.mark.parametrize("group", [8, 16, 32])
def test_rows_path(monkeypatch, group):
    """EMDEE_ROWS=G: G lanes share one atom and read consecutive entries of its (row-major) neighbor row
    (k_transpose_rows + k_pair_forces_rows). Same parity bars as the default path; every model goes through it."""
    monkeypatch.setenv("EMDEE_ROWS", str(group))
    _all_model_families()


def test_cluster2_path(monkeypatch):
    """EMDEE_CLUSTER2=1: one warp per duo of consecutive entries, half a warp per atom, lane pairs share each
    entry of the duo's union row (k_merge_duos + k_transpose_rows + k_pair_forces_cluster2)."""
    monkeypatch.setenv("EMDEE_CLUSTER2", "1")
    _all_model_families()


@pytest.mark.parametrize("mode", [1, 2])
def test_texture_path(monkeypatch, mode):
    """EMDEE_TEX=1|2: plain single-type LJ gathers its neighbor positions through the texture front-end of L1TEX
    (k_pair_forces_tex; mode 2 alternates with LDG.E.256). Only the plain-LJ instantiation exists; every other model
    falls through to the default kernels, so the whole family list must still pass."""
    monkeypatch.setenv("EMDEE_TEX", str(mode))
    _all_model_families()
    # virial-only mode goes through the COMPUTE=false instantiation
    sp, so = both(lambda lib: cm.lj_sample_system(lib, _lj)[0])
    for s in (sp, so):
        s.md.Options.Compute = False
        s.upload("coordinates", s.download("coordinates") + 0.01)
        s.compute_forces()
    assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-12
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    sp.finalize(), so.finalize()
