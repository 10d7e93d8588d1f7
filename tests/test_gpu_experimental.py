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


@pytest.mark.parametrize("group", [4, 8, 16, 32])
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


def test_typed_path(monkeypatch):
    """EMDEE_TYPED=1: several atom types, pair models all pair_lj_cut (one modifier) or pair_none, Coulomb kind fixed at
    compile time (k_pair_forces_typed; with two types the table row lives in registers). SPC/E with every eligible
    Coulomb model, a three-type LJ mixture (shared-memory table), virial-only mode, ineligible systems falling back."""
    monkeypatch.setenv("EMDEE_TYPED", "1")
    for variant in ("coul_none", "coul_cut", "coul_sf", "coul_damped", "coul_damped_smoothed", "coul_damped_square_smoothed",
                    "shifted_force(coul_cut)"):                                   # the last one is not eligible: generic kernel
        sp, so = both(lambda lib: cm.spce_sample_system(lib, COUL_VARIANTS[variant])[0])
        assert_state_parity(sp, so)
        if variant == "coul_damped_square_smoothed":
            for s in (sp, so):                                                    # virial-only instantiation + rigid-body steps
                s.random_momenta(0.0005, True, 5)
                for step in range(4):
                    s.md.Options.Compute = (step == 3)
                    s.boost(1.0, 0.0, 0.5)
                    s.displace(1.0, 0.0, 1.0)
                    s.boost(1.0, 0.0, 0.5)
            assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
            assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-10
            assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-9
        sp.finalize(), so.finalize()

    def three_types(lib):
        R, L = cm.fcc_lj_box(7, rho=0.80, jitter=0.08, seed=23)
        N = R.shape[0]
        types = (np.arange(N) % 3 + 1).astype(np.int32)
        s = lib.system(2, 1, 2.5, 0.4, N, types, np.array([1.0, 2.0, 3.0]), None)
        for t, (e, sg) in enumerate(((1.0, 1.0), (0.7, 1.1), (1.3, 0.9)), start=1):
            s.set_pair_model(t, t, lib.EmDee_pair_lj_cut(e, sg), 1.0)     # plain LJ everywhere: cross pairs mix to plain LJ too
        s.set_pair_model(1, 3, lib.EmDee_pair_none(), 1.0)                # one explicit non-interacting pair
        s.set_coul_model(lib.EmDee_coul_damped(0.3))
        s.upload("charges", np.where(types == 2, 0.0, np.where(types == 1, 0.5, -0.5)))
        s.upload("box", np.array([L]))
        s.upload("coordinates", R)
        return s
    sp, so = both(three_types)
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()
    sp, so = both(lambda lib: _two_type_system(lib, COUL_VARIANTS["coul_sf"]))   # softcore present: generic kernel
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_tile_schedule_path(monkeypatch):
    """EMDEE_TILESCHED=1: plain single-type LJ; warps take the list tiles in a brick-ordered permutation built at every
    rebuild (k_tile_keys + radix sort, k_pair_forces_sched). Results are per-thread identical to the default path; the
    reductions run over a different block order, so totals agree to rounding."""
    monkeypatch.setenv("EMDEE_TILESCHED", "1")
    _all_model_families()
    sp, so = both(lambda lib: cm.lj_sample_system(lib, _lj)[0])
    for s in (sp, so):                                   # virial-only instantiation
        s.md.Options.Compute = False
        s.upload("coordinates", s.download("coordinates") + 0.01)
        s.compute_forces()
    assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-12
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    sp.finalize(), so.finalize()


def test_rec16_path(monkeypatch):
    """EMDEE_REC16=1: plain single-type LJ with 16-byte fixed-point position records (42 bits per coordinate, relative to
    the entry's build-time cell) and a cell-tagged copy of the list; separations are formed exactly in 64-bit integers
    (k_tag_list, k_refresh_rec16, k_pair_forces_rec16). The positions are quantised at 2^-41 of a cell, so forces agree
    with the oracle to ~1e-11 relative (bar 1e-10) and totals to 1e-11 (bar of this test; the default path meets 1e-12)."""
    monkeypatch.setenv("EMDEE_REC16", "1")
    sp, so = both(lambda lib: cm.lj_sample_system(lib, _lj)[0])
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-11
    assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-11
    c = cm.load_fixture("NIST_lj_sample")
    for s in (sp, so):
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
    for step in range(1, 31):                          # drift between rebuilds, rebuilds, virial-only steps
        for s in (sp, so):
            s.md.Options.Compute = (step % 5 == 0)
            s.boost(1.0, 0.0, 0.5 * c["Dt"])
            s.displace(1.0, 0.0, c["Dt"])
            s.boost(1.0, 0.0, 0.5 * c["Dt"])
    assert sp.md.Builds == so.md.Builds and sp.md.Builds >= 2
    assert np.array_equal(sp.pairs(), so.pairs())
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-9
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-9
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-8
    sp.finalize(), so.finalize()
    for variant in ("lj_shifted_force", "softcore_0.7"):   # every other model falls through to the default kernels (strict bars)
        sp, so = both(lambda lib: cm.lj_sample_system(lib, PAIR_VARIANTS[variant])[0])
        assert_state_parity(sp, so)
        sp.finalize(), so.finalize()


@pytest.mark.parametrize("group", [4, 8])
def test_rows16_path(monkeypatch, group):
    """EMDEE_REC16=1 together with EMDEE_ROWS=G: G lanes per atom over the row-major TAGGED list, 16-byte records
    (k_pair_forces_rows16). Same bars as test_rec16_path; non-LJ layers keep using the plain rows / default kernels."""
    monkeypatch.setenv("EMDEE_REC16", "1")
    monkeypatch.setenv("EMDEE_ROWS", str(group))
    sp, so = both(lambda lib: cm.lj_sample_system(lib, _lj)[0])
    assert np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-10
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-11
    assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-11
    c = cm.load_fixture("NIST_lj_sample")
    for s in (sp, so):
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
    for step in range(1, 21):
        for s in (sp, so):
            s.md.Options.Compute = (step % 5 == 0)
            s.boost(1.0, 0.0, 0.5 * c["Dt"])
            s.displace(1.0, 0.0, c["Dt"])
            s.boost(1.0, 0.0, 0.5 * c["Dt"])
    assert sp.md.Builds == so.md.Builds and np.array_equal(sp.pairs(), so.pairs())
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-9
    assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < 1e-8
    sp.finalize(), so.finalize()
    for variant in ("lj_shifted_force", "softcore_0.7"):   # these go through the plain rows path
        sp, so = both(lambda lib: cm.lj_sample_system(lib, PAIR_VARIANTS[variant])[0])
        assert_state_parity(sp, so)
        sp.finalize(), so.finalize()
