"""Pins the CPU oracle against every live known answer the reference's own tests hold for the hot
path, plus the survey's independent O(N^2) cross-check values.

Tolerance: the reference's own (test/common/declarations.f90:16) -- 1e-8 absolute.
"""
import numpy as np
import pytest

import common as cm

TOL = 1.0e-8  # reference test/common/declarations.f90:16


def _lj(lib, eps, sig):
    return lib.EmDee_pair_lj_cut(eps, sig)


def _lj_sf(lib, eps, sig):
    return lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps, sig))


def _replay(pair_factory, threads, fast=False):
    lib = cm.oracle(fast)
    s, c = cm.lj_sample_system(lib, pair_factory, threads=threads)
    s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
    step0 = c["mvv2e"] * np.array([s.md.Energy.Potential, s.md.Virial.Total])
    out = cm.run_nve(s, c, 100)
    builds = s.md.Builds
    s.finalize()
    return step0, out, builds


@pytest.mark.parametrize("threads", [1, 2, 3])
def test_kat_pair_lj_cut(threads):
    """reference test/test_pair_lj_cut.f90:46"""
    step0, out, builds = _replay(_lj, threads)
    # step-0 single point: NIST SRSW values -4.3515E+03 / -5.6867E+02 and the survey's O(N^2) probe
    assert abs(step0[0] - (-4351.5401945438725)) < TOL
    assert abs(step0[1] - (-568.6654653181746)) < TOL
    assert np.abs(out - cm.kats()["lj_cut"]).max() < TOL
    assert builds >= 2


@pytest.mark.parametrize("threads", [1, 2])
def test_kat_pair_lj_sf(threads):
    """reference test/test_pair_lj_sf.f90:46"""
    step0, out, _ = _replay(_lj_sf, threads)
    assert abs(step0[0] - (-3870.9248857840166)) < TOL
    assert abs(step0[1] - 317.5383460124453) < TOL
    assert np.abs(out - cm.kats()["lj_sf"]).max() < TOL


def test_kat_pair_lj_square_smoothed_skin1():
    """reference test/test_pair_lj_smoothed.f90:46 -- the pinned triple corresponds to a smoothing
    width of 1.0 (Rm = 2); the shipped call passes Rc-1 = 2.0."""
    _, out, _ = _replay(lambda lib, e, s: lib.EmDee_square_smoothed(lib.EmDee_pair_lj_cut(e, s), 1.0), 2)
    assert np.abs(out - cm.kats()["lj_square_smoothed_skin1"]).max() < TOL


def test_as_shipped_square_smoothed_matches_survey_probe():
    """The test AS SHIPPED (skin = Rc-1 = 2 => Rm = 1): survey's independent restatement value."""
    _, out, _ = _replay(lambda lib, e, s: lib.EmDee_square_smoothed(lib.EmDee_pair_lj_cut(e, s), 2.0), 2)
    expect = np.array([-3992.2912839461433, -207.5362125488233, -3005.2103224771195])
    assert np.abs(out - expect).max() < 1e-7


def test_fast_build_agrees_with_kat():
    """The -Ofast build (CPU timing baseline) still reproduces the pinned triple."""
    _, out, _ = _replay(_lj, 2, fast=True)
    assert np.abs(out - cm.kats()["lj_cut"]).max() < 1e-7


def test_single_point_rc4_and_softcore_lambda1():
    lib = cm.oracle()
    s, c = cm.lj_sample_system(lib, _lj, Rc=4.0)
    assert abs(s.md.Energy.Potential - (-4467.4957249479785)) < TOL
    assert abs(s.md.Virial.Total - (-1263.883371872137)) < TOL
    s.finalize()
    # pair_softcore_cut(lambda=1) has shift=0 and must equal pair_lj_cut (reference pair_softcore_cut.f90:80)
    s, c = cm.lj_sample_system(lib, lambda l, e, sg: l.EmDee_pair_softcore_cut(e, sg, 1.0))
    assert abs(s.md.Energy.Potential - (-4351.5401945438725)) < 1e-7
    assert abs(s.md.Virial.Total - (-568.6654653181746)) < 1e-7
    s.finalize()
