"""The CPU oracle against golden numbers that did NOT come from it: tests/golden/model_golden.json is produced by an
independent O(N^2) numpy restatement of the reference's model files and setter order (tests/golden/numpy_models.py,
generator tests/golden/make_model_fixtures.py; that generator also reproduces the five SPC/E rows of SURVEY.md's appendix).
Covers every pair modifier, pair_softcore_cut, every cutoff Coulomb model (rows a11/a12 of SURVEY section 8) and the
setter-order quirks Q1 (coul_sf loses its shifts), Q1b (coul_damped_smoothed loses its switch), Q3b (mixing drops modifiers).
The CUDA product is held to the same rows in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import common as cm
import golden_cases as gc


@pytest.mark.parametrize("name", gc.ALL_CASES)
def test_oracle_matches_numpy_golden(name):
    gc.check(cm.oracle(), name)


def test_generator_is_deterministic_and_matches_the_committed_file():
    """re-runs two cheap rows of the generator: the committed JSON is what the script produces"""
    import make_model_fixtures as mk
    lj = mk.load("NIST_lj_sample")
    for name in ("lj_shifted_force", "coul_damped_square_smoothed"):
        row = mk.lj_case(lj, mk.LJ_CASES[name])
        for k in ("Epair", "Ecoul", "W"):
            assert row[k] == gc.GOLDEN[name][k]
        assert np.array_equal(np.array(row["F"]), np.array(gc.GOLDEN[name]["F"]))


def test_quirks_are_visible_in_the_golden_numbers():
    g = gc.GOLDEN
    # Q1: coul_sf set through the plain setter evaluates like the bare truncated 1/r ...
    assert g["coul_sf_literal"]["Ecoul"] == g["coul_cut"]["Ecoul"]
    # ... and like shifted_force(coul_cut) once EmDee_layer_based_parameters has re-run cutoff_setup
    assert abs(g["coul_sf_relayered"]["Ecoul"] - g["shifted_force_coul_cut"]["Ecoul"]) < 1e-12
    # Q1b: coul_damped_smoothed as set by the plain setter never switches
    assert g["coul_damped_smoothed_literal"]["Ecoul"] == g["coul_damped"]["Ecoul"]
    assert g["coul_damped_smoothed_relayered"]["Ecoul"] != g["coul_damped"]["Ecoul"]
    # Q3b: the auto-mixed cross pair has no modifier; setting it explicitly changes the dispersion energy
    assert g["spce_two_lj_types_mixed"]["Epair"] != g["spce_two_lj_types_explicit_cross"]["Epair"]
