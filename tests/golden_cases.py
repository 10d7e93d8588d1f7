"""Turns a row of tests/golden/make_model_fixtures.py's case tables into C-ABI calls on any library (CPU oracle, CUDA
product, emulator), so that every library is asserted against the SAME committed numbers of tests/golden/model_golden.json
-- numbers produced by the independent numpy evaluator (tests/golden/numpy_models.py), not by either C++ restatement."""
import json
import os
import sys

import numpy as np

import common as cm

sys.path.insert(0, cm.GOLDEN)
from make_model_fixtures import LJ_CASES, LJ_KCOUL, SPCE_CASES  # noqa: E402

GOLDEN = json.load(open(os.path.join(cm.GOLDEN, "model_golden.json")))
ALL_CASES = list(LJ_CASES) + list(SPCE_CASES)


def _model(lib, spec):
    return getattr(lib, "EmDee_" + spec[0])(*spec[1:])


def _wrap(lib, model, mod):
    return model if mod is None else getattr(lib, "EmDee_" + mod[0])(model, *mod[1:])


def _relayer_once(s):
    """EmDee_layer_based_parameters must come after the model setters and before the first upload"""
    s.layer_based_parameters(s_Rc[id(s)], [0], [1])


s_Rc = {}


def build(lib, name, threads=2):
    """-> (system, fixture dict, mvv2e-free golden row)"""
    if name in LJ_CASES:
        case = LJ_CASES[name]
        c = cm.load_fixture("NIST_lj_sample")
        eps, sig = float(c["epsilon"][0]) / c["mvv2e"], float(c["sigma"][0])
        s = lib.system(threads, 1, c["Rc"], c["Rs"], c["N"], c["atomType"], c["mass"], None)
        base = lib.EmDee_pair_lj_cut(eps, sig) if case["pair"][0] == "lj" else lib.EmDee_pair_softcore_cut(eps, sig, case["pair"][1])
        charged = "coul" in case
        s.set_pair_model(1, 1, _wrap(lib, base, case.get("mod")), LJ_KCOUL if charged else 0.0)
        if charged:
            s.set_coul_model(_wrap(lib, _model(lib, case["coul"]), case.get("cmod")))
        if case.get("relayer"):
            s.layer_based_parameters(c["Rc"], [0], [1])
        N = c["N"]
        s.upload("charges", np.where(np.arange(N) % 2 == 0, 0.5, -0.5) if charged else np.zeros(N))
        s.upload("box", np.array([c["L"]]))
        s.upload("coordinates", c["R"])
        return s, c
    case = SPCE_CASES[name]
    c = cm.load_fixture("NIST_spce_sample")
    eps = c["epsilon"] / c["mvv2e"]
    s = lib.system(threads, 1, c["Rc"], c["Rs"], c["N"], c["atomType"], c["mass"], c["molecule"])
    params = []
    for i in range(2):
        if i == 1 and "h_lj" in case:
            params.append(case["h_lj"])
            model = lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(*case["h_lj"]))
        elif eps[i] == 0.0:
            params.append(None)
            model = lib.EmDee_pair_none()
        else:
            params.append((eps[i], c["sigma"][i]))
            model = lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[i], c["sigma"][i]))
        s.set_pair_model(i + 1, i + 1, model, c["kCoul"])
    if case.get("cross"):   # the Lorentz-Berthelot cross pair, set explicitly WITH the modifier
        (e0, s0), (e1, s1) = params
        s.set_pair_model(1, 2, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(float(np.sqrt(e0 * e1)), 0.5 * (s0 + s1))), c["kCoul"])
    s.set_coul_model(_wrap(lib, _model(lib, case["coul"]), case.get("cmod")))
    if case.get("relayer"):
        s.layer_based_parameters(c["Rc"], [0], [1])
    s.upload("charges", c["Q"])
    s.upload("coordinates", c["R"])
    s.upload("box", np.array([c["L"]]))
    return s, c


def check(lib, name, tol=2e-10, ftol=1e-9):
    """asserts the library's energies, virials and the forces of the first five atoms against the golden row"""
    g = GOLDEN[name]
    s, c = build(lib, name)
    md = s.md
    scale = max(abs(g["Epair"]), abs(g["Ecoul"]), abs(g["W"]), abs(g["Wbody"]), 1e-300)
    got = dict(Epair=md.Energy.Dispersion, Ecoul=md.Energy.Coulomb, Wbody=md.Virial.Body if name in SPCE_CASES else 0.0)
    got["W"] = md.Virial.Total - got["Wbody"]
    for k in ("Epair", "Ecoul", "W", "Wbody"):
        assert abs(got[k] - g[k]) <= tol * scale, f"{name}: {k} = {got[k]!r}, golden {g[k]!r}"
    F = s.download("forces")[:5]
    Fg = np.array(g["F"])
    assert np.abs(F - Fg).max() <= ftol * max(np.abs(Fg).max(), 1e-300), f"{name}: forces {F} vs {Fg}"
    s.finalize()
