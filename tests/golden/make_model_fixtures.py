"""Generates tests/golden/model_golden.json with the independent numpy evaluator (numpy_models.py): one row per model /
modifier / setter-order case on the two NIST samples the reference ships, each row = {Epair, Ecoul, W (pair + coul),
Wbody, F of five atoms}. Run here (CPU, ~2 min); the JSON is committed and asserted against BOTH the CPU oracle
(tests/test_numpy_golden.py) and the CUDA product (tests/test_gpu_parity.py::test_numpy_golden_*).

Before writing anything the evaluator is checked against the five SPC/E single-point rows of SURVEY.md's appendix
(numbers produced by the survey's own probe, independent of this file and of the C++ restatements).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy_models as nm  # noqa: E402


def load(name):
    d = np.load(os.path.join(HERE, name + ".npz"))
    return {k: (d[k] if d[k].ndim else d[k].item()) for k in d.files}


# case name -> (pair wrapper applied to pair_lj_cut / softcore, coulomb model factory or None, relayer?)
# The SAME table drives the tests (tests/golden_cases.py turns a row into C-ABI calls).
LJ_CASES = {
    # --- pair models and modifiers on the NIST LJ sample (800 atoms, Rc = 3, one type), no charges
    "lj_cut": dict(pair=("lj",), mod=None),
    "lj_shifted": dict(pair=("lj",), mod=("shifted",)),
    "lj_shifted_force": dict(pair=("lj",), mod=("shifted_force",)),
    "lj_smoothed": dict(pair=("lj",), mod=("smoothed", 0.5)),
    "lj_shifted_smoothed": dict(pair=("lj",), mod=("shifted_smoothed", 0.5)),
    "lj_square_smoothed": dict(pair=("lj",), mod=("square_smoothed", 0.5)),
    "lj_shifted_square_smoothed": dict(pair=("lj",), mod=("shifted_square_smoothed", 0.5)),
    "softcore_cut": dict(pair=("softcore", 0.7), mod=None),
    "softcore_shifted_force": dict(pair=("softcore", 0.4), mod=("shifted_force",)),
    # --- every cutoff Coulomb model on the same sample with charges +-0.5 alternating by atom index, kCoul = 1.3
    "coul_cut": dict(pair=("lj",), mod=None, coul=("coul_cut",)),
    "coul_sf_literal": dict(pair=("lj",), mod=None, coul=("coul_sf",)),
    "coul_sf_relayered": dict(pair=("lj",), mod=None, coul=("coul_sf",), relayer=True),
    "coul_damped": dict(pair=("lj",), mod=None, coul=("coul_damped", 0.6)),
    "coul_damped_smoothed_literal": dict(pair=("lj",), mod=None, coul=("coul_damped_smoothed", 0.6, 0.5)),
    "coul_damped_smoothed_relayered": dict(pair=("lj",), mod=None, coul=("coul_damped_smoothed", 0.6, 0.5), relayer=True),
    "coul_damped_square_smoothed": dict(pair=("lj",), mod=None, coul=("coul_damped_square_smoothed", 0.6, 0.5)),
    "coul_square_smoothed": dict(pair=("lj",), mod=None, coul=("coul_square_smoothed", 0.5)),
    "coul_shifted_square_smoothed_literal": dict(pair=("lj",), mod=None, coul=("coul_shifted_square_smoothed", 0.5)),
    "coul_shifted_square_smoothed_relayered": dict(pair=("lj",), mod=None, coul=("coul_shifted_square_smoothed", 0.5), relayer=True),
    "shifted_force_coul_cut": dict(pair=("lj",), mod=None, coul=("coul_cut",), cmod=("shifted_force",)),
    "shifted_coul_cut": dict(pair=("lj",), mod=None, coul=("coul_cut",), cmod=("shifted",)),
    "smoothed_coul_damped": dict(pair=("lj",), mod=None, coul=("coul_damped", 0.6), cmod=("smoothed", 0.5)),
    "shifted_square_smoothed_coul_cut": dict(pair=("lj",), mod=None, coul=("coul_cut",), cmod=("shifted_square_smoothed", 0.5)),
}
LJ_KCOUL = 1.3

SPCE_CASES = {
    # --- SURVEY appendix rows (validated against the survey's numbers below) + setter-order quirks on SPC/E (2250 atoms)
    "spce_coul_sf_literal": dict(coul=("coul_sf",)),
    "spce_shifted_force_coul_cut": dict(coul=("coul_cut",), cmod=("shifted_force",)),
    "spce_coul_sf_relayered": dict(coul=("coul_sf",), relayer=True),
    "spce_coul_damped_square_smoothed": dict(coul=("coul_damped_square_smoothed", 0.2, 1.0)),
    "spce_coul_damped_smoothed_relayered": dict(coul=("coul_damped_smoothed", 0.2, 1.0), relayer=True),
    "spce_coul_damped_smoothed_literal": dict(coul=("coul_damped_smoothed", 0.2, 1.0)),
    # Q3b: both types Lennard-Jones wrapped in shifted_force; the auto-mixed O-H pair loses the modifier
    "spce_two_lj_types_mixed": dict(coul=("coul_damped", 0.2), h_lj=(0.02, 1.2)),
    # ... and gets it back when the cross pair is set explicitly
    "spce_two_lj_types_explicit_cross": dict(coul=("coul_damped", 0.2), h_lj=(0.02, 1.2), cross=True),
}

SURVEY_ROWS = {   # SURVEY.md appendix / tests/test_oracle_spce.py: mvv2e * (Coulomb E, Virial%Body, Virial%Total), Dispersion
    "spce_coul_sf_literal": (-28686.256823039643, -18718.386631384867, -23779.17571552381),
    "spce_shifted_force_coul_cut": (-6826.749756368041, -18108.414766379356, -1944.2104162198348),
    "spce_coul_sf_relayered": (-6826.749756368041, -18108.414766379356, -1944.2104162198348),
    "spce_coul_damped_square_smoothed": (-6856.821617251877, -18279.392679472843, -3783.053650262019),
    "spce_coul_damped_smoothed_relayered": (-6854.440663505941, -18277.973792886394, -3700.865208564038),
}
SURVEY_DISP = 957.9773289867705


def build_model(spec):
    return getattr(nm, spec[0])(*spec[1:])


def apply_mod(model, mod):
    return model if mod is None else getattr(nm, mod[0])(model, *mod[1:])


def lj_case(c, case):
    eps, sig = float(c["epsilon"][0]) / c["mvv2e"], float(c["sigma"][0])
    base = nm.pair_lj_cut(eps, sig) if case["pair"][0] == "lj" else nm.pair_softcore_cut(eps, sig, case["pair"][1])
    s = nm.System(c["Rc"], 1)
    charged = "coul" in case
    s.set_pair_model(1, 1, apply_mod(base, case.get("mod")), LJ_KCOUL if charged else 0.0)
    if charged:
        s.set_coul_model(apply_mod(build_model(case["coul"]), case.get("cmod")))
    if case.get("relayer"):
        s.layer_based_parameters()
    N = c["N"]
    Q = np.where(np.arange(N) % 2 == 0, 0.5, -0.5) if charged else np.zeros(N)
    r = s.evaluate(c["R"], c["L"], c["atomType"], Q)
    return dict(Epair=r["Epair"], Ecoul=r["Ecoul"], W=r["Wpair"] + r["Wcoul"], Wbody=0.0, F=r["F"][:5].tolist())


def spce_case(c, case):
    eps = c["epsilon"] / c["mvv2e"]
    s = nm.System(c["Rc"], 2)
    models = []
    for i in range(2):
        if i == 1 and "h_lj" in case:
            models.append(nm.shifted_force(nm.pair_lj_cut(*case["h_lj"])))
        elif eps[i] == 0.0:
            models.append(nm.pair_none())
        else:
            models.append(nm.shifted_force(nm.pair_lj_cut(eps[i], c["sigma"][i])))
        s.set_pair_model(i + 1, i + 1, models[i], c["kCoul"])
    if case.get("cross"):
        s.set_pair_model(1, 2, nm.shifted_force(nm.mix(models[0], models[1])), c["kCoul"])
    s.set_coul_model(apply_mod(build_model(case["coul"]), case.get("cmod")))
    if case.get("relayer"):
        s.layer_based_parameters()
    r = s.evaluate(c["R"], c["L"], c["atomType"], c["Q"], molecule=c["molecule"])
    mass = c["mass"][np.asarray(c["atomType"]) - 1]
    Wb = nm.body_virial(c["R"], c["L"], r["F"], c["molecule"], mass)
    return dict(Epair=r["Epair"], Ecoul=r["Ecoul"], W=r["Wpair"] + r["Wcoul"], Wbody=Wb, F=r["F"][:5].tolist())


def main():
    out = {}
    lj = load("NIST_lj_sample")
    for name, case in LJ_CASES.items():
        out[name] = lj_case(lj, case)
        print(f"{name:45s} Epair {out[name]['Epair']: .10e} Ecoul {out[name]['Ecoul']: .10e} W {out[name]['W']: .10e}", flush=True)
    sp = load("NIST_spce_sample")
    m = sp["mvv2e"]
    for name, case in SPCE_CASES.items():
        out[name] = spce_case(sp, case)
        print(f"{name:45s} Epair {out[name]['Epair']: .10e} Ecoul {out[name]['Ecoul']: .10e} W {out[name]['W']: .10e} "
              f"Wbody {out[name]['Wbody']: .10e}", flush=True)
        if name in SURVEY_ROWS:   # the survey's own probe: Coulomb E, Virial%Body, Virial%Total (= pair + coul + body), Dispersion
            ec, wb, wt = SURVEY_ROWS[name]
            got = (m * out[name]["Ecoul"], m * out[name]["Wbody"], m * (out[name]["W"] + out[name]["Wbody"]))
            assert abs(got[0] - ec) < 1e-6 and abs(got[1] - wb) < 1e-6 and abs(got[2] - wt) < 1e-6, (name, got, (ec, wb, wt))
            assert abs(m * out[name]["Epair"] - SURVEY_DISP) < 1e-7
    with open(os.path.join(HERE, "model_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("written", len(out), "rows; the five survey rows reproduced")


if __name__ == "__main__":
    main()
