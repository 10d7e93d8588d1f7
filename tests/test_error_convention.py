"""Error behaviour of the boundary: the reference has no return codes; misuse prints `Error in <task>: <msg>.` on
stderr and exits with status 1 (reference src/global.f90:51-56), and ordering constraints are enforced that way
(src/EmDeeCode.f90:323-325, 460-462, 853, 872, 879-880). Each misuse below is issued, in a subprocess, to the
product's host shim (abi.cpp on the device-free stub engine) and to the oracle: both must exit 1 with the SAME line.
"""
import subprocess
import sys
import textwrap

import pytest

import common as cm

PRELUDE = """
import sys
sys.path.insert(0, {tests!r})
import numpy as np
import common as cm
lib = cm.hoststub() if {which!r} == "stub" else cm.oracle()
R, L = cm.fcc_lj_box(4, rho=0.7)
N = R.shape[0]
types = (np.arange(N) % 2 + 1).astype(np.int32)
def fresh(layers=1):
    return lib.system(1, layers, 2.5, 0.3, N, types, None, None)
def ready(layers=1):
    s = fresh(layers)
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
    s.upload("box", [L]); s.upload("coordinates", R)
    return s
"""

CASES = {
    "pair model after initialisation": "s = ready(); s.set_pair_model(2, 2, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)",
    "coul model after initialisation": "s = ready(); s.set_coul_model(lib.EmDee_coul_sf())",
    "charges after initialisation": "s = ready(); s.upload('charges', np.zeros(N))",
    "layer parameters after initialisation": "s = ready(); s.layer_based_parameters(2.0, [1], [1])",
    "momenta before initialisation": "s = fresh(); s.upload('momenta', np.zeros((N, 3)))",
    "forces before initialisation": "s = fresh(); s.upload('forces', np.zeros((N, 3)))",
    "type index out of range": "s = fresh(); s.set_pair_model(1, 3, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)",
    "type index zero": "s = fresh(); s.set_pair_model(0, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)",
    "coul model where a pair model is due": "s = fresh(); s.set_pair_model(1, 1, lib.EmDee_coul_sf(), 0.0)",
    "pair model where a coul model is due": "s = fresh(); s.set_coul_model(lib.EmDee_pair_lj_cut(1.0, 1.0))",
    "null model": "s = fresh(); s.set_pair_model(1, 1, None, 0.0)",
    "null coul model": "s = fresh(); s.set_coul_model(None)",
    "kspace model of the wrong family": "s = fresh(); s.set_kspace_model(lib.EmDee_coul_cut())",
    "modifier on a non-model": "lib.EmDee_shifted(lib.EmDee_kspace_ewald(1e-4))",
    "softcore lambda out of range": "s = fresh(); s.set_pair_model(1, 1, lib.EmDee_pair_softcore_cut(1.0, 1.0, 1.5), 0.0)",
    "invalid upload option": "s = ready(); s.upload('velocities', np.zeros((N, 3)))",
    "invalid download option": "s = ready(); s.download('velocities', (N, 3))",
    "download coordinates before upload": "s = fresh(); s.download('coordinates')",
    "layer out of range": "s = ready(); s.switch_model_layer(2)",
    "internal cutoff above Rc": "s = fresh(2); s.layer_based_parameters(3.0, [1, 0], [1, 1])",
    "internal cutoff not positive": "s = fresh(2); s.layer_based_parameters(0.0, [0, 1], [1, 1])",
    "multimodel with a missing layer": "s = fresh(2); s.set_pair_multimodel(1, 1, [lib.EmDee_pair_lj_cut(1.0, 1.0), None], [0.0, 0.0])",
    "coul_long without kspace": ("s = fresh(); s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 1.0); "
                                 "s.set_coul_model(lib.EmDee_coul_long()); s.upload('charges', np.ones(N)); "
                                 "s.upload('box', [L]); s.upload('coordinates', R)"),
    "atom types not starting at 1": "lib.system(1, 1, 2.5, 0.3, N, types + 1, None, None)",
    "bond atom out of range": "s = fresh(); lib.EmDee_add_bond(s.md, 1, N + 1, lib.EmDee_bond_harmonic(1.0, 1.0))",
    "bond with a null model": "s = fresh(); lib.EmDee_add_bond(s.md, 1, 2, None)",
    "bond with an angle model": "s = fresh(); lib.EmDee_add_bond(s.md, 1, 2, lib.EmDee_angle_harmonic(1.0, 1.0))",
    "angle atom out of range": "s = fresh(); lib.EmDee_add_angle(s.md, 0, 1, 2, lib.EmDee_angle_harmonic(1.0, 1.0))",
    "angle with a bond model": "s = fresh(); lib.EmDee_add_angle(s.md, 1, 2, 3, lib.EmDee_bond_none())",
    "dihedral with a pair model": "s = fresh(); lib.EmDee_add_dihedral(s.md, 1, 2, 3, 4, lib.EmDee_pair_lj_cut(1.0, 1.0))",
    "dihedral atom out of range": "s = fresh(); lib.EmDee_add_dihedral(s.md, 1, 2, 3, N + 7, lib.EmDee_dihedral_none())",
    "sharing before initialisation": "a = ready(); b = fresh(); a.share_phase_space(b)",
    "sharing with different types": ("a = ready(); b = lib.system(1, 1, 2.5, 0.3, N, np.ones(N, dtype=np.int32), None, None); "
                                     "b.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0); b.upload('box', [L]); "
                                     "b.upload('coordinates', R); a.share_phase_space(b)"),
    "sharing with different masses": ("a = ready(); b = lib.system(1, 1, 2.5, 0.3, N, types, np.array([1.0, 2.0]), None); "
                                      "b.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0); b.upload('box', [L]); "
                                      "b.upload('coordinates', R); a.share_phase_space(b)"),
    "memory address of an unknown array": "s = ready(); lib.EmDee_memory_address(s.md, b'velocities')",
    "random momenta of bodies before initialisation": ("s = lib.system(1, 1, 2.5, 0.3, N, types, None, (np.arange(N) // 2 + 1).astype(np.int32)); "
                                                       "s.random_momenta(1.0, True, 1)"),
    "ewald without charges": ("s = fresh(); s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 1.0); "
                              "s.set_coul_model(lib.EmDee_coul_long()); s.set_kspace_model(lib.EmDee_kspace_ewald(1e-4)); "
                              "s.upload('box', [L]); s.upload('coordinates', R)"),
}


def start(which, body):
    code = textwrap.dedent(PRELUDE).format(tests=cm.ROOT + "/tests", which=which) + body + "\nprint('NO ERROR RAISED')\n"
    return subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cm.ROOT)


def finish(p):
    out, err = p.communicate(timeout=120)
    lines = [l for l in err.splitlines() if l.startswith("Error in")]
    return p.returncode, (lines[-1] if lines else ""), out, err


@pytest.fixture(scope="module", autouse=True)
def _built():
    cm.hoststub()
    cm.oracle()


@pytest.mark.parametrize("name", list(CASES))
def test_misuse_exits_with_the_same_message(name):
    ps, po = start("stub", CASES[name]), start("oracle", CASES[name])   # the two libraries side by side
    rc_s, msg_s, out_s, err_s = finish(ps)
    rc_o, msg_o, out_o, err_o = finish(po)
    assert rc_o == 1 and msg_o.startswith("Error in ") and msg_o.endswith("."), (rc_o, out_o, err_o)
    assert rc_s == 1, (rc_s, out_s, err_s)
    assert msg_s == msg_o
