"""The same C client (examples/lj_client.c, role of the reference's src/testc.c) is compiled once per library
and linked against it through include/emdee.h only: the drop-in claim at the link level."""
import os
import subprocess

import pytest

import common as cm

SRC = os.path.join(cm.ROOT, "examples", "lj_client.c")


def _build(libdir, libname, out):
    cmd = ["/usr/bin/gcc", "-O2", "-I", os.path.join(cm.ROOT, "include"), SRC, "-o", out, "-L", libdir,
           "-l" + libname, "-lm", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)


def _run(exe, *args):
    env = dict(os.environ, EMDEE_QUIET="1")
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    return [ln.split() for ln in r.stdout.strip().splitlines()]


def test_c_client_links_against_both_libraries_and_runs_on_the_oracle(tmp_path):
    cm.oracle()  # make sure it is built
    exe_o = str(tmp_path / "client_oracle")
    _build(os.path.dirname(cm.ORACLE_STRICT), "emdee_oracle", exe_o)
    exe_p = str(tmp_path / "client_product")
    _build(os.path.dirname(cm.api.PRODUCT_LIB), "emdee", exe_p)      # link check only (needs a GPU to run)
    rows = _run(exe_o, 1000, 20)
    assert rows[-1][:4] == ["neighbor", "list", "builds", "="] and int(rows[-1][4]) >= 1
    e0, e2 = float(rows[0][3]), float(rows[2][3])
    assert abs(e2 - e0) < 2e-3 * abs(e0)          # shifted-force LJ: total energy is conserved


@pytest.mark.gpu
def test_c_client_same_numbers_on_gpu(tmp_path):
    exe_o = str(tmp_path / "client_oracle")
    _build(os.path.dirname(cm.ORACLE_STRICT), "emdee_oracle", exe_o)
    exe_p = str(tmp_path / "client_product")
    _build(os.path.dirname(cm.api.PRODUCT_LIB), "emdee", exe_p)
    ro, rp = _run(exe_o, 4000, 40), _run(exe_p, 4000, 40)
    assert ro[-1] == rp[-1]                          # same number of list builds
    for a, b in zip(ro[:-1], rp[:-1]):
        for x, y in zip(a[1:], b[1:]):
            assert abs(float(x) - float(y)) <= 1e-9 * max(1.0, abs(float(y)))
