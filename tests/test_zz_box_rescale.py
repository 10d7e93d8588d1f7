"""Box changes after initialisation (what a barostat does): reference src/EmDeeCode.f90:820-829 only marks the
results stale; the cell grid is re-derived from the new box at the next list rebuild
(src/neighbor_lists.f90:184-195). The scenario rescales box and coordinates so that the number of cells per
dimension goes 7 -> 8 -> 6, with dynamics in between, and also changes the box alone (coordinates untouched,
list NOT rebuilt: the stale list is used with the new box length, as in the reference).

The CPU test runs the scenario on the two oracle builds (sanity of the scenario itself); the GPU test holds the
product to the usual bars against the strict oracle. File name sorts last on purpose.
"""
import numpy as np
import pytest

import common as cm


def scenario(lib, replay=None):
    """replay: snapshots of an earlier run; its rescaled coordinates are uploaded instead of this run's own, so
    that both libraries see bit-identical configurations right after each rescale."""
    R, L = cm.fcc_lj_box(6, rho=0.8, jitter=0.05, seed=77)
    N = R.shape[0]
    s = lib.system(2, 1, 2.5, 0.3, N, None, None, None)
    s.set_pair_model(1, 1, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(1.0, 1.0)), 0.0)
    s.upload("box", [L])
    s.upload("coordinates", R)
    s.random_momenta(1.0, True, 4321)
    snaps = []

    def snap(tag):
        s.compute_forces()         # EmDee_download takes md by value: only this call reports energies/Builds
        F = s.download("forces")
        snaps.append(dict(tag=tag, U=s.md.Energy.Potential, W=s.md.Virial.Total, F=F, pairs=s.pairs(),
                          builds=s.md.Builds, R=s.download("coordinates")))

    def run(n, dt=0.004):
        for _ in range(n):
            s.boost(1.0, 0.0, 0.5 * dt)
            s.displace(1.0, 0.0, dt)
            s.boost(1.0, 0.0, 0.5 * dt)

    snap("initial (M=7)")
    run(15)
    snap("after 15 steps")
    for f, tag in ((1.1, "expanded x1.1 (M=8)"), (0.95 / 1.1, "compressed to x0.95 (M=6)")):
        Rn = s.download("coordinates") * f
        if replay is not None:
            Rn = next(x["R"] for x in replay if x["tag"] == tag)
        L = L * f
        s.upload("box", [L])
        s.upload("coordinates", Rn)
        snap(tag)
        run(10)
        snap(tag + " + 10 steps")
    L = L * 1.001                      # box alone: no atom moved, so no rebuild; same list, new box length
    s.upload("box", [L])
    snap("box alone x1.001")
    s.finalize()
    return snaps


def test_scenario_on_the_oracle_builds():
    a = scenario(cm.oracle())
    b = scenario(cm.oracle(fast=True), replay=a)
    assert all(np.array_equal(x["R"], y["R"]) for x, y in zip(a, b) if "M=" in x["tag"] and "steps" not in x["tag"])
    assert [x["tag"] for x in a] == [x["tag"] for x in b]
    assert a[2]["builds"] > a[1]["builds"] and a[4]["builds"] > a[3]["builds"]   # each rescale forces a rebuild
    assert a[-1]["builds"] == a[-2]["builds"]                                     # the box-alone change does not
    assert a[-1]["U"] != a[-2]["U"]
    for x, y in zip(a, b):
        assert x["builds"] == y["builds"], x["tag"]
        assert x["U"] == pytest.approx(y["U"], rel=1e-9), x["tag"]
        assert x["W"] == pytest.approx(y["W"], rel=1e-9), x["tag"]


@pytest.mark.gpu
def test_box_rescale_matches_oracle_on_gpu():
    b = scenario(cm.oracle())
    a = scenario(cm.product(), replay=b)
    for x, y in zip(a, b):
        tag = x["tag"]
        assert x["builds"] == y["builds"], tag
        # trajectories of the two libraries drift apart at rounding level: compare pair sets only while the
        # coordinates are still identical, forces/totals at the usual bars scaled by that drift
        drift = np.abs(x["R"] - y["R"]).max()
        if drift == 0.0:
            assert np.array_equal(x["pairs"], y["pairs"]), tag
        tol = 1e-10 + 1e3 * drift
        assert cm.rel_force_error(x["F"], y["F"]) <= tol, tag
        assert abs(x["U"] - y["U"]) <= max(1e-12, 1e2 * drift) * abs(y["U"]), tag
        assert abs(x["W"] - y["W"]) <= max(1e-12, 1e2 * drift) * max(abs(y["W"]), abs(y["U"])), tag
