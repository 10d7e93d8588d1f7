"""EmDee_rdf in the oracle (reference src/EmDeeCode.f90:1281-1395): counts over the existing half neighbor list,
normalised per shell. The reference's tests hold no expected values for it ("parity unpinned"); here the
restatement is checked against a brute-force O(N^2) histogram of the same configuration, and the list-based
quirks (excluded pairs are invisible, nothing beyond the list range) are pinned down as behaviour."""
import numpy as np
import pytest

import common as cm


def brute_counts(R, L, types, ti, tj, Rc, bins):
    d = R[:, None, :] - R[None, :, :]
    d -= L * np.rint(d / L)
    r = np.sqrt((d * d).sum(-1))
    iu = np.triu_indices(len(R), 1)
    a, b, rr = types[iu[0]], types[iu[1]], r[iu]
    sel = (((a == ti) & (b == tj)) | ((a == tj) & (b == ti))) & (rr < Rc)
    return np.bincount((rr[sel] * bins / Rc).astype(int), minlength=bins)[:bins]


def normalise(counts, Rc, L, bins, Ni, Nj, same):
    b = np.arange(1, bins + 1)
    g = counts / (4.188790204786391 * (Rc / L / bins) ** 3) / (3 * b * (b - 1) + 1) / (Ni * Nj)
    return 2 * g if same else g


def two_type_system(lib):
    R, L = cm.fcc_lj_box(5, rho=0.7, jitter=0.12, seed=9)
    N = R.shape[0]
    types = (np.arange(N) % 2 + 1).astype(np.int32)
    s = lib.system(2, 1, 2.5, 0.5, N, types, None, None)
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
    s.set_pair_model(2, 2, lib.EmDee_pair_lj_cut(0.8, 0.9), 0.0)
    return s, R, L, types


def test_rdf_matches_brute_force_histogram():
    lib = cm.oracle()
    s, R, L, types = two_type_system(lib)
    s.upload("box", [L])
    s.upload("coordinates", R)
    bins, Rc = 40, 2.9                      # inside the list range Rc + skin = 3.0
    g = s.rdf(bins, Rc, [1, 1, 2], [1, 2, 2])
    n1, n2 = int((types == 1).sum()), int((types == 2).sum())
    for row, (ti, tj, Ni, Nj) in enumerate([(1, 1, n1, n1), (1, 2, n1, n2), (2, 2, n2, n2)]):
        ref = normalise(brute_counts(R, L, types, ti, tj, Rc, bins), Rc, L, bins, Ni, Nj, ti == tj)
        assert np.allclose(g[row], ref, rtol=1e-12, atol=0.0), (ti, tj)
    assert g[:, -5:].mean() == pytest.approx(1.0, abs=0.25)      # g -> 1 at large r for a disordered box
    s.finalize()


def test_rdf_sees_only_listed_pairs():
    lib = cm.oracle()
    s, R, L, types = two_type_system(lib)
    for i in range(1, 60, 2):
        s.ignore_pair(i, i + 2)             # same-type pairs (i, i+2): excluded from the list, hence from g
    s.upload("box", [L])
    s.upload("coordinates", R)
    bins, Rc = 30, 2.9
    g = s.rdf(bins, Rc, [1], [1])
    n1 = int((types == 1).sum())
    full = brute_counts(R, L, types, 1, 1, Rc, bins)
    d = R[0:59:2] - R[2:61:2]
    d -= L * np.rint(d / L)
    rr = np.sqrt((d * d).sum(-1))
    missing = np.bincount((rr[rr < Rc] * bins / Rc).astype(int), minlength=bins)[:bins]
    assert missing.sum() > 0
    assert np.allclose(g[0], normalise(full - missing, Rc, L, bins, n1, n1, True), rtol=1e-12)
    # beyond the list range nothing is counted
    g_far = s.rdf(10, 4.0, [1], [1])
    counts_far = g_far[0] * (4.188790204786391 * (4.0 / L / 10) ** 3) * (3 * np.arange(1, 11) * np.arange(0, 10) + 1) * n1 * n1 / 2
    brute_far = brute_counts(R, L, types, 1, 1, 4.0, 10)
    assert counts_far[-1] < brute_far[-1]
    s.finalize()


def test_rdf_rejects_bad_type():
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys; sys.path.insert(0, 'tests')
        import common as cm
        lib = cm.oracle()
        R, L = cm.fcc_lj_box(4, rho=0.7)
        s = lib.system(1, 1, 2.5, 0.5, R.shape[0], None, None, None)
        s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
        s.upload('box', [L]); s.upload('coordinates', R)
        s.rdf(10, 2.0, [2], [1])
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=cm.ROOT)
    assert r.returncode == 1 and "radial distribution calculation" in r.stderr
