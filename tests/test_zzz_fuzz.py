"""Seeded random scenarios mixing the features that the focused tests exercise one at a time: 1-3 atom types, several
pair models and modifiers, an optional Coulomb model (including coul_long + Ewald), rigid bodies of 3-4 atoms next to
free atoms, harmonic bonds / angles between free atoms, exclusions, one or two layers with an inner cutoff, and a short
run that alternates EmDee_verlet_step with boost / displace / boost, layer switches and a coordinates re-upload.
Every scenario runs on the CPU through the kernel emulator and -- marked `gpu` -- on the device, against the oracle.
"""
import numpy as np
import pytest

import common as cm

SEEDS = list(range(24))


def build(lib, seed):
    rng = np.random.default_rng(1000 + seed)
    nside = int(rng.integers(5, 8))
    a = float(rng.uniform(1.15, 1.4))
    L = nside * a
    Rc, skin = 2.5, float(rng.uniform(0.2, 0.5))
    while L < 2.5 * (Rc + skin) + 0.01:
        nside += 1
        L = nside * a
    g = np.arange(nside) * a
    R = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + 0.5 * a + rng.uniform(-0.08, 0.08, (nside ** 3, 3))
    N = R.shape[0]
    nt = int(rng.integers(1, 4))
    types = rng.integers(1, nt + 1, N).astype(np.int32)
    types[:nt] = np.arange(1, nt + 1)                       # every type present, type 1 first
    masses = rng.uniform(1.0, 3.0, nt)
    # rigid bodies: L- or square-shaped groups of 3-4 neighboring lattice sites (never collinear: the reference's bodies
    # have six degrees of freedom), about a third of the atoms
    bodies = np.zeros(N, dtype=np.int32)
    if rng.random() < 0.7:
        b = 0
        for i in range(N - nside - 2):
            iz, iy = i % nside, (i // nside) % nside
            if iz + 1 >= nside or iy + 1 >= nside or rng.random() > 0.12:
                continue
            members = [i, i + 1, i + nside] + ([i + nside + 1] if rng.random() < 0.5 else [])
            if any(bodies[m] for m in members):
                continue
            b += 1
            bodies[members] = b
    layers = int(rng.integers(1, 3))
    s = lib.system(int(rng.integers(1, 4)), layers, Rc, skin, N, types, masses, bodies if bodies.any() else None)
    inner = float(rng.uniform(1.6, 2.2)) if layers == 2 and rng.random() < 0.6 else None
    if layers == 2:
        s.layer_based_parameters(inner if inner else Rc, [1 if inner else 0, 0], [1, int(rng.integers(0, 2))])

    def pair_model():
        e, sg = float(rng.uniform(0.5, 1.2)), float(rng.uniform(0.9, 1.1))
        base = lib.EmDee_pair_softcore_cut(e, sg, float(rng.uniform(0.5, 1.0))) if rng.random() < 0.25 else lib.EmDee_pair_lj_cut(e, sg)
        m = int(rng.integers(0, 5))
        return [base, lib.EmDee_shifted(base), lib.EmDee_shifted_force(base), lib.EmDee_smoothed(base, 0.4),
                lib.EmDee_shifted_square_smoothed(base, 0.5)][m]

    kc = float(rng.uniform(0.5, 2.0))
    for t in range(1, nt + 1):
        if layers == 2 and rng.random() < 0.5:
            s.set_pair_multimodel(t, t, [pair_model(), pair_model()], [kc, kc])
        else:
            s.set_pair_model(t, t, pair_model() if rng.random() < 0.85 else lib.EmDee_pair_none(), kc)
    if nt >= 2 and rng.random() < 0.5:
        s.set_pair_model(1, 2, pair_model(), kc)
    coul = int(rng.integers(0, 6))
    ewald = False
    if coul == 1:
        s.set_coul_model(lib.EmDee_coul_sf())
    elif coul == 2:
        s.set_coul_model(lib.EmDee_coul_damped_smoothed(0.4, 0.5))
    elif coul == 3:
        s.set_coul_model(lib.EmDee_shifted_force(lib.EmDee_coul_cut()))
    elif coul == 4 and inner is None:
        s.set_coul_model(lib.EmDee_coul_long())
        s.set_kspace_model(lib.EmDee_kspace_ewald(1e-3))
        ewald = True
    elif coul == 5:
        s.set_coul_model(lib.EmDee_coul_square_smoothed(0.6))
    if coul:
        q = rng.choice([-0.5, 0.0, 0.5], N)
        q[0] = 0.5                                           # at least one charge (Ewald refuses a neutral-by-absence system)
        s.upload("charges", q)
    # bonded structures between FREE atoms that are lattice neighbors
    free = np.where(bodies == 0)[0]
    nb_ = 0
    for i in free[:: max(1, len(free) // 12)]:
        j, k = i + 1, i + 2
        if k < N and bodies[j] == 0 and bodies[k] == 0 and (i % nside) + 2 < nside:
            s.add_bond(i + 1, j + 1, lib.EmDee_bond_harmonic(float(rng.uniform(5, 20)), a))
            if rng.random() < 0.5:
                s.add_angle(i + 1, j + 1, k + 1, lib.EmDee_angle_harmonic(float(rng.uniform(1, 4)), 2.8))
            nb_ += 1
    for _ in range(int(rng.integers(0, 6))):
        i, j = rng.integers(1, N + 1, 2)
        s.ignore_pair(int(i), int(j))
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    s.random_momenta(float(rng.uniform(0.3, 0.8)), bool(rng.integers(0, 2)), int(rng.integers(1, 10 ** 6)))
    s.md.Options.RotationMode = int(rng.integers(0, 3))
    return s, dict(layers=layers, N=N, bodies=bool(bodies.any()), ewald=ewald, script=rng.integers(0, 4, 8), dt=0.002)


def drive(s, info):
    for k, action in enumerate(info["script"]):
        s.md.Options.Compute = bool(k % 2)
        if action == 0:
            s.verlet_step(info["dt"])
        elif action == 1 and info["layers"] == 2:
            s.switch_model_layer(1 + k % 2)
            s.compute_forces()
        elif action == 2:
            s.upload("coordinates", s.download("coordinates") + 0.01 * (k + 1))
            s.compute_forces()
        else:
            s.boost(1.0, 0.0, 0.5 * info["dt"])
            s.displace(1.0, 0.0, info["dt"])
            s.boost(1.0, 0.0, 0.5 * info["dt"])
    s.md.Options.Compute = True
    s.compute_forces()


def check(product_lib, seed):
    sp, info = build(product_lib, seed)
    so, _ = build(cm.oracle(), seed)
    for stage in ("initial", "driven"):
        if stage == "driven":
            drive(sp, info)
            drive(so, info)
        assert np.array_equal(sp.pairs(), so.pairs()), (seed, stage)
        tol = 1e-10 if stage == "initial" else 1e-7
        assert cm.rel_force_error(sp.download("forces"), so.download("forces")) < tol, (seed, stage)
        ref = max(abs(so.md.Energy.Potential), abs(so.md.Virial.Total), 1.0)
        for grp, names in (("Energy", ("Potential", "Dispersion", "Coulomb", "Bond", "Angle")), ("Virial", ("Total", "Body")),
                           ("Kinetic", ("Total", "Rotational"))):
            for n in names:
                x, y = getattr(getattr(sp.md, grp), n), getattr(getattr(so.md, grp), n)
                assert abs(x - y) <= (1e-11 if stage == "initial" else 1e-7) * max(ref, abs(y)), (seed, stage, grp, n, x, y)
        assert sp.md.Builds == so.md.Builds and sp.md.DoF == so.md.DoF and sp.md.RotDoF == so.md.RotDoF
    assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-8
    assert np.abs(sp.download("momenta") - so.download("momenta")).max() < 1e-7 * max(1.0, np.abs(so.download("momenta")).max())
    sp.finalize(), so.finalize()


@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_on_emulator(seed):
    check(cm.emulated(), seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_on_gpu(seed):
    check(cm.product(), seed)
