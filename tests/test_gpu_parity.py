"""Parity of the CUDA path (through the C ABI of emdee_b200/lib/libemdee.so) against the CPU oracle
on the same inputs, and against the reference's own known-answer triples.

Bars (north_star): neighbor pair SETS bit-exact; forces <= 1e-10 relative per atom (see
common.rel_force_error for the denominator); energy / virial totals <= 1e-12 relative.
"""
import numpy as np
import pytest

import common as cm

pytestmark = pytest.mark.gpu

FTOL = 1.0e-10   # per-atom relative force tolerance
STOL = 1.0e-12   # relative tolerance on totals
KTOL = 1.0e-8    # reference test tolerance (absolute), test/common/declarations.f90:16


def assert_state_parity(sp, so, check_pairs=True, stol=STOL, ftol=FTOL, energies=True):
    """sp: product system, so: oracle system, both already evaluated on the same configuration."""
    if check_pairs:
        pp, po = sp.pairs(), so.pairs()
        assert pp.shape == po.shape, f"pair count differs: {pp.shape[0]} vs {po.shape[0]}"
        assert np.array_equal(pp, po), "neighbor pair sets differ"
    Fp, Fo = sp.download("forces"), so.download("forces")
    err = cm.rel_force_error(Fp, Fo)
    assert err <= ftol, f"per-atom relative force error {err:.3e}"
    mp, mo = sp.md, so.md
    scale_e = max(abs(mo.Energy.Dispersion), abs(mo.Energy.Coulomb), abs(mo.Energy.Potential), 1e-300)
    scale_w = max(abs(mo.Virial.Total), abs(mo.Virial.Body), 1e-300)
    if energies:
        for name in ("Potential", "Dispersion", "Coulomb"):
            a, b = getattr(mp.Energy, name), getattr(mo.Energy, name)
            assert abs(a - b) <= stol * scale_e, f"Energy.{name}: {a!r} vs {b!r}"
    for name in ("Total", "Body"):
        a, b = getattr(mp.Virial, name), getattr(mo.Virial, name)
        assert abs(a - b) <= stol * max(scale_w, scale_e), f"Virial.{name}: {a!r} vs {b!r}"
    assert mp.DoF == mo.DoF and mp.RotDoF == mo.RotDoF


def both(builder):
    sp = builder(cm.product())
    so = builder(cm.oracle())
    return sp, so


def _lj(lib, e, s):
    return lib.EmDee_pair_lj_cut(e, s)


PAIR_VARIANTS = {
    "lj_cut": _lj,
    "lj_shifted": lambda l, e, s: l.EmDee_shifted(l.EmDee_pair_lj_cut(e, s)),
    "lj_shifted_force": lambda l, e, s: l.EmDee_shifted_force(l.EmDee_pair_lj_cut(e, s)),
    "lj_smoothed": lambda l, e, s: l.EmDee_smoothed(l.EmDee_pair_lj_cut(e, s), 0.5),
    "lj_shifted_smoothed": lambda l, e, s: l.EmDee_shifted_smoothed(l.EmDee_pair_lj_cut(e, s), 0.5),
    "lj_square_smoothed": lambda l, e, s: l.EmDee_square_smoothed(l.EmDee_pair_lj_cut(e, s), 1.0),
    "lj_shifted_square_smoothed": lambda l, e, s: l.EmDee_shifted_square_smoothed(l.EmDee_pair_lj_cut(e, s), 1.0),
    "softcore_0.7": lambda l, e, s: l.EmDee_pair_softcore_cut(e, s, 0.7),
    "softcore_0.4_shifted_force": lambda l, e, s: l.EmDee_shifted_force(l.EmDee_pair_softcore_cut(e, s, 0.4)),
}


@pytest.mark.parametrize("variant", list(PAIR_VARIANTS))
def test_nist_lj_single_point(variant):
    sp, so = both(lambda lib: cm.lj_sample_system(lib, PAIR_VARIANTS[variant])[0])
    assert_state_parity(sp, so)
    assert sp.md.Builds == so.md.Builds == 1
    sp.finalize(), so.finalize()


def test_nist_lj_reference_step0_values():
    s, c = cm.lj_sample_system(cm.product(), _lj)
    assert abs(s.md.Energy.Potential - (-4351.5401945438725)) < KTOL   # NIST SRSW: -4.3515E+03
    assert abs(s.md.Virial.Total - (-568.6654653181746)) < KTOL        # NIST SRSW: -5.6867E+02
    s.finalize()


@pytest.mark.parametrize("name,factory", [("lj_cut", _lj), ("lj_sf", PAIR_VARIANTS["lj_shifted_force"]),
                                          ("lj_square_smoothed_skin1", PAIR_VARIANTS["lj_square_smoothed"])])
def test_reference_kat_replay_on_gpu(name, factory):
    """reference test/test_pair_lj_cut.f90:46, test_pair_lj_sf.f90:46, test_pair_lj_smoothed.f90:46 (skin=1):
    100 velocity-Verlet steps with the state resident on the device, against the reference's pinned triples."""
    outs, builds = [], []
    for lib in (cm.product(), cm.oracle()):
        s, c = cm.lj_sample_system(lib, factory)
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        outs.append(cm.run_nve(s, c, 100))
        builds.append(s.md.Builds)
        if lib is cm.product():
            sp = s
        else:
            so = s
    assert np.abs(outs[0] - cm.kats()[name]).max() < KTOL, f"{outs[0]} vs {cm.kats()[name]}"
    assert np.abs(outs[0] - outs[1]).max() < KTOL
    assert builds[0] == builds[1], f"rebuild counts differ: {builds}"
    assert_state_parity(sp, so, stol=1e-10, ftol=1e-8)   # after 100 steps of chaotic dynamics
    sp.finalize(), so.finalize()


COUL_VARIANTS = {
    "coul_none": lambda l: None,
    "coul_cut": lambda l: l.EmDee_coul_cut(),
    "coul_sf": lambda l: l.EmDee_coul_sf(),
    "shifted_force(coul_cut)": lambda l: l.EmDee_shifted_force(l.EmDee_coul_cut()),
    "shifted(coul_cut)": lambda l: l.EmDee_shifted(l.EmDee_coul_cut()),
    "coul_damped": lambda l: l.EmDee_coul_damped(0.2),
    "coul_damped_smoothed": lambda l: l.EmDee_coul_damped_smoothed(0.2, 1.0),
    "coul_damped_square_smoothed": lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0),
    "coul_square_smoothed": lambda l: l.EmDee_coul_square_smoothed(1.0),
    "coul_shifted_square_smoothed": lambda l: l.EmDee_coul_shifted_square_smoothed(1.0),
    "square_smoothed(coul_damped)": lambda l: l.EmDee_square_smoothed(l.EmDee_coul_damped(0.25), 1.5),
    "smoothed(coul_cut)": lambda l: l.EmDee_smoothed(l.EmDee_coul_cut(), 2.0),
}


@pytest.mark.parametrize("variant", list(COUL_VARIANTS))
def test_spce_single_point(variant):
    """reference test/test_coul_*.f90: SPC/E, rigid bodies, LJ-sf(O) + pair_none(H) + Coulomb model."""
    sp, so = both(lambda lib: cm.spce_sample_system(lib, COUL_VARIANTS[variant])[0])
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_spce_survey_probe_value():
    s, c = cm.spce_sample_system(cm.product(), COUL_VARIANTS["coul_damped_square_smoothed"])
    m = c["mvv2e"]
    assert abs(m * s.md.Energy.Dispersion - 957.9773289867705) < 1e-7
    assert abs(m * s.md.Energy.Coulomb - (-6856.821617251877)) < 1e-6
    assert abs(m * s.md.Virial.Body - (-18279.392679472843)) < 1e-6
    assert abs(m * s.md.Virial.Total - (-3783.053650262019)) < 1e-6
    s.finalize()


def test_spce_layer_based_parameters_restores_shifts():
    def build(lib):
        orig = cm.api.System.upload
        state = {"done": False}

        def patched(self, option, array):
            if not state["done"]:
                self.layer_based_parameters(10.0, [0], [1])
                state["done"] = True
            return orig(self, option, array)

        cm.api.System.upload = patched
        try:
            return cm.spce_sample_system(lib, COUL_VARIANTS["coul_damped_smoothed"])[0]
        finally:
            cm.api.System.upload = orig

    sp, so = both(build)
    assert_state_parity(sp, so)
    assert abs(2390.057364 * sp.md.Energy.Coulomb - (-6854.440663505941)) < 1e-6
    sp.finalize(), so.finalize()


def test_spce_without_bodies_uses_exclusions():
    sp, so = both(lambda lib: cm.spce_sample_system(lib, COUL_VARIANTS["coul_damped"], bodies=False)[0])
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_spce_virial_only_mode_and_recompute():
    """Options.Compute = false (virial-only instantiation of compute.f90), then a moved configuration."""
    def build(lib):
        s, c = cm.spce_sample_system(lib, COUL_VARIANTS["coul_damped_square_smoothed"], jitter=0.0)
        s.md.Options.Compute = False
        rng = np.random.default_rng(7)
        mol = c["molecule"]
        R = c["R"] + rng.uniform(-0.3, 0.3, size=(int(mol.max()), 3))[mol - 1]
        s.upload("coordinates", R)
        s.compute_forces()
        return s

    sp, so = both(build)
    assert sp.md.Energy.UpToDate is False or sp.md.Energy.UpToDate == 0
    assert_state_parity(sp, so, energies=False)
    sp.finalize(), so.finalize()


def test_q4_virial_only_coul_none_quirk():
    """reference make_virial_compute.sh:24-29: with charges, kCoul != 0, no Coulomb model and Compute=false the
    pair virial is re-used as the Coulomb virial. Reproduced literally."""
    def build(lib):
        s, c = cm.spce_sample_system(lib, COUL_VARIANTS["coul_none"],
                                     pair_factory=lambda l, i, e, sg: l.EmDee_pair_lj_cut(max(e, 1e-4), max(sg, 1.0)))
        s.md.Options.Compute = False
        s.upload("coordinates", c["R"] * (1.0 + 1e-9))
        s.compute_forces()
        return s

    sp, so = both(build)
    assert_state_parity(sp, so, energies=False)
    sp.finalize(), so.finalize()


def _two_type_system(lib, coul=None, layers=1, multimodel=False, inner=None):
    """reference test/testfortran.f90:60-101 in miniature: two types, charges, exclusions |i-j| < 4."""
    R, L = cm.fcc_lj_box(7, rho=0.80, jitter=0.08, seed=11)   # 1372 atoms
    N = R.shape[0]
    types = np.where(np.arange(N) < N // 2, 1, 2).astype(np.int32)
    Q = np.where(types == 1, 1.0, -1.0)
    Q[: N // 5] = 0.0
    Q[N // 2: N // 2 + N // 5] = 0.0
    s = lib.system(2, layers, 2.5, 0.4, N, types, np.array([1.0, 2.5]), None)
    if inner is not None:
        s.layer_based_parameters(inner, [1 if k == 0 else 0 for k in range(layers)], [1] * layers)
    a = lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(1.0, 1.0))
    b = lib.EmDee_pair_softcore_cut(0.8, 1.1, 0.9)
    if multimodel:
        s.set_pair_multimodel(1, 1, [a, lib.EmDee_pair_lj_cut(0.5, 1.0)][:layers], [1.0] * layers)
        s.set_pair_multimodel(2, 2, [b, lib.EmDee_pair_lj_cut(0.7, 1.05)][:layers], [1.0] * layers)
    else:
        s.set_pair_model(1, 1, a, 1.0)
        s.set_pair_model(2, 2, b, 1.0)   # (1,2) is auto-mixed: softcore x lj rule, modifier dropped (Q3b)
    cmodel = coul(lib) if coul is not None else None
    if cmodel is not None:
        if multimodel:
            s.set_coul_multimodel([cmodel, lib.EmDee_coul_cut()][:layers])
        else:
            s.set_coul_model(cmodel)
    for i in range(1, N):
        for j in range(i + 1, min(i + 4, N + 1)):
            s.ignore_pair(i, j)
    s.upload("charges", Q)
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    return s


@pytest.mark.parametrize("coul", ["coul_none", "coul_sf", "coul_damped_smoothed", "shifted_force(coul_cut)"])
def test_two_types_mixing_exclusions(coul):
    sp, so = both(lambda lib: _two_type_system(lib, COUL_VARIANTS[coul]))
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_explicit_cross_pair_and_none_cross():
    def build(lib):
        R, L = cm.fcc_lj_box(6, rho=0.75, jitter=0.05, seed=5)
        N = R.shape[0]
        types = (np.arange(N) % 3 + 1).astype(np.int32)
        s = lib.system(1, 1, 2.5, 0.3, N, types, np.array([1.0, 1.0, 2.0]), None)
        s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
        s.set_pair_model(2, 2, lib.EmDee_pair_lj_cut(0.5, 0.9), 0.0)
        s.set_pair_model(3, 3, lib.EmDee_pair_none(), 0.0)
        s.set_pair_model(1, 2, lib.EmDee_smoothed(lib.EmDee_pair_lj_cut(0.3, 1.2), 0.4), 0.0)
        s.upload("box", np.array([L]))
        s.upload("coordinates", R)
        return s

    sp, so = both(build)
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_multilayer_inner_cutoff_and_layer_switch():
    def build(lib):
        return _two_type_system(lib, COUL_VARIANTS["coul_damped"], layers=2, multimodel=True, inner=1.8)

    sp, so = both(build)
    for layer in (1, 2, 1):
        sp.switch_model_layer(layer)
        so.switch_model_layer(layer)
        assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_dynamics_with_rebuilds_two_types():
    """NVE through the ABI with resident state: rebuild decisions (Builds) and trajectories must track."""
    def run(lib):
        s = _two_type_system(lib, COUL_VARIANTS["coul_sf"])
        s.random_momenta(1.2, True, 4321)
        for step in range(60):
            s.md.Options.Compute = (step % 10 == 9)
            s.boost(1.0, 0.0, 0.002)
            s.displace(1.0, 0.0, 0.004)
            s.boost(1.0, 0.0, 0.002)
        return s

    sp, so = both(run)
    assert sp.md.Builds == so.md.Builds and sp.md.Builds > 2
    assert_state_parity(sp, so, stol=1e-9, ftol=1e-7)
    assert abs(sp.md.Kinetic.Total - so.md.Kinetic.Total) <= 1e-9 * abs(so.md.Kinetic.Total)
    Rp, Ro = sp.download("coordinates"), so.download("coordinates")
    assert np.abs(Rp - Ro).max() < 1e-9
    sp.finalize(), so.finalize()


def test_upload_download_roundtrip_and_flags():
    s, c = cm.lj_sample_system(cm.product(), _lj)
    assert np.array_equal(s.download("coordinates"), c["R"])
    assert s.download("box") == c["L"]
    P = np.random.default_rng(3).normal(size=(c["N"], 3))
    s.upload("momenta", P)
    assert np.array_equal(s.download("momenta"), P)
    ke = 0.5 * (P ** 2).sum()
    assert abs(s.md.Kinetic.Total - ke) < 1e-9 * ke
    F = s.download("forces")
    s.upload("forces", 2.0 * F)
    assert np.array_equal(s.download("forces"), 2.0 * F)
    assert bool(s.md.Energy.UpToDate)
    s.upload("coordinates", c["R"])
    assert not bool(s.md.Energy.UpToDate)
    s.finalize()


def test_degenerate_inputs():
    """Edge cases: atoms sitting exactly on cell/box boundaries, coordinates far outside the box (the
    library never wraps, reference Q10), an atom without neighbours, and a tiny-negative coordinate (Q5)."""
    def build(lib):
        R, L = cm.fcc_lj_box(5, rho=0.6, jitter=0.0, seed=1)
        R[:, 0] -= 0.25 * L / 5          # put lattice planes exactly on x = 0
        R[0] = [-1e-18, 0.0, L]          # Q5 + exact upper boundary
        R[1] += np.array([3 * L, -2 * L, 7 * L])   # unwrapped far image
        s = lib.system(2, 1, 2.0, 0.3, R.shape[0], None, None, None)
        s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
        s.upload("box", np.array([L]))
        s.upload("coordinates", R)
        return s

    sp, so = both(build)
    assert_state_parity(sp, so)
    sp.finalize(), so.finalize()


def test_inert_system_has_zero_forces():
    def build(lib):
        R, L = cm.fcc_lj_box(5, rho=0.6)
        s = lib.system(1, 1, 2.0, 0.3, R.shape[0], None, None, None)
        s.upload("box", np.array([L]))
        s.upload("coordinates", R)
        return s

    sp, so = both(build)
    assert np.all(sp.download("forces") == 0.0)
    assert sp.md.Energy.Potential == 0.0 and sp.md.Virial.Total == 0.0
    assert sp.pairs().shape[0] == so.pairs().shape[0] == 0
    sp.finalize(), so.finalize()


def test_kernels_actually_ran():
    s, c = cm.lj_sample_system(cm.product(), _lj)
    st = s.stats()
    assert st.force_launches >= 1 and st.build_launches >= 1 and st.launches >= 8
    assert st.list_entries == 2 * s.pairs().shape[0]
    assert st.cells_per_dim == 6
    s.finalize()


def test_typed_path():
    """Several atom types, pair models all pair_lj_cut (one modifier) or pair_none, Coulomb kind fixed at
    compile time (k_pair_forces_typed; with two types the table row lives in registers). SPC/E with every eligible
    Coulomb model, a three-type LJ mixture (shared-memory table), virial-only mode, ineligible systems falling back."""
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


def test_kicks_with_energies_every_step():
    """Options.Compute on at every call, as bench.py runs: the kick EmDee_boost issues after a force evaluation is
    launched behind the pair kernel (Engine::plan_kick) and the first kick of the next step is answered from the sums
    the previous kick predicted. Every value a call returns must be what the oracle returns at that call."""
    def lj(lib, e, s):
        return lib.EmDee_pair_lj_cut(e, s)
    sp, c = cm.lj_sample_system(cm.product(), lj)
    so, _ = cm.lj_sample_system(cm.oracle(), lj)
    for s in (sp, so):
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        s.md.Options.Compute = True
    dt = c["Dt"]
    for step in range(25):
        for call in range(3):
            for s in (sp, so):
                if call == 1:
                    s.displace(1.0, 0.0, dt)
                else:
                    s.boost(1.0, 0.0, 0.5 * dt)
            if call != 1:
                assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-11, (step, call)
                for x in range(3):
                    assert cm.rel(sp.md.Kinetic.TransPart[x], so.md.Kinetic.TransPart[x]) < 1e-11
        assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-10, step
        assert cm.rel(sp.md.Virial.Total, so.md.Virial.Total) < 1e-9, step
    # a kick with other coefficients, a momentum upload and a thermostat-like scaling in between: predictions must not be reused
    for s in (sp, so):
        s.boost(1.0, 0.0, 0.25 * dt)
        s.boost(1.0, 0.1, 0.25 * dt)
        s.boost(1.0, 0.1, 0.25 * dt)
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-11
    P = so.download("momenta")
    for s in (sp, so):
        s.upload("momenta", 0.5 * P)
        s.boost(1.0, 0.0, 0.5 * dt)
        s.boost(1.0, 0.0, 0.5 * dt)
    assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-11
    assert sp.md.Builds == so.md.Builds
    assert cm.rel_force_error(sp.download("momenta"), so.download("momenta")) < 1e-9
    sp.finalize()
    so.finalize()


def test_deferred_kick_reaches_every_reader():
    """A kick whose kinetic sums need no reduction (predicted by the previous kick, or Options.Compute off) is not launched:
    the drift that follows applies it (Engine::boost / k_displace<true>), and every other call that reads or writes momenta
    or forces must execute it first (Engine::flush_kick). Same numbers as the oracle whatever comes between kick and drift."""
    def lj(lib, e, s):
        return lib.EmDee_pair_lj_cut(e, s)
    sp, c = cm.lj_sample_system(cm.product(), lj)
    so, _ = cm.lj_sample_system(cm.oracle(), lj)
    dt = c["Dt"]
    for s in (sp, so):
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        s.md.Options.Compute = True

    def step(s):
        s.boost(1.0, 0.0, 0.5 * dt)
        s.displace(1.0, 0.0, dt)
        s.boost(1.0, 0.0, 0.5 * dt)

    def same():
        assert cm.rel(sp.md.Kinetic.Total, so.md.Kinetic.Total) < 1e-11
        assert cm.rel_force_error(sp.download("momenta"), so.download("momenta")) < 1e-10
        assert np.abs(sp.download("coordinates") - so.download("coordinates")).max() < 1e-10

    for s in (sp, so):
        step(s)
        s.boost(1.0, 0.0, 0.5 * dt)        # predicted -> deferred in the product
    same()                                  # the momentum download comes before any drift
    for s in (sp, so):
        s.displace(1.0, 0.0, dt)
        s.boost(1.0, 0.0, 0.5 * dt)
        s.boost(1.0, 0.0, 0.5 * dt)        # deferred ...
        s.boost(1.0, 0.2, 0.3 * dt)        # ... then a kick with other coefficients instead of a drift
    same()
    for s in (sp, so):
        step(s)
        s.boost(1.0, 0.0, 0.5 * dt)        # deferred ...
    P = so.download("momenta")
    for s in (sp, so):
        s.upload("momenta", 0.9 * P)       # ... then overwritten by an upload
        step(s)
    same()
    for s in (sp, so):                      # sums not wanted: every kick with up-to-date forces is deferred
        s.md.Options.Compute = False
        for _ in range(8):
            step(s)
        s.boost(1.0, 0.0, 0.5 * dt)        # deferred, then forces of this layer are recomputed with energies
        s.md.Options.Compute = True
        s.compute_forces()
        s.displace(1.0, 0.0, dt)
        s.boost(1.0, 0.0, 0.5 * dt)
    same()
    assert cm.rel(sp.md.Energy.Potential, so.md.Energy.Potential) < 1e-10
    assert sp.md.Builds == so.md.Builds
    sp.finalize()
    so.finalize()


def test_deferred_kick_changes_no_bit(monkeypatch):
    """The drift kernel that absorbs a deferred kick does the kick's and the drift's operations in the same order as the
    two kernels it replaces: trajectories with and without deferral (EMDEE_NO_DEFER_KICK=1, read when a system is created)
    must agree bit for bit -- coordinates, momenta, forces, energies, number of list builds."""
    def lj(lib, e, s):
        return lib.EmDee_pair_lj_cut(e, s)
    lib = cm.product()

    def run(no_defer):
        if no_defer:
            monkeypatch.setenv("EMDEE_NO_DEFER_KICK", "1")
        else:
            monkeypatch.delenv("EMDEE_NO_DEFER_KICK", raising=False)
        s, c = cm.lj_sample_system(lib, lj)
        s.random_momenta(c["kB"] * c["Temp"], True, c["seed"])
        s.md.Options.Compute = True
        dt = c["Dt"]
        ke = []
        for _ in range(25):
            s.boost(1.0, 0.0, 0.5 * dt)
            ke.append(s.md.Kinetic.Total)
            s.displace(1.0, 0.0, dt)
            s.boost(1.0, 0.0, 0.5 * dt)
            ke.append(s.md.Kinetic.Total)
        out = (s.download("coordinates"), s.download("momenta"), s.download("forces"), s.md.Energy.Potential,
               s.md.Virial.Total, np.array(ke), int(s.md.Builds), s.stats().launches)
        s.finalize()
        return out
    a, b = run(False), run(True)
    for x, y in zip(a[:6], b[:6]):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    assert a[6] == b[6] and a[6] > 1
    assert a[7] < b[7]   # and the deferral did remove launches


# ---- golden numbers that came from neither C++ restatement (tests/golden/numpy_models.py) -----------------------------
import golden_cases as gc  # noqa: E402


@pytest.mark.parametrize("name", gc.ALL_CASES)
def test_numpy_golden_on_gpu(name):
    """every pair modifier, pair_softcore_cut, every cutoff Coulomb model and the Q1 / Q1b / Q3b setter orders: the CUDA
    product against the committed output of the independent numpy evaluator (the oracle is held to the same rows in
    tests/test_numpy_golden.py)"""
    gc.check(cm.product(), name)
