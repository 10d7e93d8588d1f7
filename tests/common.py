"""Shared helpers for the test-suite: library handles, fixtures and reference-style drivers.

The drivers below read like the reference's own test programs (reference test/test_pair_lj_cut.f90:36-48,
test/common/contained.f90:54-89): the same call sequence is issued to whichever library is passed in
(CPU oracle or CUDA product), so parity tests compare like with like.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from emdee_b200 import api  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_STRICT = os.path.join(ROOT, "oracle", "_build", "libemdee_oracle.so")
ORACLE_FAST = os.path.join(ROOT, "oracle", "_build", "libemdee_oracle_fast.so")

os.environ.setdefault("EMDEE_QUIET", "1")

_libs = {}


def oracle(fast: bool = False) -> api.EmDeeLib:
    path = ORACLE_FAST if fast else ORACLE_STRICT
    if path not in _libs:
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _libs[path] = api.EmDeeLib(path)
    return _libs[path]


HOSTSTUB = os.path.join(ROOT, "tests", "_build", "libemdee_hoststub.so")


def hoststub() -> api.EmDeeLib:
    """The product's host shim (emdee_b200/csrc/abi.cpp, unchanged) linked against a device-free stub of the
    engine (tests/hoststub/engine_stub.cpp): host-side semantics without a GPU. Test infrastructure only."""
    src = [os.path.join(ROOT, "emdee_b200", "csrc", "abi.cpp"), os.path.join(ROOT, "tests", "hoststub", "engine_stub.cpp")]
    deps = src + [os.path.join(ROOT, "emdee_b200", "csrc", h) for h in ("engine.h", "nb_math.h")]
    if HOSTSTUB not in _libs:
        os.makedirs(os.path.dirname(HOSTSTUB), exist_ok=True)
        if not os.path.exists(HOSTSTUB) or any(os.path.getmtime(f) > os.path.getmtime(HOSTSTUB) for f in deps):
            subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                                   "-Wl,-Bsymbolic", "-Wl,--no-undefined", "-I" + os.path.join(ROOT, "include"),
                                   "-x", "c++", *src, "-o", HOSTSTUB])
        _libs[HOSTSTUB] = api.EmDeeLib(HOSTSTUB)
    return _libs[HOSTSTUB]


def emulated() -> api.EmDeeLib:
    """The product's sources (abi.cpp, engine.cu, engine_*.cuh) compiled for the HOST against the CUDA
    execution-model emulator tests/cusim/cusim.h: kernel logic without a GPU. Test infrastructure only."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "cusim"))
    import build as cusim_build
    path = cusim_build.build()
    if path not in _libs:
        _libs[path] = api.EmDeeLib(path)
    return _libs[path]


def product() -> api.EmDeeLib:
    return api.load()


def load_fixture(name: str) -> dict:
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (d[k] if d[k].ndim else d[k].item()) for k in d.files}


def kats() -> dict:
    d = np.load(os.path.join(GOLDEN, "reference_kats.npz"))
    return {k: d[k] for k in d.files}


def lj_sample_system(lib, pair_factory, threads=2, Rc=None, skin=None, kCoul=None):
    """reference test/test_pair_lj_*.f90:36-44 up to (excluding) EmDee_random_momenta."""
    c = load_fixture("NIST_lj_sample")
    Rc = c["Rc"] if Rc is None else Rc
    skin = c["Rs"] if skin is None else skin
    kCoul = c["kCoul"] if kCoul is None else kCoul
    s = lib.system(threads, 1, Rc, skin, c["N"], c["atomType"], c["mass"], None)
    eps = c["epsilon"] / c["mvv2e"]
    for i in range(len(c["mass"])):
        s.set_pair_model(i + 1, i + 1, pair_factory(lib, eps[i], c["sigma"][i]), kCoul)
    s.upload("charges", c["Q"])
    s.upload("box", np.array([c["L"]]))
    s.upload("coordinates", c["R"])
    return s, c


def run_nve(s, c, nsteps, nprop=20):
    """reference test/common/contained.f90:54-89 (`run`): returns mvv2e*[U, W, U+K]."""
    dt = c["Dt"]
    for step in range(1, nsteps + 1):
        s.md.Options.Compute = (step % nprop == 0)
        s.boost(1.0, 0.0, 0.5 * dt)
        s.displace(1.0, 0.0, dt)
        s.boost(1.0, 0.0, 0.5 * dt)
    md = s.md
    return c["mvv2e"] * np.array([md.Energy.Potential, md.Virial.Total, md.Energy.Potential + md.Kinetic.Total])


def spce_sample_system(lib, coul_factory, threads=2, replicas=1, pair_factory=None, bodies=True,
                       Rc=None, skin=None, jitter=0.0, seed=1):
    """reference test/test_coul_*.f90:36-50: SPC/E water, LJ-sf on O, pair_none on H, rigid bodies."""
    c = load_fixture("NIST_spce_sample")
    n = replicas
    R0, L0 = c["R"].copy(), c["L"]
    if n > 1:
        # make molecules whole (minimum image w.r.t. their first atom) so that replicas tile exactly
        first_of = {}
        for a, m in enumerate(c["molecule"]):
            first_of.setdefault(m, a)
        ref = R0[[first_of[m] for m in c["molecule"]]]
        R0 = R0 - L0 * np.round((R0 - ref) / L0)
    shifts = np.array([[i, j, k] for i in range(n) for j in range(n) for k in range(n)], dtype=float) * L0
    R = (R0[None, :, :] + shifts[:, None, :]).reshape(-1, 3)
    nmol0 = int(c["molecule"].max())
    mol = (c["molecule"][None, :] + (np.arange(n ** 3) * nmol0)[:, None]).reshape(-1).astype(np.int32)
    typ = np.tile(c["atomType"], n ** 3).astype(np.int32)
    Q = np.tile(c["Q"], n ** 3)
    N = R.shape[0]
    L = L0 * n
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        per_mol = rng.uniform(-jitter, jitter, size=(int(mol.max()), 3))
        R = R + per_mol[mol - 1]
    Rc = c["Rc"] if Rc is None else Rc
    skin = c["Rs"] if skin is None else skin
    s = lib.system(threads, 1, Rc, skin, N, typ, c["mass"], mol if bodies else None)
    eps = c["epsilon"] / c["mvv2e"]
    for i in range(len(c["mass"])):
        if pair_factory is not None:
            model = pair_factory(lib, i, eps[i], c["sigma"][i])
        elif eps[i] == 0.0:
            model = lib.EmDee_pair_none()
        else:
            model = lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[i], c["sigma"][i]))
        s.set_pair_model(i + 1, i + 1, model, c["kCoul"])
    coul = coul_factory(lib)
    if coul is not None:
        s.set_coul_model(coul)
    if not bodies:
        # without rigid bodies, intramolecular pairs must be excluded explicitly
        first = {}
        for a in range(N):
            first.setdefault(mol[a], []).append(a + 1)
        for atoms in first.values():
            for x in range(len(atoms)):
                for y in range(x + 1, len(atoms)):
                    s.ignore_pair(atoms[x], atoms[y])
    s.upload("charges", Q)
    s.upload("coordinates", R)
    s.upload("box", np.array([L]))
    info = dict(c)
    info.update(N=N, L=L, R=R, Q=Q, atomType=typ, molecule=mol)
    return s, info


def fcc_lj_box(ncell: int, rho: float = 0.8442, jitter: float = 0.05, seed: int = 86245):
    """Synthetic LJ fluid start: fcc lattice ncell^3 x 4 atoms at reduced density rho, each coordinate
    jittered by uniform(-jitter, jitter)."""
    N = 4 * ncell ** 3
    L = (N / rho) ** (1.0 / 3.0)
    a = L / ncell
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    g = np.arange(ncell)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    R = ((cells[:, None, :] + base[None, :, :]) * a).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    R = R + rng.uniform(-jitter, jitter, size=R.shape)
    return R, L


def rel_force_error(F, Fref):
    """Per-atom relative force error: |dF_i|_inf / max(|Fref_i|_inf, rms(|Fref|)); returns the max.
    The rms floor keeps atoms whose net force nearly cancels from inflating the ratio."""
    d = np.abs(F - Fref).max(axis=1)
    mag = np.abs(Fref).max(axis=1)
    floor = np.sqrt((Fref ** 2).mean()) if Fref.size else 1.0
    return float((d / np.maximum(mag, max(floor, 1e-300))).max())


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
