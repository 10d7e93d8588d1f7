"""numpy_models.py -- an INDEPENDENT O(N^2) numpy evaluator of EmDee's nonbonded energies, virials and forces, written
straight from the reference's Fortran model files and setters (paths relative to the reference tree; nothing here is
derived from oracle/emdee_oracle.cpp or emdee_b200/csrc/abi.cpp). TEST INFRASTRUCTURE: it produces the committed golden
numbers of tests/golden/model_golden.json (make_model_fixtures.py), against which BOTH the CPU oracle and the CUDA
product are asserted, so that a misreading shared by those two C++ restatements cannot hide.

What is restated, and from where:
  pair_lj_cut            src/pair_lj_cut.f90:57-86         pair_softcore_cut     src/pair_softcore_cut.f90:27-29,60-101
  coul_cut               src/coul_cut.f90:62-70            coul_sf               src/coul_sf.f90:50-72
  coul_damped            src/coul_damped.f90:52-83         coul_damped_smoothed  src/coul_damped_smoothed.f90:54-115
  coul_damped_square_smoothed  src/coul_damped_square_smoothed.f90:54-114
  coul_square_smoothed   src/coul_square_smoothed.f90:52-104   coul_shifted_square_smoothed  src/coul_shifted_square_smoothed.f90:52-107
  uerfc                  src/math.f90:35-40,685-691
  modifiers              src/apply_modifier.f90:1-60, constructors src/modelClass_nonbonded.f90:60-238
  modifier_setup         src/modelClass_nonbonded.f90:243-296   cutoff_setup  src/modelClass_coul.f90:62-93
  setter order           src/EmDeeCode.f90:448-482 (set_coul_model), 283-303 (layer_based_parameters),
                         src/EmDeeData.f90:227-264 (set_pair_type, mixing), src/modelClass_pair.f90:120-140 (mix)
  pair loop              src/compute.f90:20-100 (cutoff test, QiQj, force = Wsum/r^2 * Rij)
  body virial            src/EmDeeData.f90:926-953, body offsets src/ArBee.f90 (delta = R - Rcm of the unwrapped body)
Model objects carry STATE exactly as the Fortran objects do (eshift, fshift, Rm, factor, ... persist between the setup
calls), because several reference quirks are consequences of that state.
"""
import copy

import numpy as np

NONE, SHIFTED, SHIFTED_FORCE, SMOOTHED, SHIFTED_SMOOTHED, SQUARE_SMOOTHED, SHIFTED_SQUARE_SMOOTHED = range(7)


class Model:
    """One nonbonded model object: kind + parameters + the inherited cNonBondedModel fields."""

    def __init__(self, kind, **p):
        self.kind = kind
        self.p = dict(p)
        self.modifier = NONE
        self.skin = 0.0          # modifier skin (modelClass_nonbonded.f90:41)
        self.eshift = 0.0
        self.fshift = 0.0
        self.Rm = 0.0
        self.RmSq = 0.0
        self.factor = 0.0
        self.Rm2fac = 0.0
        # coulomb-class flags (modelClass_coul.f90) and model-specific state
        self.shifted = kind == "coul_shifted_square_smoothed"
        self.shifted_force = kind == "coul_sf"
        self.Rm2 = 0.0
        self.invRm = 0.0
        if kind in ("coul_damped", "coul_damped_smoothed", "coul_damped_square_smoothed"):
            self.alpha = p["damp"]
            self.beta = 2.0 * self.alpha / np.sqrt(np.pi)
        if kind == "pair_lj_cut":
            self.eps4 = 4.0 * p["epsilon"]
            self.eps24 = 24.0 * p["epsilon"]
            self.sigSq = p["sigma"] ** 2
        if kind == "pair_softcore_cut":
            lam = p["lam"]
            self.prefactor = 4.0 * p["epsilon"] * lam ** 1.0        # exponent_n = 1
            self.prefactor6 = 6.0 * self.prefactor
            self.invSigSq = 1.0 / p["sigma"] ** 2
            self.shift = 0.5 * (1.0 - lam) ** 1.0                   # alpha = 1/2, exponent_p = 1

    # ---- E(r), W(r) = -r dE/dr of the bare model (the `compute` procedure of each model file) ----
    def compute(self, invR, invR2):
        invR = np.asarray(invR, dtype=float)
        invR2 = np.asarray(invR2, dtype=float)
        k = self.kind
        if k in ("pair_none", "coul_none"):
            return np.zeros_like(invR), np.zeros_like(invR)
        if k == "pair_lj_cut":
            sr2 = self.sigSq * invR2
            sr6 = sr2 * sr2 * sr2
            sr12 = sr6 * sr6
            return self.eps4 * (sr12 - sr6), self.eps24 * (sr12 + sr12 - sr6)
        if k == "pair_softcore_cut":
            rsig2 = self.invSigSq / invR2
            rsig6 = rsig2 * rsig2 * rsig2
            sinv = 1.0 / (rsig6 + self.shift)
            sinvSq = sinv * sinv
            sinvCb = sinv * sinvSq
            return self.prefactor * (sinvSq - sinv), self.prefactor6 * rsig6 * (sinvCb + sinvCb - sinvSq)
        if k == "coul_cut":
            return invR.copy(), invR.copy()
        if k == "coul_sf":
            rFc = self.fshift / invR
            return invR + self.eshift + rFc, invR - rFc
        if k in ("coul_damped", "coul_damped_smoothed", "coul_damped_square_smoothed"):
            x = self.alpha / invR
            expmx2 = np.exp(-x * x)
            E = uerfc(x, expmx2) * invR
            W = E + self.beta * expmx2
            if k == "coul_damped_smoothed":
                r = 1.0 / invR
                sw = r > self.Rm                       # the INHERITED Rm (coul_damped_smoothed.f90:104)
                u = self.factor * (r - self.Rm)
                G, WG = quintic(u, -30.0)
                WG = WG * self.factor * r
                return np.where(sw, E * G, E), np.where(sw, W * G + E * WG, W)
            if k == "coul_damped_square_smoothed":
                return square_switch(E, W, invR, invR2, self)
            return E, W
        if k == "coul_square_smoothed":
            return square_switch(invR.copy(), invR.copy(), invR, invR2, self)
        if k == "coul_shifted_square_smoothed":
            return square_switch(invR + self.eshift, invR.copy(), invR, invR2, self)
        raise ValueError(k)

    # ---- cCoulModel_cutoff_setup + each model's apply_cutoff (modelClass_coul.f90:62-93) ----
    def cutoff_setup(self, cutoff):
        self.fshift = 0.0
        self.eshift = 0.0
        if self.shifted or self.shifted_force:
            E, W = self.compute(1.0 / cutoff, (1.0 / cutoff) * (1.0 / cutoff))
            E, W = float(E), float(W)
            if self.shifted_force:
                self.fshift = W / cutoff
                self.eshift = -(E + W)
            else:
                self.fshift = 0.0
                self.eshift = -E
        k = self.kind
        if k == "coul_damped_smoothed":
            sw = self.p["skin"]
            self.Rm = cutoff - sw
            self.Rm2 = self.Rm ** 2
            self.invRm = 1.0 / self.Rm
            self.factor = 1.0 / (cutoff - self.Rm)
        elif k in ("coul_damped_square_smoothed", "coul_square_smoothed", "coul_shifted_square_smoothed"):
            sw = self.p["skin"]
            self.Rm2 = (cutoff - sw) ** 2
            self.invRm = 1.0 / (cutoff - sw)
            self.factor = 1.0 / (cutoff ** 2 - self.Rm2)

    # ---- cNonBondedModel_modifier_setup (modelClass_nonbonded.f90:243-296) ----
    def modifier_setup(self, cutoff):
        self.fshift = 0.0
        self.eshift = 0.0
        self.Rm = cutoff - self.skin
        self.RmSq = self.Rm ** 2
        shifting = self.modifier in (SHIFTED, SHIFTED_FORCE, SHIFTED_SMOOTHED, SHIFTED_SQUARE_SMOOTHED)
        if shifting:
            Ec, Wc = (float(v) for v in self.compute(1.0 / cutoff, 1.0 / cutoff ** 2))
        if self.modifier == SHIFTED:
            self.eshift = -Ec
        elif self.modifier == SHIFTED_FORCE:
            self.eshift = -(Ec + Wc)
            self.fshift = Wc / cutoff
        elif self.modifier in (SMOOTHED, SHIFTED_SMOOTHED):
            if shifting:
                Es, _ = (float(v) for v in self.compute(1.0 / self.Rm, 1.0 / self.RmSq))
                self.eshift = -0.5 * (Es + Ec)
            self.factor = 1.0 / (cutoff - self.Rm)
            self.Rm2fac = self.factor * self.Rm
        elif self.modifier in (SQUARE_SMOOTHED, SHIFTED_SQUARE_SMOOTHED):
            if shifting:
                Es, _ = (float(v) for v in self.compute(1.0 / self.Rm, 1.0 / self.RmSq))
                self.eshift = -0.5 * (Es + Ec)
            self.factor = 1.0 / (cutoff ** 2 - self.RmSq)
            self.Rm2fac = self.factor * self.RmSq

    # ---- model body followed by apply_modifier.f90 (the `compute` instantiation: energies and virials) ----
    def evaluate(self, invR, invR2):
        E, W = self.compute(invR, invR2)
        m = self.modifier
        if m == SHIFTED:
            E = E + self.eshift
        elif m == SHIFTED_FORCE:
            rFc = self.fshift / invR
            W = W - rFc
            E = E + self.eshift + rFc
        elif m in (SMOOTHED, SHIFTED_SMOOTHED, SQUARE_SMOOTHED, SHIFTED_SQUARE_SMOOTHED):
            square = m in (SQUARE_SMOOTHED, SHIFTED_SQUARE_SMOOTHED)
            E = E + self.eshift
            r2fac = self.factor / invR2 if square else self.factor / invR
            sw = r2fac > self.Rm2fac
            G, WG = quintic(r2fac - self.Rm2fac, -60.0 if square else -30.0)
            WG = WG * r2fac
            W = np.where(sw, W * G + E * WG, W)
            E = np.where(sw, E * G, E)
        return E, W


def uerfc(x, expmx2):
    a1, a2, a3, a4, a5, p = 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.327591100
    t = 1.0 / (1.0 + p * x)
    return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * expmx2


def quintic(u, coef):
    u2 = u * u
    u3 = u * u2
    return 1.0 + u3 * (15.0 * u - 6.0 * u2 - 10.0), coef * u2 * (2.0 * u - u2 - 1.0)


def square_switch(E, W, invR, invR2, m):
    """the `if (invR < model%invRm)` block shared by the *_square_smoothed Coulomb models"""
    sw = invR < m.invRm
    r2 = 1.0 / invR2
    u = m.factor * (r2 - m.Rm2)
    G, WG = quintic(u, -60.0)
    WG = WG * m.factor * r2
    return np.where(sw, E * G, E), np.where(sw, W * G + E * WG, W)


# ---- constructors, named after the C ABI ----
def pair_none():
    return Model("pair_none")


def coul_none():
    return Model("coul_none")


def pair_lj_cut(epsilon, sigma):
    return Model("pair_lj_cut", epsilon=epsilon, sigma=sigma)


def pair_softcore_cut(epsilon, sigma, lam):
    return Model("pair_softcore_cut", epsilon=epsilon, sigma=sigma, lam=lam)


def coul_cut():
    return Model("coul_cut")


def coul_sf():
    return Model("coul_sf")


def coul_damped(damp):
    return Model("coul_damped", damp=damp)


def coul_damped_smoothed(damp, skin):
    return Model("coul_damped_smoothed", damp=damp, skin=skin)


def coul_damped_square_smoothed(damp, skin):
    return Model("coul_damped_square_smoothed", damp=damp, skin=skin)


def coul_square_smoothed(skin):
    return Model("coul_square_smoothed", skin=skin)


def coul_shifted_square_smoothed(skin):
    return Model("coul_shifted_square_smoothed", skin=skin)


def _wrap(model, modifier, skin=0.0):   # EmDee_shifted ... EmDee_shifted_square_smoothed: a copy with modifier (+ skin) set
    new = copy.deepcopy(model)
    new.modifier = modifier
    if modifier in (SMOOTHED, SHIFTED_SMOOTHED, SQUARE_SMOOTHED, SHIFTED_SQUARE_SMOOTHED):
        new.skin = skin
    return new


def shifted(m):
    return _wrap(m, SHIFTED)


def shifted_force(m):
    return _wrap(m, SHIFTED_FORCE)


def smoothed(m, skin):
    return _wrap(m, SMOOTHED, skin)


def shifted_smoothed(m, skin):
    return _wrap(m, SHIFTED_SMOOTHED, skin)


def square_smoothed(m, skin):
    return _wrap(m, SQUARE_SMOOTHED, skin)


def shifted_square_smoothed(m, skin):
    return _wrap(m, SHIFTED_SQUARE_SMOOTHED, skin)


def mix(a, b):
    """pairContainer_mix -> the models' own mixing rules; a FRESH model: modifier NONE, skin 0 (quirk Q3b)"""
    kinds = {a.kind, b.kind}
    if kinds == {"pair_lj_cut"}:
        return pair_lj_cut(np.sqrt(a.p["epsilon"] * b.p["epsilon"]), 0.5 * (a.p["sigma"] + b.p["sigma"]))
    if kinds == {"pair_softcore_cut"}:
        return pair_softcore_cut(np.sqrt(a.p["epsilon"] * b.p["epsilon"]), 0.5 * (a.p["sigma"] + b.p["sigma"]), a.p["lam"] * b.p["lam"])
    if kinds == {"pair_softcore_cut", "pair_lj_cut"}:
        s = a if a.kind == "pair_softcore_cut" else b
        return pair_softcore_cut(np.sqrt(a.p["epsilon"] * b.p["epsilon"]), 0.5 * (a.p["sigma"] + b.p["sigma"]), s.p["lam"])
    return pair_none()


class System:
    """One-layer system as the setters build it (EmDee_system + EmDee_set_pair_model + EmDee_set_coul_model
    [+ EmDee_layer_based_parameters]); evaluation is a plain double loop over all pairs, vectorised over j."""

    def __init__(self, Rc, ntypes):
        self.Rc = Rc
        self.nt = ntypes
        self.pair = [[pair_none() for _ in range(ntypes)] for _ in range(ntypes)]
        self.coulomb = np.zeros((ntypes, ntypes), dtype=bool)
        self.kCoul = np.zeros((ntypes, ntypes))
        self.overridable = np.ones((ntypes, ntypes), dtype=bool)
        self.coul = coul_none()

    def set_pair_model(self, i, j, model, kCoul):      # 1-based types; EmDeeData.f90:227-264, EmDeeCode.f90:309-360
        i -= 1
        j -= 1
        if i == j:
            self.pair[i][i] = copy.deepcopy(model)
            self.coulomb[i, i] = kCoul != 0.0
            if self.coulomb[i, i]:
                self.kCoul[i, i] = kCoul
            self.pair[i][i].modifier_setup(self.Rc)
            for k in range(self.nt):
                if k != i and self.overridable[i, k]:
                    mixed = mix(self.pair[k][k], self.pair[i][i])
                    mixed.modifier_setup(self.Rc)
                    self.pair[i][k] = self.pair[k][i] = mixed
                    c = self.coulomb[k, k] and self.coulomb[i, i]
                    self.coulomb[i, k] = self.coulomb[k, i] = c
                    if c:
                        self.kCoul[i, k] = self.kCoul[k, i] = np.sqrt(self.kCoul[k, k] * self.kCoul[i, i])
        else:
            m = copy.deepcopy(model)
            m.modifier_setup(self.Rc)
            self.pair[i][j] = self.pair[j][i] = m
            self.coulomb[i, j] = self.coulomb[j, i] = kCoul != 0.0
            if kCoul != 0.0:
                self.kCoul[i, j] = self.kCoul[j, i] = kCoul
            self.overridable[i, j] = self.overridable[j, i] = False

    def set_coul_model(self, model):                    # EmDeeCode.f90:448-482
        self.coul = copy.deepcopy(model)
        self.coul.cutoff_setup(self.Rc)
        self.coul.modifier_setup(self.Rc)

    def layer_based_parameters(self):                   # EmDeeCode.f90:299-303 (one layer, no inner cutoff)
        self.coul.cutoff_setup(self.Rc)

    def evaluate(self, R, L, types, Q, molecule=None):
        """-> dict(Epair, Ecoul, Wpair, Wcoul, F). types 1-based; molecule: rigid-body ids (pairs inside a body are
        skipped, as the reference's list build does), or None. The body virial is body_virial() below."""
        N = R.shape[0]
        t = np.asarray(types) - 1
        Q = np.asarray(Q, dtype=float)
        charged = np.abs(Q) > np.finfo(float).eps
        F = np.zeros((N, 3))
        Ep = Ec = Wp = Wc = 0.0
        Rc2 = self.Rc ** 2
        for i in range(N - 1):
            d = R[i] - R[i + 1:]
            d -= L * np.rint(d / L)                     # minimum image (pbc of compute.f90:40 in real units)
            r2 = (d * d).sum(axis=1)
            ok = r2 < Rc2
            if molecule is not None:
                ok &= molecule[i + 1:] != molecule[i]
            if not ok.any():
                continue
            js = np.nonzero(ok)[0]
            dj, r2j = d[js], r2[js]
            invR2 = 1.0 / r2j
            invR = np.sqrt(invR2)
            tj = t[i + 1:][js]
            Wsum = np.zeros(len(js))
            for jt in np.unique(tj):
                sel = tj == jt
                E, W = self.pair[t[i]][jt].evaluate(invR[sel], invR2[sel])
                Ep += E.sum()
                Wp += W.sum()
                Wsum[sel] += W
                if self.coulomb[t[i], jt] and charged[i] and self.coul.kind != "coul_none":
                    qj = Q[i + 1:][js][sel]
                    cj = charged[i + 1:][js][sel]
                    Eq, Wq = self.coul.evaluate(invR[sel], invR2[sel])
                    QiQj = np.where(cj, self.kCoul[t[i], jt] * Q[i] * qj, 0.0)
                    Ec += (QiQj * Eq).sum()
                    Wc += (QiQj * Wq).sum()
                    tmp = Wsum[sel]
                    tmp += QiQj * Wq
                    Wsum[sel] = tmp
            Fij = (Wsum * invR2)[:, None] * dj
            F[i] += Fij.sum(axis=0)
            np.subtract.at(F, i + 1 + js, Fij)
        return dict(Epair=Ep, Ecoul=Ec, Wpair=Wp, Wcoul=Wc, F=F)


def body_virial(R, L, F, molecule, mass_of_atom):
    """-sum_bodies sum_members F . delta, delta = member position - centre of mass of the body made whole around its
    first member (EmDeeData.f90:926-953; body construction src/ArBee.f90 tBody_update)."""
    W = 0.0
    for b in np.unique(molecule):
        idx = np.nonzero(molecule == b)[0]
        Rb = R[idx].copy()
        Rb -= L * np.rint((Rb - Rb[0]) / L)
        m = mass_of_atom[idx]
        rcm = (m[:, None] * Rb).sum(axis=0) / m.sum()
        W += (F[idx] * (Rb - rcm)).sum()
    return -W
