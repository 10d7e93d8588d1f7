// =================================================================================================
// emdee_oracle.cpp -- TEST INFRASTRUCTURE. NOT PART OF THE PRODUCT.
//
// CPU (C++17 + OpenMP) restatement of the nonbonded hot path of atoms-ufrj/EmDee ("15 Oct 2018"):
// the cell-list Verlet neighbor-list build and the pairwise force/energy/virial loop, behind the
// same C ABI (include/emdee.h) -- and, since the widening of SURVEY.md section 8(f), of the callers around
// it: the rigid-body integrator (src/ArBee.f90), EmDee_verlet_step, harmonic bonds / angles, the Ewald
// reciprocal sum (src/kspace_ewald.f90, src/modelClass_kspace.f90), EmDee_rdf, EmDee_memory_address and
// EmDee_share_phase_space. The reference is Fortran 2008 and no Fortran compiler exists in the
// build image, so the reference itself cannot be compiled (oracle/_ref is therefore absent); this
// file follows the reference sources function by function and every function cites the file:line
// it restates (paths relative to the reference tree).
//
// Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs -- as the checker or as the CPU baseline, never as the thing shipped. The product library
// (emdee_b200/lib/libemdee.so) neither links nor calls anything in this directory.
//
// Parity pins (tests/test_oracle_pins.py): the reference's own live known-answer tests
//   test/test_pair_lj_cut.f90:46, test/test_pair_lj_sf.f90:46 (100-step NVE, 800-atom NIST LJ),
//   test/test_pair_lj_smoothed.f90:46 (valid for skin = 1.0),
// NIST SRSW reference energies for the LJ sample, and the survey's independent O(N^2) probes.
// Coulomb models, soft-core, rigid-body dynamics, bonded terms and Ewald are NOT pinned by any reference
// test (the reference's programs for them assert nothing): "parity unpinned" for those rows. They are
// pinned by physics instead -- conservation laws and the group property of the exact free-rotor map,
// finite-difference forces and virials, the Madelung constant of rock salt (tests/test_oracle_*.py).
//
// Two builds (oracle/Makefile): strict (-O2 -ffp-contract=off, defines bit-level truth for list
// membership) and fast (-Ofast -march=native -fopenmp, mirrors reference Makefile:13,21; used for
// CPU timing only).
// =================================================================================================

#define EMDEE_ORACLE_BUILD 1
#include "../include/emdee.h"
#include "../include/emdee_ext.h"

#include <omp.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

static_assert(sizeof(tEmDee) == 240, "tEmDee must match the reference layout (240 bytes)");
static_assert(offsetof(tEmDee, Energy) == 40 && offsetof(tEmDee, Kinetic) == 104, "layout");
static_assert(offsetof(tEmDee, Virial) == 192 && offsetof(tEmDee, Data) == 216, "layout");
static_assert(offsetof(tEmDee, Options) == 224, "layout");

namespace {

// ---- src/global.f90:51-64 ----------------------------------------------------------------------
[[noreturn]] void error(const char* task, const std::string& msg) {
  std::fprintf(stderr, "Error in %s: %s.\n", task, msg.c_str());
  std::fflush(stderr);
  std::exit(1);
}
void warning(const std::string& msg) { std::fprintf(stderr, "WARNING: %s.\n", msg.c_str()); }

// src/EmDeeData.f90:31-34
constexpr int extra = 2000;
constexpr int ndiv = 2;
constexpr int nbcells = 62;

// src/neighbor_lists.f90:26-35 -- half shell of the 5x5x5 stencil, as (x,y,z) triples
constexpr int nb[nbcells][3] = {
    {0, 0, 1},   {0, 1, 0},   {1, 0, 0},   {-1, 0, 1},  {-1, 1, 0},  {0, -1, 1},  {0, 1, 1},
    {1, 0, 1},   {1, 1, 0},   {-1, -1, 1}, {-1, 1, 1},  {1, -1, 1},  {1, 1, 1},   {0, 0, 2},
    {0, 2, 0},   {2, 0, 0},   {-2, 0, 1},  {-2, 1, 0},  {-1, 0, 2},  {-1, 2, 0},  {0, -2, 1},
    {0, -1, 2},  {0, 1, 2},   {0, 2, 1},   {1, 0, 2},   {1, 2, 0},   {2, 0, 1},   {2, 1, 0},
    {-2, -1, 1}, {-2, 1, 1},  {-1, -2, 1}, {-1, -1, 2}, {-1, 1, 2},  {-1, 2, 1},  {1, -2, 1},
    {1, -1, 2},  {1, 1, 2},   {1, 2, 1},   {2, -1, 1},  {2, 1, 1},   {-2, 0, 2},  {-2, 2, 0},
    {0, -2, 2},  {0, 2, 2},   {2, 0, 2},   {2, 2, 0},   {-2, -2, 1}, {-2, -1, 2}, {-2, 1, 2},
    {-2, 2, 1},  {-1, -2, 2}, {-1, 2, 2},  {1, -2, 2},  {1, 2, 2},   {2, -2, 1},  {2, -1, 2},
    {2, 1, 2},   {2, 2, 1},   {-2, -2, 2}, {-2, 2, 2},  {2, -2, 2},  {2, 2, 2}};

// ---- models -------------------------------------------------------------------------------------
// src/modelClass_nonbonded.f90:27-33
enum Modifier { NONE = 0, SHIFTED, SHIFTED_FORCE, SMOOTHED, SHIFTED_SMOOTHED, SQUARE_SMOOTHED,
                SHIFTED_SQUARE_SMOOTHED };

enum Kind {
  PAIR_NONE, PAIR_LJ_CUT, PAIR_SOFTCORE_CUT,
  COUL_NONE, COUL_CUT, COUL_SF, COUL_DAMPED, COUL_LONG, COUL_DAMPED_SMOOTHED,
  COUL_DAMPED_SQUARE_SMOOTHED, COUL_SQUARE_SMOOTHED, COUL_SHIFTED_SQUARE_SMOOTHED,
  BOND_NONE, BOND_HARMONIC, ANGLE_NONE, ANGLE_HARMONIC, DIHEDRAL_NONE, KSPACE_EWALD
};
inline bool is_pair(Kind k) { return k <= PAIR_SOFTCORE_CUT; }
inline bool is_coul(Kind k) { return k >= COUL_NONE && k <= COUL_SHIFTED_SQUARE_SMOOTHED; }
inline bool is_nonbonded(Kind k) { return is_pair(k) || is_coul(k); }

constexpr uint64_t MODEL_MAGIC = 0x4d4f44454c4f5243ull;

// One flat record holds the union of the fields of the reference's model class hierarchy:
// cModel (src/modelClass.f90:25-30), cNonBondedModel (src/modelClass_nonbonded.f90:36-50),
// cCoulModel (src/modelClass_coul.f90:28-37) and the concrete models. Sharing one set of
// (eshift, fshift, Rm, factor) between a Coulomb model and the modifier machinery is deliberate:
// the reference does the same through inheritance, and results depend on it (see Q1 in DESIGN.md).
struct Model {
  uint64_t magic = MODEL_MAGIC;
  Kind kind = PAIR_NONE;
  const char* name = "none";
  // cNonBondedModel
  int modifier = NONE;
  double eshift = 0, fshift = 0, skin = 0, Rm = 0, RmSq = 0, factor = 0, Rm2fac = 0;
  // cCoulModel
  bool shifted = false, shifted_force = false, requires_kspace = false;
  double alpha = 0;
  // pair_lj_cut / pair_softcore_cut
  double epsilon = 0, sigma = 0, lambda = 0;
  double eps4 = 0, eps24 = 0, sigsq = 0;
  double prefactor = 0, prefactor6 = 0, invSigSq = 0, shift = 0;
  // coulomb models
  double damp = 0, skinWidth = 0, beta = 0, Rm2 = 0, invRm = 0;
  // bonded / kspace (kept only so that handles can be created and validated)
  double p1 = 0, p2 = 0, accuracy = 0;
};

constexpr double Pi = 3.14159265358979323846;

// uerfc: src/math.f90:35-40, 685-691 (Abramowitz-Stegun 7.1.26 times a supplied exp(-x^2))
inline double uerfc(double x, double expmx2) {
  constexpr double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027,
                   a5 = 1.061405429, p = 0.327591100;
  double t = 1.0 / (1.0 + p * x);
  return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * expmx2;
}

// quintic switch shared by the smoothed Coulomb models and apply_modifier
inline void switch_GW(double u, double c, double& G, double& WGu) {
  double u2 = u * u;
  double u3 = u * u2;
  G = 1.0 + u3 * (15.0 * u - 6.0 * u2 - 10.0);
  WGu = c * u2 * (2.0 * u - u2 - 1.0);   // caller multiplies by (factor * r) or (factor * r2)
}

// ---- model bodies: *_compute / *_energy / *_virial of each model file ----------------------------
// MODE: 0 = compute (E and W), 1 = energy only, 2 = virial only.
template <int MODE>
inline void model_eval(const Model& m, double& Eij, double& Wij, double invR, double invR2) {
  switch (m.kind) {
    case PAIR_NONE:   // src/modelClass_pair.f90:165-190
      if (MODE != 2) Eij = 0.0;
      if (MODE != 1) Wij = 0.0;
      break;
    case COUL_NONE:   // src/modelClass_coul.f90:150-177 (coul_none_virial leaves Wij untouched)
      if (MODE == 0) { Eij = 0.0; Wij = 0.0; }
      if (MODE == 1) Eij = 0.0;
      break;
    case PAIR_LJ_CUT: {   // src/pair_lj_cut.f90:73-118
      double sr2 = m.sigsq * invR2;
      double sr6 = sr2 * sr2 * sr2;
      double sr12 = sr6 * sr6;
      if (MODE != 2) Eij = m.eps4 * (sr12 - sr6);
      if (MODE != 1) Wij = m.eps24 * (sr12 + sr12 - sr6);
      break;
    }
    case PAIR_SOFTCORE_CUT: {   // src/pair_softcore_cut.f90:86-135
      double rsig2 = m.invSigSq / invR2;
      double rsig6 = rsig2 * rsig2 * rsig2;
      double sinv = 1.0 / (rsig6 + m.shift);
      double sinvSq = sinv * sinv;
      double sinvCb = sinv * sinvSq;
      if (MODE != 2) Eij = m.prefactor * (sinvSq - sinv);
      if (MODE != 1) Wij = m.prefactor6 * rsig6 * (sinvCb + sinvCb - sinvSq);
      break;
    }
    case COUL_CUT:   // src/coul_cut.f90:62-94
      if (MODE != 2) Eij = invR;
      if (MODE != 1) Wij = invR;
      break;
    case COUL_SF: {   // src/coul_sf.f90:61-94
      double rFc = m.fshift / invR;
      if (MODE != 2) Eij = invR + m.eshift + rFc;
      if (MODE != 1) Wij = invR - rFc;
      break;
    }
    case COUL_DAMPED:   // src/coul_damped.f90:71-114
    case COUL_LONG: {   // src/coul_long.f90:80-123
      double x = m.alpha / invR;
      double expmx2 = std::exp(-x * x);
      double E = uerfc(x, expmx2) * invR;
      if (MODE != 2) Eij = E;
      if (MODE != 1) Wij = E + m.beta * expmx2;
      break;
    }
    case COUL_DAMPED_SMOOTHED: {   // src/coul_damped_smoothed.f90:92-166
      double r = 1.0 / invR;
      double x = m.alpha * r;
      double expmx2 = std::exp(-x * x);
      double E = uerfc(x, expmx2) * invR;
      double W = E + m.beta * expmx2;
      if (r > m.Rm) {
        double u = m.factor * (r - m.Rm);
        double G, WG;
        switch_GW(u, -30.0, G, WG);
        WG = WG * m.factor * r;
        W = W * G + E * WG;
        E = E * G;
      }
      if (MODE != 2) Eij = E;
      if (MODE != 1) Wij = W;
      break;
    }
    case COUL_DAMPED_SQUARE_SMOOTHED: {   // src/coul_damped_square_smoothed.f90:91-165
      double x = m.alpha / invR;
      double expmx2 = std::exp(-x * x);
      double E = uerfc(x, expmx2) * invR;
      double W = E + m.beta * expmx2;
      if (invR < m.invRm) {
        double r2 = 1.0 / invR2;
        double u = m.factor * (r2 - m.Rm2);
        double G, WG;
        switch_GW(u, -60.0, G, WG);
        WG = WG * m.factor * r2;
        W = W * G + E * WG;
        E = E * G;
      }
      if (MODE != 2) Eij = E;
      if (MODE != 1) Wij = W;
      break;
    }
    case COUL_SQUARE_SMOOTHED: {   // src/coul_square_smoothed.f90:83-151
      double W = invR, E = invR;
      if (invR < m.invRm) {
        double r2 = 1.0 / invR2;
        double u = m.factor * (r2 - m.Rm2);
        double G, WG;
        switch_GW(u, -60.0, G, WG);
        WG = WG * m.factor * r2;
        if (MODE == 2) W = W * (G + WG);       // coul_square_smoothed_virial, line 148
        else W = W * G + E * WG;
        E = E * G;
      }
      if (MODE != 2) Eij = E;
      if (MODE != 1) Wij = W;
      break;
    }
    case COUL_SHIFTED_SQUARE_SMOOTHED: {   // src/coul_shifted_square_smoothed.f90:86-154
      double W = invR;
      double E = W + m.eshift;
      if (invR < m.invRm) {
        double r2 = 1.0 / invR2;
        double u = m.factor * (r2 - m.Rm2);
        double G, WG;
        switch_GW(u, -60.0, G, WG);
        WG = WG * m.factor * r2;
        if (MODE == 2) W = W * G + (W + m.eshift) * WG;   // ..._virial, line 151
        else W = W * G + E * WG;
        E = E * G;
      }
      if (MODE != 2) Eij = E;
      if (MODE != 1) Wij = W;
      break;
    }
    default:
      break;
  }
}

// src/apply_modifier.f90:1-60
template <bool COMPUTE>
inline void apply_modifier(const Model& m, double& Eij, double& Wij, double invR, double invR2) {
  switch (m.modifier) {
    case SHIFTED:
      if (COMPUTE) Eij = Eij + m.eshift;
      break;
    case SHIFTED_FORCE: {
      double rFc = m.fshift / invR;
      Wij = Wij - rFc;
      if (COMPUTE) Eij = Eij + m.eshift + rFc;
      break;
    }
    case SMOOTHED:
    case SHIFTED_SMOOTHED:
    case SQUARE_SMOOTHED:
    case SHIFTED_SQUARE_SMOOTHED: {
      const bool square = (m.modifier == SQUARE_SMOOTHED || m.modifier == SHIFTED_SQUARE_SMOOTHED);
      if (COMPUTE) Eij = Eij + m.eshift;
      double r2fac = square ? m.factor / invR2 : m.factor / invR;
      if (r2fac > m.Rm2fac) {
        if (!COMPUTE) {
          double dummy = 0.0;
          model_eval<1>(m, Eij, dummy, invR, invR2);
          Eij = Eij + m.eshift;
        }
        double u = r2fac - m.Rm2fac;
        double G, WG;
        switch_GW(u, square ? -60.0 : -30.0, G, WG);
        WG = WG * r2fac;
        Wij = Wij * G + Eij * WG;
        Eij = Eij * G;
      }
      break;
    }
    default:
      break;
  }
}

// src/modelClass_nonbonded.f90:245-296
void modifier_setup(Model& m, double cutoff) {
  double Ec = 0, Wc = 0, Es = 0, Ws = 0;
  m.fshift = 0.0;
  m.eshift = 0.0;
  m.Rm = cutoff - m.skin;
  m.RmSq = m.Rm * m.Rm;
  bool shifting = m.modifier == SHIFTED || m.modifier == SHIFTED_FORCE ||
                  m.modifier == SHIFTED_SMOOTHED || m.modifier == SHIFTED_SQUARE_SMOOTHED;
  if (shifting) model_eval<0>(m, Ec, Wc, 1.0 / cutoff, 1.0 / (cutoff * cutoff));
  switch (m.modifier) {
    case SHIFTED:
      m.eshift = -Ec;
      break;
    case SHIFTED_FORCE:
      m.eshift = -(Ec + Wc);
      m.fshift = Wc / cutoff;
      break;
    case SMOOTHED:
    case SHIFTED_SMOOTHED:
      if (shifting) {
        model_eval<0>(m, Es, Ws, 1.0 / m.Rm, 1.0 / m.RmSq);
        m.eshift = -0.5 * (Es + Ec);
      }
      m.factor = 1.0 / (cutoff - m.Rm);
      m.Rm2fac = m.factor * m.Rm;
      break;
    case SQUARE_SMOOTHED:
    case SHIFTED_SQUARE_SMOOTHED:
      if (shifting) {
        model_eval<0>(m, Es, Ws, 1.0 / m.Rm, 1.0 / m.RmSq);
        m.eshift = -0.5 * (Es + Ec);
      }
      m.factor = 1.0 / (cutoff * cutoff - m.RmSq);
      m.Rm2fac = m.factor * m.RmSq;
      break;
    default:
      break;
  }
}

// *_apply_cutoff of the smoothed Coulomb models
void apply_cutoff(Model& m, double Rc) {
  switch (m.kind) {
    case COUL_DAMPED_SMOOTHED:   // src/coul_damped_smoothed.f90:76-85
      m.Rm = Rc - m.skinWidth;
      m.Rm2 = m.Rm * m.Rm;
      m.invRm = 1.0 / m.Rm;
      m.factor = 1.0 / (Rc - m.Rm);
      break;
    case COUL_DAMPED_SQUARE_SMOOTHED:    // src/coul_damped_square_smoothed.f90:76-84
    case COUL_SQUARE_SMOOTHED:           // src/coul_square_smoothed.f90:69-76
    case COUL_SHIFTED_SQUARE_SMOOTHED:   // src/coul_shifted_square_smoothed.f90:72-79
      m.Rm2 = (Rc - m.skinWidth) * (Rc - m.skinWidth);
      m.invRm = 1.0 / (Rc - m.skinWidth);
      m.factor = 1.0 / (Rc * Rc - m.Rm2);
      break;
    default:   // src/modelClass_coul.f90:97-101
      break;
  }
}

// src/modelClass_coul.f90:62-93
void cutoff_setup(Model& m, double cutoff) {
  m.fshift = 0.0;
  m.eshift = 0.0;
  if (m.shifted || m.shifted_force) {
    double invR = 1.0 / cutoff;
    double invR2 = invR * invR;
    double E = 0, W = 0;
    model_eval<0>(m, E, W, invR, invR2);
    if (m.shifted_force) {
      m.fshift = W / cutoff;
      m.eshift = -(E + W);
    } else {
      m.fshift = 0.0;
      m.eshift = -E;
    }
  }
  apply_cutoff(m, cutoff);
}

// model setup routines
void setup_lj(Model& m, double epsilon, double sigma) {   // src/pair_lj_cut.f90:52-69
  m.kind = PAIR_LJ_CUT;
  m.name = "lj_cut";
  m.epsilon = epsilon;
  m.sigma = sigma;
  m.eps4 = 4.0 * epsilon;
  m.eps24 = 24.0 * epsilon;
  m.sigsq = sigma * sigma;
}
void setup_softcore(Model& m, double epsilon, double sigma, double lambda) {   // src/pair_softcore_cut.f90:58-82
  m.kind = PAIR_SOFTCORE_CUT;
  m.name = "softcore_cut";
  m.epsilon = epsilon;
  m.sigma = sigma;
  m.lambda = lambda;
  if (lambda < 0.0 || lambda > 1.0) error("pair_softcore_cut setup", "out-of-range parameter lambda");
  m.prefactor = 4.0 * epsilon * lambda;          // lambda**exponent_n, exponent_n = 1
  m.prefactor6 = 6.0 * m.prefactor;
  m.invSigSq = 1.0 / (sigma * sigma);
  m.shift = 0.5 * (1.0 - lambda);                // alpha*(1-lambda)**exponent_p, alpha = 1/2
}

// *_mix: returns true and fills `mixed` when a rule exists (src/pair_lj_cut.f90:122-138,
// src/pair_softcore_cut.f90:139-163, src/modelClass_pair.f90:194-200)
bool model_mix(const Model& self, const Model& other, Model& mixed) {
  mixed = Model();
  switch (self.kind) {
    case PAIR_NONE:
      mixed.kind = PAIR_NONE;
      mixed.name = "none";
      return true;
    case PAIR_LJ_CUT:
      if (other.kind == PAIR_LJ_CUT) {
        setup_lj(mixed, std::sqrt(self.epsilon * other.epsilon), 0.5 * (self.sigma + other.sigma));
        return true;
      }
      return false;
    case PAIR_SOFTCORE_CUT:
      if (other.kind == PAIR_SOFTCORE_CUT) {
        setup_softcore(mixed, std::sqrt(self.epsilon * other.epsilon), 0.5 * (self.sigma + other.sigma),
                       self.lambda * other.lambda);
        return true;
      }
      if (other.kind == PAIR_LJ_CUT) {
        setup_softcore(mixed, std::sqrt(self.epsilon * other.epsilon), 0.5 * (self.sigma + other.sigma),
                       self.lambda);
        return true;
      }
      return false;
    default:
      return false;
  }
}

// src/modelClass_pair.f90:66-74
struct PairContainer {
  Model model;
  bool coulomb = false;
  double kCoul = 0.0;
};

// src/modelClass_pair.f90:120-140
PairContainer container_mix(const PairContainer& a, const PairContainer& b) {
  PairContainer c;
  if (!model_mix(b.model, a.model, c.model)) {
    if (!model_mix(a.model, b.model, c.model)) {
      c.model = Model();
      warning(std::string("no mixing rule found for models ") + a.model.name + " and " + b.model.name);
    }
  }
  c.coulomb = a.coulomb && b.coulomb;
  if (c.coulomb) c.kCoul = std::sqrt(a.kCoul * b.kCoul);
  return c;
}

// ---- lists (src/lists.f90:26-99) ----------------------------------------------------------------
struct List {
  int nitems = 0, nobjects = 0, count = 0;
  std::vector<int> first, middle, last, item;
  std::vector<double> value;
  bool has_value = false;
  void allocate(int nitems_, int nobjects_, bool middle_ = false, bool value_ = false) {
    nobjects = nobjects_;
    nitems = nitems_;
    count = 0;
    first.assign(nobjects, 1);
    last.assign(nobjects, 0);
    item.assign(nitems, 0);
    if (middle_) middle.assign(nobjects, 0);
    has_value = value_;
    if (value_) value.assign(nitems, 0.0);
  }
  void resize(int size) {
    item.resize(size);
    if (has_value) value.resize(size);
    nitems = size;
  }
};

// ---- RNG: KISS + ziggurat normal (src/math.f90:42-177) -------------------------------------------
struct Kiss {
  bool seeding_required = true;
  int32_t kn[128];
  double wn[128], fn[128];
  uint32_t x, y, z, w;   // unsigned storage, two's-complement wraparound like gfortran's integer(4)
  static uint32_t m(uint32_t k, int n) { return k ^ (n >= 0 ? (k << n) : (k >> (-n))); }
  void init(int32_t seed) {   // src/math.f90:150-163
    x = m(m(m((uint32_t)seed, 13), -17), 5);
    y = m(m(m(x, 13), -17), 5);
    z = m(m(m(y, 13), -17), 5);
    w = m(m(m(z, 13), -17), 5);
  }
  int32_t i32() {   // src/math.f90:167-177
    x = 69069u * x + 1327217885u;
    y ^= (y << 13);
    y ^= (y >> 17);
    y ^= (y << 5);
    z = 18000u * (z & 65535u) + (z >> 16);
    w = 30903u * (w & 65535u) + (w >> 16);
    return (int32_t)(x + y + (z << 16) + w);
  }
  void setup(int32_t seed) {   // src/math.f90:80-105
    const double m1 = 2147483648.0;
    init(seed);
    seeding_required = false;
    double dn = 3.442619855899, tn = 3.442619855899, vn = 0.00991256303526217;
    double q = vn * std::exp(0.5 * dn * dn);
    kn[0] = (int32_t)((dn / q) * m1);
    kn[1] = 0;
    wn[0] = q / m1;
    wn[127] = dn / m1;
    fn[0] = 1.0;
    fn[127] = std::exp(-0.5 * dn * dn);
    for (int i = 126; i >= 1; --i) {
      dn = std::sqrt(-2.0 * std::log(vn / dn + std::exp(-0.5 * dn * dn)));
      kn[i + 1] = (int32_t)((dn / tn) * m1);
      tn = dn;
      fn[i] = std::exp(-0.5 * dn * dn);
      wn[i] = dn / m1;
    }
  }
  static int32_t iabs(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }
  double normal() {   // src/math.f90:109-146
    const double r = 3.442620, s = 0.2328306e-9;
    int32_t hz = i32();
    int32_t iz = hz & 127;
    if (iabs(hz) < kn[iz]) return hz * wn[iz];
    for (;;) {
      if (iz == 0) {
        double xx, yy;
        do {
          xx = -0.2904764 * std::log(s * i32() + 0.5);
          yy = -std::log(s * i32() + 0.5);
        } while (!(yy + yy >= xx * xx));
        double rnor = r + xx;
        if (hz <= 0) rnor = -rnor;
        return rnor;
      }
      double xx = hz * wn[iz];
      if (fn[iz] + (s * i32() + 0.5) * (fn[iz - 1] - fn[iz]) < std::exp(-0.5 * xx * xx)) return xx;
      hz = i32();
      iz = hz & 127;
      if (iabs(hz) < kn[iz]) return hz * wn[iz];
    }
  }
};

// src/math.f90:230-237
inline double phi(double x) {
  if (std::fabs(x) > 1e-4) return (1.0 - std::exp(-x)) / x;
  return 1.0 + 0.5 * x * ((1.0 / 3.0) * x * (1.0 - 0.25 * x) - 1.0);
}

// src/math.f90:642-657
double inverse_of_x_plus_ln_x(double y) {
  const double tol = 1.0e-12;
  double x = (y > 0.5671432904097839) ? y - std::log(y) : std::exp(y);
  double x0 = x + 1.0;
  while (std::fabs(x - x0) > tol * x0) {
    x0 = x;
    x = x * (y + 1.0 - std::log(x)) / (x + 1.0);
  }
  return x;
}

// ---- helpers for the rigid-body integrator (src/math.f90:181-226, 428-436) ------------------------
inline void cross3(const double a[3], const double b[3], double c[3]) {   // src/math.f90:181-185
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double sign1(double x) { return std::signbit(x) ? -1.0 : 1.0; }   // sign(one, x)
inline int staircase(double x) {   // src/math.f90:428-436
  return x > 0.0 ? (int)std::ceil(x - 0.5) : (int)std::floor(x + 0.5);
}

// The four quaternion matrices of src/ArBee.f90:134-176, applied to a vector (Fortran reshape is
// column-major: column k of B / C multiplies v(k); Bt / Ct are their transposes).
inline void mulB(const double q[4], const double v[3], double o[4]) {
  o[0] = -q[1] * v[0] - q[2] * v[1] - q[3] * v[2];
  o[1] = q[0] * v[0] - q[3] * v[1] + q[2] * v[2];
  o[2] = q[3] * v[0] + q[0] * v[1] - q[1] * v[2];
  o[3] = -q[2] * v[0] + q[1] * v[1] + q[0] * v[2];
}
inline void mulC(const double q[4], const double v[3], double o[4]) {
  o[0] = -q[1] * v[0] - q[2] * v[1] - q[3] * v[2];
  o[1] = q[0] * v[0] + q[3] * v[1] - q[2] * v[2];
  o[2] = -q[3] * v[0] + q[0] * v[1] + q[1] * v[2];
  o[3] = q[2] * v[0] - q[1] * v[1] + q[0] * v[2];
}
inline void mulBt(const double q[4], const double p[4], double o[3]) {
  o[0] = -q[1] * p[0] + q[0] * p[1] + q[3] * p[2] - q[2] * p[3];
  o[1] = -q[2] * p[0] - q[3] * p[1] + q[0] * p[2] + q[1] * p[3];
  o[2] = -q[3] * p[0] + q[2] * p[1] - q[1] * p[2] + q[0] * p[3];
}
inline void mulCt(const double q[4], const double p[4], double o[3]) {
  o[0] = -q[1] * p[0] + q[0] * p[1] - q[3] * p[2] + q[2] * p[3];
  o[1] = -q[2] * p[0] + q[3] * p[1] + q[0] * p[2] - q[1] * p[3];
  o[2] = -q[3] * p[0] - q[2] * p[1] + q[1] * p[2] + q[0] * p[3];
}

// src/math.f90:189-218 (Shepperd's method); A is row-major here, A[i][j] = A(i+1,j+1)
void quaternion_from_matrix(const double A[3][3], double Q[4]) {
  const double a11 = A[0][0], a22 = A[1][1], a33 = A[2][2];
  const double Q2[4] = {1.0 + a11 + a22 + a33, 1.0 + a11 - a22 - a33, 1.0 - a11 + a22 - a33, 1.0 - a11 - a22 + a33};
  int imax = 0;
  for (int k = 1; k < 4; ++k)
    if (Q2[k] > Q2[imax]) imax = k;   // maxloc: first maximum
  const double Q2max = Q2[imax], f = std::sqrt(1.0 / Q2max);
  double v[4];
  switch (imax) {
    case 0: v[0] = Q2max; v[1] = A[1][2] - A[2][1]; v[2] = A[2][0] - A[0][2]; v[3] = A[0][1] - A[1][0]; break;
    case 1: v[0] = A[1][2] - A[2][1]; v[1] = Q2max; v[2] = A[0][1] + A[1][0]; v[3] = A[0][2] + A[2][0]; break;
    case 2: v[0] = A[2][0] - A[0][2]; v[1] = A[0][1] + A[1][0]; v[2] = Q2max; v[3] = A[1][2] + A[2][1]; break;
    default: v[0] = A[0][1] - A[1][0]; v[1] = A[0][2] + A[2][0]; v[2] = A[1][2] + A[2][1]; v[3] = Q2max; break;
  }
  for (int k = 0; k < 4; ++k) Q[k] = 0.5 * v[k] * f;
}

// src/math.f90:244-424 (Kopp's analytical 3x3 symmetric eigen-solver: Cardano eigenvalues sorted by
// decreasing magnitude, eigenvectors from cross products). `m` holds the upper triangle; q[i][k] is
// component i of eigenvector k, as in the reference's q(i,k).
void diagonalization(const double m[3][3], double q[3][3], double w[3]) {
  const double eps = 2.220446049250313e-16, sqrt3 = std::sqrt(3.0);
  double a[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = m[i][j];
  const double de = a[0][1] * a[1][2], dd = a[0][1] * a[0][1], ee = a[1][2] * a[1][2], ff = a[0][2] * a[0][2];
  const double tr = a[0][0] + a[1][1] + a[2][2];
  const double c1 = (a[0][0] * a[1][1] + a[0][0] * a[2][2] + a[1][1] * a[2][2]) - (dd + ee + ff);
  const double c0 = 27.0 * (a[2][2] * dd + a[0][0] * ee + a[1][1] * ff - a[0][0] * a[1][1] * a[2][2] - 2.0 * a[0][2] * de);
  double p = tr * tr - 3.0 * c1;
  const double r = tr * (p - 1.5 * c1) - 0.5 * c0;
  const double sqrtp = std::sqrt(std::fabs(p));
  const double ang = (1.0 / 3.0) * std::atan2(std::sqrt(std::fabs(6.75 * c1 * c1 * (p - c1) + c0 * (r + 0.25 * c0))), r);
  const double c = sqrtp * std::cos(ang), s = (1.0 / sqrt3) * sqrtp * std::sin(ang);
  p = (1.0 / 3.0) * (tr - c);
  double w1 = p + c, w2 = p - s, w3 = p + s;
  if (std::fabs(w1) < std::fabs(w3)) std::swap(w1, w3);
  if (std::fabs(w1) < std::fabs(w2)) std::swap(w1, w2);
  if (std::fabs(w2) < std::fabs(w3)) std::swap(w2, w3);
  w[0] = w1; w[1] = w2; w[2] = w3;
  const double wmax8eps = 8.0 * eps * std::fabs(w1), thresh = wmax8eps * wmax8eps;

  // compute_eigenvector (379-416): normalise v = (A - w).e1 x (A - w).e2, with the degenerate-column exits
  auto finish_vector = [&](double v[3], double n1tmp, double n2tmp) {
    const double norm = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double n1 = n1tmp + a[0][0] * a[0][0], n2 = n2tmp + a[1][1] * a[1][1], err = n1 * n2;
    if (n1 <= thresh) {
      v[0] = 1.0; v[1] = 0.0; v[2] = 0.0;
    } else if (n2 <= thresh) {
      v[0] = 0.0; v[1] = 1.0; v[2] = 0.0;
    } else if (norm < (64.0 * eps) * (64.0 * eps) * err) {
      double t = std::fabs(a[0][1]), f = -a[0][0] / a[0][1];
      if (std::fabs(a[1][1]) > t) { t = std::fabs(a[1][1]); f = -a[0][1] / a[1][1]; }
      if (std::fabs(a[1][2]) > t) f = -a[0][2] / a[1][2];
      const double nn = 1.0 / std::sqrt(1.0 + f * f);
      v[0] = nn; v[1] = f * nn; v[2] = 0.0;
    } else {
      const double sc = std::sqrt(1.0 / norm);
      for (int i = 0; i < 3; ++i) v[i] *= sc;
    }
  };

  double n1 = a[0][1] * a[0][1] + a[0][2] * a[0][2], n2 = a[0][1] * a[0][1] + a[1][2] * a[1][2];
  const double q12 = a[0][1] * a[1][2] - a[0][2] * a[1][1];   // q(1,2) = q(1,1) before the shift of the diagonal
  const double q22 = a[0][2] * a[0][1] - a[1][2] * a[0][0];
  const double q32 = a[0][1] * a[0][1];
  a[0][0] -= w1;
  a[1][1] -= w1;
  double v1[3] = {q12 + a[0][2] * w1, q22 + a[1][2] * w1, a[0][0] * a[1][1] - q32};
  finish_vector(v1, n1, n2);
  double v2[3] = {0, 0, 0};
  const double t = w1 - w2;
  if (std::fabs(t) > wmax8eps) {
    a[0][0] += t;
    a[1][1] += t;
    v2[0] = q12 + a[0][2] * w2; v2[1] = q22 + a[1][2] * w2; v2[2] = a[0][0] * a[1][1] - q32;
    finish_vector(v2, n1, n2);
  } else {   // degenerate pair (340-375)
    a[1][0] = a[0][1];
    a[2][0] = a[0][2];
    a[2][1] = a[1][2];
    a[0][0] += w1;
    a[1][1] += w1;
    bool success = false;
    for (int i = 0; i < 3 && !success; ++i) {
      a[i][i] -= w2;
      const double col[3] = {a[0][i], a[1][i], a[2][i]};
      n1 = col[0] * col[0] + col[1] * col[1] + col[2] * col[2];
      success = n1 > thresh;
      if (success) {
        cross3(v1, col, v2);
        const double norm = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
        success = norm > (256.0 * eps) * (256.0 * eps) * n1;
        if (success) {
          const double sc = std::sqrt(1.0 / norm);
          for (int x = 0; x < 3; ++x) v2[x] *= sc;
        }
      }
    }
    if (!success) {   // any vector orthogonal to v1 will do
      int i = 0;
      while (v1[i] == 0.0) ++i;
      const int j = (i + 1) % 3;
      const double nn = 1.0 / std::sqrt(v1[i] * v1[i] + v1[j] * v1[j]);
      v2[i] = v1[j] * nn;
      v2[j] = -v1[i] * nn;
      v2[(i + 2) % 3] = 0.0;
    }
  }
  double v3[3];
  cross3(v1, v2, v3);
  for (int i = 0; i < 3; ++i) { q[i][0] = v1[i]; q[i][1] = v2[i]; q[i][2] = v3[i]; }
}

// src/math.f90:505-546, 550-577, 581-638: Carlson's symmetric elliptic integrals (duplication method)
const double quiet_NaN = std::numeric_limits<double>::quiet_NaN();
double Carlson_RF(double x, double y, double z) {
  const double errtol = 0.001, tiny = 2.2250738585072014e-308, huge = 1.7976931348623157e308;
  const double lolim = std::cbrt(5.0 * tiny), uplim = 0.3 * std::cbrt(0.2 * huge);
  if (std::min({x, y, z}) < 0.0 || std::max({x, y, z}) > uplim || std::min({x + y, x + z, y + z}) < lolim) return quiet_NaN;
  double xn = x, yn = y, zn = z, mu, xd, yd, zd;
  for (;;) {
    mu = (xn + yn + zn) / 3.0;
    xd = 2.0 - (mu + xn) / mu;
    yd = 2.0 - (mu + yn) / mu;
    zd = 2.0 - (mu + zn) / mu;
    if (std::max({std::fabs(xd), std::fabs(yd), std::fabs(zd)}) < errtol) break;
    const double xr = std::sqrt(xn), yr = std::sqrt(yn), zr = std::sqrt(zn);
    const double lam = xr * (yr + zr) + yr * zr;
    xn = (xn + lam) * 0.25;
    yn = (yn + lam) * 0.25;
    zn = (zn + lam) * 0.25;
  }
  const double e2 = xd * yd - zd * zd, e3 = xd * yd * zd;
  const double s = 1.0 + ((1.0 / 24.0) * e2 - 0.10 - (3.0 / 44.0) * e3) * e2 + (1.0 / 14.0) * e3;
  return s / std::sqrt(mu);
}
double Carlson_RC(double x, double y) {
  const double errtol = 0.001, tiny = 2.2250738585072014e-308, huge = 1.7976931348623157e308;
  if (x < 0.0 || y <= 0.0 || std::max(x, y) > 0.2 * huge || x + y < 5.0 * tiny) return quiet_NaN;
  double xn = x, yn = y, mu, sn;
  for (;;) {
    mu = (xn + yn + yn) / 3.0;
    sn = (yn + mu) / mu - 2.0;
    if (std::fabs(sn) < errtol) break;
    const double lam = 2.0 * std::sqrt(xn) * std::sqrt(yn) + yn;
    xn = (xn + lam) * 0.25;
    yn = (yn + lam) * 0.25;
  }
  const double s = sn * sn * (0.30 + sn * ((1.0 / 7.0) + sn * (0.375 + sn * (9.0 / 22.0))));
  return (1.0 + s) / std::sqrt(mu);
}
double Carlson_RJ(double x, double y, double z, double p) {
  const double errtol = 0.001, tiny = 2.2250738585072014e-308, huge = 1.7976931348623157e308;
  const double lolim = std::cbrt(5.0 * tiny), uplim = 0.30 * std::cbrt(0.2 * huge);
  const double c1 = 3.0 / 14.0, c2 = 1.0 / 3.0, c3 = 3.0 / 22.0, c4 = 3.0 / 26.0;
  if (std::min({x, y, z}) < 0.0 || std::max({x, y, z, p}) > uplim || std::min({x + y, x + z, y + z, p}) < lolim) return quiet_NaN;
  double xn = x, yn = y, zn = z, pn = p, sigma = 0.0, power4 = 1.0, mu, xd, yd, zd, pd;
  for (;;) {
    mu = (xn + yn + zn + pn + pn) * 0.20;
    xd = (mu - xn) / mu;
    yd = (mu - yn) / mu;
    zd = (mu - zn) / mu;
    pd = (mu - pn) / mu;
    if (std::max({std::fabs(xd), std::fabs(yd), std::fabs(zd), std::fabs(pd)}) < errtol) break;
    const double xr = std::sqrt(xn), yr = std::sqrt(yn), zr = std::sqrt(zn);
    const double lam = xr * (yr + zr) + yr * zr;
    double alfa = pn * (xr + yr + zr) + xr * yr * zr;
    alfa = alfa * alfa;
    const double beta = pn * (pn + lam) * (pn + lam);
    sigma = sigma + power4 * Carlson_RC(alfa, beta);
    power4 = power4 * 0.25;
    xn = (xn + lam) * 0.25;
    yn = (yn + lam) * 0.25;
    zn = (zn + lam) * 0.25;
    pn = (pn + lam) * 0.25;
  }
  const double ea = xd * (yd + zd) + yd * zd, eb = xd * yd * zd, ec = pd * pd;
  const double e2 = ea - 3.0 * ec, e3 = eb + 2.0 * pd * (ea - ec);
  const double s1 = 1.0 + e2 * (-c1 + 0.75 * c3 * e2 - 1.5 * c4 * e3);
  const double s2 = eb * (0.5 * c2 + pd * (-c3 - c3 + pd * c4));
  const double s3 = pd * ea * (c2 - pd * c3) - c2 * pd * ec;
  return 3.0 * sigma + power4 * (s1 + s2 + s3) / (mu * std::sqrt(mu));
}

// src/math.f90:440-501: Jacobi elliptic functions sn, cn, dn by descending Landen (AGM) steps
void jacobi(double u, double m, double jac[3]) {
  const int NN = 16;
  const double eps = 2.220446049250313e-16;
  if (std::fabs(m) > 1.0) {
    jac[0] = jac[1] = jac[2] = quiet_NaN;
  } else if (std::fabs(m) < 2.0 * eps) {
    jac[0] = std::sin(u); jac[1] = std::cos(u); jac[2] = 1.0;
  } else if (std::fabs(m - 1.0) < 2.0 * eps) {
    jac[0] = std::tanh(u);
    jac[1] = jac[2] = 1.0 / std::cosh(u);
  } else {
    double mu[NN], nu[NN], c[NN], d[NN];
    int n = 0;
    mu[0] = 1.0;
    nu[0] = std::sqrt(1.0 - m);
    while (std::fabs(mu[n] - nu[n]) > 4.0 * eps * std::fabs(mu[n] + nu[n])) {
      mu[n + 1] = 0.5 * (mu[n] + nu[n]);
      nu[n + 1] = std::sqrt(mu[n] * nu[n]);
      ++n;
      if (n >= NN - 1) { jac[0] = jac[1] = jac[2] = quiet_NaN; return; }
    }
    const double sin_umu = std::sin(u * mu[n]), cos_umu = std::cos(u * mu[n]);
    const bool small_sin = std::fabs(sin_umu) < std::fabs(cos_umu);
    const double t = small_sin ? sin_umu / cos_umu : cos_umu / sin_umu;
    c[n] = mu[n] * t;
    d[n] = 1.0;
    while (n > 0) {
      c[n - 1] = d[n] * c[n];
      const double r = (c[n] * c[n]) / mu[n];
      --n;
      d[n] = (r + nu[n]) / (r + mu[n]);
    }
    double sn, cn, dn;
    if (small_sin) {
      dn = std::sqrt(1.0 - m) / d[n];
      cn = dn * sign1(cos_umu) / std::hypot(1.0, c[n]);
      sn = cn * c[n] / std::sqrt(1.0 - m);
    } else {
      dn = d[n];
      sn = sign1(sin_umu) / std::hypot(1.0, c[n]);
      cn = c[n] * sn;
    }
    jac[0] = sn; jac[1] = cn; jac[2] = dn;
  }
}

// ---- rigid bodies (src/ArBee.f90) ------------------------------------------------------------------
struct Body {
  int NP = 0;
  int dof = 6;
  double mass = 0, invMass = 0;
  double MoI[3] = {0, 0, 0}, invMoI[3] = {0, 0, 0};
  double rcm[3] = {0, 0, 0}, pcm[3] = {0, 0, 0};
  double q[4] = {0, 0, 0, 0}, pi[4] = {0, 0, 0, 0};
  double omega[3] = {0, 0, 0}, F[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
  double I113 = 0, I313 = 0, I223 = 0, I212 = 0, m312 = 0;
  std::vector<int> index;        // 0-based atom indices
  std::vector<double> M;         // masses
  std::vector<double> d;         // 3*NP, body-frame positions
  std::vector<double> delta;     // 3*NP, space-frame positions relative to rcm

  // tBody_update (97-130): centre of mass, inertia tensor -> principal frame, quaternion, body-frame coordinates
  void update(const std::vector<double>& coords) {
    for (int x = 0; x < 3; ++x) {
      double s = 0.0;
      for (int i = 0; i < NP; ++i) s += M[i] * coords[3 * i + x];
      rcm[x] = s * invMass;
    }
    for (int i = 0; i < NP; ++i)
      for (int x = 0; x < 3; ++x) delta[3 * i + x] = coords[3 * i + x] - rcm[x];
    double inertia[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < NP; ++i) {
      const double dx = delta[3 * i], dy = delta[3 * i + 1], dz = delta[3 * i + 2];
      inertia[0][0] += M[i] * (dy * dy + dz * dz);
      inertia[1][1] += M[i] * (dx * dx + dz * dz);
      inertia[2][2] += M[i] * (dx * dx + dy * dy);
      inertia[0][1] += M[i] * dx * dy;
      inertia[0][2] += M[i] * dx * dz;
      inertia[1][2] += M[i] * dy * dz;
    }
    inertia[0][1] = -inertia[0][1];
    inertia[0][2] = -inertia[0][2];
    inertia[1][2] = -inertia[1][2];
    double vec[3][3], A[3][3];
    diagonalization(inertia, vec, MoI);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i][j] = vec[j][i];   // A = transpose(eigenvector matrix): rows are the principal axes
    for (int x = 0; x < 3; ++x) invMoI[x] = 1.0 / MoI[x];
    I113 = 1.0 / (MoI[0] * (MoI[0] - MoI[2]));
    I313 = 1.0 / (MoI[2] * (MoI[0] - MoI[2]));
    I223 = 1.0 / (MoI[1] * (MoI[1] - MoI[2]));
    I212 = 1.0 / (MoI[1] * (MoI[0] - MoI[1]));
    m312 = (MoI[2] - MoI[0]) / MoI[1];
    quaternion_from_matrix(A, q);
    for (int i = 0; i < NP; ++i)
      for (int r = 0; r < 3; ++r)
        d[3 * i + r] = A[r][0] * delta[3 * i] + A[r][1] * delta[3 * i + 1] + A[r][2] * delta[3 * i + 2];
  }

  // delta = Ct(q) B(q) d, the closing line of both rotation routines (190, 274)
  void refresh_delta() {
    for (int i = 0; i < NP; ++i) {
      double t4[4];
      mulB(q, &d[3 * i], t4);
      mulCt(q, t4, &delta[3 * i]);
    }
  }

  // tBody_rotate_uniaxial (194-217)
  void rotate_uniaxial(int k, double dt) {
    double BkQ[4], BkPi[4];
    auto perm = [&](const double v[4], double o[4]) {
      if (k == 1) { o[0] = -v[1]; o[1] = v[0]; o[2] = v[3]; o[3] = -v[2]; }
      else if (k == 2) { o[0] = -v[2]; o[1] = -v[3]; o[2] = v[0]; o[3] = v[1]; }
      else { o[0] = -v[3]; o[1] = v[2]; o[2] = -v[1]; o[3] = v[0]; }
    };
    perm(q, BkQ);
    perm(pi, BkPi);
    double dot = 0.0;
    for (int x = 0; x < 4; ++x) dot += pi[x] * BkQ[x];
    const double ang = dt * dot / (4.0 * MoI[k - 1]);
    const double vs = std::sin(ang), vc = std::cos(ang);
    for (int x = 0; x < 4; ++x) {
      q[x] = vc * q[x] + vs * BkQ[x];
      pi[x] = vc * pi[x] + vs * BkPi[x];
    }
  }

  // tBody_rotate_no_squish (178-192): Miller et al. splitting, n sub-steps
  void rotate_no_squish(double delta_t, int n) {
    const double dt = delta_t / n, half_dt = 0.5 * dt;
    for (int i = 0; i < n; ++i) {
      rotate_uniaxial(3, half_dt);
      rotate_uniaxial(2, half_dt);
      rotate_uniaxial(1, dt);
      rotate_uniaxial(2, half_dt);
      rotate_uniaxial(3, half_dt);
    }
    refresh_delta();
  }

  // tBody_rotate_exact (221-313): torque-free rotation of an asymmetric top in closed form
  // (Jacobi elliptic functions for omega, Carlson integrals for the precession angle)
  void rotate_exact(double dt) {
    const double eps = 2.220446049250313e-16, Pi = 3.14159265358979324;
    const double w0[3] = {omega[0], omega[1], omega[2]};
    double Iw[3] = {MoI[0] * w0[0], MoI[1] * w0[1], MoI[2] * w0[2]};
    double Lsq = Iw[1] * Iw[1] + Iw[2] * Iw[2];
    if (Lsq < eps) {
      rotate_uniaxial(1, dt);   // returns WITHOUT refreshing delta, as the reference does
      return;
    }
    Lsq = Iw[0] * Iw[0] + Lsq;
    const double L = std::sqrt(Lsq);
    const double TwoKr = Iw[0] * w0[0] + Iw[1] * w0[1] + Iw[2] * w0[2];
    const double r1 = Lsq - TwoKr * MoI[2], r3 = TwoKr * MoI[0] - Lsq;
    const double l1 = I223 * r1, l3 = I212 * r3, lmin = std::min(l1, l3);
    double a[3] = {sign1(w0[0]) * std::sqrt(I113 * r1), std::sqrt(lmin), sign1(w0[2]) * std::sqrt(I313 * r3)};
    const double m = lmin / std::max(l1, l3);
    const double K = Carlson_RF(0.0, 1.0 - m, 1.0), inv2K = 0.5 / K;
    double s0 = w0[1] / a[1], c0, u0;
    int i0;
    if (std::fabs(s0) < 1.0) {
      c0 = (l1 < l3) ? w0[0] / a[0] : w0[2] / a[2];
      u0 = s0 * Carlson_RF(1.0 - s0 * s0, 1.0 - m * s0 * s0, 1.0);
      i0 = staircase(u0 * inv2K);
    } else {
      a[1] = std::fabs(w0[1]);
      s0 = sign1(s0);
      c0 = 0.0;
      u0 = std::copysign(K, s0);
      i0 = 0;
    }
    const double wp = m312 * a[0] * a[2] / a[1];
    const double u = wp * dt + u0;
    const int jump = staircase(u * inv2K) - i0;
    double jac[3];
    jacobi(u, m, jac);
    const double sn = jac[0], cn = jac[1], dn = jac[2];
    auto Theta = [](double x, double n, double mm) {
      const double x2 = x * x;
      return -(1.0 / 3.0) * n * x * x2 * Carlson_RJ(1.0 - x2, 1.0 - mm * x2, 1.0, 1.0 + n * x2);
    };
    const double alpha = MoI[0] * a[0] / L;
    double deltaF;
    if (l1 < l3) {
      omega[0] = a[0] * cn; omega[1] = a[1] * sn; omega[2] = a[2] * dn;
      const double d0 = w0[2] / a[2];
      const double eta = alpha * alpha / (1.0 - alpha * alpha), C = std::sqrt(m + eta);
      deltaF = u - u0 + sign1(cn) * Theta(sn, eta, m) - sign1(c0) * Theta(s0, eta, m) +
               (alpha / C) * (std::atan(C * sn / dn) - std::atan(C * s0 / d0));
      if (jump != 0) deltaF = deltaF + jump * 2.0 * Theta(1.0, eta, m);
      deltaF = (eta + 1.0) * deltaF;
    } else {
      omega[0] = a[0] * dn; omega[1] = a[1] * sn; omega[2] = a[2] * cn;
      double eta = alpha * alpha;
      eta = eta / (1.0 - eta);
      const double k2eta = m * eta, C = std::sqrt(1.0 + k2eta);
      deltaF = u - u0 + sign1(cn) * Theta(sn, k2eta, m) - sign1(c0) * Theta(s0, k2eta, m) +
               (alpha / C) * (std::atan(C * sn / cn) - std::atan(C * s0 / c0));
      if (jump != 0) deltaF = deltaF + jump * (2.0 * Theta(1.0, k2eta, m) + (alpha / C) * Pi);
      deltaF = (eta + 1.0) * deltaF;
    }
    const double ang = (Lsq * (u - u0) + r3 * deltaF) / (2.0 * L * MoI[0] * wp);
    const double z0[4] = {Iw[2], Iw[1], L - Iw[0], 0.0};
    for (int x = 0; x < 3; ++x) Iw[x] = MoI[x] * omega[x];
    const double cph = std::cos(ang), sph = std::sin(ang);
    const double z[4] = {Iw[2] * cph - Iw[1] * sph, Iw[1] * cph + Iw[2] * sph, (L - Iw[0]) * cph, (L - Iw[0]) * sph};
    double z0q = 0.0;
    for (int x = 0; x < 4; ++x) z0q += z0[x] * q[x];
    double t3[3], t4[4], nq[4];
    mulCt(z0, q, t3);
    mulC(z, t3, t4);
    double nrm = 0.0;
    for (int x = 0; x < 4; ++x) { nq[x] = z[x] * z0q + t4[x]; nrm += nq[x] * nq[x]; }
    nrm = std::sqrt(nrm);
    for (int x = 0; x < 4; ++x) q[x] = nq[x] / nrm;
    const double twoIw[3] = {2.0 * Iw[0], 2.0 * Iw[1], 2.0 * Iw[2]};
    mulB(q, twoIw, pi);
    refresh_delta();
  }

  // tBody_particle_momenta (317-326)
  void particle_momenta(double* P) const {   // 3*NP
    double t4[4], w[3];
    mulB(q, omega, t4);
    mulCt(q, t4, w);   // space-frame angular velocity
    for (int k = 0; k < NP; ++k) {
      double c[3];
      cross3(w, &delta[3 * k], c);
      for (int x = 0; x < 3; ++x) P[3 * k + x] = M[k] * (invMass * pcm[x] + c[x]);
    }
  }

  // tBody_force_and_torque (330-343)
  void force_and_torque(const double* Fall) {
    for (int x = 0; x < 3; ++x) F[x] = tau[x] = 0.0;
    for (int j = 0; j < NP; ++j) {
      const double* Fj = Fall + 3 * (size_t)index[j];
      double c[3];
      cross3(&delta[3 * j], Fj, c);
      for (int x = 0; x < 3; ++x) { F[x] = F[x] + Fj[x]; tau[x] = tau[x] + c[x]; }
    }
  }

  // tBody_assign_momenta (347-357): from angular velocities (3) or from quaternion momenta (4)
  void assign_omega(const double w[3]) {
    const double v[3] = {2.0 * MoI[0] * w[0], 2.0 * MoI[1] * w[1], 2.0 * MoI[2] * w[2]};
    for (int x = 0; x < 3; ++x) omega[x] = w[x];
    mulB(q, v, pi);
  }
  void assign_pi(const double p[4]) {
    for (int x = 0; x < 4; ++x) pi[x] = p[x];
    double t3[3];
    mulBt(q, pi, t3);
    for (int x = 0; x < 3; ++x) omega[x] = 0.5 * invMoI[x] * t3[x];
  }
};

// ---- system state (src/EmDeeData.f90:66-151) -----------------------------------------------------
struct Cell { int neighbor[nbcells]; };

// What EmDee_share_phase_space aliases between two systems (src/EmDeeCode.f90:260-263: R, P, body, Lbox pointers)
struct Phase {
  bool hasL = false, hasR = false;
  double Lbox = 0;
  std::vector<double> R, P;
  std::vector<Body> body;
};

// Reciprocal-space Ewald solver: cKspaceModel (src/modelClass_kspace.f90:38-72) + kspace_ewald (src/kspace_ewald.f90:37-58)
struct Ewald {
  double alpha = 0, beta = 0, kmax = 0, Volume = 0;
  int ntypes = 0, nTypePairs = 0, nvecs = 0, nlayers = 1;
  int nmax[3] = {-1, -1, -1};
  std::vector<int> type;                     // distinct charged types, ascending (1-based type ids)
  std::vector<int> first, last, item;        // atoms of those types grouped by type (0-based items / atoms)
  std::vector<double> value;                 // their charges
  std::vector<double> Erigid;                // (nTypePairs) intrabody + self terms, unscaled by the Coulomb constant
  std::vector<double> lambda, lambda1D;      // (ntypes, ntypes, nlayers), (nTypePairs, nlayers)
  std::vector<int> n;                        // (3, nvecs)
  std::vector<double> prefac, vec;           // (nvecs), (3, nvecs)
  static int symm1D(int i, int j) {          // src/math.f90:695-702 (1-based)
    const int x = std::min(i, j) - 1, y = std::max(i, j) - 1;
    return x + (y + 1) * y / 2 + 1;
  }
  // cKspaceModel_discount (src/modelClass_kspace.f90:320-334): minus the smooth (erf) part of a pair interaction
  void discount(double& E, double& W, double rsq, double QiQj) const {
    const double r = std::sqrt(rsq), x = alpha * r, expmx2 = std::exp(-x * x);
    E = -QiQj * (1.0 - uerfc(x, expmx2)) / r;
    W = E + QiQj * beta * expmx2;
  }
};

struct System {
  int natoms = 0, mcells = 0, ncells = 0, maxcells = 0, maxatoms = 0, maxpairs = 0, ntypes = 1;
  int nbodies = 0, nfree = 0, nthreads = 1, threadAtoms = 0, threadFreeAtoms = 0, threadBodies = 0;
  int nlayers = 1, layer = 1;
  std::shared_ptr<Phase> ph = std::make_shared<Phase>();   // R, P, bodies, box (shared after EmDee_share_phase_space)
  double Rc = 0, skin = 0, RcSq = 0, xRc = 0, xRcSq = 0, skinSq = 0, InRc = 0, InRcSq = 0, xInRcSq = 0;
  double totalMass = 0, startTime = 0;
  bool initialized = false;
  Kiss random;
  List cellAtom, threadCell, excluded;
  std::vector<int> atomType, atomCell, atomBody, free_, atomsInCell;
  std::vector<double> charge, mass, invMass, R0;
  std::vector<char> charged;
  std::vector<Cell> cell;
  std::vector<List> neighbor;
  std::vector<PairContainer> pair;   // (ntypes, ntypes, nlayers)
  std::vector<Model> coul;           // (nlayers)
  Model kspace;
  Ewald ewald;
  std::vector<char> multilayer, overridable, interact, pairs_exist;
  std::vector<char> bonded, useInRc, forcesUpToDate;
  struct Struct { int i, j, k, l; Model model; };          // src/structs.f90:27-30 (1-based atom indices)
  std::vector<Struct> bonds, angles, dihedrals;
  std::vector<double> layerF;        // (3, N, nlayers)
  std::vector<tEnergy> layerEnergy;
  std::vector<tVirial> layerVirial;
  bool multilayer_coulomb = false, kspace_active = false;

  PairContainer& pr(int i, int j, int l) { return pair[(size_t)(l - 1) * ntypes * ntypes + (size_t)(j - 1) * ntypes + (i - 1)]; }
  char& tt(std::vector<char>& a, int i, int j) { return a[(size_t)(j - 1) * ntypes + (i - 1)]; }
  double* F() { return layerF.data() + (size_t)(layer - 1) * 3 * natoms; }
  double layerRc(int l) const { return useInRc[l - 1] ? InRc : Rc; }
};

System* sys(const tEmDee& md) { return static_cast<System*>(md.Data); }

std::string option_string(const char* option) {   // src/global.f90:120-128
  std::string s;
  for (int i = 0; i < 256 && option[i] != '\0'; ++i) s.push_back(option[i]);
  return s;
}

bool ranged(std::initializer_list<int> idx, int imax) {   // src/global.f90:68-78
  for (int i : idx)
    if (!(i > 0 && i <= imax)) return false;
  return true;
}

Model* as_model(void* handle) {
  if (handle == nullptr) return nullptr;
  Model* m = static_cast<Model*>(handle);
  if (m->magic != MODEL_MAGIC) return nullptr;
  return m;
}
void* deliver(const Model& m) { return new Model(m); }   // src/modelClass.f90:51-60 (never freed)

// src/EmDeeData.f90:268-351
void allocate_rigid_bodies(System& me, const int* bodies) {
  const int N = me.natoms;
  me.atomBody.assign(N, 0);
  if (bodies != nullptr) {
    // clean_body_indices (310-349): bodies with a single atom or id <= 0 become free atoms;
    // the others are renumbered 1..nbodies in order of first appearance.
    // (the reference's linear search through `saved` is replaced by a hash map: same result, O(N) instead of
    // O(N x bodies), so that million-atom water boxes set up in seconds when this port is the CPU baseline)
    std::vector<int> index(N, 0), saved, amount, first;
    std::unordered_map<int, int> slotOf;
    slotOf.reserve((size_t)N);
    for (int i = 0; i < N; ++i) {
      int ibody = bodies[i];
      if (ibody > 0) {
        auto it = slotOf.find(ibody);
        int j = it == slotOf.end() ? -1 : it->second;
        if (j < 0) {
          slotOf.emplace(ibody, (int)saved.size());
          saved.push_back(ibody);
          amount.push_back(1);
          first.push_back(i);
          index[i] = 0;
        } else {
          amount[j] += 1;
          index[i] = j + 1;
          index[first[j]] = j + 1;
        }
      }
    }
    int nb_ = 0;
    std::vector<int> renum(saved.size(), 0);
    for (size_t j = 0; j < saved.size(); ++j)
      if (amount[j] > 1) renum[j] = ++nb_;
    for (int i = 0; i < N; ++i)
      if (index[i] > 0) index[i] = renum[index[i] - 1];
    me.atomBody = index;
    me.nbodies = nb_;
    me.free_.clear();
    for (int i = 0; i < N; ++i)
      if (me.atomBody[i] == 0) me.free_.push_back(i);
    me.nfree = (int)me.free_.size();
    me.ph->body.assign(me.nbodies, Body());
    for (int i = 0; i < N; ++i) {
      int b = me.atomBody[i];
      if (b > 0) {
        me.ph->body[b - 1].index.push_back(i);
        me.ph->body[b - 1].M.push_back(me.mass[i]);
      }
    }
    for (auto& b : me.ph->body) {   // tBody_setup, src/ArBee.f90:78-93
      b.NP = (int)b.index.size();
      b.mass = 0.0;
      for (double mm : b.M) b.mass += mm;
      b.invMass = 1.0 / b.mass;
      b.delta.assign(3 * b.NP, 0.0);
      b.d.assign(3 * b.NP, 0.0);
    }
    int k = me.nbodies;
    for (int j = 0; j < N; ++j)
      if (me.atomBody[j] == 0) me.atomBody[j] = ++k;
  } else {
    me.nbodies = 0;
    me.free_.resize(N);
    for (int i = 0; i < N; ++i) {
      me.free_[i] = i;
      me.atomBody[i] = i + 1;
    }
    me.nfree = N;
    me.ph->body.clear();
  }
  me.threadFreeAtoms = (me.nfree + me.nthreads - 1) / me.nthreads;
  me.threadBodies = (me.nbodies + me.nthreads - 1) / me.nthreads;
}

// src/EmDeeData.f90:420-439: make each body whole with respect to its first atom, then tBody_update
void update_rigid_bodies(System& me) {
  const double L = me.ph->Lbox, invL = 1.0 / L;
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
  for (int j = 0; j < me.nbodies; ++j) {
    Body& b = me.ph->body[j];
    std::vector<double> R(3 * b.NP);
    for (int i = 0; i < b.NP; ++i)
      for (int x = 0; x < 3; ++x) R[3 * i + x] = me.ph->R[3 * (size_t)b.index[i] + x];
    for (int i = 1; i < b.NP; ++i)
      for (int x = 0; x < 3; ++x) R[3 * i + x] = R[3 * i + x] - L * std::round(invL * (R[3 * i + x] - R[x]));
    b.update(R);
    for (int i = 0; i < b.NP; ++i)
      for (int x = 0; x < 3; ++x) me.ph->R[3 * (size_t)b.index[i] + x] = R[3 * i + x];
  }
}

// src/EmDeeData.f90:823-860
void move(System& me, double R_factor, double P_factor, double dt, bool translate, bool rotate, int mode) {
  if (translate) {
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int i = 0; i < me.nbodies; ++i) {
      Body& b = me.ph->body[i];
      for (int x = 0; x < 3; ++x) b.rcm[x] = R_factor * b.rcm[x] + P_factor * b.invMass * b.pcm[x];
    }
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int f = 0; f < me.nfree; ++f) {
      const int j = me.free_[f];
      for (int x = 0; x < 3; ++x)
        me.ph->R[3 * (size_t)j + x] = R_factor * me.ph->R[3 * (size_t)j + x] + P_factor * me.ph->P[3 * (size_t)j + x] * me.invMass[j];
    }
  }
  if (rotate) {
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int i = 0; i < me.nbodies; ++i) {
      Body& b = me.ph->body[i];
      if (mode == 0) b.rotate_exact(dt);
      else b.rotate_no_squish(dt, mode);
      for (int k = 0; k < b.NP; ++k)
        for (int x = 0; x < 3; ++x) me.ph->R[3 * (size_t)b.index[k] + x] = b.rcm[x] + b.delta[3 * k + x];
    }
  }
}

// src/EmDeeData.f90:864-896
void boost(System& me, double P_factor, double F_factor, const double* F, bool translate, bool rotate) {
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
  for (int i = 0; i < me.nbodies; ++i) me.ph->body[i].force_and_torque(F);
  if (translate) {
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int i = 0; i < me.nbodies; ++i) {
      Body& b = me.ph->body[i];
      for (int x = 0; x < 3; ++x) b.pcm[x] = P_factor * b.pcm[x] + F_factor * b.F[x];
    }
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int f = 0; f < me.nfree; ++f) {
      const int j = me.free_[f];
      for (int x = 0; x < 3; ++x) me.ph->P[3 * (size_t)j + x] = P_factor * me.ph->P[3 * (size_t)j + x] + F_factor * F[3 * (size_t)j + x];
    }
  }
  if (rotate) {
    const double Ctau = 2.0 * F_factor;
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int i = 0; i < me.nbodies; ++i) {
      Body& b = me.ph->body[i];
      const double t3[3] = {Ctau * b.tau[0], Ctau * b.tau[1], Ctau * b.tau[2]};
      double t4[4], np[4];
      mulC(b.q, t3, t4);
      for (int x = 0; x < 4; ++x) np[x] = P_factor * b.pi[x] + t4[x];
      b.assign_pi(np);
    }
  }
}

// src/EmDeeData.f90:900-922; thread partials (blocks of threadBodies / threadFreeAtoms) summed in thread order
void kinetic_energies(const System& me, bool translate, bool rotate, double twoKEt[3], double twoKEr[3]) {
  const int T = me.nthreads;
  if (translate) {
    for (int x = 0; x < 3; ++x) twoKEt[x] = 0.0;
    for (int t = 1; t <= T; ++t) {
      double k[3] = {0, 0, 0};
      for (int i = (t - 1) * me.threadBodies; i < std::min(t * me.threadBodies, me.nbodies); ++i)
        for (int x = 0; x < 3; ++x) k[x] = k[x] + me.ph->body[i].invMass * me.ph->body[i].pcm[x] * me.ph->body[i].pcm[x];
      for (int f = (t - 1) * me.threadFreeAtoms; f < std::min(t * me.threadFreeAtoms, me.nfree); ++f) {
        const int j = me.free_[f];
        for (int x = 0; x < 3; ++x) k[x] = k[x] + me.invMass[j] * me.ph->P[3 * (size_t)j + x] * me.ph->P[3 * (size_t)j + x];
      }
      for (int x = 0; x < 3; ++x) twoKEt[x] += k[x];
    }
  }
  if (rotate) {
    for (int x = 0; x < 3; ++x) twoKEr[x] = 0.0;
    for (int t = 1; t <= T; ++t) {
      double k[3] = {0, 0, 0};
      for (int i = (t - 1) * me.threadBodies; i < std::min(t * me.threadBodies, me.nbodies); ++i)
        for (int x = 0; x < 3; ++x) k[x] = k[x] + me.ph->body[i].MoI[x] * me.ph->body[i].omega[x] * me.ph->body[i].omega[x];
      for (int x = 0; x < 3; ++x) twoKEr[x] += k[x];
    }
  }
}

// src/EmDeeData.f90:157-189
void assign_momenta(System& me, const double* P, double twoKEt[3], double twoKEr[3]) {
  for (int x = 0; x < 3; ++x) twoKEt[x] = twoKEr[x] = 0.0;
  for (int f = 0; f < me.nfree; ++f) {
    const int i = me.free_[f];
    for (int x = 0; x < 3; ++x) {
      me.ph->P[3 * (size_t)i + x] = P[3 * (size_t)i + x];
      twoKEt[x] += me.invMass[i] * P[3 * (size_t)i + x] * P[3 * (size_t)i + x];
    }
  }
  for (int i = 0; i < me.nbodies; ++i) {
    Body& b = me.ph->body[i];
    double L[3] = {0, 0, 0};
    for (int x = 0; x < 3; ++x) b.pcm[x] = 0.0;
    for (int j = 0; j < b.NP; ++j) {
      const double* Pj = P + 3 * (size_t)b.index[j];
      double c[3];
      cross3(&b.delta[3 * j], Pj, c);
      for (int x = 0; x < 3; ++x) { b.pcm[x] = b.pcm[x] + Pj[x]; L[x] = L[x] + c[x]; }
    }
    for (int x = 0; x < 3; ++x) twoKEt[x] += b.invMass * b.pcm[x] * b.pcm[x];
    const double twoL[3] = {2.0 * L[0], 2.0 * L[1], 2.0 * L[2]};
    double np[4];
    mulC(b.q, twoL, np);
    b.assign_pi(np);
    for (int x = 0; x < 3; ++x) twoKEr[x] += b.MoI[x] * b.omega[x] * b.omega[x];
  }
}

void set_kinetic_from_sums(tEmDee* md, const double twoKEt[3], const double twoKEr[3]) {   // src/EmDeeCode.f90:862-868, 984-990
  for (int x = 0; x < 3; ++x) { md->Kinetic.RotPart[x] = 0.5 * twoKEr[x]; md->Kinetic.TransPart[x] = 0.5 * twoKEt[x]; }
  md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
  md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
  md->Kinetic.ShadowKinetic = md->Kinetic.Total;
  md->Kinetic.ShadowRotational = md->Kinetic.Rotational;
  md->Kinetic.UpToDate = true;
}

// src/EmDeeData.f90:193-223
void check_actual_interactions(System& me) {
  const int nt = me.ntypes;
  std::vector<char> inter((size_t)nt * nt * me.nlayers, 0);
  std::vector<char> neutral(nt, 1);
  for (int i = 1; i <= nt; ++i) {
    int cnt = 0;
    for (int a = 0; a < me.natoms; ++a)
      if (me.atomType[a] == i && me.charged[a]) ++cnt;
    neutral[i - 1] = (cnt == 0);
    for (int j = 1; j <= i; ++j) {
      bool any = false;
      for (int k = 1; k <= me.nlayers; ++k) {
        PairContainer& p = me.pr(i, j, k);
        bool no_pair = p.model.kind == PAIR_NONE;
        bool no_coul = me.coul[k - 1].kind == COUL_NONE || !p.coulomb;
        bool coul_only = no_pair && !no_coul;
        bool inert = (no_pair && no_coul) || (coul_only && neutral[i - 1] && neutral[j - 1]);
        inter[(size_t)(k - 1) * nt * nt + (size_t)(j - 1) * nt + (i - 1)] = !inert;
        inter[(size_t)(k - 1) * nt * nt + (size_t)(i - 1) * nt + (j - 1)] = !inert;
        any = any || !inert;
      }
      me.tt(me.interact, i, j) = any;
      me.tt(me.interact, j, i) = any;
    }
  }
  for (int k = 1; k <= me.nlayers; ++k) {
    bool any = false;
    // the reference's local `interact(i,j,k)` is only filled for j <= i; `any` over it is the same
    for (size_t q = 0; q < (size_t)nt * nt; ++q) any = any || inter[(size_t)(k - 1) * nt * nt + q];
    me.pairs_exist[k - 1] = any;
  }
}

// src/EmDeeData.f90:227-264
void set_pair_type(System& me, int itype, int jtype, int layer, const Model& model, double kCoul) {
  const double cutoff = me.layerRc(layer);
  if (!is_pair(model.kind)) error("pair model setup", "a valid pair model must be provided");
  if (itype == jtype) {
    PairContainer& ii = me.pr(itype, itype, layer);
    ii.model = model;   // pairContainer = modelContainer copies the model only (modelClass_pair.f90:101-116)
    ii.coulomb = kCoul != 0.0;
    if (ii.coulomb) ii.kCoul = kCoul;
    modifier_setup(ii.model, cutoff);
    for (int ktype = 1; ktype <= me.ntypes; ++ktype) {
      if (ktype != itype && me.tt(me.overridable, itype, ktype)) {
        PairContainer mixed = container_mix(me.pr(ktype, ktype, layer), me.pr(itype, itype, layer));
        modifier_setup(mixed.model, cutoff);
        me.pr(itype, ktype, layer) = mixed;
        me.pr(ktype, itype, layer) = mixed;
      }
    }
  } else {
    PairContainer& ij = me.pr(itype, jtype, layer);
    ij.model = model;
    ij.coulomb = kCoul != 0.0;
    if (ij.coulomb) ij.kCoul = kCoul;
    modifier_setup(ij.model, cutoff);
    me.pr(jtype, itype, layer) = ij;
  }
}

// kspace_ewald_update (src/kspace_ewald.f90:103-184): the half-space of wave vectors inside the cutoff sphere,
// their Gaussian prefactors and force vectors. Called once, at initialization (the reference never calls it again).
void ewald_update(System& me, double L) {
  Ewald& k = me.ewald;
  const double unit = 2.0 * Pi / L, unitSq = unit * unit;
  const int nm = (int)std::ceil(k.kmax / unit);
  for (int x = 0; x < 3; ++x) k.nmax[x] = nm;
  const int M = 2 * nm + 1, maxnvecs = (M * M * M) / 2, M2 = M, M2M3 = M * M;
  const double kmaxSq = unitSq * nm * nm;
  k.n.clear();
  for (int i = maxnvecs + 1; i <= 2 * maxnvecs; ++i) {
    const int n1 = i / M2M3, j = i - n1 * M2M3, n2 = j / M2, n3 = j - n2 * M2;
    const int kv[3] = {n1 - nm, n2 - nm, n3 - nm};
    if (unitSq * kv[0] * kv[0] + unitSq * kv[1] * kv[1] + unitSq * kv[2] * kv[2] <= kmaxSq) k.n.insert(k.n.end(), kv, kv + 3);
  }
  k.nvecs = (int)k.n.size() / 3;
  const double B = -0.25 / (k.alpha * k.alpha);
  k.Volume = L * L * L;
  const double fourPiByV = 4.0 * Pi / k.Volume;
  k.prefac.assign(k.nvecs, 0.0);
  k.vec.assign(3 * (size_t)k.nvecs, 0.0);
  for (int i = 0; i < k.nvecs; ++i) {
    const double kk[3] = {unit * k.n[3 * i], unit * k.n[3 * i + 1], unit * k.n[3 * i + 2]};
    const double ksq = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
    k.prefac[i] = fourPiByV * std::exp(ksq * B) / ksq;
    for (int x = 0; x < 3; ++x) k.vec[3 * (size_t)i + x] = 2.0 * k.prefac[i] * kk[x];
  }
}

// cKspaceModel_initialize (src/modelClass_kspace.f90:108-273) + kspace_ewald_set_parameters (src/kspace_ewald.f90:82-99)
void ewald_initialize(System& me, double Rc) {
  const char* task = "kspace model initialization";
  Ewald& k = me.ewald;
  const int N = me.natoms;
  const double L = me.ph->Lbox;
  int ncharged = 0;
  for (int i = 0; i < N; ++i) ncharged += me.charged[i] ? 1 : 0;
  if (ncharged == 0) error(task, "system has no charged atoms");
  // distinct types of charged atoms, ascending
  std::vector<char> seen(me.ntypes + 1, 0);
  for (int i = 0; i < N; ++i)
    if (me.charged[i]) seen[me.atomType[i]] = 1;
  k.type.clear();
  for (int t = 1; t <= me.ntypes; ++t)
    if (seen[t]) k.type.push_back(t);
  k.ntypes = (int)k.type.size();
  // every atom of those types, grouped by type (143-161: pack(seq, types == type(i)), charged or not)
  k.first.assign(k.ntypes, 0);
  k.last.assign(k.ntypes, 0);
  k.item.clear();
  k.value.clear();
  for (int t = 0; t < k.ntypes; ++t) {
    k.first[t] = (int)k.item.size();
    for (int i = 0; i < N; ++i)
      if (me.atomType[i] == k.type[t]) { k.item.push_back(i); k.value.push_back(me.charge[i]); }
    k.last[t] = (int)k.item.size() - 1;
  }
  // accuracy = exp(-s^2)/s^2, alpha = s/Rc, kmax = 2*alpha*s
  const double sacc = std::sqrt(inverse_of_x_plus_ln_x(-std::log(me.kspace.accuracy)));
  k.alpha = sacc / Rc;
  k.kmax = 2.0 * k.alpha * sacc;
  std::printf("KSPACE PARAMETERS: alpha = %10.5f and kmax = %10.5f\n", k.alpha, k.kmax);
  k.beta = 2.0 * k.alpha / std::sqrt(Pi);
  // intrabody pairs of charged atoms: their smooth part is taken back out, as is the self energy (169-236)
  k.nTypePairs = k.ntypes * (k.ntypes - 1) / 2 + k.ntypes;
  k.Erigid.assign(k.nTypePairs, 0.0);
  std::vector<int> local(me.ntypes + 1, 0);
  for (int t = 0; t < k.ntypes; ++t) local[k.type[t]] = t + 1;
  const double invL = 1.0 / L;
  for (const Body& b : me.ph->body)
    for (int ii = 0; ii < b.NP - 1; ++ii) {
      const int i = b.index[ii];
      if (!me.charged[i]) continue;
      for (int jj = ii + 1; jj < b.NP; ++jj) {
        const int j = b.index[jj];
        if (!me.charged[j]) continue;
        double rsq = 0.0;
        for (int x = 0; x < 3; ++x) {
          double d = me.ph->R[3 * (size_t)i + x] - me.ph->R[3 * (size_t)j + x];
          d = d - L * std::round(invL * d);
          rsq += d * d;
        }
        double Eij, Wij;
        k.discount(Eij, Wij, rsq, me.charge[i] * me.charge[j]);
        k.Erigid[Ewald::symm1D(local[me.atomType[i]], local[me.atomType[j]]) - 1] += Eij;
      }
    }
  for (int t = 0; t < k.ntypes; ++t) {
    double q2 = 0.0;
    for (int q = k.first[t]; q <= k.last[t]; ++q) q2 += k.value[q] * k.value[q];
    k.Erigid[Ewald::symm1D(t + 1, t + 1) - 1] -= k.alpha * q2 / std::sqrt(Pi);
  }
  // Coulomb constants of the type pairs, per layer (238-252)
  k.nlayers = me.nlayers;
  k.lambda.assign((size_t)k.ntypes * k.ntypes * k.nlayers, 0.0);
  k.lambda1D.assign((size_t)k.nTypePairs * k.nlayers, 0.0);
  for (int l = 1; l <= me.nlayers; ++l)
    for (int i = 0; i < k.ntypes; ++i)
      for (int j = i; j < k.ntypes; ++j) {
        const double lam = me.pr(k.type[i], k.type[j], l).kCoul;
        k.lambda[((size_t)(l - 1) * k.ntypes + j) * k.ntypes + i] = lam;
        k.lambda[((size_t)(l - 1) * k.ntypes + i) * k.ntypes + j] = lam;
        k.lambda1D[(size_t)(l - 1) * k.nTypePairs + Ewald::symm1D(i + 1, j + 1) - 1] = lam;
      }
  ewald_update(me, L);
}

// compute_kspace (src/EmDeeData.f90:689-700): kspace_ewald_prepare + kspace_ewald_compute (src/kspace_ewald.f90:188-314)
// + the energy part of discount_rigid_pairs (src/modelClass_kspace.f90:277-316; forces are not discounted there either)
void compute_kspace(System& me, const double* Rs, double& Elong, double* F) {
  const Ewald& k = me.ewald;
  const int nv = k.nvecs, nt = k.ntypes, nm = k.nmax[0], nch = (int)k.item.size();
  const int layer = me.layer;
  // q_j exp(i k.r_j) for every listed atom: powers of exp(2 pi i s) along each axis (recursion, 208-219, 241-255)
  std::vector<double> qre((size_t)nv * nch), qim((size_t)nv * nch);
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
  for (int j = 0; j < nch; ++j) {
    const int i = k.item[j];
    std::vector<double> gre(3 * (size_t)(2 * nm + 1)), gim(3 * (size_t)(2 * nm + 1));
    for (int x = 0; x < 3; ++x) {
      double* re = gre.data() + (size_t)x * (2 * nm + 1) + nm;   // index -nm..nm
      double* im = gim.data() + (size_t)x * (2 * nm + 1) + nm;
      const double theta = 2.0 * Pi * Rs[3 * (size_t)i + x];
      const double zr = std::cos(theta), zi = std::sin(theta);
      re[0] = 1.0; im[0] = 0.0;
      double br = zr, bi = zi;
      for (int p = 1; p <= nm; ++p) {
        re[p] = br; im[p] = bi;
        const double tr = br * zr - bi * zi, ti = br * zi + bi * zr;
        br = tr; bi = ti;
      }
      for (int p = 1; p <= nm; ++p) { re[-p] = re[p]; im[-p] = -im[p]; }
    }
    const int W1 = 2 * nm + 1;
    for (int v = 0; v < nv; ++v) {
      const int a = k.n[3 * v] + nm, b = k.n[3 * v + 1] + nm, c = k.n[3 * v + 2] + nm;
      const double ar = gre[a], ai = gim[a], br = gre[W1 + b], bi = gim[W1 + b], cr = gre[2 * W1 + c], ci = gim[2 * W1 + c];
      const double abr = ar * br - ai * bi, abi = ar * bi + ai * br;
      qre[(size_t)j * nv + v] = k.value[j] * (abr * cr - abi * ci);
      qim[(size_t)j * nv + v] = k.value[j] * (abr * ci + abi * cr);
    }
  }
  // type-specific structure factors S(k, t) and sigma = S . lambda(layer)
  std::vector<double> Sre((size_t)nv * nt, 0.0), Sim((size_t)nv * nt, 0.0), gre_((size_t)nv * nt, 0.0), gim_((size_t)nv * nt, 0.0);
  for (int t = 0; t < nt; ++t)
    for (int j = k.first[t]; j <= k.last[t]; ++j)
      for (int v = 0; v < nv; ++v) {
        Sre[(size_t)t * nv + v] += qre[(size_t)j * nv + v];
        Sim[(size_t)t * nv + v] += qim[(size_t)j * nv + v];
      }
  const double* lam = k.lambda.data() + (size_t)(layer - 1) * nt * nt;
  for (int t = 0; t < nt; ++t)
    for (int u = 0; u < nt; ++u)
      for (int v = 0; v < nv; ++v) {
        gre_[(size_t)t * nv + v] += Sre[(size_t)u * nv + v] * lam[(size_t)t * nt + u];
        gim_[(size_t)t * nv + v] += Sim[(size_t)u * nv + v] * lam[(size_t)t * nt + u];
      }
  double E = 0.0;
  for (int v = 0; v < nv; ++v) {
    double d = 0.0;
    for (int t = 0; t < nt; ++t) d += Sre[(size_t)t * nv + v] * gre_[(size_t)t * nv + v] + Sim[(size_t)t * nv + v] * gim_[(size_t)t * nv + v];
    E += k.prefac[v] * d;
  }
  Elong = Elong + E;
  // forces: F_i += sum_k vec(k) * (Re sigma Im q - Re q Im sigma)
  for (int t = 0; t < nt; ++t)
#pragma omp parallel for num_threads(me.nthreads) schedule(static)
    for (int j = k.first[t]; j <= k.last[t]; ++j) {
      double f[3] = {0, 0, 0};
      for (int v = 0; v < nv; ++v) {
        const double c = gre_[(size_t)t * nv + v] * qim[(size_t)j * nv + v] - qre[(size_t)j * nv + v] * gim_[(size_t)t * nv + v];
        for (int x = 0; x < 3; ++x) f[x] += k.vec[3 * (size_t)v + x] * c;
      }
      for (int x = 0; x < 3; ++x) F[3 * (size_t)k.item[j] + x] += f[x];
    }
  // rigid pairs + self energy
  double Er = 0.0;
  for (int p = 0; p < k.nTypePairs; ++p) Er += k.lambda1D[(size_t)(layer - 1) * k.nTypePairs + p] * k.Erigid[p];
  Elong = Elong + Er;
}

// src/EmDeeData.f90:359-416
void perform_initialization(System& me, int& DoF, int& RotDoF) {
  const char* task = "system initialization";
  update_rigid_bodies(me);
  int bodyDoF = 0;
  for (auto& b : me.ph->body) bodyDoF += b.dof;
  RotDoF = bodyDoF - 3 * me.nbodies;
  DoF = 3 * me.nfree + bodyDoF - 3;
  check_actual_interactions(me);
  bool any_required = false;
  double kspaceRc = 0.0;
  bool first = true;
  for (int l = 1; l <= me.nlayers; ++l) {
    if (me.coul[l - 1].requires_kspace) {
      any_required = true;
      if (first) { kspaceRc = me.layerRc(l); first = false; }
      else if (me.layerRc(l) != kspaceRc)
        error(task, "all layers with ewald-like coulomb models must have the same cutoff");
    }
  }
  if (any_required != me.kspace_active) {
    if (me.kspace_active) me.kspace_active = false;
    else error(task, "a kspace solver is required, but has not been defined");
  }
  if (me.kspace_active) {
    ewald_initialize(me, kspaceRc);
    for (int l = 1; l <= me.nlayers; ++l) {
      Model& c = me.coul[l - 1];
      if (c.requires_kspace) {   // coul_long_kspace_setup, src/coul_long.f90:66-73
        c.alpha = me.ewald.alpha;
        c.beta = 2.0 * me.ewald.alpha / std::sqrt(Pi);
      }
    }
  }
  me.initialized = true;
}

// src/neighbor_lists.f90:41-59
double maximum_approach_sq(int N, const double* R, const double* R0) {
  auto dsq = [&](int i) {
    double a = R[3 * (size_t)i] - R0[3 * (size_t)i], b = R[3 * (size_t)i + 1] - R0[3 * (size_t)i + 1],
           c = R[3 * (size_t)i + 2] - R0[3 * (size_t)i + 2];
    return a * a + b * b + c * c;
  };
  double maximum = dsq(0);
  double next = maximum;
  for (int i = 1; i < N; ++i) {
    double deltaSq = dsq(i);
    if (deltaSq > maximum) {
      next = maximum;
      maximum = deltaSq;
    }
  }
  return maximum + 2 * std::sqrt(maximum * next) + next;
}

inline int ipbc(int x, int M) { return x < 0 ? x + M : (x >= M ? x - M : x); }

// src/neighbor_lists.f90:63-171
void distribute_atoms(System& me, int M, const double* Rs) {
  const int MM = M * M;
  const int T = me.nthreads;
  bool make_cells = M != me.mcells;
  int cells_per_thread = 0;
  if (make_cells) {
    me.mcells = M;
    me.ncells = M * MM;
    if (me.ncells > me.maxcells) {
      me.cell.assign(me.ncells, Cell());
      me.cellAtom.first.assign(me.ncells, 1);
      me.cellAtom.last.assign(me.ncells, 0);
      me.atomsInCell.assign(me.ncells, 0);
      me.threadCell.allocate(0, T);
      me.maxcells = me.ncells;
    }
    cells_per_thread = (me.ncells + T - 1) / T;
  }
  std::vector<int> maxNatoms(T, 0), threadNatoms(T, 0);
  std::vector<int> next(me.natoms, 0);
#pragma omp parallel num_threads(T)
  {
    const int thread = omp_get_thread_num() + 1;
    int first, last;
    if (make_cells) {
      first = (thread - 1) * cells_per_thread + 1;
      last = std::min(thread * cells_per_thread, me.ncells);
      for (int icell = first; icell <= last; ++icell) {
        int k = icell - 1;
        int iz = k / MM;
        int j = k - iz * MM;
        int iy = j / M;
        int ix = j - iy * M;
        for (int q = 0; q < nbcells; ++q)
          me.cell[icell - 1].neighbor[q] =
              1 + ipbc(ix + nb[q][0], M) + ipbc(iy + nb[q][1], M) * M + ipbc(iz + nb[q][2], M) * MM;
      }
      me.threadCell.first[thread - 1] = first;
      me.threadCell.last[thread - 1] = last;
    } else {
      first = me.threadCell.first[thread - 1];
      last = me.threadCell.last[thread - 1];
    }
    const int a1 = (thread - 1) * me.threadAtoms + 1, aN = std::min(thread * me.threadAtoms, me.natoms);
    for (int i = a1; i <= aN; ++i) {
      int ic[3];
      for (int x = 0; x < 3; ++x) {
        double r = Rs[3 * (size_t)(i - 1) + x];
        ic[x] = (int)(M * (r - std::floor(r)));
        if (ic[x] >= M) ic[x] = M - 1;   // Q5: result-neutral clamp (the reference would index out of range)
      }
      me.atomCell[i - 1] = 1 + ic[0] + M * ic[1] + MM * ic[2];
    }
#pragma omp barrier
    std::vector<int> head(std::max(last - first + 1, 0), 0);
    for (int c = first; c <= last; ++c) me.atomsInCell[c - 1] = 0;
    for (int i = 1; i <= me.natoms; ++i) {
      int icell = me.atomCell[i - 1];
      if (icell >= first && icell <= last) {
        next[i - 1] = head[icell - first];
        head[icell - first] = i;
        me.atomsInCell[icell - 1] += 1;
      }
    }
    int s = 0;
    for (int c = first; c <= last; ++c) s += me.atomsInCell[c - 1];
    threadNatoms[thread - 1] = s;
#pragma omp barrier
    int mx = 0;
    int k = 0;
    for (int t = 0; t < thread - 1; ++t) k += threadNatoms[t];
    for (int icell = first; icell <= last; ++icell) {
      me.cellAtom.first[icell - 1] = k + 1;
      int i = head[icell - first];
      while (i != 0) {
        k += 1;
        me.cellAtom.item[k - 1] = i;
        i = next[i - 1];
      }
      me.cellAtom.last[icell - 1] = k;
      if (me.atomsInCell[icell - 1] > mx) mx = me.atomsInCell[icell - 1];
    }
    maxNatoms[thread - 1] = mx;
  }
  me.maxatoms = *std::max_element(maxNatoms.begin(), maxNatoms.end());
  me.maxpairs = (me.maxatoms * ((2 * nbcells + 1) * me.maxatoms - 1)) / 2;
}

inline double pbc(double x) { return x - std::round(x); }   // x - anint(x)

// src/neighbor_lists.f90:199-300
void build_neighbor_lists(System& me, int thread, const double* Rs) {
  const double invL2 = 1.0 / (me.ph->Lbox * me.ph->Lbox);
  const double xRc2 = me.xRcSq * invL2;
  const double xInRc2 = me.xInRcSq * invL2;
  std::vector<char> include(me.natoms, 1);
  std::vector<int> atom((size_t)(nbcells + 1) * std::max(me.maxatoms, 1));
  std::vector<double> Ratom;
  int npairs = 0;
  List& neighbor = me.neighbor[thread - 1];
  const std::vector<int>& c1 = me.cellAtom.first;
  const std::vector<int>& cN = me.cellAtom.last;
  const int nt = me.ntypes;
  for (int icell = me.threadCell.first[thread - 1]; icell <= me.threadCell.last[thread - 1]; ++icell) {
    const int* neigh = me.cell[icell - 1].neighbor;
    int nlocal = me.atomsInCell[icell - 1];
    int ntotal = 0;
    for (int k = c1[icell - 1]; k <= cN[icell - 1]; ++k) atom[ntotal++] = me.cellAtom.item[k - 1];
    for (int q = 0; q < nbcells; ++q)
      for (int k = c1[neigh[q] - 1]; k <= cN[neigh[q] - 1]; ++k) atom[ntotal++] = me.cellAtom.item[k - 1];
    if (neighbor.nitems < npairs + nlocal * ntotal) neighbor.resize(npairs + nlocal * ntotal + extra);
    Ratom.resize(3 * (size_t)ntotal);
    for (int q = 0; q < ntotal; ++q)
      for (int x = 0; x < 3; ++x) Ratom[3 * (size_t)q + x] = Rs[3 * (size_t)(atom[q] - 1) + x];
    for (int k = 0; k < nlocal; ++k) {
      int i = atom[k];
      int first = npairs + 1;
      neighbor.first[i - 1] = first;
      int* item = neighbor.item.data() + (first - 1);
      double* value = neighbor.value.data() + (first - 1);
      int ipairs = 0, middle = 0;
      int itype = me.atomType[i - 1];
      int ibody = me.atomBody[i - 1];
      const int x1 = me.excluded.first[i - 1], xN = me.excluded.last[i - 1];
      for (int q = x1; q <= xN; ++q) include[me.excluded.item[q - 1] - 1] = 0;
      for (int m = k + 1; m < ntotal; ++m) {
        double dx = pbc(Ratom[3 * (size_t)k] - Ratom[3 * (size_t)m]);
        double dy = pbc(Ratom[3 * (size_t)k + 1] - Ratom[3 * (size_t)m + 1]);
        double dz = pbc(Ratom[3 * (size_t)k + 2] - Ratom[3 * (size_t)m + 2]);
        double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < xRc2) {
          int j = atom[m];
          if (include[j - 1]) {
            if (me.atomBody[j - 1] != ibody) {
              if (me.interact[(size_t)(itype - 1) * nt + (me.atomType[j - 1] - 1)]) {
                // insert_neighbor (283-298): keep the atom's neighbors sorted by r2 ascending
                int q;
                for (q = ipairs; q >= 1; --q) {
                  if (value[q - 1] < r2) break;
                  value[q] = value[q - 1];
                  item[q] = item[q - 1];
                }
                value[q] = r2;
                item[q] = j;
                ipairs += 1;
                if (r2 < xInRc2) middle += 1;
              }
            }
          }
        }
      }
      for (int q = x1; q <= xN; ++q) include[me.excluded.item[q - 1] - 1] = 1;
      neighbor.middle[i - 1] = npairs + middle;
      npairs += ipairs;
      neighbor.last[i - 1] = npairs;
    }
  }
  neighbor.count = npairs;
}

// src/neighbor_lists.f90:175-195
void handle_neighbor_lists(System& me, int& builds, double& time, const double* Rs) {
  time -= omp_get_wtime();
  if (maximum_approach_sq(me.natoms, me.ph->R.data(), me.R0.data()) > me.skinSq) {
    int M = (int)std::floor(ndiv * me.ph->Lbox / me.xRc);
    distribute_atoms(me, std::max(M, 2 * ndiv + 1), Rs);
    me.R0 = me.ph->R;
    builds += 1;
#pragma omp parallel num_threads(me.nthreads)
    build_neighbor_lists(me, omp_get_thread_num() + 1, Rs);
  }
  time += omp_get_wtime();
}

// src/EmDeeData.f90:644-685 wrapping src/compute.f90:20-100
template <bool COMPUTE>
void compute_pairs(System& me, int thread, const double* Rs, double* F, double& Epair, double& Ecoul,
                   double& Wpair, double& Wcoul) {
  const int N = me.natoms;
  std::fill(F, F + 3 * (size_t)N, 0.0);
  Wpair = 0.0;
  Wcoul = 0.0;
  Epair = 0.0;
  // NOTE: Ecoul is intent(out) in the reference but only ever accumulated; the caller zeroes E(:).
  if (!me.pairs_exist[me.layer - 1]) return;
  const double L2 = me.ph->Lbox * me.ph->Lbox;
  const double invL2 = 1.0 / L2;
  double Rc2;
  const std::vector<int>* upper;
  List& neighbor = me.neighbor[thread - 1];
  if (me.useInRc[me.layer - 1]) {
    Rc2 = me.InRcSq * invL2;
    upper = &neighbor.middle;
  } else {
    Rc2 = me.RcSq * invL2;
    upper = &neighbor.last;
  }
  const int tfirst = me.threadCell.first[thread - 1], tlast = me.threadCell.last[thread - 1];
  if (tlast < tfirst) return;
  const int firstAtom = me.cellAtom.first[tfirst - 1];
  const int lastAtom = me.cellAtom.last[tlast - 1];
  const Model& coul = me.coul[me.layer - 1];
  for (int k = firstAtom; k <= lastAtom; ++k) {
    const int i = me.cellAtom.item[k - 1];
    const int itype = me.atomType[i - 1];
    const double Qi = me.charge[i - 1];
    const bool icharged = me.charged[i - 1];
    const double Ri[3] = {Rs[3 * (size_t)(i - 1)], Rs[3 * (size_t)(i - 1) + 1], Rs[3 * (size_t)(i - 1) + 2]};
    double Fi[3] = {0.0, 0.0, 0.0};
    for (int m = neighbor.first[i - 1]; m <= (*upper)[i - 1]; ++m) {
      const int j = neighbor.item[m - 1];
      const double Rij[3] = {pbc(Ri[0] - Rs[3 * (size_t)(j - 1)]), pbc(Ri[1] - Rs[3 * (size_t)(j - 1) + 1]),
                             pbc(Ri[2] - Rs[3 * (size_t)(j - 1) + 2])};
      const double r2 = Rij[0] * Rij[0] + Rij[1] * Rij[1] + Rij[2] * Rij[2];
      if (r2 < Rc2) {
        const double invR2 = invL2 / r2;
        const double invR = std::sqrt(invR2);
        const int jtype = me.atomType[j - 1];
        const bool ijcharged = icharged && me.charged[j - 1];
        const PairContainer& pair = me.pr(jtype, itype, me.layer);
        double Eij = 0.0, Wij = 0.0;
        model_eval<COMPUTE ? 0 : 2>(pair.model, Eij, Wij, invR, invR2);
        apply_modifier<COMPUTE>(pair.model, Eij, Wij, invR, invR2);
        if (COMPUTE) Epair = Epair + Eij;
        Wpair = Wpair + Wij;
        double Wsum = Wij;
        if (ijcharged && pair.coulomb) {
          model_eval<COMPUTE ? 0 : 2>(coul, Eij, Wij, invR, invR2);   // Q4: coul_none leaves Wij in virial mode
          apply_modifier<COMPUTE>(coul, Eij, Wij, invR, invR2);
          const double QiQj = pair.kCoul * Qi * me.charge[j - 1];
          if (COMPUTE) Ecoul = Ecoul + QiQj * Eij;
          Wij = QiQj * Wij;
          Wcoul = Wcoul + Wij;
          Wsum = Wsum + Wij;
        }
        const double s = Wsum * invR2;
        for (int x = 0; x < 3; ++x) {
          const double Fij = s * Rij[x];
          Fi[x] = Fi[x] + Fij;
          F[3 * (size_t)(j - 1) + x] = F[3 * (size_t)(j - 1) + x] - Fij;
        }
      }
    }
    for (int x = 0; x < 3; ++x) F[3 * (size_t)(i - 1) + x] = F[3 * (size_t)(i - 1) + x] + Fi[x];
  }
  for (size_t q = 0; q < 3 * (size_t)N; ++q) F[q] = me.ph->Lbox * F[q];
}

// src/EmDeeData.f90:443-487 with bond_harmonic_compute (src/bond_harmonic.f90:68-80); R holds SCALED coordinates.
// With an Ewald-type Coulomb model the smooth part of each bonded (excluded) pair is taken back out (470-478).
void compute_bonds(System& me, const double* Rs, double* F, double& Potential, double& Virial, double& Ecoul) {
  const bool bonded = me.bonded[me.layer - 1], kspace = me.coul[me.layer - 1].requires_kspace;
  if (me.bonds.empty() || !(bonded || kspace)) return;
  const double L = me.ph->Lbox, invL2 = 1.0 / (L * L);
  for (const System::Struct& b : me.bonds) {
    double Rij[3], r2 = 0.0;
    for (int x = 0; x < 3; ++x) {
      Rij[x] = Rs[3 * (size_t)(b.i - 1) + x] - Rs[3 * (size_t)(b.j - 1) + x];
      Rij[x] = Rij[x] - std::round(Rij[x]);
      r2 += Rij[x] * Rij[x];
    }
    const double invR2 = invL2 / r2;
    double E = 0.0, W = 0.0;
    if (bonded) {
      if (b.model.kind == BOND_HARMONIC) {
        const double r = 1.0 / std::sqrt(invR2), delta = r - b.model.p2;
        E = (0.5 * b.model.p1) * delta * delta;
        W = (-b.model.p1) * delta * r;
      }
      Potential = Potential + E;
      Virial = Virial + W;
    }
    if (kspace) {
      const double QiQj = me.pr(me.atomType[b.i - 1], me.atomType[b.j - 1], me.layer).kCoul * me.charge[b.i - 1] * me.charge[b.j - 1];
      double EL, WL;
      me.ewald.discount(EL, WL, 1.0 / invR2, QiQj);
      Ecoul = Ecoul + EL;
      W = W + WL;
    }
    for (int x = 0; x < 3; ++x) {
      const double Fij = W * invR2 * L * Rij[x];
      F[3 * (size_t)(b.i - 1) + x] += Fij;
      F[3 * (size_t)(b.j - 1) + x] -= Fij;
    }
  }
}

// src/EmDeeData.f90:491-550 with angle_harmonic_compute (src/angle_harmonic.f90:68-78)
void compute_angles(System& me, const double* Rs, double* F, double& Potential, double& Virial, double& Ecoul) {
  const bool bonded = me.bonded[me.layer - 1], kspace = me.coul[me.layer - 1].requires_kspace;
  if (me.angles.empty() || !(bonded || kspace)) return;
  const double L = me.ph->Lbox;
  for (const System::Struct& a : me.angles) {
    const size_t i = a.i - 1, j = a.j - 1, k = a.k - 1;
    double av[3], bv[3], aa = 0.0, bb = 0.0, ab = 0.0;
    for (int x = 0; x < 3; ++x) {
      av[x] = Rs[3 * i + x] - Rs[3 * j + x];
      bv[x] = Rs[3 * k + x] - Rs[3 * j + x];
      av[x] = L * (av[x] - std::round(av[x]));
      bv[x] = L * (bv[x] - std::round(bv[x]));
      aa += av[x] * av[x];
      bb += bv[x] * bv[x];
      ab += av[x] * bv[x];
    }
    if (bonded) {
      const double theta = std::acos(ab / std::sqrt(aa * bb));
      double Ea = 0.0, Fa = 0.0;
      if (a.model.kind == ANGLE_HARMONIC) {
        const double delta = theta - a.model.p2;
        Ea = (0.5 * a.model.p1) * delta * delta;
        Fa = (-a.model.p1) * delta;
      }
      const double factor = Fa / std::sqrt(aa * bb - ab * ab);
      double w = 0.0;
      for (int x = 0; x < 3; ++x) {
        const double Fi = ((ab / aa) * av[x] - bv[x]) * factor, Fk = ((ab / bb) * bv[x] - av[x]) * factor;
        F[3 * i + x] += Fi;
        F[3 * k + x] += Fk;
        F[3 * j + x] -= (Fi + Fk);
        w += Fi * av[x] + Fk * bv[x];
      }
      Potential = Potential + Ea;
      Virial = Virial + w;
    }
    if (kspace) {   // the 1-3 pair of the angle (532-546)
      double Rik[3], RikSq = 0.0;
      for (int x = 0; x < 3; ++x) { Rik[x] = av[x] - bv[x]; RikSq += Rik[x] * Rik[x]; }
      const double QiQk = me.pr(me.atomType[i], me.atomType[k], me.layer).kCoul * me.charge[i] * me.charge[k];
      double EL, WL;
      me.ewald.discount(EL, WL, RikSq, QiQk);
      Ecoul = Ecoul + EL;
      for (int x = 0; x < 3; ++x) {
        const double Fik = WL * Rik[x] / RikSq;
        F[3 * i + x] += Fik;
        F[3 * k + x] -= Fik;
      }
    }
  }
}

// src/EmDeeData.f90:926-953
double rigid_body_virial(System& me) {
  std::vector<double> W(me.nthreads, 0.0);
#pragma omp parallel num_threads(me.nthreads)
  {
    const int thread = omp_get_thread_num() + 1;
    double w = 0.0;
    const double* F = me.F();
    for (int i = (thread - 1) * me.threadBodies + 1; i <= std::min(thread * me.threadBodies, me.nbodies); ++i) {
      const Body& b = me.ph->body[i - 1];
      double s = 0.0;
      for (int a = 0; a < b.NP; ++a)
        for (int x = 0; x < 3; ++x) s += F[3 * (size_t)b.index[a] + x] * b.delta[3 * a + x];
      w = w + s;
    }
    W[thread - 1] = w;
  }
  double tot = 0.0;
  for (double w : W) tot += w;
  return -tot;
}

void invalidate(System& me, tEmDee* md) {
  std::fill(me.forcesUpToDate.begin(), me.forcesUpToDate.end(), 0);
  for (auto& e : me.layerEnergy) e.UpToDate = false;
  md->Energy.UpToDate = false;
}

}  // namespace

// =================================================================================================
//                                     C   A B I
// =================================================================================================
extern "C" {

const char* EmDeeX_backend(void) { return "oracle-cpu"; }

// src/EmDeeCode.f90:69-208
tEmDee EmDee_system(int threads, int layers, double rc, double skin, int N, int* types, double* masses,
                    int* bodies) {
  if (std::getenv("EMDEE_QUIET") == nullptr) std::printf("EmDee (version: %11s)\n", "15 Oct 2018");
  System* me = new System();
  me->nthreads = threads;
  me->nlayers = layers;
  me->Rc = rc;
  me->skin = skin;
  me->RcSq = rc * rc;
  me->xRc = rc + skin;
  me->xRcSq = me->xRc * me->xRc;
  me->InRc = me->Rc;
  me->InRcSq = me->RcSq;
  me->xInRcSq = me->xRcSq;
  me->skinSq = skin * skin;
  me->natoms = N;
  me->threadAtoms = (N + threads - 1) / threads;
  me->kspace_active = false;
  if (types != nullptr) {
    if (*std::min_element(types, types + N) != 1) error("system setup", "wrong specification of atom types");
    me->ntypes = *std::max_element(types, types + N);
    me->atomType.assign(types, types + N);
  } else {
    me->ntypes = 1;
    me->atomType.assign(N, 1);
  }
  if (masses != nullptr) {
    me->mass.resize(N);
    me->invMass.resize(N);
    me->totalMass = 0.0;
    for (int i = 0; i < N; ++i) {
      me->mass[i] = masses[me->atomType[i] - 1];
      me->invMass[i] = 1.0 / masses[me->atomType[i] - 1];
      me->totalMass += masses[me->atomType[i] - 1];
    }
  } else {
    me->mass.assign(N, 1.0);
    me->invMass.assign(N, 1.0);
    me->totalMass = (double)N;
  }
  me->startTime = omp_get_wtime();
  me->ph->P.assign(3 * (size_t)N, 0.0);
  me->R0.assign(3 * (size_t)N, 0.0);
  me->charge.assign(N, 0.0);
  me->charged.assign(N, 0);
  me->atomCell.assign(N, 0);
  allocate_rigid_bodies(*me, bodies);
  me->cellAtom.allocate(N, 0);
  me->neighbor.resize(threads);
  for (auto& l : me->neighbor) l.allocate(extra, N, true, true);
  me->excluded.allocate(extra, N);
  me->pair.assign((size_t)me->ntypes * me->ntypes * layers, PairContainer());
  me->multilayer.assign((size_t)me->ntypes * me->ntypes, 0);
  me->overridable.assign((size_t)me->ntypes * me->ntypes, 1);
  me->interact.assign((size_t)me->ntypes * me->ntypes, 0);
  Model noCoul;
  noCoul.kind = COUL_NONE;
  me->multilayer_coulomb = false;
  me->coul.assign(layers, noCoul);
  me->pairs_exist.assign(layers, 0);
  me->useInRc.assign(layers, 0);
  me->bonded.assign(layers, 1);
  me->forcesUpToDate.assign(layers, 0);
  me->layerF.assign(3 * (size_t)N * layers, 0.0);
  me->layer = 1;
  me->initialized = false;

  tEmDee md;
  std::memset(&md, 0, sizeof(md));
  md.Builds = 0;
  md.Energy.UpToDate = false;
  me->layerEnergy.assign(layers, md.Energy);
  me->layerVirial.assign(layers, md.Virial);
  md.DoF = 3 * (N - 1);
  md.RotDoF = 0;
  md.Data = me;
  md.Options.Translate = true;
  md.Options.Rotate = true;
  md.Options.RotationMode = 0;
  md.Options.AutoBodyUpdate = true;
  md.Options.Compute = true;
  return md;
}

// src/EmDeeCode.f90:212-235
void* EmDee_memory_address(tEmDee md, const char* option) {
  System* me = sys(md);
  std::string item = option_string(option);
  if (item == "coordinates") return me->ph->R.data();
  if (item == "momenta") return me->ph->P.data();
  if (item == "forces") return me->F();
  if (item == "layerForces") return me->layerF.data();
  error("memory address retrieving", "invalid option " + item);
}

// src/EmDeeCode.f90:239-269
void EmDee_share_phase_space(tEmDee mdkeep, tEmDee* mdlose) {
  const char* task = "phase space sharing";
  System* keep = sys(mdkeep);
  System* lose = sys(*mdlose);
  if (!(keep->initialized && lose->initialized)) error(task, "EmDee system 1 has not been initialized");
  if (keep->natoms != lose->natoms) error(task, "different numbers of atoms");
  if (keep->atomType != lose->atomType) error(task, "atom types do not match");
  if (keep->mass != lose->mass) error(task, "atom masses do not match");
  if (keep->atomBody != lose->atomBody) error(task, "rigid bodies do not match");
  lose->ph = keep->ph;
  mdlose->Kinetic.Total = mdkeep.Kinetic.Total;
  mdlose->Kinetic.Rotational = mdkeep.Kinetic.Rotational;
  for (int x = 0; x < 3; ++x) {
    mdlose->Kinetic.TransPart[x] = mdkeep.Kinetic.TransPart[x];
    mdlose->Kinetic.RotPart[x] = mdkeep.Kinetic.RotPart[x];
  }
}

// src/EmDeeCode.f90:273-305
void EmDee_layer_based_parameters(tEmDee md, double InternalRc, int* Apply, int* Bonded) {
  const char* task = "layer-based parameter setting";
  System* me = sys(md);
  if (me->initialized) error(task, "system has already been initialized");
  bool any = false;
  for (int l = 0; l < me->nlayers; ++l) {
    me->useInRc[l] = Apply[l] != 0;
    any = any || me->useInRc[l];
  }
  if (any && (InternalRc <= 0.0 || InternalRc > me->Rc)) error(task, "invalid internal cutoff specification");
  me->InRc = InternalRc;
  me->InRcSq = InternalRc * InternalRc;
  me->xInRcSq = (InternalRc + me->skin) * (InternalRc + me->skin);
  for (int l = 0; l < me->nlayers; ++l) me->bonded[l] = Bonded[l] != 0;
  for (int l = 1; l <= me->nlayers; ++l) cutoff_setup(me->coul[l - 1], me->layerRc(l));
}

// src/EmDeeCode.f90:309-359
void EmDee_set_pair_model(tEmDee md, int itype, int jtype, void* model, double kCoul) {
  const char* task = "pair model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  if (!ranged({itype, jtype}, me->ntypes)) error(task, "provided type index is out of range");
  Model* m = as_model(model);
  if (m == nullptr || !is_pair(m->kind)) error(task, "a valid pair model must be provided");
  for (int layer = 1; layer <= me->nlayers; ++layer) set_pair_type(*me, itype, jtype, layer, *m, kCoul);
  me->tt(me->multilayer, itype, jtype) = 0;
  me->tt(me->multilayer, jtype, itype) = 0;
  if (itype == jtype) {
    for (int k = 1; k <= me->ntypes; ++k)
      if (k != itype && me->tt(me->overridable, itype, k)) {
        me->tt(me->multilayer, itype, k) = me->tt(me->multilayer, k, k);
        me->tt(me->multilayer, k, itype) = me->tt(me->multilayer, k, k);
      }
  } else {
    me->tt(me->overridable, itype, jtype) = 0;
    me->tt(me->overridable, jtype, itype) = 0;
  }
}

// src/EmDeeCode.f90:363-412
void EmDee_set_pair_multimodel(tEmDee md, int itype, int jtype, void* model[], double kCoul[]) {
  const char* task = "pair multimodel setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  if (!ranged({itype, jtype}, me->ntypes)) error(task, "provided type index is out of range");
  for (int layer = 1; layer <= me->nlayers; ++layer) {
    Model* m = as_model(model[layer - 1]);
    if (m == nullptr) error(task, std::to_string(me->nlayers) + " valid pair models must be provided");
    if (!is_pair(m->kind)) error(task, "a valid pair model must be provided");
    set_pair_type(*me, itype, jtype, layer, *m, kCoul[layer - 1]);
  }
  me->tt(me->multilayer, itype, jtype) = 1;
  me->tt(me->multilayer, jtype, itype) = 1;
  if (itype == jtype) {
    for (int k = 1; k <= me->ntypes; ++k)
      if (k != itype && me->tt(me->overridable, itype, k)) {
        me->tt(me->multilayer, itype, k) = 1;
        me->tt(me->multilayer, k, itype) = 1;
      }
  } else {
    me->tt(me->overridable, itype, jtype) = 0;
    me->tt(me->overridable, jtype, itype) = 0;
  }
}

// src/EmDeeCode.f90:416-444 (only the Ewald accuracy is kept; see perform_initialization)
void EmDee_set_kspace_model(tEmDee md, void* model) {
  const char* task = "kspace model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  Model* m = as_model(model);
  if (m == nullptr || m->kind != KSPACE_EWALD) error(task, "a valid kspace model must be provided");
  me->kspace = *m;
  me->kspace_active = true;
}

// src/EmDeeCode.f90:448-482
void EmDee_set_coul_model(tEmDee md, void* model) {
  const char* task = "coulomb model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  Model* m = as_model(model);
  if (m == nullptr || !is_coul(m->kind)) error(task, "a valid coulomb model must be provided");
  for (int layer = 1; layer <= me->nlayers; ++layer) {
    double layerRc = me->layerRc(layer);
    me->coul[layer - 1] = *m;
    cutoff_setup(me->coul[layer - 1], layerRc);
    modifier_setup(me->coul[layer - 1], layerRc);   // zeroes eshift/fshift again and resets Rm: Q1
  }
  me->multilayer_coulomb = false;
}

// src/EmDeeCode.f90:486-520
void EmDee_set_coul_multimodel(tEmDee md, void* model[]) {
  const char* task = "coulomb multimodel setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  for (int layer = 1; layer <= me->nlayers; ++layer) {
    double layerRc = me->layerRc(layer);
    Model* m = as_model(model[layer - 1]);
    if (m == nullptr) error(task, std::to_string(me->nlayers) + " valid coulomb models must be provided");
    if (!is_coul(m->kind)) error(task, "a valid coulomb model must be provided");
    me->coul[layer - 1] = *m;
    cutoff_setup(me->coul[layer - 1], layerRc);
    modifier_setup(me->coul[layer - 1], layerRc);
  }
  me->multilayer_coulomb = true;
}

// src/EmDeeCode.f90:524-570. The reference keeps, per atom, a sorted duplicate-free CSR row.
void EmDee_ignore_pair(tEmDee md, int i, int j) {
  System* me = sys(md);
  if (i == j || !ranged({i, j}, me->natoms)) return;
  List& ex = me->excluded;
  int n = ex.count;
  if (n + 2 > ex.nitems) ex.resize(n + extra);   // reference resizes when n == nitems; same effect, no overrun
  auto add_item = [&](int a, int b) {
    int start = ex.first[a - 1];
    int end = ex.last[a - 1];
    int pos;   // 1-based insertion position
    if (end < start) pos = end + 1;
    else if (b > ex.item[end - 1]) pos = end + 1;
    else {
      while (b > ex.item[start - 1]) start += 1;
      if (b == ex.item[start - 1]) return;
      pos = start;
    }
    for (int q = n; q >= pos; --q) ex.item[q] = ex.item[q - 1];
    ex.item[pos - 1] = b;
    for (int q = a; q < me->natoms; ++q) ex.first[q] += 1;
    for (int q = a - 1; q < me->natoms; ++q) ex.last[q] += 1;
    n += 1;
  };
  add_item(i, j);
  add_item(j, i);
  ex.count = n;
}

// src/EmDeeCode.f90:574-655
void EmDee_add_bond(tEmDee md, int i, int j, void* model) {
  System* me = sys(md);
  if (!ranged({i, j}, me->natoms)) error("add_bond", "atom index out of range");
  if (model == nullptr) error("add_bond", "a valid model must be provided");
  Model* m = as_model(model);
  if (m == nullptr || !(m->kind == BOND_NONE || m->kind == BOND_HARMONIC)) error("add_bond", "the provided model must be a bond model");
  me->bonds.push_back({i, j, 0, 0, *m});
  EmDee_ignore_pair(md, i, j);
}
void EmDee_add_angle(tEmDee md, int i, int j, int k, void* model) {
  System* me = sys(md);
  if (!ranged({i, j, k}, me->natoms)) error("add_angle", "atom index out of range");
  if (model == nullptr) error("add_angle", "a valid model must be provided");
  Model* m = as_model(model);
  if (m == nullptr || !(m->kind == ANGLE_NONE || m->kind == ANGLE_HARMONIC)) error("add_angle", "the provided model must be an angle model");
  me->angles.push_back({i, j, k, 0, *m});
  EmDee_ignore_pair(md, i, j);
  EmDee_ignore_pair(md, i, k);
  EmDee_ignore_pair(md, j, k);
}
void EmDee_add_dihedral(tEmDee md, int i, int j, int k, int l, void* model) {
  System* me = sys(md);
  if (!ranged({i, j, k, l}, me->natoms)) error("add_dihedral", "atom index out of range");
  if (model == nullptr) error("add_dihedral", "a valid model must be provided");
  Model* m = as_model(model);
  if (m == nullptr || m->kind != DIHEDRAL_NONE) error("add_dihedral", "the provided model must be a dihedral model");
  me->dihedrals.push_back({i, j, k, l, *m});   // stored and excluded; EmDee_compute_forces never evaluates dihedrals (1240-1241)
  const int a[4] = {i, j, k, l};
  for (int x = 0; x < 4; ++x)
    for (int y = x + 1; y < 4; ++y) EmDee_ignore_pair(md, a[x], a[y]);
}

void EmDee_compute_forces(tEmDee* md);

// src/EmDeeCode.f90:659-801
void EmDee_download(tEmDee md, const char* option, double* address) {
  System* me = sys(md);
  std::string item = option_string(option);
  if (address == nullptr) error("download", "provided address is invalid");
  const size_t n3 = 3 * (size_t)me->natoms;
  if (item == "box") {
    *address = me->ph->Lbox;
  } else if (item == "coordinates") {
    if (!me->ph->hasR) error("download", "coordinates have not been allocated");
    std::copy(me->ph->R.begin(), me->ph->R.end(), address);
  } else if (item == "momenta") {   // get_momenta (726-736)
    for (int f = 0; f < me->nfree; ++f)
      for (int x = 0; x < 3; ++x) address[3 * (size_t)me->free_[f] + x] = me->ph->P[3 * (size_t)me->free_[f] + x];
    for (const Body& b : me->ph->body) {
      std::vector<double> Pb(3 * (size_t)b.NP);
      b.particle_momenta(Pb.data());
      for (int k = 0; k < b.NP; ++k)
        for (int x = 0; x < 3; ++x) address[3 * (size_t)b.index[k] + x] = Pb[3 * k + x];
    }
  } else if (item == "forces") {
    if (!me->forcesUpToDate[me->layer - 1]) EmDee_compute_forces(&md);
    std::copy(me->F(), me->F() + n3, address);
  } else if (item == "centersOfMass") {
    for (int b = 0; b < me->nbodies; ++b)
      for (int x = 0; x < 3; ++x) address[3 * (size_t)b + x] = me->ph->body[b].rcm[x];
    for (int f = 0; f < me->nfree; ++f)
      for (int x = 0; x < 3; ++x) address[3 * (size_t)(me->nbodies + f) + x] = me->ph->R[3 * (size_t)me->free_[f] + x];
  } else if (item == "quaternions" || item == "quatmom" || item == "quattau") {   // get_quaternions (759-775)
    for (int i = 0; i < me->nbodies; ++i) {
      const Body& b = me->ph->body[i];
      double v[4];
      if (item == "quaternions") std::copy(b.q, b.q + 4, v);
      else if (item == "quatmom") std::copy(b.pi, b.pi + 4, v);
      else {
        const double t2[3] = {2.0 * b.tau[0], 2.0 * b.tau[1], 2.0 * b.tau[2]};
        mulC(b.q, t2, v);
      }
      std::copy(v, v + 4, address + 4 * (size_t)i);
    }
  } else if (item == "angmom" || item == "bodycoord" || item == "bodymom" || item == "bodyforces" ||
             item == "torques" || item == "inertia") {   // get_body_properties (777-799)
    for (int i = 0; i < me->nbodies; ++i) {
      const Body& b = me->ph->body[i];
      const double* v = item == "angmom" ? b.omega : item == "bodycoord" ? b.rcm : item == "bodymom" ? b.pcm
                      : item == "bodyforces" ? b.F : item == "torques" ? b.tau : b.MoI;
      std::copy(v, v + 3, address + 3 * (size_t)i);
    }
  } else {
    error("download", "invalid option");
  }
}

// src/EmDeeCode.f90:929-946
void EmDee_switch_model_layer(tEmDee* md, int layer) {
  System* me = sys(*md);
  if (layer != me->layer) {
    if (layer < 1 || layer > me->nlayers) error("model layer switch", "selected layer is out of range");
    me->layer = layer;
    md->Energy = me->layerEnergy[layer - 1];
    md->Virial = me->layerVirial[layer - 1];
  }
}

// src/EmDeeCode.f90:805-925
void EmDee_upload(tEmDee* md, const char* option, double* address) {
  System* me = sys(*md);
  std::string item = option_string(option);
  if (address == nullptr) error("upload", "provided address is invalid");
  auto initialize_system = [&]() {   // 916-923
    perform_initialization(*me, md->DoF, md->RotDoF);
    for (int layer = me->nlayers; layer >= 1; --layer) {
      EmDee_switch_model_layer(md, layer);
      EmDee_compute_forces(md);
    }
  };
  const size_t n3 = 3 * (size_t)me->natoms;
  if (item == "box") {
    me->ph->hasL = true;
    me->ph->Lbox = *address;
    if (me->initialized) invalidate(*me, md);
    else if (me->ph->hasR) initialize_system();
  } else if (item == "coordinates") {
    if (!me->ph->hasR) { me->ph->R.assign(n3, 0.0); me->ph->hasR = true; }
    std::copy(address, address + n3, me->ph->R.begin());
    if (me->initialized) {
      invalidate(*me, md);
      if (md->Options.AutoBodyUpdate) update_rigid_bodies(*me);
    } else if (me->ph->hasL) {
      initialize_system();
    }
  } else if (item == "momenta") {
    if (!me->initialized) error("upload", "box and coordinates have not been defined");
    double twoKEt[3], twoKEr[3];
    assign_momenta(*me, address, twoKEt, twoKEr);
    set_kinetic_from_sums(md, twoKEt, twoKEr);
  } else if (item == "forces") {
    if (!me->initialized) error("upload", "box and coordinates have not been defined");
    std::copy(address, address + n3, me->F());
  } else if (item == "charges") {
    if (me->initialized) error("upload", "cannot set charges after box and coordinates initialization");
    for (int i = 0; i < me->natoms; ++i) {
      me->charge[i] = address[i];
      me->charged[i] = std::fabs(address[i]) > 2.220446049250313e-16;   // epsilon(one)
    }
    invalidate(*me, md);
  } else {
    error("upload", "invalid option");
  }
}

// src/EmDeeCode.f90:950-1020
void EmDee_random_momenta(tEmDee* md, double kT, bool adjust, int seed) {
  System* me = sys(*md);
  if (me->random.seeding_required) me->random.setup(seed);
  double twoKEt[3] = {0, 0, 0}, twoKEr[3] = {0, 0, 0};
  Kiss& rng = me->random;
  if (me->nbodies != 0) {
    if (!me->initialized) error("random_momenta", "coordinates have not defined");
    for (Body& b : me->ph->body) {
      const double s = std::sqrt(b.mass * kT);
      for (int x = 0; x < 3; ++x) b.pcm[x] = s * rng.normal();
      double w[3];
      for (int x = 0; x < 3; ++x) w[x] = std::sqrt(b.invMoI[x] * kT) * rng.normal();
      b.assign_omega(w);
      for (int x = 0; x < 3; ++x) {
        twoKEt[x] = twoKEt[x] + b.invMass * b.pcm[x] * b.pcm[x];
        twoKEr[x] = twoKEr[x] + b.MoI[x] * b.omega[x] * b.omega[x];
      }
    }
  }
  for (int f = 0; f < me->nfree; ++f) {
    int i = me->free_[f];
    double s = std::sqrt(me->mass[i] * kT);
    for (int x = 0; x < 3; ++x) me->ph->P[3 * (size_t)i + x] = s * rng.normal();
    for (int x = 0; x < 3; ++x) twoKEt[x] += me->invMass[i] * me->ph->P[3 * (size_t)i + x] * me->ph->P[3 * (size_t)i + x];
  }
  if (adjust) {   // adjust_momenta, 994-1018
    double vcm[3];
    int bodyDoF = 0;
    for (const Body& b : me->ph->body) bodyDoF += b.dof;
    for (int x = 0; x < 3; ++x) {
      double s = 0.0, sb = 0.0;
      for (int f = 0; f < me->nfree; ++f) s += me->ph->P[3 * (size_t)me->free_[f] + x];
      for (const Body& b : me->ph->body) sb += b.pcm[x];
      vcm[x] = (s + sb) / me->totalMass;
    }
    for (int x = 0; x < 3; ++x) twoKEt[x] = 0.0;
    for (int f = 0; f < me->nfree; ++f) {
      int i = me->free_[f];
      for (int x = 0; x < 3; ++x) {
        me->ph->P[3 * (size_t)i + x] = me->ph->P[3 * (size_t)i + x] - me->mass[i] * vcm[x];
        twoKEt[x] += me->invMass[i] * me->ph->P[3 * (size_t)i + x] * me->ph->P[3 * (size_t)i + x];
      }
    }
    for (Body& b : me->ph->body)
      for (int x = 0; x < 3; ++x) {
        b.pcm[x] = b.pcm[x] - b.mass * vcm[x];
        twoKEt[x] += b.invMass * b.pcm[x] * b.pcm[x];
      }
    double total = 0.0;
    for (int x = 0; x < 3; ++x) total += twoKEt[x] + twoKEr[x];
    double factor = std::sqrt((3 * me->nfree + bodyDoF - 3) * kT / total);
    for (int f = 0; f < me->nfree; ++f) {
      int i = me->free_[f];
      for (int x = 0; x < 3; ++x) me->ph->P[3 * (size_t)i + x] = factor * me->ph->P[3 * (size_t)i + x];
    }
    for (Body& b : me->ph->body) {
      for (int x = 0; x < 3; ++x) b.pcm[x] = factor * b.pcm[x];
      const double w[3] = {factor * b.omega[0], factor * b.omega[1], factor * b.omega[2]};
      b.assign_omega(w);
    }
    for (int x = 0; x < 3; ++x) { twoKEt[x] = factor * factor * twoKEt[x]; twoKEr[x] = factor * factor * twoKEr[x]; }
  }
  set_kinetic_from_sums(md, twoKEt, twoKEr);
}

// src/EmDeeCode.f90:1024-1065 with boost / kinetic_energies (src/EmDeeData.f90:864-922)
void EmDee_boost(tEmDee* md, double lambda, double alpha, double dt) {
  System* me = sys(*md);
  double CF = phi(alpha * dt) * dt;
  double CP = 1.0 - alpha * CF;
  CF = lambda * CF;
  if (lambda != 0.0 && !me->forcesUpToDate[me->layer - 1]) EmDee_compute_forces(md);
  const bool compute = md->Options.Compute, translate = md->Options.Translate, rotate = md->Options.Rotate;
  boost(*me, CP, CF, me->F(), translate, rotate);
  if (compute) {
    double twoKEt[3], twoKEr[3];
    kinetic_energies(*me, translate, rotate, twoKEt, twoKEr);
    if (translate)
      for (int x = 0; x < 3; ++x) md->Kinetic.TransPart[x] = 0.5 * twoKEt[x];
    if (rotate) {
      for (int x = 0; x < 3; ++x) md->Kinetic.RotPart[x] = 0.5 * twoKEr[x];
      md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
    }
    md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] +
                        md->Kinetic.Rotational;
  }
  md->Kinetic.UpToDate = compute;
}

// src/EmDeeCode.f90:1069-1103 with move (src/EmDeeData.f90:823-860)
void EmDee_displace(tEmDee* md, double lambda, double alpha, double dt) {
  System* me = sys(*md);
  md->Time.Motion -= omp_get_wtime();
  double CR, CP;
  if (alpha != 0.0) {
    CP = phi(alpha * dt) * dt;
    CR = 1.0 - alpha * CP;
    me->ph->Lbox = CR * me->ph->Lbox;
  } else {
    CP = dt;
    CR = 1.0;
  }
  CP = lambda * CP;
  move(*me, CR, CP, dt, md->Options.Translate, md->Options.Rotate, md->Options.RotationMode);
  invalidate(*me, md);
  md->Time.Motion += omp_get_wtime();
}

// src/EmDeeCode.f90:1107-1211: one velocity-Verlet step with the shadow-Hamiltonian bookkeeping
void EmDee_verlet_step(tEmDee* md, double dt) {
  System* me = sys(*md);
  const double dt_2 = 0.5 * dt;
  const bool compute = md->Options.Compute;
  const int mode = md->Options.RotationMode;
  double Us = 0.0, Ks_t = 0.0, Ks_r = 0.0;
  std::vector<double> r0, q0, s0;
  auto virtual_rotation = [&](const Body& b, double tstep, double q[4]) {   // 1150-1163
    Body c = b;
    const double t3[3] = {tstep * c.tau[0], tstep * c.tau[1], tstep * c.tau[2]};
    double t4[4], np[4];
    mulC(c.q, t3, t4);
    for (int x = 0; x < 4; ++x) np[x] = c.pi[x] + t4[x];
    c.assign_pi(np);
    if (mode != 0) c.rotate_no_squish(tstep, mode);
    else c.rotate_exact(tstep);
    std::copy(c.q, c.q + 4, q);
  };
  // pre_force (1165-1182)
  if (compute) {
    r0.assign(3 * (size_t)me->nbodies, 0.0);
    q0.assign(4 * (size_t)me->nbodies, 0.0);
    s0.assign(3 * (size_t)me->nfree, 0.0);
    for (int i = 0; i < me->nbodies; ++i) {
      const Body& b = me->ph->body[i];
      for (int x = 0; x < 3; ++x) r0[3 * (size_t)i + x] = 2.5 * b.rcm[x] + dt_2 * b.invMass * (b.pcm[x] - dt_2 * b.F[x]);
      double vq[4];
      virtual_rotation(b, -dt, vq);
      for (int x = 0; x < 4; ++x) q0[4 * (size_t)i + x] = 0.5 * vq[x] - 3.0 * b.q[x];
    }
    const double* F = me->F();
    for (int f = 0; f < me->nfree; ++f) {
      const int j = me->free_[f];
      for (int x = 0; x < 3; ++x)
        s0[3 * (size_t)f + x] = 2.5 * me->ph->R[3 * (size_t)j + x] +
                                dt_2 * me->invMass[j] * (me->ph->P[3 * (size_t)j + x] - dt_2 * F[3 * (size_t)j + x]);
    }
  }
  boost(*me, 1.0, dt_2, me->F(), true, true);
  move(*me, 1.0, dt, dt, true, true, mode);

  EmDee_compute_forces(md);

  // post_force (1184-1209)
  boost(*me, 1.0, dt_2, me->F(), true, true);
  if (compute) {
    double twoKEt[3], twoKEr[3];
    kinetic_energies(*me, true, true, twoKEt, twoKEr);
    for (int i = 0; i < me->nbodies; ++i) {
      const Body& b = me->ph->body[i];
      double rdot[3], vq[4], qdot[4];
      for (int x = 0; x < 3; ++x) {
        rdot[x] = 2.5 * b.rcm[x] + dt * b.invMass * (b.pcm[x] + dt_2 * b.F[x]) - r0[3 * (size_t)i + x];
        Ks_t = Ks_t + rdot[x] * b.pcm[x];   // sum(rdot*pcm), accumulated term by term
      }
      virtual_rotation(b, dt, vq);
      double qq = 0.0;
      for (int x = 0; x < 4; ++x) { qdot[x] = q0[4 * (size_t)i + x] + 1.5 * b.q[x] + vq[x]; }
      for (int x = 0; x < 4; ++x) qq += qdot[x] * b.q[x];
      double acc = 0.0;
      for (int x = 0; x < 4; ++x) acc += (qdot[x] - qq * b.q[x]) * b.pi[x];
      Ks_r = Ks_r + acc;
      double t4[4], tau_b[3];
      mulC(b.q, b.tau, t4);
      mulBt(b.q, t4, tau_b);
      double ff = 0.0, tt = 0.0;
      for (int x = 0; x < 3; ++x) { ff += b.F[x] * b.F[x]; tt += b.invMoI[x] * tau_b[x] * tau_b[x]; }
      Us = Us + b.invMass * ff + tt;
    }
    const double* F = me->F();
    for (int f = 0; f < me->nfree; ++f) {
      const int j = me->free_[f];
      double ff = 0.0, rp = 0.0;
      for (int x = 0; x < 3; ++x) {
        const double rdot = 2.5 * me->ph->R[3 * (size_t)j + x] +
                            dt * me->invMass[j] * (me->ph->P[3 * (size_t)j + x] + dt_2 * F[3 * (size_t)j + x]) - s0[3 * (size_t)f + x];
        rp += rdot * me->ph->P[3 * (size_t)j + x];
        ff += F[3 * (size_t)j + x] * F[3 * (size_t)j + x];
      }
      Ks_t = Ks_t + rp;
      Us = Us + me->invMass[j] * ff;
    }
    for (int x = 0; x < 3; ++x) { md->Kinetic.TransPart[x] = 0.5 * twoKEt[x]; md->Kinetic.RotPart[x] = 0.5 * twoKEr[x]; }
    md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
    md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
    md->Energy.ShadowPotential = md->Energy.ShadowPotential - dt * dt * Us / 24.0;
    md->Kinetic.ShadowRotational = Ks_r / (6.0 * dt);
    md->Kinetic.ShadowKinetic = (Ks_t + Ks_r) / (6.0 * dt);
  }
  md->Kinetic.UpToDate = compute;
}

// src/EmDeeCode.f90:1215-1277
void EmDee_compute_forces(tEmDee* md) {
  System* me = sys(*md);
  const int N = me->natoms, T = me->nthreads;
  enum { pair = 0, coul = 1, long_ = 2, bond = 3, angle = 4 };
  std::vector<double> Rs(3 * (size_t)N), Fs(3 * (size_t)N * T);
  for (size_t q = 0; q < 3 * (size_t)N; ++q) Rs[q] = me->ph->R[q] / me->ph->Lbox;

  handle_neighbor_lists(*me, md->Builds, md->Time.Neighbor, Rs.data());

  md->Time.Pair -= omp_get_wtime();
  const bool compute = md->Options.Compute;
  double E[5] = {0, 0, 0, 0, 0}, W[5] = {0, 0, 0, 0, 0};
  std::vector<double> Et(4 * (size_t)T, 0.0);
#pragma omp parallel num_threads(T)
  {
    const int thread = omp_get_thread_num() + 1;
    double* F = Fs.data() + 3 * (size_t)N * (thread - 1);
    double Ep = 0, Ec = 0, Wp = 0, Wc = 0;
    if (compute) compute_pairs<true>(*me, thread, Rs.data(), F, Ep, Ec, Wp, Wc);
    else compute_pairs<false>(*me, thread, Rs.data(), F, Ep, Ec, Wp, Wc);
    Et[4 * (size_t)(thread - 1) + 0] = Ep;
    Et[4 * (size_t)(thread - 1) + 1] = Ec;
    Et[4 * (size_t)(thread - 1) + 2] = Wp;
    Et[4 * (size_t)(thread - 1) + 3] = Wc;
  }
  for (int t = 0; t < T; ++t) {   // reduction(+:E,W)
    E[pair] += Et[4 * (size_t)t + 0];
    E[coul] += Et[4 * (size_t)t + 1];
    W[pair] += Et[4 * (size_t)t + 2];
    W[coul] += Et[4 * (size_t)t + 3];
  }
  double* Fm = me->F();   // me%F = sum(Fs,3)
#pragma omp parallel for num_threads(T) schedule(static)
  for (long long q = 0; q < 3LL * N; ++q) {
    double s = 0.0;
    for (int t = 0; t < T; ++t) s += Fs[(size_t)t * 3 * N + q];
    Fm[q] = s;
  }
  compute_bonds(*me, Rs.data(), Fm, E[bond], W[bond], E[coul]);      // 1240: after the thread sum here (addition commutes)
  compute_angles(*me, Rs.data(), Fm, E[angle], W[angle], E[coul]);   // 1241
  if (me->coul[me->layer - 1].requires_kspace) {
    compute_kspace(*me, Rs.data(), E[long_], Fm);
    W[long_] = E[coul] + E[long_] - W[coul];
  }
  md->Virial.Total = W[0] + W[1] + W[2] + W[3] + W[4];
  if (me->nbodies != 0) {
    md->Virial.Body = rigid_body_virial(*me);
    md->Virial.Total = md->Virial.Total + md->Virial.Body;
  }
  if (compute) {
    md->Energy.Dispersion = E[pair];
    md->Energy.Coulomb = E[coul] + E[long_];
    md->Energy.Bond = E[bond];
    md->Energy.Angle = E[angle];
    md->Energy.Potential = E[0] + E[1] + E[2] + E[3] + E[4];
    md->Energy.ShadowPotential = md->Energy.Potential;
  }
  md->Energy.UpToDate = compute;
  me->forcesUpToDate[me->layer - 1] = 1;
  me->layerEnergy[me->layer - 1] = md->Energy;
  me->layerVirial[me->layer - 1] = md->Virial;
  double time = omp_get_wtime();
  md->Time.Pair += time;
  md->Time.Total = time - me->startTime;
}

// reference src/EmDeeCode.f90:1281-1395 (pair counting over the EXISTING half neighbor list, no rebuild) and
// src/math.f90:695-702 (symm1D). Restated literally, including: pairs the list does not hold (excluded, same
// body, non-interacting types, beyond Rc+skin) are never counted; `pairOn(itype,jtype) = .true.` with vector
// subscripts switches on every (itype(k), jtype(l)) combination; the `middle` split is used when
// Rc < 1.0001*InRc. One guard added: a distance that rounds onto Rc exactly would index bin `bins+1`
// (out of bounds in the reference); it is dropped here.
void EmDee_rdf(tEmDee md, int bins, double Rc, int pairs, int* itype, int* jtype, double* g) {
  System* me = sys(md);
  const char* task = "radial distribution calculation";
  for (int k = 0; k < pairs; ++k)
    if (itype[k] < 1 || itype[k] > me->ntypes || jtype[k] < 1 || jtype[k] > me->ntypes)
      error(task, "at least one provided type index is out of range");
  auto symm1D = [](int i, int j) {
    const int x = std::min(i, j) - 1, y = std::max(i, j) - 1;
    return x + (y + 1) * y / 2 + 1;
  };
  const int nt = me->ntypes;
  std::vector<char> hasPair(nt + 1, 0), pairOn((size_t)(nt + 1) * (nt + 1), 0);
  int maxtype = 0;
  for (int k = 0; k < pairs; ++k) {
    hasPair[itype[k]] = hasPair[jtype[k]] = 1;
    maxtype = std::max(maxtype, std::max(itype[k], jtype[k]));
    for (int l = 0; l < pairs; ++l) {
      pairOn[(size_t)itype[k] * (nt + 1) + jtype[l]] = 1;
      pairOn[(size_t)jtype[l] * (nt + 1) + itype[k]] = 1;
    }
  }
  const int nsym = symm1D(maxtype, maxtype);
  std::vector<long long> pairCount((size_t)bins * nsym, 0);
  const double invL = 1.0 / me->ph->Lbox, invL2 = invL * invL;
  const int N = me->natoms;
  std::vector<double> Rs(3 * (size_t)N);
  for (size_t q = 0; q < Rs.size(); ++q) Rs[q] = invL * me->ph->R[q];
  const double Rc2 = Rc * Rc * invL2;
  const double binsByRc = bins / (Rc * invL);
  const bool useMiddle = Rc < 1.0001 * me->InRc;
  for (int thread = 1; thread <= me->nthreads; ++thread) {   // serial over the per-thread lists: integer counts
    List& neighbor = me->neighbor[thread - 1];
    const std::vector<int>& last = useMiddle ? neighbor.middle : neighbor.last;
    const int tfirst = me->threadCell.first[thread - 1], tlast = me->threadCell.last[thread - 1];
    if (tlast < tfirst) continue;
    for (int k = me->cellAtom.first[tfirst - 1]; k <= me->cellAtom.last[tlast - 1]; ++k) {
      const int i = me->cellAtom.item[k - 1];
      const int it = me->atomType[i - 1];
      if (!hasPair[it]) continue;
      const double* Ri = &Rs[3 * (size_t)(i - 1)];
      for (int m = neighbor.first[i - 1]; m <= last[i - 1]; ++m) {
        const int j = neighbor.item[m - 1];
        const int jt = me->atomType[j - 1];
        if (!pairOn[(size_t)it * (nt + 1) + jt]) continue;
        const double* Rj = &Rs[3 * (size_t)(j - 1)];
        const double d0 = pbc(Ri[0] - Rj[0]), d1 = pbc(Ri[1] - Rj[1]), d2 = pbc(Ri[2] - Rj[2]);
        const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        if (r2 < Rc2) {
          const int bin = (int)(std::sqrt(r2) * binsByRc) + 1;
          if (bin <= bins) pairCount[(size_t)(symm1D(it, jt) - 1) * bins + (bin - 1)] += 1;
        }
      }
    }
  }
  const double Pi4_3 = 4.188790204786391;
  const double w = Rc * invL / bins;
  const double shell0 = Pi4_3 * (w * w * w);
  std::vector<long long> count(maxtype + 1, 0);
  for (int a = 0; a < N; ++a)
    if (me->atomType[a] <= maxtype) count[me->atomType[a]] += 1;
  for (int p = 0; p < pairs; ++p) {
    const int i = itype[p], j = jtype[p];
    const double NiNj = (double)(count[i] * count[j]);
    for (int b = 1; b <= bins; ++b) {
      double rdf = (double)pairCount[(size_t)(symm1D(i, j) - 1) * bins + (b - 1)] / shell0;
      rdf = rdf / (double)(3 * b * (b - 1) + 1);
      g[(size_t)p * bins + (b - 1)] = (i == j) ? 2.0 * rdf / NiNj : rdf / NiNj;
    }
  }
}

// ---- modifiers (src/modelClass_nonbonded.f90:83-241) ---------------------------------------------
static void* wrap_modifier(void* model, int modifier, double skin, bool has_skin, const char* task) {
  Model* m = as_model(model);
  if (m == nullptr || !is_nonbonded(m->kind)) error(task, "a valid pair model must be provided");
  Model n = *m;
  n.modifier = modifier;
  if (has_skin) n.skin = skin;
  return deliver(n);
}
void* EmDee_shifted(void* model) { return wrap_modifier(model, SHIFTED, 0, false, "shifted potential assignment"); }
void* EmDee_shifted_force(void* model) { return wrap_modifier(model, SHIFTED_FORCE, 0, false, "shifted-force potential assignment"); }
void* EmDee_smoothed(void* model, double skin) { return wrap_modifier(model, SMOOTHED, skin, true, "smoothed potential assignment"); }
void* EmDee_shifted_smoothed(void* model, double skin) { return wrap_modifier(model, SHIFTED_SMOOTHED, skin, true, "shifted-smoothed potential assignment"); }
void* EmDee_square_smoothed(void* model, double skin) { return wrap_modifier(model, SQUARE_SMOOTHED, skin, true, "square-smoothed potential assignment"); }
void* EmDee_shifted_square_smoothed(void* model, double skin) { return wrap_modifier(model, SHIFTED_SQUARE_SMOOTHED, skin, true, "shifted-square-smoothed potential assignment"); }

// ---- constructors --------------------------------------------------------------------------------
static void* make(Kind k, const char* name) { Model m; m.kind = k; m.name = name; return deliver(m); }
void* EmDee_pair_none(void) { return make(PAIR_NONE, "none"); }
void* EmDee_coul_none(void) { return make(COUL_NONE, "none"); }
void* EmDee_bond_none(void) { return make(BOND_NONE, "none"); }
void* EmDee_angle_none(void) { return make(ANGLE_NONE, "none"); }
void* EmDee_dihedral_none(void) { return make(DIHEDRAL_NONE, "none"); }
void* EmDee_pair_lj_cut(double epsilon, double sigma) { Model m; setup_lj(m, epsilon, sigma); return deliver(m); }
void* EmDee_pair_softcore_cut(double epsilon, double sigma, double lambda) { Model m; setup_softcore(m, epsilon, sigma, lambda); return deliver(m); }
void* EmDee_coul_cut(void) { return make(COUL_CUT, "cut"); }
void* EmDee_coul_sf(void) { Model m; m.kind = COUL_SF; m.name = "sf"; m.shifted_force = true; return deliver(m); }
void* EmDee_coul_damped(double damp) {   // src/coul_damped.f90:50-64
  Model m; m.kind = COUL_DAMPED; m.name = "damped"; m.damp = damp; m.alpha = damp; m.beta = 2.0 * m.alpha / std::sqrt(Pi);
  return deliver(m);
}
void* EmDee_coul_long(void) { Model m; m.kind = COUL_LONG; m.name = "long"; m.requires_kspace = true; return deliver(m); }
void* EmDee_coul_damped_smoothed(double damp, double skinWidth) {   // src/coul_damped_smoothed.f90:54-72
  Model m; m.kind = COUL_DAMPED_SMOOTHED; m.name = "damped_openmm_smoothed"; m.damp = damp; m.skinWidth = skinWidth;
  m.alpha = damp; m.beta = 2.0 * m.alpha / std::sqrt(Pi);
  return deliver(m);
}
void* EmDee_coul_damped_square_smoothed(double damp, double skinWidth) {   // src/coul_damped_square_smoothed.f90:54-72
  Model m; m.kind = COUL_DAMPED_SQUARE_SMOOTHED; m.name = "damped_smoothed"; m.damp = damp; m.skinWidth = skinWidth;
  m.alpha = damp; m.beta = 2.0 * m.alpha / std::sqrt(Pi);
  return deliver(m);
}
void* EmDee_coul_square_smoothed(double skinWidth) { Model m; m.kind = COUL_SQUARE_SMOOTHED; m.name = "smoothed"; m.skinWidth = skinWidth; return deliver(m); }
void* EmDee_coul_shifted_square_smoothed(double skinWidth) {
  Model m; m.kind = COUL_SHIFTED_SQUARE_SMOOTHED; m.name = "shifted_smoothed"; m.skinWidth = skinWidth; m.shifted = true;
  return deliver(m);
}
void* EmDee_bond_harmonic(double k, double r0) { Model m; m.kind = BOND_HARMONIC; m.name = "harmonic"; m.p1 = k; m.p2 = r0; return deliver(m); }
void* EmDee_angle_harmonic(double k, double theta0) { Model m; m.kind = ANGLE_HARMONIC; m.name = "harmonic"; m.p1 = k; m.p2 = theta0; return deliver(m); }
void* EmDee_kspace_ewald(double accuracy) { Model m; m.kind = KSPACE_EWALD; m.name = "ewald"; m.accuracy = accuracy; return deliver(m); }

// ---- extensions (include/emdee_ext.h) ------------------------------------------------------------
long long EmDeeX_pair_count(tEmDee md) {
  System* me = sys(md);
  long long n = 0;
  for (auto& l : me->neighbor) n += l.count;
  return n;
}

long long EmDeeX_download_pairs(tEmDee md, int* pairs, long long capacity) {
  System* me = sys(md);
  long long n = 0;
  for (int t = 0; t < me->nthreads; ++t) {
    List& nl = me->neighbor[t];
    const int tfirst = me->threadCell.first[t], tlast = me->threadCell.last[t];
    if (tlast < tfirst) continue;
    for (int k = me->cellAtom.first[tfirst - 1]; k <= me->cellAtom.last[tlast - 1]; ++k) {
      int i = me->cellAtom.item[k - 1];
      for (int m = nl.first[i - 1]; m <= nl.last[i - 1]; ++m) {
        if (n >= capacity) return n;
        int j = nl.item[m - 1];
        pairs[2 * n] = std::min(i, j) - 1;
        pairs[2 * n + 1] = std::max(i, j) - 1;
        ++n;
      }
    }
  }
  return n;
}

// Oracle-only (not part of the boundary in include/): flat dump of one layer's interaction tables, which
// tests/test_host_tables.py compares with what the product's host shim hands to its kernels.
// Records of 13 doubles {kind, modifier, eshift, fshift, Rm, factor, Rm2fac, a, b, c, d, kCoul, coulomb} for
// the ntypes x ntypes pair entries (row-major) and then the Coulomb model, followed by the ntypes x ntypes
// `interact` mask, `pairs_exist` and `useInRc`. Kind codes: 0 pair_none, 1 lj_cut, 2 softcore_cut, 10 coul_none,
// 11 cut, 12 sf, 13 damped/long, 14 damped_smoothed, 15 damped_square_smoothed, 16 square_smoothed,
// 17 shifted_square_smoothed. Slots: LJ a=4eps b=24eps c=sigma^2; softcore a=prefactor b=6*prefactor
// c=1/sigma^2 d=shift; Coulomb a=alpha b=beta c=Rm^2 d=1/Rm. Returns the number of doubles written.
int EmDeeX_dump_tables(tEmDee md, int layer, double* out, int cap) {
  System* me = sys(md);
  if (layer < 1 || layer > me->nlayers) return -1;
  const int nt = me->ntypes;
  const int need = 13 * (nt * nt + 1) + nt * nt + 2;
  if (cap < need) return -need;
  int k = 0;
  auto code = [](Kind kd) {
    switch (kd) {
      case PAIR_NONE: return 0;
      case PAIR_LJ_CUT: return 1;
      case PAIR_SOFTCORE_CUT: return 2;
      case COUL_NONE: return 10;
      case COUL_CUT: return 11;
      case COUL_SF: return 12;
      case COUL_DAMPED: case COUL_LONG: return 13;
      case COUL_DAMPED_SMOOTHED: return 14;
      case COUL_DAMPED_SQUARE_SMOOTHED: return 15;
      case COUL_SQUARE_SMOOTHED: return 16;
      case COUL_SHIFTED_SQUARE_SMOOTHED: return 17;
      default: return -1;
    }
  };
  auto put = [&](const Model& m, double kCoul, double coulomb) {
    double a = 0, b = 0, c = 0, d = 0;
    if (m.kind == PAIR_LJ_CUT) { a = m.eps4; b = m.eps24; c = m.sigsq; }
    else if (m.kind == PAIR_SOFTCORE_CUT) { a = m.prefactor; b = m.prefactor6; c = m.invSigSq; d = m.shift; }
    else if (is_coul(m.kind)) { a = m.alpha; b = m.beta; c = m.Rm2; d = m.invRm; }
    const double v[13] = {(double)code(m.kind), (double)m.modifier, m.eshift, m.fshift, m.Rm, m.factor, m.Rm2fac,
                          a, b, c, d, kCoul, coulomb};
    for (double x : v) out[k++] = x;
  };
  for (int i = 1; i <= nt; ++i)
    for (int j = 1; j <= nt; ++j) {
      PairContainer& p = me->pr(i, j, layer);
      put(p.model, p.coulomb ? p.kCoul : 0.0, p.coulomb ? 1.0 : 0.0);
    }
  put(me->coul[layer - 1], 0.0, 0.0);
  for (int i = 1; i <= nt; ++i)
    for (int j = 1; j <= nt; ++j) out[k++] = me->tt(me->interact, i, j) ? 1.0 : 0.0;
  out[k++] = me->pairs_exist[layer - 1] ? 1.0 : 0.0;
  out[k++] = me->useInRc[layer - 1] ? 1.0 : 0.0;
  return k;
}

void EmDeeX_finalize(tEmDee* md) {
  delete sys(*md);
  md->Data = nullptr;
}

}  // extern "C"
