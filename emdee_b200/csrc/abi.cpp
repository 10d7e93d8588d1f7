// abi.cpp -- the drop-in boundary: every `bind(C)` entry point of the reference library
// (src/EmDeeCode.f90, generated src/models.f90) re-hosted in C++ on top of the CUDA engine.
//
// What lives here: the reference's host-side semantics for the nonbonded path -- model objects and
// the order-dependent setup rules (set_pair_type + mixing + modifier_setup, cutoff_setup; reference
// src/EmDeeData.f90:193-264, src/modelClass_nonbonded.f90:245-296, src/modelClass_coul.f90:62-93,
// src/modelClass_pair.f90:120-140), lazy initialisation and up-to-date flags (src/EmDeeCode.f90:805-946),
// the error convention (src/global.f90:51-56). What does NOT live here: any arithmetic over atoms --
// that is all in engine.cu, on the device. There is no CPU fallback.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/emdee.h"
#include "../../include/emdee_ext.h"
#include "engine.h"

static_assert(sizeof(tEmDee) == 240, "tEmDee must be 240 bytes (reference src/EmDeeCode.f90:51-61)");
static_assert(offsetof(tEmDee, Data) == 216 && offsetof(tEmDee, Options) == 224, "tEmDee layout");

namespace {

using emdee::Engine;
using nb::DevModel;

[[noreturn]] void error(const char* task, const std::string& msg) {   // src/global.f90:51-56
  std::fprintf(stderr, "Error in %s: %s.\n", task, msg.c_str());
  std::fflush(stderr);
  std::exit(1);
}
void warning(const std::string& msg) { std::fprintf(stderr, "WARNING: %s.\n", msg.c_str()); }   // :60-64

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- model handles ------------------------------------------------------------------------------
enum Family { F_PAIR, F_COUL, F_BOND, F_ANGLE, F_DIHEDRAL, F_KSPACE };
constexpr uint64_t HANDLE_TAG = 0x456d446565423230ull;   // "EmDeeB20"

struct HostModel {
  uint64_t tag = HANDLE_TAG;
  Family family = F_PAIR;
  const char* name = "none";
  DevModel dev{};            // what the kernels see
  // host-only state of the reference's class hierarchy
  double skin = 0.0;         // cNonBondedModel%skin (modifier smoothing width)
  double RmSq = 0.0;
  bool shifted = false, shifted_force = false, requires_kspace = false;   // cCoulModel flags
  double epsilon = 0.0, sigma = 0.0, lambda = 0.0;                        // raw pair parameters (for mixing)
  double skinWidth = 0.0;                                                  // smoothed Coulomb models
  double accuracy = 0.0;                                                   // kspace_ewald
  bool is_long = false;                                                    // coul_long (kind shares K_COUL_DAMPED)
};

HostModel blank(Family fam, int kind, const char* name) {
  HostModel m;
  m.family = fam;
  m.name = name;
  std::memset(&m.dev, 0, sizeof(m.dev));
  m.dev.kind = kind;
  m.dev.modifier = nb::M_NONE;
  return m;
}

HostModel* handle(void* p) {
  if (p == nullptr) return nullptr;
  HostModel* m = static_cast<HostModel*>(p);
  return m->tag == HANDLE_TAG ? m : nullptr;
}
void* deliver(const HostModel& m) { return new HostModel(m); }   // never freed, like src/modelClass.f90:51-60

void eval(const HostModel& m, double invR, double invR2, double& E, double& W) {
  nb::eval_kind<nb::K_DYNAMIC>(m.dev, nb::make_dist(invR, invR2), E, W);
}

HostModel make_lj(double epsilon, double sigma) {   // src/pair_lj_cut.f90:52-69
  HostModel m = blank(F_PAIR, nb::K_PAIR_LJ_CUT, "lj_cut");
  m.epsilon = epsilon;
  m.sigma = sigma;
  m.dev.a = 4.0 * epsilon;
  m.dev.b = 24.0 * epsilon;
  m.dev.c = sigma * sigma;
  return m;
}
HostModel make_softcore(double epsilon, double sigma, double lambda) {   // src/pair_softcore_cut.f90:58-82
  HostModel m = blank(F_PAIR, nb::K_PAIR_SOFTCORE_CUT, "softcore_cut");
  m.epsilon = epsilon;
  m.sigma = sigma;
  m.lambda = lambda;
  if (lambda < 0.0 || lambda > 1.0) error("pair_softcore_cut setup", "out-of-range parameter lambda");
  m.dev.a = 4.0 * epsilon * lambda;
  m.dev.b = 6.0 * m.dev.a;
  m.dev.c = 1.0 / (sigma * sigma);
  m.dev.d = 0.5 * (1.0 - lambda);
  return m;
}

// src/modelClass_nonbonded.f90:245-296
void modifier_setup(HostModel& m, double cutoff) {
  DevModel& d = m.dev;
  d.fshift = 0.0;
  d.eshift = 0.0;
  d.Rm = cutoff - m.skin;
  m.RmSq = d.Rm * d.Rm;
  const int mod = d.modifier;
  const bool shifting = mod == nb::M_SHIFTED || mod == nb::M_SHIFTED_FORCE || mod == nb::M_SHIFTED_SMOOTHED ||
                        mod == nb::M_SHIFTED_SQUARE_SMOOTHED;
  double Ec = 0, Wc = 0, Es = 0, Ws = 0;
  if (shifting) eval(m, 1.0 / cutoff, 1.0 / (cutoff * cutoff), Ec, Wc);
  if (mod == nb::M_SHIFTED) {
    d.eshift = -Ec;
  } else if (mod == nb::M_SHIFTED_FORCE) {
    d.eshift = -(Ec + Wc);
    d.fshift = Wc / cutoff;
  } else if (mod == nb::M_SMOOTHED || mod == nb::M_SHIFTED_SMOOTHED) {
    if (shifting) {
      eval(m, 1.0 / d.Rm, 1.0 / m.RmSq, Es, Ws);
      d.eshift = -0.5 * (Es + Ec);
    }
    d.factor = 1.0 / (cutoff - d.Rm);
    d.Rm2fac = d.factor * d.Rm;
  } else if (mod == nb::M_SQUARE_SMOOTHED || mod == nb::M_SHIFTED_SQUARE_SMOOTHED) {
    if (shifting) {
      eval(m, 1.0 / d.Rm, 1.0 / m.RmSq, Es, Ws);
      d.eshift = -0.5 * (Es + Ec);
    }
    d.factor = 1.0 / (cutoff * cutoff - m.RmSq);
    d.Rm2fac = d.factor * m.RmSq;
  }
}

// src/modelClass_coul.f90:62-93 + the *_apply_cutoff overrides of the smoothed Coulomb models
void cutoff_setup(HostModel& m, double cutoff) {
  DevModel& d = m.dev;
  d.fshift = 0.0;
  d.eshift = 0.0;
  if (m.shifted || m.shifted_force) {
    double invR = 1.0 / cutoff, E = 0, W = 0;
    eval(m, invR, invR * invR, E, W);
    if (m.shifted_force) {
      d.fshift = W / cutoff;
      d.eshift = -(E + W);
    } else {
      d.fshift = 0.0;
      d.eshift = -E;
    }
  }
  switch (d.kind) {
    case nb::K_COUL_DAMPED_SMOOTHED:   // src/coul_damped_smoothed.f90:76-85
      d.Rm = cutoff - m.skinWidth;
      d.c = d.Rm * d.Rm;
      d.d = 1.0 / d.Rm;
      d.factor = 1.0 / (cutoff - d.Rm);
      break;
    case nb::K_COUL_DAMPED_SQUARE_SMOOTHED:    // src/coul_damped_square_smoothed.f90:76-84
    case nb::K_COUL_SQUARE_SMOOTHED:           // src/coul_square_smoothed.f90:69-76
    case nb::K_COUL_SHIFTED_SQUARE_SMOOTHED:   // src/coul_shifted_square_smoothed.f90:72-79
      d.c = (cutoff - m.skinWidth) * (cutoff - m.skinWidth);
      d.d = 1.0 / (cutoff - m.skinWidth);
      d.factor = 1.0 / (cutoff * cutoff - d.c);
      break;
    default:
      break;
  }
}

// mixing rules: src/pair_lj_cut.f90:122-138, src/pair_softcore_cut.f90:139-163, src/modelClass_pair.f90:194-200
bool mix_rule(const HostModel& self, const HostModel& other, HostModel& out) {
  const double eps = std::sqrt(self.epsilon * other.epsilon), sig = 0.5 * (self.sigma + other.sigma);
  switch (self.dev.kind) {
    case nb::K_PAIR_NONE:
      out = blank(F_PAIR, nb::K_PAIR_NONE, "none");
      return true;
    case nb::K_PAIR_LJ_CUT:
      if (other.dev.kind != nb::K_PAIR_LJ_CUT) return false;
      out = make_lj(eps, sig);
      return true;
    case nb::K_PAIR_SOFTCORE_CUT:
      if (other.dev.kind == nb::K_PAIR_SOFTCORE_CUT) out = make_softcore(eps, sig, self.lambda * other.lambda);
      else if (other.dev.kind == nb::K_PAIR_LJ_CUT) out = make_softcore(eps, sig, self.lambda);
      else return false;
      return true;
    default:
      return false;
  }
}

struct PairSlot {   // reference pairContainer
  HostModel model = blank(F_PAIR, nb::K_PAIR_NONE, "none");
  bool coulomb = false;
  double kCoul = 0.0;
};

PairSlot mix_slots(const PairSlot& a, const PairSlot& b) {   // src/modelClass_pair.f90:120-140
  PairSlot c;
  if (!mix_rule(b.model, a.model, c.model) && !mix_rule(a.model, b.model, c.model)) {
    c.model = blank(F_PAIR, nb::K_PAIR_NONE, "none");
    warning(std::string("no mixing rule found for models ") + a.model.name + " and " + b.model.name);
  }
  c.coulomb = a.coulomb && b.coulomb;
  if (c.coulomb) c.kCoul = std::sqrt(a.kCoul * b.kCoul);
  return c;
}

// ---- KISS + ziggurat (src/math.f90:42-177): needed by EmDee_random_momenta --------------------------
struct Rng {
  bool seeding_required = true;
  int32_t kn[128];
  double wn[128], fn[128];
  uint32_t x = 0, y = 0, z = 0, w = 0;
  static uint32_t mixbits(uint32_t k, int n) { return k ^ (n > 0 ? k << n : k >> -n); }
  static uint32_t scramble(uint32_t v) { return mixbits(mixbits(mixbits(v, 13), -17), 5); }
  void seed(int32_t s) {
    x = scramble((uint32_t)s);
    y = scramble(x);
    z = scramble(y);
    w = scramble(z);
    seeding_required = false;
    const double m1 = 2147483648.0, vn = 0.00991256303526217;
    double dn = 3.442619855899, tn = dn;
    const double q = vn * std::exp(0.5 * dn * dn);
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
  int32_t next_i32() {
    x = 69069u * x + 1327217885u;
    y ^= y << 13;
    y ^= y >> 17;
    y ^= y << 5;
    z = 18000u * (z & 65535u) + (z >> 16);
    w = 30903u * (w & 65535u) + (w >> 16);
    return (int32_t)(x + y + (z << 16) + w);
  }
  static bool inside(int32_t hz, int32_t bound) {   // abs(hz) < bound with 32-bit wraparound abs
    int32_t a = hz < 0 ? (int32_t)(0u - (uint32_t)hz) : hz;
    return a < bound;
  }
  double uni() { return 0.2328306e-9 * next_i32() + 0.5; }
  double normal() {
    int32_t hz = next_i32();
    int iz = hz & 127;
    if (inside(hz, kn[iz])) return hz * wn[iz];
    for (;;) {
      if (iz == 0) {
        double a, b;
        do {
          a = -0.2904764 * std::log(uni());
          b = -std::log(uni());
        } while (b + b < a * a);
        return hz <= 0 ? -(3.442620 + a) : 3.442620 + a;
      }
      const double v = hz * wn[iz];
      if (fn[iz] + uni() * (fn[iz - 1] - fn[iz]) < std::exp(-0.5 * v * v)) return v;
      hz = next_i32();
      iz = hz & 127;
      if (inside(hz, kn[iz])) return hz * wn[iz];
    }
  }
};

double phi(double x) {   // src/math.f90:230-237
  return std::fabs(x) > 1e-4 ? (1.0 - std::exp(-x)) / x : 1.0 + 0.5 * x * ((1.0 / 3.0) * x * (1.0 - 0.25 * x) - 1.0);
}

double inverse_of_x_plus_ln_x(double y) {   // src/math.f90:642-657
  double x = y > 0.5671432904097839 ? y - std::log(y) : std::exp(y);
  double x0 = x + 1.0;
  while (std::fabs(x - x0) > 1.0e-12 * x0) {
    x0 = x;
    x = x * (y + 1.0 - std::log(x)) / (x + 1.0);
  }
  return x;
}

// ---- system -------------------------------------------------------------------------------------
struct RigidBody {   // the part of reference tBody (src/ArBee.f90:29-106) the hot path reads
  std::vector<int> atoms;   // 0-based
  std::vector<double> m;
  double mass = 0.0;
};

struct System {
  int N = 0, ntypes = 1, nlayers = 1, layer = 1, threads = 1;
  double Rc = 0, skin = 0, InRc = 0, totalMass = 0, startTime = 0;
  std::shared_ptr<double> box = std::make_shared<double>(0.0);   // me%Lbox is a pointer in the reference: shared by EmDee_share_phase_space
  bool hasL = false, hasR = false, initialized = false, kspace_active = false;
  std::vector<int> type;                 // 1-based type of each atom
  std::vector<double> mass, invMass, charge;
  std::vector<char> charged;
  std::vector<int> atomBody, freeAtoms;
  std::vector<RigidBody> bodies;
  std::vector<std::vector<int>> excluded;   // per atom, sorted unique 1-based partners
  bool exclusions_dirty = true;
  std::vector<double> ewaldRigid;                // per layer: sum over type pairs of kCoul * (intrabody erf terms + self energy)
  std::vector<emdee::BondedTerm> bondedTerms;   // bonds and angles in the order they were added (reference structList)
  int ndihedrals = 0;
  bool bonded_dirty = false;
  std::vector<PairSlot> pair;            // (ntypes, ntypes, nlayers)
  std::vector<HostModel> coul;           // per layer
  HostModel kspace = blank(F_KSPACE, 0, "ewald");
  // bonded models keep their two parameters in dev.a, dev.b (k, r0 | k, theta0); dev.kind: 0 = none, 1 = harmonic
  std::vector<char> overridable, multilayer, interact, pairs_exist, useInRc, bonded, forcesUpToDate;
  std::vector<tEnergy> layerEnergy;
  std::vector<tVirial> layerVirial;
  Rng random;
  Engine* engine = nullptr;

  PairSlot& slot(int i, int j, int l) { return pair[((size_t)(l - 1) * ntypes + (j - 1)) * ntypes + (i - 1)]; }
  char& flag(std::vector<char>& a, int i, int j) { return a[(size_t)(j - 1) * ntypes + (i - 1)]; }
  double layerRc(int l) const { return useInRc[l - 1] ? InRc : Rc; }
  int nbodies() const { return (int)bodies.size(); }
};

System* sys(const tEmDee& md) { return static_cast<System*>(md.Data); }

std::string opt(const char* s) {   // src/global.f90:120-128
  size_t n = 0;
  while (n < 256 && s[n] != '\0') ++n;
  return std::string(s, n);
}
bool ranged(std::initializer_list<int> v, int imax) {
  for (int i : v)
    if (i <= 0 || i > imax) return false;
  return true;
}

// src/EmDeeData.f90:227-264
void set_pair_type(System& me, int itype, int jtype, int layer, const HostModel& model, double kCoul) {
  const double cutoff = me.layerRc(layer);
  auto assign = [&](PairSlot& s) {
    s.model = model;            // container assignment copies the model only
    s.coulomb = kCoul != 0.0;
    if (s.coulomb) s.kCoul = kCoul;
    modifier_setup(s.model, cutoff);
  };
  if (itype == jtype) {
    assign(me.slot(itype, itype, layer));
    for (int k = 1; k <= me.ntypes; ++k) {
      if (k == itype || !me.flag(me.overridable, itype, k)) continue;
      PairSlot mixed = mix_slots(me.slot(k, k, layer), me.slot(itype, itype, layer));
      modifier_setup(mixed.model, cutoff);
      me.slot(itype, k, layer) = mixed;
      me.slot(k, itype, layer) = mixed;
    }
  } else {
    assign(me.slot(itype, jtype, layer));
    me.slot(jtype, itype, layer) = me.slot(itype, jtype, layer);
  }
}

// src/EmDeeData.f90:268-351 (allocate_rigid_bodies + clean_body_indices)
void setup_bodies(System& me, const int* bodies) {
  const int N = me.N;
  me.atomBody.assign(N, 0);
  me.bodies.clear();
  me.freeAtoms.clear();
  if (bodies == nullptr) {
    for (int i = 0; i < N; ++i) {
      me.atomBody[i] = i + 1;
      me.freeAtoms.push_back(i);
    }
    return;
  }
  // ids that occur more than once become bodies, numbered by first appearance
  // (the reference searches its list of ids linearly for every atom, O(N x bodies); a hash map gives the same numbering)
  std::vector<int> ids, count, which(N, -1);
  std::unordered_map<int, int> slotOf;
  slotOf.reserve((size_t)N);
  for (int i = 0; i < N; ++i) {
    if (bodies[i] <= 0) continue;
    auto it = slotOf.find(bodies[i]);
    int k;
    if (it == slotOf.end()) {
      k = (int)ids.size();
      slotOf.emplace(bodies[i], k);
      ids.push_back(bodies[i]);
      count.push_back(1);
    } else {
      k = it->second;
      count[k] += 1;
    }
    which[i] = k;
  }
  std::vector<int> bodyIndex(ids.size(), 0);
  int nb_ = 0;
  for (size_t k = 0; k < ids.size(); ++k)
    if (count[k] > 1) bodyIndex[k] = ++nb_;
  me.bodies.resize(nb_);
  for (int i = 0; i < N; ++i) {
    int b = which[i] >= 0 ? bodyIndex[which[i]] : 0;
    me.atomBody[i] = b;
    if (b > 0) {
      me.bodies[b - 1].atoms.push_back(i);
      me.bodies[b - 1].m.push_back(me.mass[i]);
      me.bodies[b - 1].mass += me.mass[i];
    } else {
      me.freeAtoms.push_back(i);
    }
  }
  int next = nb_;
  for (int i = 0; i < N; ++i)
    if (me.atomBody[i] == 0) me.atomBody[i] = ++next;
}

// src/EmDeeData.f90:193-223
void check_actual_interactions(System& me) {
  const int nt = me.ntypes;
  std::vector<char> neutral(nt, 1);
  for (int a = 0; a < me.N; ++a)
    if (me.charged[a]) neutral[me.type[a] - 1] = 0;
  std::fill(me.pairs_exist.begin(), me.pairs_exist.end(), 0);
  for (int i = 1; i <= nt; ++i)
    for (int j = 1; j <= i; ++j) {
      bool any = false;
      for (int k = 1; k <= me.nlayers; ++k) {
        const PairSlot& p = me.slot(i, j, k);
        const bool no_pair = p.model.dev.kind == nb::K_PAIR_NONE;
        const bool no_coul = me.coul[k - 1].dev.kind == nb::K_COUL_NONE || !p.coulomb;
        const bool inert = (no_pair && no_coul) || (no_pair && !no_coul && neutral[i - 1] && neutral[j - 1]);
        if (!inert) {
          any = true;
          me.pairs_exist[k - 1] = 1;
        }
      }
      me.flag(me.interact, i, j) = any;
      me.flag(me.interact, j, i) = any;
    }
}

void push_tables(System& me) {
  std::vector<char> inter0((size_t)me.ntypes * me.ntypes);
  for (int i = 0; i < me.ntypes; ++i)
    for (int j = 0; j < me.ntypes; ++j) inter0[(size_t)i * me.ntypes + j] = me.flag(me.interact, i + 1, j + 1);
  me.engine->set_interact(inter0);
  for (int l = 1; l <= me.nlayers; ++l) {
    emdee::LayerTable t;
    t.pair.resize((size_t)me.ntypes * me.ntypes);
    for (int i = 1; i <= me.ntypes; ++i)
      for (int j = 1; j <= me.ntypes; ++j) {
        const PairSlot& s = me.slot(i, j, l);
        emdee::PairEntry e;
        e.model = s.model.dev;
        e.kCoul = s.kCoul;
        e.coulomb = s.coulomb ? 1 : 0;
        e.pad = 0;
        t.pair[(size_t)(i - 1) * me.ntypes + (j - 1)] = e;
      }
    t.coul = me.coul[l - 1].dev;
    t.pairs_exist = me.pairs_exist[l - 1];
    t.useInRc = me.useInRc[l - 1];
    me.engine->set_layer(l - 1, t);
  }
}

void push_exclusions(System& me) {
  std::vector<int> first(me.N), last(me.N), item;
  for (int i = 0; i < me.N; ++i) {
    first[i] = (int)item.size() + 1;
    item.insert(item.end(), me.excluded[i].begin(), me.excluded[i].end());
    last[i] = (int)item.size();
  }
  me.engine->set_exclusions(first, last, item);
  me.exclusions_dirty = false;
}

// cKspaceModel_initialize (src/modelClass_kspace.f90:108-273) and kspace_ewald_set_parameters / kspace_ewald_update
// (src/kspace_ewald.f90:82-184): everything about the reciprocal-space solver that is fixed at initialization -- the
// splitting parameter, the half space of wave vectors with their prefactors, the charged types, the Coulomb constants
// of the type pairs, and the constant energy that takes the smooth part of intrabody pairs and the self term back out.
// The per-step sums run on the device (engine_ewald.cuh). Like the reference, the wave vectors are NOT rebuilt when
// the box changes later.
void setup_ewald(System& me, double Rc) {
  const char* task = "kspace model initialization";
  const double PI = 3.14159265358979324, L = *me.box;
  bool any = false;
  std::vector<char> hasCharge(me.ntypes + 1, 0);
  for (int a = 0; a < me.N; ++a)
    if (me.charged[a]) { any = true; hasCharge[me.type[a]] = 1; }
  if (!any) error(task, "system has no charged atoms");
  std::vector<int> local(me.ntypes + 1, -1), kinds;
  for (int t = 1; t <= me.ntypes; ++t)
    if (hasCharge[t]) { local[t] = (int)kinds.size(); kinds.push_back(t); }
  const int ntk = (int)kinds.size();
  emdee::EwaldSetup e;
  e.ntk = ntk;
  e.atomKType.resize(me.N);
  for (int a = 0; a < me.N; ++a) e.atomKType[a] = local[me.type[a]];
  const double s = std::sqrt(inverse_of_x_plus_ln_x(-std::log(me.kspace.accuracy)));   // accuracy = exp(-s^2)/s^2
  e.alpha = s / Rc;
  e.beta = 2.0 * e.alpha / std::sqrt(PI);
  const double kmax = 2.0 * e.alpha * s;
  std::printf("KSPACE PARAMETERS: alpha = %10.5f and kmax = %10.5f\n", e.alpha, kmax);
  for (HostModel& c : me.coul)
    if (c.requires_kspace) {   // src/coul_long.f90:66-73
      c.dev.a = e.alpha;
      c.dev.b = e.beta;
    }
  // wave vectors: upper half of the (2 nmax + 1)^3 lattice (third index fastest), inside the sphere of radius nmax * 2pi/L
  const double unit = 2.0 * PI / L, unitSq = unit * unit;
  const int nmax = (int)std::ceil(kmax / unit), M = 2 * nmax + 1, half = (M * M * M) / 2;
  const double kmaxSq = unitSq * nmax * nmax, B = -0.25 / (e.alpha * e.alpha), fourPiByV = 4.0 * PI / (L * L * L);
  for (int i = half + 1; i <= 2 * half; ++i) {
    const int n1 = i / (M * M), rem = i - n1 * M * M, n2 = rem / M, n3 = rem - n2 * M;
    const int kv[3] = {n1 - nmax, n2 - nmax, n3 - nmax};
    if (unitSq * kv[0] * kv[0] + unitSq * kv[1] * kv[1] + unitSq * kv[2] * kv[2] > kmaxSq) continue;
    const double k[3] = {unit * kv[0], unit * kv[1], unit * kv[2]};
    const double ksq = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
    e.n.insert(e.n.end(), kv, kv + 3);
    e.prefac.push_back(fourPiByV * std::exp(ksq * B) / ksq);
  }
  // Coulomb constants of the charged type pairs, per layer
  e.lambda.assign((size_t)me.nlayers * ntk * ntk, 0.0);
  for (int l = 1; l <= me.nlayers; ++l)
    for (int i = 0; i < ntk; ++i)
      for (int j = 0; j < ntk; ++j) {
        const int lo = std::min(i, j), hi = std::max(i, j);   // the reference reads kCoul(itype, jtype) with i <= j
        e.lambda[((size_t)(l - 1) * ntk + i) * ntk + j] = me.slot(kinds[lo], kinds[hi], l).kCoul;
      }
  // constant energy: -erf(alpha r)/r of every intrabody charged pair and -alpha q^2/sqrt(pi) of every listed atom,
  // grouped by type pair so that each layer can weight them with its own Coulomb constants
  std::vector<double> Erigid((size_t)ntk * ntk, 0.0);   // [lo*ntk + hi]
  const double invL = 1.0 / L;
  std::vector<double> hostR;
  if (!me.bodies.empty()) {
    hostR.resize(3 * (size_t)me.N);
    me.engine->download_coordinates(hostR.data());
  }
  for (const RigidBody& b : me.bodies)
    for (size_t x = 0; x + 1 < b.atoms.size(); ++x) {
      const int i = b.atoms[x];
      if (!me.charged[i]) continue;
      for (size_t y = x + 1; y < b.atoms.size(); ++y) {
        const int j = b.atoms[y];
        if (!me.charged[j]) continue;
        double rsq = 0.0;
        for (int c = 0; c < 3; ++c) {
          double d = hostR[3 * (size_t)i + c] - hostR[3 * (size_t)j + c];
          d -= L * std::round(invL * d);
          rsq += d * d;
        }
        const double r = std::sqrt(rsq), xx = e.alpha * r;
        const double Eij = -(me.charge[i] * me.charge[j]) * (1.0 - nb::uerfc(xx, std::exp(-xx * xx))) / r;
        const int ti = local[me.type[i]], tj = local[me.type[j]];
        Erigid[(size_t)std::min(ti, tj) * ntk + std::max(ti, tj)] += Eij;
      }
    }
  for (int a = 0; a < me.N; ++a)
    if (local[me.type[a]] >= 0) {
      const int t = local[me.type[a]];
      Erigid[(size_t)t * ntk + t] -= e.alpha * (me.charge[a] * me.charge[a]) / std::sqrt(PI);
    }
  me.ewaldRigid.assign(me.nlayers, 0.0);
  for (int l = 1; l <= me.nlayers; ++l)
    for (int i = 0; i < ntk; ++i)
      for (int j = i; j < ntk; ++j)
        me.ewaldRigid[l - 1] += me.slot(kinds[i], kinds[j], l).kCoul * Erigid[(size_t)i * ntk + j];
  me.engine->set_ewald(e);
}

// src/EmDeeData.f90:359-416
void perform_initialization(System& me, tEmDee* md) {
  const char* task = "system initialization";
  // update_rigid_bodies (src/EmDeeData.f90:420-439) + tBody_update: unwrapping, centres of mass, member offsets,
  // principal frames and quaternions, all on the device, in place on the coordinates uploaded earlier
  if (me.nbodies() != 0) me.engine->update_body_frames(*me.box);
  const int bodyDoF = 6 * me.nbodies();
  md->RotDoF = bodyDoF - 3 * me.nbodies();
  md->DoF = 3 * (int)me.freeAtoms.size() + bodyDoF - 3;
  check_actual_interactions(me);
  bool required = false;
  double kspaceRc = 0.0;
  for (int l = 1; l <= me.nlayers; ++l)
    if (me.coul[l - 1].requires_kspace) {
      if (!required) kspaceRc = me.layerRc(l);
      else if (me.layerRc(l) != kspaceRc)
        error(task, "all layers with ewald-like coulomb models must have the same cutoff");
      required = true;
    }
  if (required != me.kspace_active) {
    if (me.kspace_active) me.kspace_active = false;
    else error(task, "a kspace solver is required, but has not been defined");
  }
  if (me.kspace_active) setup_ewald(me, kspaceRc);
  me.engine->set_charges(me.charge.data());
  push_tables(me);
  push_exclusions(me);
  me.initialized = true;
}

void invalidate(System& me, tEmDee* md) {
  std::fill(me.forcesUpToDate.begin(), me.forcesUpToDate.end(), 0);
  for (tEnergy& e : me.layerEnergy) e.UpToDate = false;
  md->Energy.UpToDate = false;
}

void set_kinetic(tEmDee* md, const double twoKE[3], const double* twoKEr = nullptr) {
  for (int x = 0; x < 3; ++x) {
    md->Kinetic.TransPart[x] = 0.5 * twoKE[x];
    md->Kinetic.RotPart[x] = twoKEr ? 0.5 * twoKEr[x] : 0.0;
  }
  md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
  md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
  md->Kinetic.ShadowKinetic = md->Kinetic.Total;
  md->Kinetic.ShadowRotational = md->Kinetic.Rotational;
  md->Kinetic.UpToDate = true;
}

void* with_modifier(void* model, int modifier, double skin, bool set_skin, const char* task) {
  HostModel* m = handle(model);
  if (m == nullptr || (m->family != F_PAIR && m->family != F_COUL)) error(task, "a valid pair model must be provided");
  HostModel n = *m;
  n.dev.modifier = modifier;
  if (set_skin) n.skin = skin;
  return deliver(n);
}

}  // namespace

// =================================================================================================
extern "C" {

const char* EmDeeX_backend(void) { return "b200-cuda"; }

tEmDee EmDee_system(int threads, int layers, double rc, double skin, int N, int* types, double* masses,
                    int* bodies) {
  if (std::getenv("EMDEE_QUIET") == nullptr) std::printf("EmDee (version: %11s)\n", "15 Oct 2018");
  System* me = new System();
  me->threads = threads;   // host threads: kept for ABI compatibility, the device parallelises the work
  me->nlayers = layers;
  me->Rc = rc;
  me->skin = skin;
  me->InRc = rc;
  me->N = N;
  if (types != nullptr) {
    int lo = types[0], hi = types[0];
    for (int i = 1; i < N; ++i) {
      lo = types[i] < lo ? types[i] : lo;
      hi = types[i] > hi ? types[i] : hi;
    }
    if (lo != 1) error("system setup", "wrong specification of atom types");
    me->ntypes = hi;
    me->type.assign(types, types + N);
  } else {
    me->ntypes = 1;
    me->type.assign(N, 1);
  }
  me->mass.assign(N, 1.0);
  me->invMass.assign(N, 1.0);
  me->totalMass = (double)N;
  if (masses != nullptr) {
    me->totalMass = 0.0;
    for (int i = 0; i < N; ++i) {
      me->mass[i] = masses[me->type[i] - 1];
      me->invMass[i] = 1.0 / masses[me->type[i] - 1];
      me->totalMass += me->mass[i];
    }
  }
  me->startTime = now();
  me->charge.assign(N, 0.0);
  me->charged.assign(N, 0);
  setup_bodies(*me, bodies);
  me->excluded.assign(N, {});
  const size_t nt2 = (size_t)me->ntypes * me->ntypes;
  me->pair.assign(nt2 * layers, PairSlot());
  me->coul.assign(layers, blank(F_COUL, nb::K_COUL_NONE, "none"));
  me->overridable.assign(nt2, 1);
  me->multilayer.assign(nt2, 0);
  me->interact.assign(nt2, 0);
  me->pairs_exist.assign(layers, 0);
  me->useInRc.assign(layers, 0);
  me->bonded.assign(layers, 1);
  me->forcesUpToDate.assign(layers, 0);
  me->engine = new Engine(N, me->ntypes, layers, rc, skin, me->type.data(), me->mass.data(), me->invMass.data(),
                          me->atomBody.data(), me->nbodies());
  if (me->nbodies() != 0) {
    std::vector<int> first(1, 0), members;
    std::vector<double> memberMass;
    for (const RigidBody& b : me->bodies) {
      members.insert(members.end(), b.atoms.begin(), b.atoms.end());
      memberMass.insert(memberMass.end(), b.m.begin(), b.m.end());
      first.push_back((int)members.size());
    }
    me->engine->set_bodies(first, members, memberMass);
  }

  tEmDee md;
  std::memset(&md, 0, sizeof(md));
  me->layerEnergy.assign(layers, md.Energy);
  me->layerVirial.assign(layers, md.Virial);
  md.DoF = 3 * (N - 1);
  md.Data = me;
  md.Options.Translate = true;
  md.Options.Rotate = true;
  md.Options.RotationMode = 0;
  md.Options.AutoBodyUpdate = true;
  md.Options.Compute = true;
  return md;
}

// src/EmDeeCode.f90:212-235. The returned pointer addresses the library's own array (moved into CUDA managed memory
// on first request, see Engine::expose): valid on the host between library calls, for reading and for writing.
void* EmDee_memory_address(tEmDee md, const char* option) {
  System* me = sys(md);
  const std::string item = opt(option);
  if (item == "coordinates") {
    if (!me->hasR) error("memory address retrieving", "coordinates have not been allocated");   // unassociated pointer in the reference
    return me->engine->expose(Engine::EXPOSE_R, 0);
  }
  if (item == "momenta") return me->engine->expose(Engine::EXPOSE_P, 0);
  if (item == "forces") return me->engine->expose(Engine::EXPOSE_F, me->layer - 1);
  if (item == "layerForces") return me->engine->expose(Engine::EXPOSE_LAYER_F, 0);
  error("memory address retrieving", "invalid option " + item);
}

// src/EmDeeCode.f90:239-269
void EmDee_share_phase_space(tEmDee mdkeep, tEmDee* mdlose) {
  const char* task = "phase space sharing";
  System* keep = sys(mdkeep);
  System* lose = sys(*mdlose);
  if (!(keep->initialized && lose->initialized)) error(task, "EmDee system 1 has not been initialized");
  if (keep->N != lose->N) error(task, "different numbers of atoms");
  if (keep->type != lose->type) error(task, "atom types do not match");
  if (keep->mass != lose->mass) error(task, "atom masses do not match");
  if (keep->atomBody != lose->atomBody) error(task, "rigid bodies do not match");
  if (keep != lose) {
    lose->engine->share_phase_space(*keep->engine);
    lose->box = keep->box;
  }
  mdlose->Kinetic.Total = mdkeep.Kinetic.Total;
  mdlose->Kinetic.Rotational = mdkeep.Kinetic.Rotational;
  for (int x = 0; x < 3; ++x) {
    mdlose->Kinetic.TransPart[x] = mdkeep.Kinetic.TransPart[x];
    mdlose->Kinetic.RotPart[x] = mdkeep.Kinetic.RotPart[x];
  }
}

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
  me->engine->set_inner_cutoff(InternalRc);
  for (int l = 0; l < me->nlayers; ++l) me->bonded[l] = Bonded[l] != 0;
  for (int l = 1; l <= me->nlayers; ++l) cutoff_setup(me->coul[l - 1], me->layerRc(l));   // cutoff_setup ONLY (Q1)
}

static void after_pair_setting(System* me, int itype, int jtype, char multi) {
  me->flag(me->multilayer, itype, jtype) = multi;
  me->flag(me->multilayer, jtype, itype) = multi;
  if (itype == jtype) {
    for (int k = 1; k <= me->ntypes; ++k)
      if (k != itype && me->flag(me->overridable, itype, k)) {
        char v = multi ? 1 : me->flag(me->multilayer, k, k);
        me->flag(me->multilayer, itype, k) = v;
        me->flag(me->multilayer, k, itype) = v;
      }
  } else {
    me->flag(me->overridable, itype, jtype) = 0;
    me->flag(me->overridable, jtype, itype) = 0;
  }
}

void EmDee_set_pair_model(tEmDee md, int itype, int jtype, void* model, double kCoul) {
  const char* task = "pair model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  if (!ranged({itype, jtype}, me->ntypes)) error(task, "provided type index is out of range");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_PAIR) error(task, "a valid pair model must be provided");
  for (int layer = 1; layer <= me->nlayers; ++layer) set_pair_type(*me, itype, jtype, layer, *m, kCoul);
  after_pair_setting(me, itype, jtype, 0);
}

void EmDee_set_pair_multimodel(tEmDee md, int itype, int jtype, void* model[], double kCoul[]) {
  const char* task = "pair multimodel setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  if (!ranged({itype, jtype}, me->ntypes)) error(task, "provided type index is out of range");
  for (int layer = 1; layer <= me->nlayers; ++layer) {
    HostModel* m = handle(model[layer - 1]);
    if (m == nullptr) error(task, std::to_string(me->nlayers) + " valid pair models must be provided");
    if (m->family != F_PAIR) error(task, "a valid pair model must be provided");
    set_pair_type(*me, itype, jtype, layer, *m, kCoul[layer - 1]);
  }
  after_pair_setting(me, itype, jtype, 1);
}

void EmDee_set_kspace_model(tEmDee md, void* model) {
  const char* task = "kspace model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_KSPACE) error(task, "a valid kspace model must be provided");
  me->kspace = *m;
  me->kspace_active = true;
}

static void set_coul_layer(System* me, int layer, const HostModel& m) {
  const double cutoff = me->layerRc(layer);
  me->coul[layer - 1] = m;
  cutoff_setup(me->coul[layer - 1], cutoff);
  modifier_setup(me->coul[layer - 1], cutoff);   // literal call order of src/EmDeeCode.f90:474-475 (Q1, Q1b)
}

void EmDee_set_coul_model(tEmDee md, void* model) {
  const char* task = "coulomb model setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_COUL) error(task, "a valid coulomb model must be provided");
  for (int layer = 1; layer <= me->nlayers; ++layer) set_coul_layer(me, layer, *m);
}

void EmDee_set_coul_multimodel(tEmDee md, void* model[]) {
  const char* task = "coulomb multimodel setup";
  System* me = sys(md);
  if (me->initialized) error(task, "cannot set model after coordinates have been defined");
  for (int layer = 1; layer <= me->nlayers; ++layer) {
    HostModel* m = handle(model[layer - 1]);
    if (m == nullptr) error(task, std::to_string(me->nlayers) + " valid coulomb models must be provided");
    if (m->family != F_COUL) error(task, "a valid coulomb model must be provided");
    set_coul_layer(me, layer, *m);
  }
}

void EmDee_ignore_pair(tEmDee md, int i, int j) {   // src/EmDeeCode.f90:524-570
  System* me = sys(md);
  if (i == j || !ranged({i, j}, me->N)) return;
  auto insert = [&](int a, int b) {
    std::vector<int>& row = me->excluded[a - 1];
    size_t pos = 0;
    while (pos < row.size() && row[pos] < b) ++pos;
    if (pos < row.size() && row[pos] == b) return;
    row.insert(row.begin() + pos, b);
    me->exclusions_dirty = true;
  };
  insert(i, j);
  insert(j, i);
}

// src/EmDeeCode.f90:574-655: bonded structures are excluded from the neighbor lists as they are added
void EmDee_add_bond(tEmDee md, int i, int j, void* model) {
  System* me = sys(md);
  if (!ranged({i, j}, me->N)) error("add_bond", "atom index out of range");
  if (model == nullptr) error("add_bond", "a valid model must be provided");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_BOND) error("add_bond", "the provided model must be a bond model");
  me->bondedTerms.push_back({i - 1, j - 1, 0, m->dev.kind == 0 ? 0 : 1, m->dev.a, m->dev.b});
  me->bonded_dirty = true;
  EmDee_ignore_pair(md, i, j);
}
void EmDee_add_angle(tEmDee md, int i, int j, int k, void* model) {
  System* me = sys(md);
  if (!ranged({i, j, k}, me->N)) error("add_angle", "atom index out of range");
  if (model == nullptr) error("add_angle", "a valid model must be provided");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_ANGLE) error("add_angle", "the provided model must be an angle model");
  me->bondedTerms.push_back({i - 1, j - 1, k - 1, m->dev.kind == 0 ? 2 : 3, m->dev.a, m->dev.b});
  me->bonded_dirty = true;
  EmDee_ignore_pair(md, i, j);
  EmDee_ignore_pair(md, i, k);
  EmDee_ignore_pair(md, j, k);
}
void EmDee_add_dihedral(tEmDee md, int i, int j, int k, int l, void* model) {
  System* me = sys(md);
  if (!ranged({i, j, k, l}, me->N)) error("add_dihedral", "atom index out of range");
  if (model == nullptr) error("add_dihedral", "a valid model must be provided");
  HostModel* m = handle(model);
  if (m == nullptr || m->family != F_DIHEDRAL) error("add_dihedral", "the provided model must be a dihedral model");
  // The reference stores dihedrals but EmDee_compute_forces never evaluates them (src/EmDeeCode.f90:1240-1241 call
  // compute_bonds and compute_angles only) and dihedral_none is the only model: the exclusions are the whole effect.
  me->ndihedrals += 1;
  const int a[4] = {i, j, k, l};
  for (int x = 0; x < 4; ++x)
    for (int y = x + 1; y < 4; ++y) EmDee_ignore_pair(md, a[x], a[y]);
}

void EmDee_compute_forces(tEmDee* md);

void EmDee_download(tEmDee md, const char* option, double* address) {
  System* me = sys(md);
  const std::string item = opt(option);
  if (address == nullptr) error("download", "provided address is invalid");
  if (item == "box") {
    *address = *me->box;
  } else if (item == "coordinates") {
    if (!me->hasR) error("download", "coordinates have not been allocated");
    me->engine->download_coordinates(address);
  } else if (item == "momenta") {
    me->engine->refresh_member_momenta();   // body members: m_k (pcm/M + omega x delta_k), src/ArBee.f90:317-326
    me->engine->download_momenta(address);
  } else if (item == "forces") {
    if (!me->forcesUpToDate[me->layer - 1]) EmDee_compute_forces(&md);
    me->engine->download_forces(me->layer - 1, address);
  } else if (item == "centersOfMass") {   // bodies first, then the free atoms (src/EmDeeCode.f90:748-757)
    const size_t nb = (size_t)me->nbodies();
    me->engine->download_body(Engine::BODY_RCM, address);
    std::vector<double> R(3 * (size_t)me->N);
    me->engine->download_coordinates(R.data());
    for (size_t f = 0; f < me->freeAtoms.size(); ++f)
      for (int x = 0; x < 3; ++x) address[3 * (nb + f) + x] = R[3 * (size_t)me->freeAtoms[f] + x];
  } else if (item == "quaternions") {
    me->engine->download_body(Engine::BODY_QUATERNION, address);
  } else if (item == "quatmom") {
    me->engine->download_body(Engine::BODY_QUATMOM, address);
  } else if (item == "quattau") {   // C(q) (2 tau), src/EmDeeCode.f90:770-771
    const size_t nb = (size_t)me->nbodies();
    std::vector<double> q(4 * nb), tau(3 * nb);
    me->engine->download_body(Engine::BODY_QUATERNION, q.data());
    me->engine->download_body(Engine::BODY_TORQUE, tau.data());
    for (size_t b = 0; b < nb; ++b) {
      const double* a = &q[4 * b];
      const double v[3] = {2.0 * tau[3 * b], 2.0 * tau[3 * b + 1], 2.0 * tau[3 * b + 2]};
      double* o = address + 4 * b;
      o[0] = -a[1] * v[0] - a[2] * v[1] - a[3] * v[2];
      o[1] = a[0] * v[0] + a[3] * v[1] - a[2] * v[2];
      o[2] = -a[3] * v[0] + a[0] * v[1] + a[1] * v[2];
      o[3] = a[2] * v[0] - a[1] * v[1] + a[0] * v[2];
    }
  } else if (item == "angmom") {
    me->engine->download_body(Engine::BODY_OMEGA, address);
  } else if (item == "bodycoord") {
    me->engine->download_body(Engine::BODY_RCM, address);
  } else if (item == "bodymom") {
    me->engine->download_body(Engine::BODY_PCM, address);
  } else if (item == "bodyforces") {
    me->engine->download_body(Engine::BODY_FORCE, address);
  } else if (item == "torques") {
    me->engine->download_body(Engine::BODY_TORQUE, address);
  } else if (item == "inertia") {
    me->engine->download_body(Engine::BODY_INERTIA, address);
  } else {
    error("download", "invalid option");
  }
}

void EmDee_switch_model_layer(tEmDee* md, int layer) {   // src/EmDeeCode.f90:929-946
  System* me = sys(*md);
  if (layer == me->layer) return;
  if (layer < 1 || layer > me->nlayers) error("model layer switch", "selected layer is out of range");
  me->layer = layer;
  md->Energy = me->layerEnergy[layer - 1];
  md->Virial = me->layerVirial[layer - 1];
}

void EmDee_upload(tEmDee* md, const char* option, double* address) {   // src/EmDeeCode.f90:805-925
  System* me = sys(*md);
  const std::string item = opt(option);
  if (address == nullptr) error("upload", "provided address is invalid");
  auto initialize_system = [&]() {
    perform_initialization(*me, md);
    for (int layer = me->nlayers; layer >= 1; --layer) {
      EmDee_switch_model_layer(md, layer);
      EmDee_compute_forces(md);
    }
  };
  if (item == "box") {
    me->hasL = true;
    *me->box = *address;
    if (me->initialized) invalidate(*me, md);
    else if (me->hasR) initialize_system();
  } else if (item == "coordinates") {
    me->hasR = true;
    me->engine->upload_coordinates(address);
    if (me->initialized) {
      invalidate(*me, md);
      if (md->Options.AutoBodyUpdate && me->nbodies() != 0) me->engine->update_body_frames(*me->box);
    } else if (me->hasL) {
      initialize_system();
    }
  } else if (item == "momenta") {
    if (!me->initialized) error("upload", "box and coordinates have not been defined");
    me->engine->upload_momenta(address);
    if (me->nbodies() == 0) {
      // body-less systems keep the GPU-verified round-1 path: the three sums are formed from the client's buffer
      double twoKE[3] = {0, 0, 0};
      for (int i = 0; i < me->N; ++i)
        for (int x = 0; x < 3; ++x) twoKE[x] += me->invMass[i] * address[3 * (size_t)i + x] * address[3 * (size_t)i + x];
      set_kinetic(md, twoKE);
    } else {
      emdee::KineticAll k;
      me->engine->take_member_momenta(k);   // assign_momenta (src/EmDeeData.f90:157-189) on the device: free atoms and bodies
      set_kinetic(md, k.twoKEt, k.twoKEr);
    }
  } else if (item == "forces") {
    if (!me->initialized) error("upload", "box and coordinates have not been defined");
    me->engine->upload_forces(me->layer - 1, address);
  } else if (item == "charges") {
    if (me->initialized) error("upload", "cannot set charges after box and coordinates initialization");
    for (int i = 0; i < me->N; ++i) {
      me->charge[i] = address[i];
      me->charged[i] = std::fabs(address[i]) > 2.220446049250313e-16;
    }
    invalidate(*me, md);
  } else {
    error("upload", "invalid option");
  }
}

void EmDee_random_momenta(tEmDee* md, double kT, bool adjust, int seed) {   // src/EmDeeCode.f90:950-1020
  System* me = sys(*md);
  if (me->random.seeding_required) me->random.seed(seed);
  // The reference draws from ONE sequential generator (bodies first, then free atoms), so the stream is produced
  // on the host; the per-body algebra (pi = B(q) 2 I omega) runs on the device.
  const size_t nb = (size_t)me->nbodies();
  std::vector<double> P(3 * (size_t)me->N, 0.0), pcm(3 * nb), omega(3 * nb), MoI(3 * nb);
  double twoKE[3] = {0, 0, 0}, twoKEr[3] = {0, 0, 0};
  if (nb != 0) {
    if (!me->initialized) error("random_momenta", "coordinates have not defined");
    me->engine->download_body(Engine::BODY_INERTIA, MoI.data());
    for (size_t b = 0; b < nb; ++b) {
      const double s = std::sqrt(me->bodies[b].mass * kT);
      for (int x = 0; x < 3; ++x) pcm[3 * b + x] = s * me->random.normal();
      for (int x = 0; x < 3; ++x) omega[3 * b + x] = std::sqrt((1.0 / MoI[3 * b + x]) * kT) * me->random.normal();
      for (int x = 0; x < 3; ++x) {
        twoKE[x] += (1.0 / me->bodies[b].mass) * pcm[3 * b + x] * pcm[3 * b + x];
        twoKEr[x] += MoI[3 * b + x] * omega[3 * b + x] * omega[3 * b + x];
      }
    }
  }
  for (int i : me->freeAtoms) {
    const double s = std::sqrt(me->mass[i] * kT);
    for (int x = 0; x < 3; ++x) P[3 * (size_t)i + x] = s * me->random.normal();
    for (int x = 0; x < 3; ++x) twoKE[x] += me->invMass[i] * P[3 * (size_t)i + x] * P[3 * (size_t)i + x];
  }
  if (adjust) {
    double vcm[3] = {0, 0, 0};
    for (int x = 0; x < 3; ++x) {
      double sb = 0.0;
      for (int i : me->freeAtoms) vcm[x] += P[3 * (size_t)i + x];
      for (size_t b = 0; b < nb; ++b) sb += pcm[3 * b + x];
      vcm[x] = (vcm[x] + sb) / me->totalMass;
      twoKE[x] = 0.0;
    }
    for (int i : me->freeAtoms)
      for (int x = 0; x < 3; ++x) {
        P[3 * (size_t)i + x] -= me->mass[i] * vcm[x];
        twoKE[x] += me->invMass[i] * P[3 * (size_t)i + x] * P[3 * (size_t)i + x];
      }
    for (size_t b = 0; b < nb; ++b)
      for (int x = 0; x < 3; ++x) {
        pcm[3 * b + x] -= me->bodies[b].mass * vcm[x];
        twoKE[x] += (1.0 / me->bodies[b].mass) * pcm[3 * b + x] * pcm[3 * b + x];
      }
    const double total = (twoKE[0] + twoKEr[0]) + (twoKE[1] + twoKEr[1]) + (twoKE[2] + twoKEr[2]);
    const double factor = std::sqrt((3 * (int)me->freeAtoms.size() + 6 * (int)nb - 3) * kT / total);
    for (double& p : P) p *= factor;
    for (double& p : pcm) p *= factor;
    for (double& w : omega) w *= factor;
    for (int x = 0; x < 3; ++x) {
      twoKE[x] *= factor * factor;
      twoKEr[x] *= factor * factor;
    }
  }
  me->engine->upload_momenta(P.data());
  if (nb != 0) {
    me->engine->upload_body(Engine::BODY_PCM, pcm.data());
    me->engine->upload_body(Engine::BODY_OMEGA, omega.data());
    me->engine->derive_quaternion_momenta();
  }
  set_kinetic(md, twoKE, twoKEr);
}

void EmDee_boost(tEmDee* md, double lambda, double alpha, double dt) {   // src/EmDeeCode.f90:1024-1065
  System* me = sys(*md);
  double CF = phi(alpha * dt) * dt;
  const double CP = 1.0 - alpha * CF;
  CF = lambda * CF;
  const bool compute = md->Options.Compute;
  if (lambda != 0.0 && !me->forcesUpToDate[me->layer - 1]) {
    // free atoms whose forces come from the pair kernel alone: the kick is launched right behind that kernel and shares
    // its host wait (engine.h: plan_kick); the call of Engine::boost below then only returns the sums
    const bool extras = (!me->bondedTerms.empty() && (me->bonded[me->layer - 1] || me->coul[me->layer - 1].requires_kspace)) ||
                        me->coul[me->layer - 1].requires_kspace;
    if (me->initialized && me->nbodies() == 0 && md->Options.Translate && !extras)
      me->engine->plan_kick(me->layer - 1, CP, CF, compute);
    EmDee_compute_forces(md);
  }
  if (me->nbodies() != 0) {
    emdee::KineticAll ka;
    me->engine->boost_all(me->layer - 1, CP, CF, md->Options.Translate, md->Options.Rotate, compute, ka);
    if (compute) {
      if (md->Options.Translate)
        for (int x = 0; x < 3; ++x) md->Kinetic.TransPart[x] = 0.5 * ka.twoKEt[x];
      if (md->Options.Rotate) {
        for (int x = 0; x < 3; ++x) md->Kinetic.RotPart[x] = 0.5 * ka.twoKEr[x];
        md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
      }
      md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
    }
    md->Kinetic.UpToDate = compute;
    return;
  }
  emdee::KineticScalars ke;
  if (md->Options.Translate) me->engine->boost(me->layer - 1, CP, CF, compute, ke);
  if (compute) {
    if (md->Options.Translate)
      for (int x = 0; x < 3; ++x) md->Kinetic.TransPart[x] = 0.5 * ke.twoKE[x];
    if (md->Options.Rotate) {
      for (int x = 0; x < 3; ++x) md->Kinetic.RotPart[x] = 0.0;
      md->Kinetic.Rotational = 0.0;
    }
    md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
  }
  md->Kinetic.UpToDate = compute;
}

void EmDee_displace(tEmDee* md, double lambda, double alpha, double dt) {   // src/EmDeeCode.f90:1069-1103
  System* me = sys(*md);
  const double t0 = now();
  double CR = 1.0, CP = dt;
  if (alpha != 0.0) {
    CP = phi(alpha * dt) * dt;
    CR = 1.0 - alpha * CP;
    *me->box = CR * *me->box;
  }
  CP = lambda * CP;
  if (me->nbodies() != 0)
    me->engine->move_all(CR, CP, dt, md->Options.Translate, md->Options.Rotate, md->Options.RotationMode);
  else if (md->Options.Translate)
    me->engine->displace(CR, CP);
  invalidate(*me, md);
  md->Time.Motion += now() - t0;
}

// src/EmDeeCode.f90:1107-1211: half kick, drift, forces, half kick -- translation and rotation always on -- plus, when
// Options%Compute is set, the shadow-Hamiltonian corrections (pre_force / post_force bookkeeping on the device)
void EmDee_verlet_step(tEmDee* md, double dt) {
  System* me = sys(*md);
  const double dt_2 = 0.5 * dt;
  const bool compute = md->Options.Compute;
  const int mode = md->Options.RotationMode;
  emdee::KineticAll ka;
  if (compute) me->engine->shadow_pre(me->layer - 1, dt, mode);
  me->engine->boost_all(me->layer - 1, 1.0, dt_2, true, true, false, ka);
  me->engine->move_all(1.0, dt, dt, true, true, mode);
  EmDee_compute_forces(md);
  me->engine->boost_all(me->layer - 1, 1.0, dt_2, true, true, compute, ka);
  if (compute) {
    double Us = 0, Ks_t = 0, Ks_r = 0;
    me->engine->shadow_post(me->layer - 1, dt, mode, Us, Ks_t, Ks_r);
    for (int x = 0; x < 3; ++x) {
      md->Kinetic.TransPart[x] = 0.5 * ka.twoKEt[x];
      md->Kinetic.RotPart[x] = 0.5 * ka.twoKEr[x];
    }
    md->Kinetic.Rotational = md->Kinetic.RotPart[0] + md->Kinetic.RotPart[1] + md->Kinetic.RotPart[2];
    md->Kinetic.Total = md->Kinetic.TransPart[0] + md->Kinetic.TransPart[1] + md->Kinetic.TransPart[2] + md->Kinetic.Rotational;
    md->Energy.ShadowPotential = md->Energy.ShadowPotential - dt * dt * Us / 24.0;
    md->Kinetic.ShadowRotational = Ks_r / (6.0 * dt);
    md->Kinetic.ShadowKinetic = (Ks_t + Ks_r) / (6.0 * dt);
  }
  md->Kinetic.UpToDate = compute;
}

void EmDee_compute_forces(tEmDee* md) {   // src/EmDeeCode.f90:1215-1277
  System* me = sys(*md);
  if (!me->initialized) error("force computation", "box and coordinates have not been defined");
  if (me->exclusions_dirty) push_exclusions(*me);
  const bool compute = md->Options.Compute;
  emdee::ForceScalars r;
  double tn = 0.0;
  const double t0 = now();
  const bool rebuilt = me->engine->compute_forces(me->layer - 1, compute, *me->box, r, tn);
  if (rebuilt) md->Builds += 1;
  emdee::BondedScalars bs;
  const bool kspace = me->coul[me->layer - 1].requires_kspace;
  if (!me->bondedTerms.empty() && (me->bonded[me->layer - 1] || kspace)) {   // compute_bonds / compute_angles, src/EmDeeData.f90:443-550
    if (me->bonded_dirty) {
      me->engine->set_bonded(me->bondedTerms);
      me->bonded_dirty = false;
    }
    me->engine->add_bonded(me->layer - 1, *me->box, me->bonded[me->layer - 1] != 0, kspace, bs);
  }
  r.Ecoul += bs.Ecoul;
  double Elong = 0.0, WbodyLong = 0.0;
  if (kspace) {   // compute_kspace, src/EmDeeData.f90:689-700
    me->engine->add_ewald(me->layer - 1, *me->box, Elong, WbodyLong);
    Elong += me->ewaldRigid[me->layer - 1];
  }
  md->Time.Neighbor += tn;
  double Wlong = 0.0;
  if (kspace) Wlong = r.Ecoul + Elong - r.Wcoul;   // W(long) = E(coul) + E(long) - W(coul), src/EmDeeCode.f90:1251
  md->Virial.Total = r.Wpair + r.Wcoul + Wlong + bs.Wbond + bs.Wangle;
  if (me->nbodies() != 0) {
    md->Virial.Body = r.Wbody + bs.Wbody + WbodyLong;
    md->Virial.Total = md->Virial.Total + md->Virial.Body;
  }
  if (compute) {
    md->Energy.Dispersion = r.Epair;
    md->Energy.Coulomb = r.Ecoul + Elong;
    md->Energy.Bond = bs.Ebond;
    md->Energy.Angle = bs.Eangle;
    md->Energy.Potential = r.Epair + r.Ecoul + Elong + bs.Ebond + bs.Eangle;
    md->Energy.ShadowPotential = md->Energy.Potential;
  }
  md->Energy.UpToDate = compute;
  me->forcesUpToDate[me->layer - 1] = 1;
  me->layerEnergy[me->layer - 1] = md->Energy;
  me->layerVirial[me->layer - 1] = md->Virial;
  const double t1 = now();
  md->Time.Pair += (t1 - t0) - tn;
  md->Time.Total = t1 - me->startTime;
}

// src/EmDeeCode.f90:1281-1395. Counting runs on the device over the resident list (Engine::rdf); the
// normalisation below follows the reference line by line (1329-1342). Two notes: (1) the reference picks the
// `middle` split of its r^2-sorted half list when Rc < 1.0001*InRc; the full list kept here has no such split
// and the current-distance test r < Rc selects the same pairs while the list is valid; (2) N(i)*N(j) is formed
// in 64 bits (the reference's default-integer product overflows above ~46k atoms per type).
// The counting kernel was written after the last GPU session of round 1: its logic is verified on the CPU through
// the kernel emulator (tests/test_emulated_kernels.py); its GPU parity test is tests/test_zz_rdf.py.
void EmDee_rdf(tEmDee md, int bins, double Rc, int pairs, int* itype, int* jtype, double* g) {
  const char* task = "radial distribution calculation";
  System* me = sys(md);
  for (int k = 0; k < pairs; ++k)
    if (!ranged({itype[k], jtype[k]}, me->ntypes)) error(task, "at least one provided type index is out of range");
  if (!me->initialized) error(task, "box and coordinates have not been defined");
  auto symm1D = [](int i, int j) {   // src/math.f90:695-702
    const int x = std::min(i, j) - 1, y = std::max(i, j) - 1;
    return x + (y + 1) * y / 2 + 1;
  };
  const int nt = me->ntypes;
  int maxtype = 0;
  for (int k = 0; k < pairs; ++k) maxtype = std::max(maxtype, std::max(itype[k], jtype[k]));
  const int nsym = symm1D(maxtype, maxtype);
  std::vector<unsigned short> pairSym((size_t)nt * nt, 0);   // pairOn(itype,jtype) with vector subscripts: all combinations
  for (int k = 0; k < pairs; ++k)
    for (int l = 0; l < pairs; ++l) {
      const int a = itype[k], b = jtype[l];
      pairSym[(size_t)(a - 1) * nt + (b - 1)] = pairSym[(size_t)(b - 1) * nt + (a - 1)] = (unsigned short)symm1D(a, b);
    }
  const double invL = 1.0 / *me->box;
  std::vector<long long> counts;
  me->engine->rdf(*me->box, bins, Rc * Rc * invL * invL, bins / (Rc * invL), pairSym, nsym, counts);
  const double Pi4_3 = 4.188790204786391;
  const double w = Rc * invL / bins;
  const double shell0 = Pi4_3 * (w * w * w);
  std::vector<long long> count(maxtype + 1, 0);
  for (int a = 0; a < me->N; ++a)
    if (me->type[a] <= maxtype) count[me->type[a]] += 1;
  for (int p = 0; p < pairs; ++p) {
    const int i = itype[p], j = jtype[p];
    const double NiNj = (double)(count[i] * count[j]);
    for (int b = 1; b <= bins; ++b) {
      double rdf = (double)counts[(size_t)(symm1D(i, j) - 1) * bins + (b - 1)] / shell0;
      rdf = rdf / (double)(3 * b * (b - 1) + 1);
      g[(size_t)p * bins + (b - 1)] = (i == j) ? 2.0 * rdf / NiNj : rdf / NiNj;
    }
  }
}

void* EmDee_shifted(void* m) { return with_modifier(m, nb::M_SHIFTED, 0, false, "shifted potential assignment"); }
void* EmDee_shifted_force(void* m) { return with_modifier(m, nb::M_SHIFTED_FORCE, 0, false, "shifted-force potential assignment"); }
void* EmDee_smoothed(void* m, double skin) { return with_modifier(m, nb::M_SMOOTHED, skin, true, "smoothed potential assignment"); }
void* EmDee_shifted_smoothed(void* m, double skin) { return with_modifier(m, nb::M_SHIFTED_SMOOTHED, skin, true, "shifted-smoothed potential assignment"); }
void* EmDee_square_smoothed(void* m, double skin) { return with_modifier(m, nb::M_SQUARE_SMOOTHED, skin, true, "square-smoothed potential assignment"); }
void* EmDee_shifted_square_smoothed(void* m, double skin) { return with_modifier(m, nb::M_SHIFTED_SQUARE_SMOOTHED, skin, true, "shifted-square-smoothed potential assignment"); }

void* EmDee_pair_none(void) { return deliver(blank(F_PAIR, nb::K_PAIR_NONE, "none")); }
void* EmDee_coul_none(void) { return deliver(blank(F_COUL, nb::K_COUL_NONE, "none")); }
void* EmDee_bond_none(void) { return deliver(blank(F_BOND, 0, "none")); }
void* EmDee_angle_none(void) { return deliver(blank(F_ANGLE, 0, "none")); }
void* EmDee_dihedral_none(void) { return deliver(blank(F_DIHEDRAL, 0, "none")); }
void* EmDee_pair_lj_cut(double epsilon, double sigma) { return deliver(make_lj(epsilon, sigma)); }
void* EmDee_pair_softcore_cut(double epsilon, double sigma, double lambda) { return deliver(make_softcore(epsilon, sigma, lambda)); }
void* EmDee_coul_cut(void) { return deliver(blank(F_COUL, nb::K_COUL_CUT, "cut")); }
void* EmDee_coul_sf(void) {
  HostModel m = blank(F_COUL, nb::K_COUL_SF, "sf");
  m.shifted_force = true;
  return deliver(m);
}
static HostModel damped_family(int kind, const char* name, double damp, double skinWidth) {
  HostModel m = blank(F_COUL, kind, name);
  m.skinWidth = skinWidth;
  m.dev.a = damp;                                            // alpha
  m.dev.b = 2.0 * damp / std::sqrt(3.14159265358979323846);  // beta
  return m;
}
void* EmDee_coul_damped(double damp) { return deliver(damped_family(nb::K_COUL_DAMPED, "damped", damp, 0.0)); }
void* EmDee_coul_long(void) {
  HostModel m = blank(F_COUL, nb::K_COUL_DAMPED, "long");
  m.requires_kspace = true;
  m.is_long = true;
  return deliver(m);
}
void* EmDee_coul_damped_smoothed(double damp, double skinWidth) {
  return deliver(damped_family(nb::K_COUL_DAMPED_SMOOTHED, "damped_openmm_smoothed", damp, skinWidth));
}
void* EmDee_coul_damped_square_smoothed(double damp, double skinWidth) {
  return deliver(damped_family(nb::K_COUL_DAMPED_SQUARE_SMOOTHED, "damped_smoothed", damp, skinWidth));
}
void* EmDee_coul_square_smoothed(double skinWidth) {
  HostModel m = blank(F_COUL, nb::K_COUL_SQUARE_SMOOTHED, "smoothed");
  m.skinWidth = skinWidth;
  return deliver(m);
}
void* EmDee_coul_shifted_square_smoothed(double skinWidth) {
  HostModel m = blank(F_COUL, nb::K_COUL_SHIFTED_SQUARE_SMOOTHED, "shifted_smoothed");
  m.skinWidth = skinWidth;
  m.shifted = true;
  return deliver(m);
}
void* EmDee_bond_harmonic(double k, double r0) {   // src/bond_harmonic.f90:48-64
  HostModel m = blank(F_BOND, 1, "harmonic");
  m.dev.a = k;
  m.dev.b = r0;
  return deliver(m);
}
void* EmDee_angle_harmonic(double k, double theta0) {   // src/angle_harmonic.f90:48-64
  HostModel m = blank(F_ANGLE, 1, "harmonic");
  m.dev.a = k;
  m.dev.b = theta0;
  return deliver(m);
}
void* EmDee_kspace_ewald(double accuracy) {
  HostModel m = blank(F_KSPACE, 1, "ewald");
  m.accuracy = accuracy;
  return deliver(m);
}

// ---- extensions -----------------------------------------------------------------------------------
long long EmDeeX_pair_count(tEmDee md) { return sys(md)->engine->pair_count(); }
long long EmDeeX_download_pairs(tEmDee md, int* pairs, long long capacity) {
  return sys(md)->engine->download_pairs(pairs, capacity);
}
void EmDeeX_finalize(tEmDee* md) {
  System* me = sys(*md);
  if (me == nullptr) return;
  delete me->engine;
  delete me;
  md->Data = nullptr;
}
void EmDeeX_stats(tEmDee md, tEmDeeXStats* out) {
  System* me = sys(md);
  if (me->initialized) me->engine->update_list_stats(me->layer - 1, *me->box);
  emdee::EngineStats s = me->engine->stats();
  out->launches = s.launches;
  out->force_launches = s.force_launches;
  out->force_ms = s.force_ms;
  out->build_launches = s.build_launches;
  out->build_ms = s.build_ms;
  out->list_entries = s.list_entries;
  out->interacting = s.interacting;
  out->cells_per_dim = s.cells_per_dim;
  out->device = s.device;
}
void EmDeeX_set_kernel_timing(tEmDee md, int enabled) { sys(md)->engine->set_kernel_timing(enabled != 0); }
void EmDeeX_kernel_times(tEmDee md, double* ms8, long long* n8) { sys(md)->engine->kernel_times(ms8, n8); }
void EmDeeX_synchronize(tEmDee md) { sys(md)->engine->synchronize(); }
int EmDeeX_comm_mode(tEmDee md) { return sys(md)->engine->comm_mode(); }
void EmDeeX_io_bytes(tEmDee md, long long* h2d, long long* d2h) { sys(md)->engine->io_bytes(*h2d, *d2h); }
void EmDeeX_tune(tEmDee md, const char* knob, int value) { sys(md)->engine->tune(knob, value); }
void* EmDeeX_stream(tEmDee md) { return sys(md)->engine->stream_handle(); }
double EmDeeX_measure_fp64_tflops(void) { return emdee::measure_fp64_fma_tflops(); }
void EmDeeX_math_probe(int what, int n, const double* in, double* out) { emdee::math_probe(what, n, in, out); }
void EmDeeX_comm_unique_id(char* out128) { emdee::comm_unique_id(out128); }
void EmDeeX_comm_init(tEmDee md, int rank, int world, const char* unique_id) {
  System* me = sys(md);
  if (me->initialized) error("multi-GPU setup", "EmDeeX_comm_init must be called before box and coordinates are uploaded");
  if (rank < 0 || rank >= world) error("multi-GPU setup", "rank is out of range");
  me->engine->comm_init(rank, world, unique_id);
}
void EmDeeX_slab_range(int M, int rank, int world, int* z0, int* z1) { emdee::slab_range(M, rank, world, *z0, *z1); }

}  // extern "C"
