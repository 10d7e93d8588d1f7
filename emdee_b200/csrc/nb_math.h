// nb_math.h -- pair / Coulomb model bodies and the modifier post-processing, usable from host setup
// code and from the CUDA kernels (one definition, so the constants the host derives at the cutoff are
// evaluated by exactly the functions the kernels run).
//
// Reference loops replaced (paths relative to the reference tree):
//   src/pair_lj_cut.f90:73-86, src/pair_softcore_cut.f90:86-101, src/coul_cut.f90:62-70,
//   src/coul_sf.f90:61-72, src/coul_damped.f90:71-83, src/coul_long.f90:80-92,
//   src/coul_damped_smoothed.f90:92-115, src/coul_damped_square_smoothed.f90:91-114,
//   src/coul_square_smoothed.f90:83-104, src/coul_shifted_square_smoothed.f90:86-107,
//   src/math.f90:685-691 (uerfc), src/apply_modifier.f90:1-60.
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

namespace nb {

// model kinds (device-visible subset: everything the pair loop can dispatch to)
enum Kind : int {
  K_PAIR_NONE = 0,
  K_PAIR_LJ_CUT = 1,
  K_PAIR_SOFTCORE_CUT = 2,
  K_COUL_NONE = 10,
  K_COUL_CUT = 11,
  K_COUL_SF = 12,
  K_COUL_DAMPED = 13,   // also coul_long's real-space term (same formula, alpha from the Ewald setup)
  K_COUL_DAMPED_SMOOTHED = 14,
  K_COUL_DAMPED_SQUARE_SMOOTHED = 15,
  K_COUL_SQUARE_SMOOTHED = 16,
  K_COUL_SHIFTED_SQUARE_SMOOTHED = 17,
  K_DYNAMIC = -1        // template marker: dispatch at run time
};

// modifiers, numbered as in src/modelClass_nonbonded.f90:27-33
enum Modifier : int {
  M_NONE = 0,
  M_SHIFTED = 1,
  M_SHIFTED_FORCE = 2,
  M_SMOOTHED = 3,
  M_SHIFTED_SMOOTHED = 4,
  M_SQUARE_SMOOTHED = 5,
  M_SHIFTED_SQUARE_SMOOTHED = 6,
  M_DYNAMIC = -1
};

// Everything a kernel needs to evaluate one model + its modifier. 80 bytes, POD.
//   kind-specific slots:  LJ        a=eps4       b=eps24       c=sigsq
//                         softcore  a=prefactor  b=prefactor6  c=invSigSq  d=shift
//                         coulomb   a=alpha      b=beta        c=Rm2       d=invRm
struct DevModel {
  int kind;
  int modifier;
  double eshift, fshift, Rm, factor, Rm2fac;
  double a, b, c, d;
};

// The pair distance in all the forms the model formulas use, so that no model body needs a division:
// the reference writes `a/invR`, `1/invR2`, ...; here r and r2 are carried along (r = r2*invR), which is
// the same number to within one rounding.
struct Dist {
  double invR, invR2, r, r2;
};
NB_HD Dist make_dist(double invR, double invR2) {   // host-side setup: the reference passes (1/r, 1/r^2)
  Dist d;
  d.invR = invR;
  d.invR2 = invR2;
  d.r = 1.0 / invR;
  d.r2 = 1.0 / invR2;
  return d;
}

// reciprocal: full-precision division on the host; on the device the 20-bit hardware seed plus one cubic
// refinement (relative error < 2^-57)
NB_HD double rcp(double a) {
#if defined(__CUDA_ARCH__)
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  double t = fma(e, e, e);
  return fma(x, t, x);
#else
  return 1.0 / a;
#endif
}

NB_HD double uerfc(double x, double expmx2) {
  const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027,
               a5 = 1.061405429, p = 0.327591100;
  double t = rcp(1.0 + p * x);
  return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * expmx2;
}

// quintic switch: G(u) and the u-dependent part of W_G, coef = -30 (switch in r) or -60 (in r^2)
NB_HD void quintic(double u, double coef, double& G, double& WGu) {
  double u2 = u * u;
  double u3 = u * u2;
  G = 1.0 + u3 * (15.0 * u - 6.0 * u2 - 10.0);
  WGu = coef * u2 * (2.0 * u - u2 - 1.0);
}

// E(r), W(r) = -r dE/dr of model `m` (unit charges for Coulomb kinds). KIND is a compile-time kind
// or K_DYNAMIC. invR = 1/r, invR2 = 1/r^2 in real (unscaled) units.
template <int KIND>
NB_HD void eval_kind(const DevModel& m, const Dist& D, double& E, double& W) {
  const double invR = D.invR, invR2 = D.invR2;
  const int kind = (KIND == K_DYNAMIC) ? m.kind : KIND;
  switch (kind) {
    case K_PAIR_LJ_CUT: {
      double sr2 = m.c * invR2;
      double sr6 = sr2 * sr2 * sr2;
      double sr12 = sr6 * sr6;
      E = m.a * (sr12 - sr6);
      W = m.b * (sr12 + sr12 - sr6);
      break;
    }
    case K_PAIR_SOFTCORE_CUT: {
      double rsig2 = m.c * D.r2;
      double rsig6 = rsig2 * rsig2 * rsig2;
      double sinv = rcp(rsig6 + m.d);
      double sinvSq = sinv * sinv;
      double sinvCb = sinv * sinvSq;
      E = m.a * (sinvSq - sinv);
      W = m.b * rsig6 * (sinvCb + sinvCb - sinvSq);
      break;
    }
    case K_COUL_CUT:
      E = invR;
      W = invR;
      break;
    case K_COUL_SF: {
      double rFc = m.fshift * D.r;
      E = invR + m.eshift + rFc;
      W = invR - rFc;
      break;
    }
    case K_COUL_DAMPED: {
      double x = m.a * D.r;
      double expmx2 = exp(-x * x);
      E = uerfc(x, expmx2) * invR;
      W = E + m.b * expmx2;
      break;
    }
    case K_COUL_DAMPED_SMOOTHED: {
      double r = D.r;
      double x = m.a * r;
      double expmx2 = exp(-x * x);
      E = uerfc(x, expmx2) * invR;
      W = E + m.b * expmx2;
      if (r > m.Rm) {   // m.Rm is the INHERITED field the modifier machinery also writes
        double G, WG;
        quintic(m.factor * (r - m.Rm), -30.0, G, WG);
        WG = WG * m.factor * r;
        W = W * G + E * WG;
        E = E * G;
      }
      break;
    }
    case K_COUL_DAMPED_SQUARE_SMOOTHED: {
      double x = m.a * D.r;
      double expmx2 = exp(-x * x);
      E = uerfc(x, expmx2) * invR;
      W = E + m.b * expmx2;
      if (invR < m.d) {
        double r2 = D.r2;
        double G, WG;
        quintic(m.factor * (r2 - m.c), -60.0, G, WG);
        WG = WG * m.factor * r2;
        W = W * G + E * WG;
        E = E * G;
      }
      break;
    }
    case K_COUL_SQUARE_SMOOTHED:
    case K_COUL_SHIFTED_SQUARE_SMOOTHED: {
      W = invR;
      E = (kind == K_COUL_SHIFTED_SQUARE_SMOOTHED) ? W + m.eshift : W;
      if (invR < m.d) {
        double r2 = D.r2;
        double G, WG;
        quintic(m.factor * (r2 - m.c), -60.0, G, WG);
        WG = WG * m.factor * r2;
        W = W * G + E * WG;
        E = E * G;
      }
      break;
    }
    default:   // K_PAIR_NONE, K_COUL_NONE
      E = 0.0;
      W = 0.0;
      break;
  }
}

// src/apply_modifier.f90: post-processing of (E, W) by the model's modifier.
template <int MOD>
NB_HD void eval_modifier(const DevModel& m, const Dist& D, double& E, double& W) {
  const int mod = (MOD == M_DYNAMIC) ? m.modifier : MOD;
  switch (mod) {
    case M_SHIFTED:
      E = E + m.eshift;
      break;
    case M_SHIFTED_FORCE: {
      double rFc = m.fshift * D.r;
      W = W - rFc;
      E = E + m.eshift + rFc;
      break;
    }
    case M_SMOOTHED:
    case M_SHIFTED_SMOOTHED:
    case M_SQUARE_SMOOTHED:
    case M_SHIFTED_SQUARE_SMOOTHED: {
      const bool square = (mod == M_SQUARE_SMOOTHED || mod == M_SHIFTED_SQUARE_SMOOTHED);
      E = E + m.eshift;
      double r2fac = square ? m.factor * D.r2 : m.factor * D.r;
      if (r2fac > m.Rm2fac) {
        double G, WG;
        quintic(r2fac - m.Rm2fac, square ? -60.0 : -30.0, G, WG);
        WG = WG * r2fac;
        W = W * G + E * WG;
        E = E * G;
      }
      break;
    }
    default:
      break;
  }
}

}  // namespace nb
