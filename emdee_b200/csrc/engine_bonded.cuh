// engine_bonded.cuh -- harmonic bonds and angles on the device (reference src/EmDeeData.f90:443-550, src/bond_harmonic.f90:68-80,
// src/angle_harmonic.f90:68-78). Part of the single translation unit engine.cu (included there, in order).
//
// Same rule as the pair kernel: every atom's force is finished inside one thread. The host builds, per atom, the list of
// bonded terms the atom takes part in (term index and the atom's role in it); a thread walks its atom's list, evaluates
// each term and keeps only its own share of the force. A term is therefore evaluated once per member (2x for a bond, 3x
// for an angle) instead of being scattered with atomics: the sums have a fixed order and the result is deterministic.
// Energy and virial of a term are counted by the member with role 0.
#pragma once

namespace emdee {
namespace {

enum { T_BOND_NONE = 0, T_BOND_HARMONIC = 1, T_ANGLE_NONE = 2, T_ANGLE_HARMONIC = 3 };

// With an Ewald-type Coulomb model the smooth (erf) part of every bonded pair -- excluded from the real-space sum but
// present in the reciprocal one -- is taken back out here (reference cKspaceModel_discount, src/modelClass_kspace.f90:320-334).
struct BondedEwald {
  int on;                  // the layer's Coulomb model requires k-space
  double alpha, beta;
  const double* q;         // charges
  const int* type;         // 0-based atom types
  const PairEntry* tab;    // the layer's type-pair table (kCoul)
  int nt;
};

__device__ __forceinline__ void ewald_discount(const BondedEwald& k, int a, int b, double rsq, double& E, double& W) {
  const double QiQj = k.tab[k.type[a] * k.nt + k.type[b]].kCoul * k.q[a] * k.q[b];
  const double r = sqrt(rsq), x = k.alpha * r, e = exp(-x * x);
  E = -QiQj * (1.0 - nb::uerfc(x, e)) / r;
  W = E + QiQj * k.beta * e;
}

// out: [0] E(bond) [1] W(bond) [2] E(angle) [3] W(angle) [4] -sum F_bonded . delta (rigid-body virial share)
//      [5] Coulomb energy taken back out of the bonded pairs (k-space layers only)
__global__ void __launch_bounds__(TPB) k_bonded(int N, const int* __restrict__ first, const int* __restrict__ ref,
                                                const BondedTerm* __restrict__ terms, const double* __restrict__ R, double L,
                                                int bonded, BondedEwald ks, const unsigned char* __restrict__ owned,
                                                double* __restrict__ F,
                                                const double* __restrict__ delta, double* __restrict__ partial,
                                                unsigned int* __restrict__ ticket, double* __restrict__ out,
                                                double reach2, int* __restrict__ toolong) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (a < N && first[a + 1] > first[a] && (owned == nullptr || owned[a])) {   // several GPUs: each rank finishes the atoms it owns
    double f[3] = {0.0, 0.0, 0.0};
    for (int t = first[a]; t < first[a + 1]; ++t) {
      const int role = ref[t] & 3;
      const BondedTerm tm = terms[ref[t] >> 2];
      if (tm.kind == T_BOND_NONE || tm.kind == T_BOND_HARMONIC) {
        // scaled separation with the reference's minimum image, then back to real units
        double d[3], r2 = 0.0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const double s = __ddiv_rn(R[3 * (size_t)tm.a0 + x], L) - __ddiv_rn(R[3 * (size_t)tm.a1 + x], L);
          d[x] = s - rint(s);
          r2 += d[x] * d[x];
        }
        const double invR2 = (1.0 / (L * L)) / r2;
        // several GPUs: a partner is current on this rank only while it lies inside the halo, i.e. closer than Rc + skin
        if (owned != nullptr && !(r2 * (L * L) < reach2)) *toolong = 1;
        double E = 0.0, W = 0.0;
        if (bonded && tm.kind == T_BOND_HARMONIC) {
          const double r = 1.0 / sqrt(invR2), dr = r - tm.p2;
          E = (0.5 * tm.p1) * dr * dr;
          W = -tm.p1 * dr * r;
        }
        if (role == 0) {
          acc[0] += E;
          acc[1] += W;   // the virial counts the bond model only; the discount below enters the force alone
        }
        if (ks.on) {
          double EL, WL;
          ewald_discount(ks, tm.a0, tm.a1, 1.0 / invR2, EL, WL);
          W += WL;
          if (role == 0) acc[5] += EL;
        }
        const double g = W * invR2 * L, sgn = role == 0 ? 1.0 : -1.0;
#pragma unroll
        for (int x = 0; x < 3; ++x) f[x] += sgn * (g * d[x]);
      } else {
        // a0 - a1 - a2 with the vertex at a1
        double av[3], bv[3], aa = 0.0, bb = 0.0, ab = 0.0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const double c = __ddiv_rn(R[3 * (size_t)tm.a1 + x], L);
          double s = __ddiv_rn(R[3 * (size_t)tm.a0 + x], L) - c, u = __ddiv_rn(R[3 * (size_t)tm.a2 + x], L) - c;
          av[x] = L * (s - rint(s));
          bv[x] = L * (u - rint(u));
          aa += av[x] * av[x];
          bb += bv[x] * bv[x];
          ab += av[x] * bv[x];
        }
        if (owned != nullptr && !(aa < reach2 && bb < reach2)) *toolong = 1;
        if (bonded) {
          double Ea = 0.0, Fa = 0.0;
          if (tm.kind == T_ANGLE_HARMONIC) {
            const double dth = acos(ab / sqrt(aa * bb)) - tm.p2;
            Ea = (0.5 * tm.p1) * dth * dth;
            Fa = -tm.p1 * dth;
          }
          const double fac = Fa / sqrt(aa * bb - ab * ab);
          double w = 0.0;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const double Fi = ((ab / aa) * av[x] - bv[x]) * fac, Fk = ((ab / bb) * bv[x] - av[x]) * fac;
            f[x] += role == 0 ? Fi : (role == 2 ? Fk : -(Fi + Fk));
            w += Fi * av[x] + Fk * bv[x];
          }
          if (role == 0) {
            acc[2] += Ea;
            acc[3] += w;
          }
        }
        if (ks.on) {   // the 1-3 pair a0 - a2
          double rik[3], rsq = 0.0, EL, WL;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            rik[x] = av[x] - bv[x];
            rsq += rik[x] * rik[x];
          }
          ewald_discount(ks, tm.a0, tm.a2, rsq, EL, WL);
          if (role != 1) {
            const double sgn = role == 0 ? 1.0 : -1.0;
#pragma unroll
            for (int x = 0; x < 3; ++x) f[x] += sgn * (WL * rik[x] / rsq);
          }
          if (role == 0) acc[5] += EL;
        }
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) F[3 * (size_t)a + x] += f[x];
    if (delta != nullptr)
      acc[4] = -(f[0] * delta[3 * (size_t)a] + f[1] * delta[3 * (size_t)a + 1] + f[2] * delta[3 * (size_t)a + 2]);
  }
  reduce_and_finish<6>(acc, partial, ticket, out);
}

}  // namespace
}  // namespace emdee
