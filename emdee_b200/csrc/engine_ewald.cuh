// engine_ewald.cuh -- reciprocal-space Ewald sum on the device (reference src/kspace_ewald.f90:188-314: kspace_ewald_prepare,
// kspace_ewald_compute). Part of the single translation unit engine.cu (included there, in order).
//
// The reference stores q_j exp(i k.r_j) for every (wave vector, atom) pair -- O(N nvecs) memory -- and builds it from
// per-axis power recursions. Here nothing of that size is stored: one block per wave vector reduces the per-type
// structure factors S(k,t) over all atoms (phases through sincospi of the reduced fraction n.s - rint(n.s), exact range
// reduction), forms sigma(k,t) = sum_u S(k,u) lambda(u,t) and the vector's energy in its epilogue; a second kernel, one
// thread per atom, walks the wave vectors again (uniform, cache-resident loads of sigma and k) and finishes the atom's
// force in registers. Deterministic: fixed reduction trees, no atomics.
#pragma once

namespace emdee {
namespace {

constexpr int EWALD_MAX_TYPES = 8;   // distinct types of charged atoms

struct EwaldView {
  int nvecs, ntk, N;
  const int* n;            // 3 per wave vector
  const double* prefac;    // (4 pi / V) exp(-k^2 / 4 alpha^2) / k^2
  const int* ktype;        // per atom: index among the charged types, or -1
  const double* q;         // per atom charge
  const unsigned char* owned;   // several GPUs: atoms this rank sums / finishes (nullptr on one GPU)
  const double* R;
  double* sigma;           // 2 * ntk per wave vector: (re, im) per type
};

// fraction of a turn of exp(i k.r): n . (R / L), reduced to [-1/2, 1/2]
__device__ __forceinline__ double turn_fraction(const int* __restrict__ n, const double* __restrict__ R, size_t a, double L) {
  const double f = n[0] * __ddiv_rn(R[3 * a], L) + n[1] * __ddiv_rn(R[3 * a + 1], L) + n[2] * __ddiv_rn(R[3 * a + 2], L);
  return f - rint(f);
}

// one block per wave vector; out[0] accumulates sum_k prefac(k) sum_t Re(S conj(sigma)).
// raw_only (several GPUs): store this rank's partial S(k,t) in the sigma buffer and stop; after the all-reduce of that
// buffer, k_ewald_sigma turns it into sigma and adds up the energy.
__global__ void __launch_bounds__(TPB) k_ewald_structure(EwaldView v, double L, const double* __restrict__ lambda, int raw_only,
                                                         double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                         double* __restrict__ out) {
  __shared__ double red[TPB / 32][2 * EWALD_MAX_TYPES];
  const int kv = blockIdx.x;
  const int nk[3] = {v.n[3 * kv], v.n[3 * kv + 1], v.n[3 * kv + 2]};
  double S[2 * EWALD_MAX_TYPES];
#pragma unroll
  for (int t = 0; t < 2 * EWALD_MAX_TYPES; ++t) S[t] = 0.0;
  for (int a = threadIdx.x; a < v.N; a += blockDim.x) {
    const int t = v.ktype[a];
    if (t < 0 || (v.owned != nullptr && !v.owned[a])) continue;
    double s, c;
    sincospi(2.0 * turn_fraction(nk, v.R, (size_t)a, L), &s, &c);
    const double qa = v.q[a];
#pragma unroll
    for (int u = 0; u < EWALD_MAX_TYPES; ++u)
      if (u == t) {
        S[2 * u] += qa * c;
        S[2 * u + 1] += qa * s;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 2 * EWALD_MAX_TYPES; ++t) {
    double x = S[t];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) red[warp][t] = x;
  }
  __syncthreads();
  double mine[1] = {0.0};
  if (threadIdx.x == 0) {
    double T[2 * EWALD_MAX_TYPES];
    for (int t = 0; t < 2 * v.ntk; ++t) {
      T[t] = 0.0;
      for (int w = 0; w < TPB / 32; ++w) T[t] += red[w][t];
    }
    if (raw_only) {
      for (int t = 0; t < 2 * v.ntk; ++t) v.sigma[(size_t)kv * v.ntk * 2 + t] = T[t];
      return;
    }
    double e = 0.0;
    for (int t = 0; t < v.ntk; ++t) {
      double gr = 0.0, gi = 0.0;
      for (int u = 0; u < v.ntk; ++u) {
        gr += T[2 * u] * lambda[u * v.ntk + t];
        gi += T[2 * u + 1] * lambda[u * v.ntk + t];
      }
      v.sigma[((size_t)kv * v.ntk + t) * 2] = gr;
      v.sigma[((size_t)kv * v.ntk + t) * 2 + 1] = gi;
      e += T[2 * t] * gr + T[2 * t + 1] * gi;
    }
    mine[0] = v.prefac[kv] * e;
  }
  if (raw_only) return;
  grid_finish<1>(mine, partial, ticket, out, 1.0);
}

// several GPUs, after the all-reduce of the raw structure factors: one thread per wave vector forms sigma in place and
// the vector's energy (identical on every rank)
__global__ void __launch_bounds__(TPB) k_ewald_sigma(EwaldView v, const double* __restrict__ lambda, double* __restrict__ partial,
                                                     unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int kv = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (kv < v.nvecs) {
    double T[2 * EWALD_MAX_TYPES];
#pragma unroll
    for (int t = 0; t < 2 * EWALD_MAX_TYPES; ++t) T[t] = t < 2 * v.ntk ? v.sigma[(size_t)kv * v.ntk * 2 + t] : 0.0;
    double e = 0.0;
    for (int t = 0; t < v.ntk; ++t) {
      double gr = 0.0, gi = 0.0;
#pragma unroll
      for (int u = 0; u < EWALD_MAX_TYPES; ++u)
        if (u < v.ntk) {
          gr += T[2 * u] * lambda[u * v.ntk + t];
          gi += T[2 * u + 1] * lambda[u * v.ntk + t];
        }
      v.sigma[((size_t)kv * v.ntk + t) * 2] = gr;
      v.sigma[((size_t)kv * v.ntk + t) * 2 + 1] = gi;
#pragma unroll
      for (int u = 0; u < EWALD_MAX_TYPES; ++u)
        if (u == t) e += T[2 * u] * gr + T[2 * u + 1] * gi;
    }
    acc[0] = v.prefac[kv] * e;
  }
  reduce_and_finish<2>(acc, partial, ticket, out);
}

// one thread per atom: F_a += sum_k 2 prefac(k) k (Re sigma Im(q e^{ikr}) - Re(q e^{ikr}) Im sigma); out[0] = -sum F.delta
__global__ void __launch_bounds__(TPB) k_ewald_forces(EwaldView v, double L, double* __restrict__ F,
                                                      const double* __restrict__ delta, double* __restrict__ partial,
                                                      unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (a < v.N && v.ktype[a] >= 0 && (v.owned == nullptr || v.owned[a])) {
    const int t = v.ktype[a];
    const double qa = v.q[a], unit = 2.0 * 3.14159265358979324 / L;
    double f[3] = {0.0, 0.0, 0.0};
    for (int kv = 0; kv < v.nvecs; ++kv) {
      const int nk[3] = {v.n[3 * kv], v.n[3 * kv + 1], v.n[3 * kv + 2]};
      double s, c;
      sincospi(2.0 * turn_fraction(nk, v.R, (size_t)a, L), &s, &c);
      const double gr = v.sigma[((size_t)kv * v.ntk + t) * 2], gi = v.sigma[((size_t)kv * v.ntk + t) * 2 + 1];
      const double w = 2.0 * v.prefac[kv] * (gr * (qa * s) - (qa * c) * gi);
#pragma unroll
      for (int x = 0; x < 3; ++x) f[x] += (unit * nk[x]) * w;
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) F[3 * (size_t)a + x] += f[x];
    if (delta != nullptr)
      acc[0] = -(f[0] * delta[3 * (size_t)a] + f[1] * delta[3 * (size_t)a + 1] + f[2] * delta[3 * (size_t)a + 2]);
  }
  reduce_and_finish<2>(acc, partial, ticket, out);
}

}  // namespace
}  // namespace emdee
