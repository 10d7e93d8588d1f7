// engine_dynamics.cuh -- device-resident velocity-Verlet pieces for free atoms (reference src/EmDeeData.f90:823-922).
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
#pragma once

namespace emdee {
namespace {

// ------------------------------------------------------------------------------------------------
// Device-resident dynamics for free atoms. Un-fused arithmetic so trajectories track the reference.
// ------------------------------------------------------------------------------------------------
// Streaming kernels: each thread owns APT consecutive atoms = 3*APT consecutive doubles, moved as 32-byte
// vectors (every sector is touched exactly once per array); one block = TPB*APT atoms, so the grid-wide
// finish folds ~N/512 partials instead of N/128.
constexpr int APT = 4;

__device__ __forceinline__ bool vec_ok(const double* p, long long a0, int n) {
  return n == APT && ((reinterpret_cast<unsigned long long>(p + 3 * a0) & 31ull) == 0ull);   // per-layer force slabs may be unaligned
}
__device__ __forceinline__ void load12(const double* __restrict__ p, long long a0, int n, double (&v)[3 * APT]) {
  if (vec_ok(p, a0, n)) {
    const double4* q = reinterpret_cast<const double4*>(p + 3 * a0);   // 96-byte stride: 32-byte aligned
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double4 t = q[k];
      v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3 * APT; ++k) v[k] = (k < 3 * n) ? p[3 * a0 + k] : 0.0;
  }
}
// masked store: only the atoms flagged in `own` are written. On several GPUs the slots of atoms another rank owns must
// not be touched at all -- not even rewritten with the value just read: the neighbor's halo push (k_push_step) may land
// between the read and the write.
__device__ __forceinline__ void store12_own(double* __restrict__ p, long long a0, const bool (&own)[APT], const double (&v)[3 * APT]) {
#pragma unroll
  for (int j = 0; j < APT; ++j)
    if (own[j]) {
      p[3 * (a0 + j)] = v[3 * j];
      p[3 * (a0 + j) + 1] = v[3 * j + 1];
      p[3 * (a0 + j) + 2] = v[3 * j + 2];
    }
}
__device__ __forceinline__ void store12(double* __restrict__ p, long long a0, int n, const double (&v)[3 * APT]) {
  if (vec_ok(p, a0, n)) {
    double4* q = reinterpret_cast<double4*>(p + 3 * a0);
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k] = make_double4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else {
    for (int k = 0; k < 3 * n; ++k) p[3 * a0 + k] = v[k];
  }
}

// The kinetic sums come in two sets: [0..2] for the momenta this kick produces and [3..5] for the momenta ONE MORE
// identical kick would produce from the same forces (velocity Verlet issues the second half kick of a step and the first
// half kick of the next back to back): Engine::boost answers that next call from these sums without a reduction or a
// host wait. Same operations, same order as the kernel that executes the next kick, so the numbers are bit-identical.
// `crit`: speculative launch behind a speculative pair kernel (Engine::compute_forces): nothing happens when the rebuild
// criterion fired.
__global__ void __launch_bounds__(TPB) k_boost(int N, double CP, double CF, double* __restrict__ P,
                                               const double* __restrict__ F, const double* __restrict__ invMass,
                                               const unsigned char* __restrict__ owned, int want_ke,
                                               double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                               double* __restrict__ out, HostSlot* hs, unsigned long long seq,
                                               const double* __restrict__ crit, double skinSq) {
  __shared__ double red[TPB / 32][6];
  if (crit != nullptr && __ldcg(crit) > skinSq) return;
  const long long a0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * APT;
  const int n = (a0 >= N) ? 0 : (int)min((long long)APT, N - a0);
  double k[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (n > 0) {
    bool own[APT];
    bool any = false;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      own[j] = j < n && (owned == nullptr || owned[a0 + j]);   // multi-GPU: each rank integrates the atoms it owns
      any = any || own[j];
    }
    if (any) {
      double p[3 * APT], f[3 * APT];
      load12(P, a0, n, p);
      load12(F, a0, n, f);
#pragma unroll
      for (int j = 0; j < APT; ++j) {
        if (own[j]) {
          const double im = invMass[a0 + j];
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const double q = __dadd_rn(__dmul_rn(CP, p[3 * j + x]), __dmul_rn(CF, f[3 * j + x]));
            p[3 * j + x] = q;
            k[x] += __dmul_rn(__dmul_rn(im, q), q);
            const double q2 = __dadd_rn(__dmul_rn(CP, q), __dmul_rn(CF, f[3 * j + x]));
            k[3 + x] += __dmul_rn(__dmul_rn(im, q2), q2);
          }
        }
      }
      if (owned == nullptr) store12(P, a0, n, p);
      else store12_own(P, a0, own, p);
    }
  }
  if (!want_ke) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int x = 0; x < 6; ++x) {
    double v = k[x];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[threadIdx.x >> 5][x] = v;
  }
  __syncthreads();
  double mine[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (threadIdx.x == 0) {
#pragma unroll
    for (int x = 0; x < 6; ++x)
      for (int w = 0; w < TPB / 32; ++w) mine[x] += red[w][x];
  }
  grid_finish<6>(mine, partial, ticket, out, 1.0, hs, seq);
}

// R = CR*R + CP*P/m, fused with the rebuild criterion on the NEW coordinates (|R - R0|^2 ordered scan).
// Multi-GPU: only owned atoms move here and the criterion is evaluated by the distributed kernels below
// (partial == nullptr skips the fused scan).
// KICK: the kick EmDee_boost issued just before this drift and whose kinetic sums were already known (Engine::boost,
// "deferred kick") is applied here first, P = kCP*P + kCF*F with the very operations of k_boost, so the momenta make one
// trip through memory per half step instead of two.
template <bool KICK>
__global__ void __launch_bounds__(TPB) k_displace(int N, double CR, double CP, double* __restrict__ R,
                                                  double* __restrict__ P, const double* __restrict__ invMass,
                                                  const unsigned char* __restrict__ owned,
                                                  const double* __restrict__ R0, MaxNext* __restrict__ partial,
                                                  unsigned int* __restrict__ ticket, double* __restrict__ result,
                                                  double kCP, double kCF, const double* __restrict__ F) {
  const long long a0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * APT;
  const int n = (a0 >= N) ? 0 : (int)min((long long)APT, N - a0);
  MaxNext s = mn_identity();
  if (n > 0) {
    bool own[APT];
    bool any = false;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      own[j] = j < n && (owned == nullptr || owned[a0 + j]);
      any = any || own[j];
    }
    if (any) {
      double r[3 * APT], p[3 * APT];
      load12(R, a0, n, r);
      load12(P, a0, n, p);
      if (KICK) {
        double f[3 * APT];
        load12(F, a0, n, f);
#pragma unroll
        for (int j = 0; j < APT; ++j)
          if (own[j]) {
#pragma unroll
            for (int x = 0; x < 3; ++x) p[3 * j + x] = __dadd_rn(__dmul_rn(kCP, p[3 * j + x]), __dmul_rn(kCF, f[3 * j + x]));
          }
        if (owned == nullptr) store12(P, a0, n, p);
        else store12_own(P, a0, own, p);
      }
#pragma unroll
      for (int j = 0; j < APT; ++j) {
        if (own[j]) {
          const double im = invMass[a0 + j];
#pragma unroll
          for (int x = 0; x < 3; ++x)
            r[3 * j + x] = __dadd_rn(__dmul_rn(CR, r[3 * j + x]), __dmul_rn(__dmul_rn(CP, p[3 * j + x]), im));
        }
      }
      if (owned == nullptr) store12(R, a0, n, r);
      else store12_own(R, a0, own, r);
      if (partial != nullptr) {
        double r0[3 * APT];
        load12(R0, a0, n, r0);
#pragma unroll
        for (int j = 0; j < APT; ++j) {
          if (j < n) {
            const double dx = __dsub_rn(r[3 * j], r0[3 * j]), dy = __dsub_rn(r[3 * j + 1], r0[3 * j + 1]),
                         dz = __dsub_rn(r[3 * j + 2], r0[3 * j + 2]);
            MaxNext e;
            e.m = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            e.n = (a0 + j == 0) ? e.m : -1.0 / 0.0;
            s = mn_combine(s, e);   // atoms in index order inside the thread, threads in order inside the block
          }
        }
      }
    }
  }
  if (partial == nullptr) return;
  s = block_ordered_reduce(s);
  check_finish(s, partial, ticket, result);
}

}  // namespace
}  // namespace emdee
