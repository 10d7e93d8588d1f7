// engine_dist.cuh -- device kernels of the multi-GPU slab decomposition (criterion, ownership masks, migration and halo packing).
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
#pragma once

namespace emdee {
namespace {

// ------------------------------------------------------------------------------------------------
// Multi-GPU pieces (one rank per GPU, z-slab decomposition; collectives are issued by the host).
// ------------------------------------------------------------------------------------------------
// Rebuild criterion over the atoms this rank owns. The reference's sequential scan equals
//   maximum = max_i d_i, i* = first index attaining it, next = (i* == 0) ? maximum : max_{i < i*} d_i,
// so it is evaluated in two phases: (1) per-rank (max, first index) -> all-gather -> global (maximum, i*);
// (2) only when the decision is not already implied by maximum (maximum <= value <= 4*maximum):
// per-rank max over owned i < i* -> all-reduce(max).
struct MaxIdx {
  double m;
  long long i;
};
__device__ __forceinline__ MaxIdx mi_better(MaxIdx a, MaxIdx b) { return (b.m > a.m || (b.m == a.m && b.i < a.i)) ? b : a; }

__global__ void __launch_bounds__(TPB) k_check_dist(const double* __restrict__ R, const double* __restrict__ R0,
                                                    const unsigned char* __restrict__ owned, int N, long long below,
                                                    MaxIdx* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                    MaxIdx* __restrict__ result) {
  __shared__ MaxIdx sm[TPB];
  __shared__ bool last;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  MaxIdx v;
  v.m = -1.0 / 0.0;
  v.i = 0x7fffffffffffffffLL;
  if (i < N && owned[i] && i < below) {
    MaxNext d = mn_atom(R, R0, i);
    v.m = d.m;
    v.i = i;
  }
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int off = TPB / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sm[threadIdx.x] = mi_better(sm[threadIdx.x], sm[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __stcg(&partial[blockIdx.x].m, sm[0].m);
    __stcg(&partial[blockIdx.x].i, sm[0].i);
    last = (take_ticket(ticket) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  MaxIdx acc;
  acc.m = -1.0 / 0.0;
  acc.i = 0x7fffffffffffffffLL;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += TPB) {
    MaxIdx p;
    p.m = __ldcg(&partial[b].m);
    p.i = __ldcg(&partial[b].i);
    acc = mi_better(acc, p);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int off = TPB / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sm[threadIdx.x] = mi_better(sm[threadIdx.x], sm[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    result[0] = sm[0];
    *ticket = 0u;
  }
}

// dst = owned ? src : 0 (three doubles per atom): the summand of the all-reduce that rebuilds a full array
__global__ void __launch_bounds__(TPB) k_mask_owned(int N, const unsigned char* __restrict__ owned,
                                                    const double* __restrict__ src, double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const bool o = owned[i];
#pragma unroll
  for (int x = 0; x < 3; ++x) dst[3 * (size_t)i + x] = o ? src[3 * (size_t)i + x] : 0.0;
}

// Rebuild-time migration (multi-GPU): instead of re-assembling the full coordinate and momentum arrays on
// every rank, each rank sends its owned atoms that now sit within three cell layers of a slab face (or just
// beyond it) to the neighbor on that side, as (id, R, P) records. After that a rank "knows" its previously
// owned atoms plus what it received -- a superset of its new slab + 2-layer halo, because nothing moves more
// than skin/2 < one layer between rebuilds -- and re-bins only those.
__global__ void __launch_bounds__(TPB) k_mig_flags(int N, double L, GridDesc g, const double* __restrict__ R,
                                                   const unsigned char* __restrict__ owned,
                                                   unsigned char* __restrict__ fl) {   // 2 flag arrays of N
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  bool up = false, dn = false;
  if (owned[i]) {
    const double rs = __ddiv_rn(R[3 * (size_t)i + 2], L);
    int cz = (int)__dmul_rn((double)g.M, __dsub_rn(rs, floor(rs)));
    if (cz >= g.M) cz = g.M - 1;
    const int z1 = g.z0 + g.nzl;
    up = ((cz - (z1 - 3)) % g.M + g.M) % g.M < 5;     // layers z1-3 .. z1+1 (periodic)
    dn = (((g.z0 + 2) - cz) % g.M + g.M) % g.M < 5;   // layers z0-2 .. z0+2
  }
  fl[i] = up;
  fl[(size_t)N + i] = dn;
}

__global__ void __launch_bounds__(TPB) k_pack7(int n, const int* __restrict__ list, const double* __restrict__ R,
                                               const double* __restrict__ P, double* __restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int a = list[k];
  double* b = buf + 7 * (size_t)k;
  b[0] = (double)a;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    b[1 + x] = R[3 * (size_t)a + x];
    b[4 + x] = P[3 * (size_t)a + x];
  }
}
__global__ void __launch_bounds__(TPB) k_unpack7(int n, const double* __restrict__ buf, double* __restrict__ R,
                                                 double* __restrict__ P, unsigned char* __restrict__ known) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* b = buf + 7 * (size_t)k;
  const int a = (int)b[0];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    R[3 * (size_t)a + x] = b[1 + x];
    P[3 * (size_t)a + x] = b[4 + x];
  }
  known[a] = 1;
}

// halo bookkeeping: which owned atoms sit in my top / bottom two layers (to send), which foreign atoms sit
// in the two layers above / below my slab (to receive). Lists are compacted in ascending atom order, so
// a sender's list and the matching receiver's list are identical without exchanging indices.
__global__ void __launch_bounds__(TPB) k_halo_flags(int N, GridDesc g, const int* __restrict__ atomCell,
                                                    unsigned char* __restrict__ fl) {   // 4 flag arrays of N
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (atomCell[i] < 0) {   // not known to this rank
    fl[i] = fl[(size_t)N + i] = fl[2 * (size_t)N + i] = fl[3 * (size_t)N + i] = 0;
    return;
  }
  const int cz = (atomCell[i] >> 20) & 1023;
  const int z1 = g.z0 + g.nzl;
  const bool own = cz >= g.z0 && cz < z1;
  const int up0 = z1 % g.M, up1 = (z1 + 1) % g.M;
  const int dn0 = (g.z0 - 2 + g.M) % g.M, dn1 = (g.z0 - 1 + g.M) % g.M;
  fl[i] = own && cz >= z1 - 2;                          // send up
  fl[(size_t)N + i] = own && cz < g.z0 + 2;             // send down
  fl[2 * (size_t)N + i] = !own && (cz == dn0 || cz == dn1);   // receive from below
  fl[3 * (size_t)N + i] = !own && (cz == up0 || cz == up1);   // receive from above
}

__global__ void __launch_bounds__(TPB) k_pack3(int n, const int* __restrict__ list, const double* __restrict__ X,
                                               double* __restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int a = list[k];
  buf[3 * (size_t)k] = X[3 * (size_t)a];
  buf[3 * (size_t)k + 1] = X[3 * (size_t)a + 1];
  buf[3 * (size_t)k + 2] = X[3 * (size_t)a + 2];
}
__global__ void __launch_bounds__(TPB) k_unpack3(int n, const int* __restrict__ list, const double* __restrict__ buf,
                                                 double* __restrict__ X) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int a = list[k];
  X[3 * (size_t)a] = buf[3 * (size_t)k];
  X[3 * (size_t)a + 1] = buf[3 * (size_t)k + 1];
  X[3 * (size_t)a + 2] = buf[3 * (size_t)k + 2];
}

}  // namespace
}  // namespace emdee
