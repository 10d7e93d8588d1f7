// engine.cu -- CUDA hot path (sm_100a): cell binning, cell sort with periodic ghost images, Verlet
// list build, pair force/energy/virial kernels, device-resident velocity-Verlet pieces.
//
// Reference loops replaced (paths relative to the reference tree):
//   k_displacement_check, k_displace (fused check)  <- src/neighbor_lists.f90:41-59,184
//   k_bin / k_fill / k_place                        <- src/neighbor_lists.f90:63-171   (distribute_atoms)
//   k_build_list                                    <- src/neighbor_lists.f90:199-300  (build_neighbor_lists)
//   k_refresh_positions                             <- src/EmDeeCode.f90:1228          (Rs = R/L)
//   k_pair_forces   <- src/compute.f90:20-100 + src/apply_modifier.f90 + model bodies,
//                      src/EmDeeData.f90:644-685, src/EmDeeCode.f90:1236-1247 (reductions),
//                      src/EmDeeData.f90:926-953 (rigid_body_virial, fused in the epilogue)
//   k_boost / k_displace                            <- src/EmDeeData.f90:823-922 (free atoms)
//
// Design: atoms are sorted by cell of an EXTENDED grid (M+4)^3 that carries explicit
// periodic ghost images in a 2-cell shell, so the force kernel needs no minimum-image arithmetic; the
// Verlet list is a FULL list (both directions of every pair) stored as 32-lane tiles (ELL-in-tile,
// coalesced 128-byte rows), so every atom's force is finished inside one thread: no atomics, no
// scatter, deterministic summation order. List MEMBERSHIP is decided with the reference's exact
// arithmetic (un-fused IEEE operations on unwrapped scaled coordinates), so pair sets are bit-identical;
// an FP32 pre-test with a rigorous error band only short-cuts candidates far from the cutoff sphere.
#include "engine.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace emdee {

namespace {

#define CUDA_CHECK(call)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "Error in CUDA runtime: %s (%s:%d).\n", cudaGetErrorString(err__),     \
                   __FILE__, __LINE__);                                                           \
      std::exit(1);                                                                               \
    }                                                                                             \
  } while (0)

[[noreturn]] void fatal(const char* task, const char* msg) {
  std::fprintf(stderr, "Error in %s: %s.\n", task, msg);
  std::exit(1);
}

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  void ensure(size_t m, double slack = 1.0) {
    if (m > n) {
      if (p) CUDA_CHECK(cudaFree(p));
      n = (size_t)(m * slack) + 16;
      CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
    }
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

constexpr int TPB = 128;           // threads per block for per-atom / per-entry kernels
constexpr int TILE = 32;           // list tile = one warp of consecutive sorted entries
constexpr int MAX_SMEM_TYPES = 16; // interaction table staged in shared memory up to this many types
constexpr double MAGIC_RINT = 6755399441055744.0;   // 1.5 * 2^52: (x + M) - M == rint(x) for |x| < 2^51
constexpr double DEPS = 2.220446049250313e-16;      // epsilon(1d0): charged = |q| > epsilon

inline int nblocks(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

// Ticket for the "last block finishes" pattern. Release semantics order this thread's earlier global
// writes before the increment WITHOUT an acquire (an acquire invalidates the SM's whole L1, which would
// throw away the position lines the other resident blocks are still gathering from; measured: L1 hit rate
// 75% -> 36% with a plain __threadfence() per block). Only the last block pays a full fence.
__device__ __forceinline__ unsigned int take_ticket(unsigned int* ticket) {
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ticket) : "memory");
  return old;
}

// ------------------------------------------------------------------------------------------------
// Grid-wide finish without a second launch: every block publishes WIDTH partial sums, takes a ticket,
// and the block drawing the last ticket folds all partials in a FIXED order (thread t sums blocks
// t, t+T, ...; then a fixed shared-memory tree), so the result never depends on which block is last.
// ------------------------------------------------------------------------------------------------
template <int WIDTH>
__device__ __forceinline__ void grid_finish(const double (&mine)[WIDTH], double* __restrict__ partial,
                                            unsigned int* __restrict__ ticket, double* __restrict__ out,
                                            double scale_first4) {
  __shared__ double fin[TPB][WIDTH];
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < WIDTH; ++q) __stcg(&partial[(size_t)blockIdx.x * WIDTH + q], mine[q]);
    last = (take_ticket(ticket) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const bool worker = threadIdx.x < TPB;   // blocks may be larger than TPB; the fold always uses TPB threads
  if (worker) {
    double acc[WIDTH];
#pragma unroll
    for (int q = 0; q < WIDTH; ++q) acc[q] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += TPB) {
#pragma unroll
      for (int q = 0; q < WIDTH; ++q) acc[q] += __ldcg(&partial[(size_t)b * WIDTH + q]);
    }
#pragma unroll
    for (int q = 0; q < WIDTH; ++q) fin[threadIdx.x][q] = acc[q];
  }
  __syncthreads();
  for (int off = TPB / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
#pragma unroll
      for (int q = 0; q < WIDTH; ++q) fin[threadIdx.x][q] += fin[threadIdx.x + off][q];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < WIDTH; ++q) out[q] = fin[0][q] * ((q < 4) ? scale_first4 : 1.0);
    *ticket = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// K0: rebuild trigger. Ordered reduction reproducing the sequential scan of maximum_approach_sq:
// state (m, n): m = running maximum, n = value `next` holds. combine(A then B) =
//   B.m > A.m ? (B.m, max(A.m, B.n)) : A.     Atom 0 contributes (d0, d0), atom i>0 (d_i, -inf).
// The criterion is evaluated where the coordinates change (k_displace) or, after an upload, by
// k_displacement_check; either way the last block folds the per-block states IN BLOCK ORDER.
// ------------------------------------------------------------------------------------------------
struct MaxNext {
  double m, n;
};
__device__ __forceinline__ MaxNext mn_combine(MaxNext a, MaxNext b) {
  if (b.m > a.m) {
    MaxNext r;
    r.m = b.m;
    r.n = fmax(a.m, b.n);
    return r;
  }
  return a;
}
__device__ __forceinline__ MaxNext mn_identity() {
  MaxNext s;
  s.m = -1.0 / 0.0;
  s.n = -1.0 / 0.0;
  return s;
}
__device__ __forceinline__ MaxNext mn_atom(const double* __restrict__ R, const double* __restrict__ R0, long long i) {
  double dx = __dsub_rn(R[3 * i], R0[3 * i]);
  double dy = __dsub_rn(R[3 * i + 1], R0[3 * i + 1]);
  double dz = __dsub_rn(R[3 * i + 2], R0[3 * i + 2]);
  double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  MaxNext e;
  e.m = d;
  e.n = (i == 0) ? d : -1.0 / 0.0;
  return e;
}

// ordered reduction over the block (thread order = atom order); result valid in thread 0
__device__ __forceinline__ MaxNext block_ordered_reduce(MaxNext s) {
  __shared__ MaxNext warp_state[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 1; off < 32; off <<= 1) {
    MaxNext o;
    o.m = __shfl_down_sync(0xffffffffu, s.m, off);
    o.n = __shfl_down_sync(0xffffffffu, s.n, off);
    if ((lane & (2 * off - 1)) == 0 && lane + off < 32) s = mn_combine(s, o);
  }
  if (lane == 0) warp_state[warp] = s;
  __syncthreads();
  MaxNext r = warp_state[0];
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) r = mn_combine(r, warp_state[w]);
  }
  __syncthreads();
  return r;
}

// publish this block's state; the last block folds all block states in block order and writes
// maximum + 2*sqrt(maximum*next) + next (reference neighbor_lists.f90:57)
__device__ __forceinline__ void check_finish(MaxNext mine, MaxNext* __restrict__ partial,
                                             unsigned int* __restrict__ ticket, double* __restrict__ result) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __stcg(&partial[blockIdx.x].m, mine.m);
    __stcg(&partial[blockIdx.x].n, mine.n);
    last = (take_ticket(ticket) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int nparts = gridDim.x;
  const int per = (nparts + blockDim.x - 1) / blockDim.x;
  MaxNext s = mn_identity();
  for (int q = 0; q < per; ++q) {
    int i = threadIdx.x * per + q;
    if (i < nparts) {
      MaxNext p;
      p.m = __ldcg(&partial[i].m);
      p.n = __ldcg(&partial[i].n);
      s = mn_combine(s, p);
    }
  }
  s = block_ordered_reduce(s);
  if (threadIdx.x == 0) {
    result[0] = __dadd_rn(__dadd_rn(s.m, __dmul_rn(2.0, __dsqrt_rn(__dmul_rn(s.m, s.n)))), s.n);
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(TPB) k_displacement_check(const double* __restrict__ R,
                                                            const double* __restrict__ R0, int N,
                                                            MaxNext* __restrict__ partial,
                                                            unsigned int* __restrict__ ticket,
                                                            double* __restrict__ result) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  MaxNext s = (i < N) ? mn_atom(R, R0, i) : mn_identity();
  s = block_ordered_reduce(s);
  check_finish(s, partial, ticket, result);
}

// ------------------------------------------------------------------------------------------------
// K1-K3: binning into the extended grid (real cells [2, M+2) per dimension + 2-cell ghost shell).
// ------------------------------------------------------------------------------------------------
struct GridDesc {
  int M;     // real cells per dimension of the WHOLE box (reference: max(floor(2L/xRc), 5))
  int Mx;    // extended cells per dimension in x and y = M + 4
  int z0;    // first global cell layer owned by this rank (0 on a single GPU)
  int nzl;   // number of owned layers (M on a single GPU)
  int Mz;    // extended layers in z = nzl + 4 (two halo layers each side: periodic images or neighbor ranks' atoms)
};

// images of an atom whose real cell coordinate is c (x or y): s in {0} U {+1 if c<=1} U {-1 if c>=M-2};
// at most two because M >= 5.
__device__ __forceinline__ int image_shifts(int c, int M, int s[2]) {
  s[0] = 0;
  if (c <= 1) {
    s[1] = 1;
    return 2;
  }
  if (c >= M - 2) {
    s[1] = -1;
    return 2;
  }
  return 1;
}

// z direction, slab aware: the atom in global layer cz appears at local layer cz - z0 + 2 + s*M for every
// s in {-1,0,1} that lands inside [0, Mz). On a single GPU (z0 = 0, Mz = M + 4) this is image_shifts().
__device__ __forceinline__ int image_shifts_z(int cz, const GridDesc& g, int s[3], int lz[3]) {
  int n = 0;
  for (int t = -1; t <= 1; ++t) {
    const int l = cz - g.z0 + 2 + t * g.M;
    if (l >= 0 && l < g.Mz) {
      s[n] = t;
      lz[n] = l;
      ++n;
    }
  }
  return n;
}

__global__ void __launch_bounds__(TPB) k_bin(const double* __restrict__ R, int N, double L, GridDesc g,
                                             double* __restrict__ Rs, int* __restrict__ atomCell,
                                             int* __restrict__ atomFloor, unsigned char* __restrict__ owned,
                                             const unsigned char* __restrict__ known, int* __restrict__ cellCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (known != nullptr && !known[i]) {   // multi-GPU: this rank holds no current position for the atom
    owned[i] = 0;
    atomCell[i] = -1;
    return;
  }
  int c[3];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    double rs = __ddiv_rn(R[3 * (size_t)i + x], L);   // Rs = R/L, IEEE division like the strict oracle
    Rs[3 * (size_t)i + x] = rs;
    double fl = floor(rs);
    int ic = (int)__dmul_rn((double)g.M, __dsub_rn(rs, fl));   // int(M*(Rs - floor(Rs)))
    if (ic >= g.M) ic = g.M - 1;                               // Q5 clamp (tiny negative Rs)
    c[x] = ic;
    atomFloor[3 * (size_t)i + x] = (int)fl;
  }
  atomCell[i] = c[0] | (c[1] << 10) | (c[2] << 20);
  owned[i] = (c[2] >= g.z0 && c[2] < g.z0 + g.nzl);
  int sx[2], sy[2], sz[3], lz[3];
  int nx = image_shifts(c[0], g.M, sx), ny = image_shifts(c[1], g.M, sy), nz = image_shifts_z(c[2], g, sz, lz);
  for (int a = 0; a < nz; ++a)
    for (int b = 0; b < ny; ++b)
      for (int d = 0; d < nx; ++d) {
        int ex = c[0] + 2 + sx[d] * g.M, ey = c[1] + 2 + sy[b] * g.M;
        atomicAdd(&cellCount[ex + g.Mx * (ey + g.Mx * lz[a])], 1);
      }
}

__global__ void __launch_bounds__(TPB) k_fill(int N, GridDesc g, const int* __restrict__ atomCell,
                                              const int* __restrict__ cellStart, int* __restrict__ cellFill,
                                              int* __restrict__ slotAtom, int* __restrict__ slotImg,
                                              int* __restrict__ slotCell) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int pc = atomCell[i];
  if (pc < 0) return;   // not known to this rank
  int c[3] = {pc & 1023, (pc >> 10) & 1023, (pc >> 20) & 1023};
  int sx[2], sy[2], sz[3], lz[3];
  int nx = image_shifts(c[0], g.M, sx), ny = image_shifts(c[1], g.M, sy), nz = image_shifts_z(c[2], g, sz, lz);
  for (int a = 0; a < nz; ++a)
    for (int b = 0; b < ny; ++b)
      for (int d = 0; d < nx; ++d) {
        int ex = c[0] + 2 + sx[d] * g.M, ey = c[1] + 2 + sy[b] * g.M;
        int cell = ex + g.Mx * (ey + g.Mx * lz[a]);
        int slot = cellStart[cell] + atomicAdd(&cellFill[cell], 1);
        slotAtom[slot] = i;
        slotImg[slot] = (sx[d] + 1) | ((sy[b] + 1) << 2) | ((sz[a] + 1) << 4);
        slotCell[slot] = cell;
      }
}

// Deterministic order inside a cell: ascending atom index (rank by counting), then materialise the
// per-entry arrays the list build and the force kernel read.
struct PlaceArgs {
  int Next;
  const int* slotAtom;
  const int* slotImg;
  const int* slotCell;
  const int* cellStart;
  const int* atomFloor;
  const double* Rs;
  const int* atomType;
  const int* atomBody;
  const unsigned char* owned;
  int4* sMeta;            // {atom, sx, sy, sz}: position = R/L + (sx,sy,sz)
  int* sCell;
  unsigned char* sGhost;
  int* sType;
  int* sBody;
  double4* sRs;           // unwrapped scaled coordinates of the underlying atom (exact membership test)
  float4* sPosF;          // ghost-shifted scaled position rounded to FP32 (pre-test only)
  int* nbrCount;
};

__global__ void __launch_bounds__(TPB) k_place(PlaceArgs a) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.Next) return;
  int cell = a.slotCell[t];
  int at = a.slotAtom[t];
  int lo = a.cellStart[cell], hi = a.cellStart[cell + 1];
  int rank = 0;
  for (int u = lo; u < hi; ++u) rank += (a.slotAtom[u] < at);
  int e = lo + rank;
  int img = a.slotImg[t];
  int sx = (img & 3) - 1 - a.atomFloor[3 * (size_t)at];
  int sy = ((img >> 2) & 3) - 1 - a.atomFloor[3 * (size_t)at + 1];
  int sz = ((img >> 4) & 3) - 1 - a.atomFloor[3 * (size_t)at + 2];
  a.sMeta[e] = make_int4(at, sx, sy, sz);
  a.sCell[e] = cell;
  a.sGhost[e] = (img != (1 | (1 << 2) | (1 << 4))) || !a.owned[at];   // real = central image of an atom this rank owns
  a.sType[e] = a.atomType[at];
  a.sBody[e] = a.atomBody[at];
  double x = a.Rs[3 * (size_t)at], y = a.Rs[3 * (size_t)at + 1], z = a.Rs[3 * (size_t)at + 2];
  a.sRs[e] = make_double4(x, y, z, 0.0);
  a.sPosF[e] = make_float4((float)(x + (double)sx), (float)(y + (double)sy), (float)(z + (double)sz),
                            __int_as_float(a.atomBody[at]));   // w = body id bits: the build's body mask needs no extra gather
  a.nbrCount[e] = 0;
}

// ------------------------------------------------------------------------------------------------
// K4: Verlet list build. One thread per real entry; candidates = the 5x5x5 block of extended cells
// around the entry's cell, walked as 25 contiguous x-runs that are first clipped against the cutoff
// sphere (a run, or its ends, that cannot hold a neighbor is skipped). Membership is the reference's,
// bit for bit:
//   d = Rs_i - Rs_j (unwrapped scaled), d -= anint(d), r2 = (dx^2 + dy^2) + dz^2, r2 < xRc^2/L^2
// with every operation individually rounded (no FMA contraction; rint replaces anint: they differ only
// at |d| = k + 1/2 exactly, where (d - round(d))^2 is the same number). An FP32 pre-test on the
// ghost-shifted positions decides candidates whose FP32 r^2 lies outside [xRc2 - band, xRc2 + band];
// `band` bounds the FP32 error rigorously (host: build_band), so only the thin shell is re-tested in FP64
// and the accepted set is exactly the reference's.
// ------------------------------------------------------------------------------------------------
struct BuildArgs {
  int Next, cap, nt;
  int all_interact;      // every type pair interacts: the per-candidate type lookup is skipped
  GridDesc g;
  double xRc2s;          // xRcSq * invL2
  double xRcs;           // sqrt of it, padded (run clipping only; conservative)
  float r2_accept, r2_reject;   // FP32 pre-test thresholds: < accept => in, > reject => out
  const int* cellStart;
  const int4* sMeta;
  const int* sCell;
  const unsigned char* sGhost;
  const int* sType;
  const int* sBody;
  const double4* sRs;
  const float4* sPosF;
  const int* exFirst;   // CSR over atoms (0-based rows), items = 0-based atom ids ascending
  const int* exItem;
  const unsigned char* interact;   // nt*nt
  int* nbr;
  int* nbrCount;
  int* flags;   // [0] = max count seen, [1] = overflow
};

__device__ __forceinline__ double strict_pbc_sq(double a, double b) {
  double d = __dsub_rn(a, b);
  double r = __dadd_rn(__dadd_rn(d, MAGIC_RINT), -MAGIC_RINT);
  d = __dsub_rn(d, r);
  return __dmul_rn(d, d);
}

__global__ void __launch_bounds__(TPB) k_build_list(const __grid_constant__ BuildArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  const int lane = threadIdx.x & 31;
  if (e < a.Next && !a.sGhost[e]) {
    int* out = a.nbr + ((size_t)(e >> 5) * a.cap) * TILE + lane;
    const int atom_i = a.sMeta[e].x;
    const int type_i = a.sType[e], body_i = a.sBody[e];
    const double4 ri = a.sRs[e];
    const float4 pf = a.sPosF[e];
    const int x0 = a.exFirst[atom_i], x1 = a.exFirst[atom_i + 1];
    const int cell = a.sCell[e];
    const int Mx = a.g.Mx;
    const int ez = cell / (Mx * Mx), ey = (cell - ez * Mx * Mx) / Mx, ex = cell - Mx * (ey + Mx * ez);
    // geometry for run clipping, in scaled units: extended cell c spans [(c-2)/M, (c-1)/M)
    const float w = 1.0f / (float)a.g.M;
    const float slack = 1.0e-5f * w + 4.0e-7f;   // covers FP32 rounding of the clip arithmetic and positions
    const float rc = (float)a.xRcs + slack;
    const float rc2 = rc * rc;
    for (int dz = -2; dz <= 2; ++dz) {
      const float zlo = (float)(ez + dz - 2 + a.g.z0) * w, zhi = zlo + w;   // local layer -> global coordinate
      const float gz = fmaxf(0.0f, fmaxf(zlo - pf.z, pf.z - zhi) - slack);
      for (int dy = -2; dy <= 2; ++dy) {
        const float ylo = (float)(ey + dy - 2) * w, yhi = ylo + w;
        const float gy = fmaxf(0.0f, fmaxf(ylo - pf.y, pf.y - yhi) - slack);
        const float rem = rc2 - gz * gz - gy * gy;
        if (rem <= 0.0f) continue;
        const float hx = sqrtf(rem) + slack;
        int cl = (int)floorf((pf.x - hx) * (float)a.g.M) + 2;   // `slack` (inside hx) exceeds the FP32 rounding here
        int ch = (int)floorf((pf.x + hx) * (float)a.g.M) + 2;
        cl = max(cl, ex - 2);
        ch = min(ch, ex + 2);
        const int row = Mx * ((ey + dy) + Mx * (ez + dz));
        const int f0 = a.cellStart[row + cl], f1 = a.cellStart[row + ch + 1];
        for (int f = f0; f < f1; ++f) {
          const float4 qf = __ldg(&a.sPosF[f]);
          const float dxf = pf.x - qf.x, dyf = pf.y - qf.y, dzf = pf.z - qf.z;
          const float r2f = fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf));
          if (r2f > a.r2_reject) continue;
          if (f == e) continue;
          if (r2f >= a.r2_accept) {   // inside the FP32 uncertainty band: decide exactly
            const double4 rj = a.sRs[f];
            const double r2 = __dadd_rn(__dadd_rn(strict_pbc_sq(ri.x, rj.x), strict_pbc_sq(ri.y, rj.y)),
                                        strict_pbc_sq(ri.z, rj.z));
            if (!(r2 < a.xRc2s)) continue;
          }
          bool ok = (__float_as_int(qf.w) != body_i) && (a.all_interact || a.interact[type_i * a.nt + a.sType[f]]);
          if (ok && x0 < x1) {
            const int atom_j = a.sMeta[f].x;
            for (int q = x0; ok && q < x1; ++q) ok = (a.exItem[q] != atom_j);
          }
          if (ok) {
            if (cnt < a.cap) out[(size_t)cnt * TILE] = f;
            ++cnt;
          }
        }
      }
    }
    a.nbrCount[e] = min(cnt, a.cap);
  }
  int mx = cnt;
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0 && mx > 0) {
    atomicMax(&a.flags[0], mx);
    if (mx > a.cap) a.flags[1] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-step refresh of sorted positions: pos = R/L + (image shift - floor at build time), w = charge.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_refresh_positions(int Next, double L, const double* __restrict__ R,
                                                           const double* __restrict__ q,
                                                           const int4* __restrict__ sMeta,
                                                           double4* __restrict__ pos) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Next) return;
  int4 m = sMeta[e];
  double4 p;
  p.x = __ddiv_rn(R[3 * (size_t)m.x], L) + (double)m.y;
  p.y = __ddiv_rn(R[3 * (size_t)m.x + 1], L) + (double)m.z;
  p.z = __ddiv_rn(R[3 * (size_t)m.x + 2], L) + (double)m.w;
  p.w = q[m.x];
  pos[e] = p;
}

// ------------------------------------------------------------------------------------------------
// K5: pair forces. One thread per real entry, full list, no atomics.
// ------------------------------------------------------------------------------------------------
struct ForceArgs {
  int Next, cap, nt;
  double Rc2s;      // cutoff^2 in scaled units (RcSq or InRcSq times invL2)
  double L, L2, invL, invL2;
  const double4* pos;
  const int* nbr;
  const int* nbrCount;
  const int4* sMeta;
  const unsigned char* sGhost;
  const int* sType;
  const double* delta;     // (3,N) offsets from the body centre of mass (rigid-body virial), or nullptr
  const PairEntry* tab;    // nt*nt (device)
  PairEntry single;        // the only entry when nt == 1
  nb::DevModel coul;
  int q4_quirk;            // virial-only + coul_none: Wij keeps the pair value (reference make_virial_compute.sh:24-29)
  double* F;               // (3,N) output, original atom order
  double* partial;         // gridDim.x * 5
  unsigned int* ticket;
  double* out;             // 5 scalars: Epair, Ecoul, Wpair, Wcoul, Wbody
};

// reciprocal to full double precision from the 20-bit hardware seed: cubic (two-term) refinement,
// relative error ~ e0^3 < 2^-57
__device__ __forceinline__ double fast_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  double t = fma(e, e, e);
  return fma(x, t, x);
}

// one 32-byte gather = one 256-bit load = one sector (LDG.E.256 on sm_100a)
__device__ __forceinline__ double4 ld_pos(const double4* p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

struct PairAcc {
  double fx = 0.0, fy = 0.0, fz = 0.0, Ep = 0.0, Ec = 0.0, Wp = 0.0, Wc = 0.0;
};

// One neighbor of atom i (position pi, type itype): cutoff test, pair model + modifier, optional Coulomb
// model + modifier, force accumulation (reference compute.f90:44-93). `f` is the neighbor's sorted entry,
// only used to look its type up when the system has several types.
template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE>
__device__ __forceinline__ void pair_term(const ForceArgs& a, const PairEntry* tab, const double4& pi, int itype,
                                          bool icharged, double c1, const double4& pj, int f, PairAcc& s) {
  constexpr bool LJ_FAST = SINGLE && PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  const bool has_coul = (CK == nb::K_DYNAMIC) ? true : (CK != nb::K_COUL_NONE);
  const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
  const double r2 = dx * dx + dy * dy + dz * dz;
  if (r2 < a.Rc2s) {
    if (LJ_FAST) {
      // plain single-type Lennard-Jones: unscaled sums, constants applied once per atom (lj_fast_scale)
      const double rinv = fast_rcp(r2);
      const double sr2 = c1 * rinv;
      const double sr6 = sr2 * sr2 * sr2;
      const double sr12 = sr6 * sr6;
      if (COMPUTE) s.Ep += sr12 - sr6;
      const double w = fma(2.0, sr12, -sr6);
      s.Wp += w;
      const double t = w * rinv;
      s.fx = fma(t, dx, s.fx);
      s.fy = fma(t, dy, s.fy);
      s.fz = fma(t, dz, s.fz);
    } else {
      nb::Dist D;
      if (NEED_INVR) {
        D.invR = rsqrt(r2) * a.invL;
        D.invR2 = D.invR * D.invR;
      } else {
        D.invR2 = fast_rcp(r2) * a.invL2;
        D.invR = 0.0;
      }
      D.r2 = r2 * a.L2;          // real-unit r^2 and r, so that no model body divides
      D.r = D.r2 * D.invR;
      const double invR2 = D.invR2;
      const PairEntry& pe = SINGLE ? a.single : tab[itype * a.nt + a.sType[f]];
      double E, W;
      nb::eval_kind<PK>(pe.model, D, E, W);
      nb::eval_modifier<PM>(pe.model, D, E, W);
      if (COMPUTE) s.Ep += E;
      s.Wp += W;
      double Wsum = W;
      if (has_coul) {
        if (icharged && fabs(pj.w) > DEPS && pe.coulomb) {
          double Eq, Wq;
          if (!COMPUTE && a.q4_quirk) {
            Eq = 0.0;
            Wq = W;
          } else {
            nb::eval_kind<CK>(a.coul, D, Eq, Wq);
            nb::eval_modifier<CM>(a.coul, D, Eq, Wq);
          }
          const double QiQj = pe.kCoul * pi.w * pj.w;
          if (COMPUTE) s.Ec += QiQj * Eq;
          Wq = QiQj * Wq;
          s.Wc += Wq;
          Wsum += Wq;
        }
      }
      const double t = Wsum * invR2;
      s.fx = fma(t, dx, s.fx);
      s.fy = fma(t, dy, s.fy);
      s.fz = fma(t, dz, s.fz);
    }
  }
}

// final per-atom scaling (F = L * sum, reference compute.f90:99) and store; returns the body-virial term
template <bool LJ_FAST>
__device__ __forceinline__ double finish_atom(const ForceArgs& a, int atom, PairAcc& s) {
  if (LJ_FAST) {
    const double fs = a.single.model.b * a.invL2 * a.L;   // eps24 * invL2 * L
    s.fx *= fs;
    s.fy *= fs;
    s.fz *= fs;
    s.Ep *= a.single.model.a;   // eps4
    s.Wp *= a.single.model.b;   // eps24
  } else {
    s.fx *= a.L;
    s.fy *= a.L;
    s.fz *= a.L;
  }
  a.F[3 * (size_t)atom] = s.fx;
  a.F[3 * (size_t)atom + 1] = s.fy;
  a.F[3 * (size_t)atom + 2] = s.fz;
  if (a.delta != nullptr)
    return -(s.fx * a.delta[3 * (size_t)atom] + s.fy * a.delta[3 * (size_t)atom + 1] + s.fz * a.delta[3 * (size_t)atom + 2]);
  return 0.0;
}

// block reduction of the five scalars (fixed shuffle tree + fixed warp order) followed by the grid finish
__device__ __forceinline__ void reduce_scalars(const ForceArgs& a, double Ep, double Ec, double Wp, double Wc, double Wb) {
  __shared__ double red[32][5];
  const int lane = threadIdx.x & 31;
  double v[5] = {Ep, Ec, Wp, Wc, Wb};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double x = v[q];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) red[threadIdx.x >> 5][q] = x;
  }
  __syncthreads();
  double mine[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q)
      for (int w = 0; w < nw; ++w) mine[q] += red[w][q];
  }
  grid_finish<5>(mine, a.partial, a.ticket, a.out, 0.5);   // pair sums halved: the full list holds i-j and j-i
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE, int UNROLL = 2, int THREADS = TPB,
          int MINBLOCKS = 1>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) k_pair_forces(const __grid_constant__ ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PairEntry* tab = a.tab;
  if (!SINGLE && a.nt <= MAX_SMEM_TYPES) {
    PairEntry* st = reinterpret_cast<PairEntry*>(smem_raw);
    const int words = a.nt * a.nt * (int)(sizeof(PairEntry) / sizeof(int));
    for (int w = threadIdx.x; w < words; w += blockDim.x)
      reinterpret_cast<int*>(st)[w] = reinterpret_cast<const int*>(a.tab)[w];
    __syncthreads();
    tab = st;
  }
  constexpr bool LJ_FAST = SINGLE && PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  PairAcc s;
  double Wb = 0.0;
  if (e < a.Next) {
    const int cnt = a.nbrCount[e];   // ghosts hold 0
    const double4 pi = a.pos[e];
    const int itype = SINGLE ? 0 : a.sType[e];
    const bool icharged = fabs(pi.w) > DEPS;
    const int* nb_ptr = a.nbr + ((size_t)(e >> 5) * a.cap) * TILE + lane;
    const double c1 = a.single.model.c * a.invL2;   // LJ_FAST: sr2 = sigsq * invL2 / r2
    int k = 0;
    for (; k + UNROLL <= cnt; k += UNROLL) {   // UNROLL gathers in flight before any is consumed
      int f[UNROLL];
      double4 p[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) f[u] = nb_ptr[(size_t)(k + u) * TILE];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) p[u] = ld_pos(a.pos + f[u]);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, p[u], f[u], s);
    }
    for (; k < cnt; ++k) {
      const int f0 = nb_ptr[(size_t)k * TILE];
      pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, ld_pos(a.pos + f0), f0, s);
    }
    if (!a.sGhost[e]) Wb = finish_atom<LJ_FAST>(a, a.sMeta[e].x, s);   // ghosts: no list (count 0), no force slot
  }
  reduce_scalars(a, s.Ep, s.Ec, s.Wp, s.Wc, Wb);
}

// ================================================================================================
// Rows path (opt-in, EMDEE_ROWS=G with G in {8,16,32}; not yet measured on a GPU): G lanes share ONE atom and
// take its neighbors G at a time, so the lanes of a gather read CONSECUTIVE entries of one row. Rows are
// ascending in the sorted entry index and the sorted order is cell-major, so consecutive row entries are mostly
// consecutive in memory: a warp-gather touches ~8-12 distinct 128-byte lines instead of ~26 when every lane
// follows its own atom (tools/lsu_probe.cu measures exactly this trade). The price is a
// G-lane shuffle reduction of the force per atom and a row-major copy of the list (k_transpose_rows, once per
// rebuild). Summation order differs from the default path, results agree to rounding.
// ================================================================================================
constexpr int ROWS_TILES_PER_BLOCK = 8;

// tile-major list (slot k of entry e at ((e/32)*cap + k)*32 + e%32) -> row-major (rows[e*pitch + k]), through
// shared memory so that both the reads and the writes are 128-byte coalesced
__global__ void __launch_bounds__(32 * ROWS_TILES_PER_BLOCK) k_transpose_rows(int Next, int cap, int pitch,
                                                                              const int* __restrict__ nbr,
                                                                              const int* __restrict__ nbrCount,
                                                                              int* __restrict__ rows) {
  __shared__ int tile[ROWS_TILES_PER_BLOCK][32][33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * ROWS_TILES_PER_BLOCK + w;   // tile of 32 entries (warp-uniform)
  const long long e = t * TILE + lane;
  const int cnt = (e < Next) ? nbrCount[e] : 0;
  int mx = cnt;
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  for (int k0 = 0; k0 < mx; k0 += 32) {
    for (int r = 0; r < 32 && k0 + r < mx; ++r)   // slot k0+r of the 32 entries: one coalesced 128-byte row
      tile[w][r][lane] = nbr[((size_t)t * cap + k0 + r) * TILE + lane];
    __syncwarp();
    for (int r = 0; r < 32; ++r) {                // entry r of the tile: its slots k0 .. k0+31
      const int c = __shfl_sync(0xffffffffu, cnt, r);
      const int k = k0 + lane;
      if (k < c) rows[(size_t)(t * TILE + r) * pitch + k] = tile[w][lane][r];
    }
    __syncwarp();
  }
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE, int G, int UNROLL>
__global__ void __launch_bounds__(256) k_pair_forces_rows(const __grid_constant__ ForceArgs a, int pitch,
                                                          const int* __restrict__ rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PairEntry* tab = a.tab;
  if (!SINGLE && a.nt <= MAX_SMEM_TYPES) {
    PairEntry* st = reinterpret_cast<PairEntry*>(smem_raw);
    const int words = a.nt * a.nt * (int)(sizeof(PairEntry) / sizeof(int));
    for (int w = threadIdx.x; w < words; w += blockDim.x)
      reinterpret_cast<int*>(st)[w] = reinterpret_cast<const int*>(a.tab)[w];
    __syncthreads();
    tab = st;
  }
  constexpr bool LJ_FAST = SINGLE && PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  constexpr int APW = 32 / G;   // atoms per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long e = warp * APW + lane / G;   // the sorted entry this lane's group works on
  const bool valid = e < a.Next;
  PairAcc s;
  double Wb = 0.0;
  if (valid) {
    const int cnt = a.nbrCount[e];   // ghosts hold 0
    if (cnt > 0) {
      const double4 pi = a.pos[e];
      const int itype = SINGLE ? 0 : a.sType[e];
      const bool icharged = fabs(pi.w) > DEPS;
      const int* row = rows + (size_t)e * pitch;
      const double c1 = a.single.model.c * a.invL2;
      int k = sub;
      for (; k + G * (UNROLL - 1) < cnt; k += G * UNROLL) {   // UNROLL gathers in flight per lane
        int f[UNROLL];
        double4 p[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) f[u] = row[k + G * u];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) p[u] = ld_pos(a.pos + f[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, p[u], f[u], s);
      }
      for (; k < cnt; k += G) {
        const int f0 = row[k];
        pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, ld_pos(a.pos + f0), f0, s);
      }
    }
  }
  // every lane of the warp arrives here: fold the G partial forces of each atom (fixed butterfly order)
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    s.fx += __shfl_xor_sync(0xffffffffu, s.fx, off);
    s.fy += __shfl_xor_sync(0xffffffffu, s.fy, off);
    s.fz += __shfl_xor_sync(0xffffffffu, s.fz, off);
  }
  if (LJ_FAST) {   // the energy / virial partials stay per lane: scale each (cf. finish_atom)
    s.Ep *= a.single.model.a;
    s.Wp *= a.single.model.b;
  }
  if (valid && sub == 0 && !a.sGhost[e]) {
    const double fs = LJ_FAST ? a.single.model.b * a.invL2 * a.L : a.L;
    const size_t atom = (size_t)a.sMeta[e].x;
    const double fx = s.fx * fs, fy = s.fy * fs, fz = s.fz * fs;
    a.F[3 * atom] = fx;
    a.F[3 * atom + 1] = fy;
    a.F[3 * atom + 2] = fz;
    if (a.delta != nullptr) Wb = -(fx * a.delta[3 * atom] + fy * a.delta[3 * atom + 1] + fz * a.delta[3 * atom + 2]);
  }
  reduce_scalars(a, s.Ep, s.Ec, s.Wp, s.Wc, Wb);
}

// ================================================================================================
// Duo path: one thread owns TWO consecutive sorted entries (same or adjacent cell) and walks the UNION of
// their neighbor rows, so a neighbor that both atoms see is gathered once. The LSU data path (one
// wavefront per distinct 32-byte sector of a divergent gather) is what binds the force kernel; the union
// of two neighboring atoms' lists is ~1.3 lists instead of 2, i.e. ~1/3 fewer gathers for the same pair
// arithmetic. Rows produced by k_build_list are ascending in the sorted entry index, so the union is a
// sorted merge (k_merge_duos, once per rebuild). Union entry = neighbor index | bit30 (first atom sees it)
// | bit31 (second atom sees it).
// ================================================================================================
constexpr unsigned int DUO_IDX = 0x3fffffffu, DUO_B0 = 0x40000000u, DUO_B1 = 0x80000000u;

__global__ void __launch_bounds__(TPB) k_merge_duos(int Next, int cap, int cap2, const int* __restrict__ nbr,
                                                    const int* __restrict__ nbrCount, unsigned int* __restrict__ duoNbr,
                                                    int* __restrict__ duoCount, int* __restrict__ flags) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int e0 = 2 * d, e1 = 2 * d + 1;
  int cnt = 0;
  if (e0 < Next) {
    const int c0 = nbrCount[e0], c1 = (e1 < Next) ? nbrCount[e1] : 0;
    const int* r0 = nbr + ((size_t)(e0 >> 5) * cap) * TILE + (e0 & 31);
    const int* r1 = nbr + ((size_t)(e1 >> 5) * cap) * TILE + (e1 & 31);
    unsigned int* out = duoNbr + ((size_t)(d >> 5) * cap2) * TILE + (d & 31);
    int k0 = 0, k1 = 0;
    int a = (k0 < c0) ? r0[0] : 0x7fffffff, b = (k1 < c1) ? r1[0] : 0x7fffffff;
    while (k0 < c0 || k1 < c1) {
      const int f = min(a, b);
      unsigned int v = (unsigned int)f;
      if (a == f) {
        v |= DUO_B0;
        ++k0;
        a = (k0 < c0) ? r0[(size_t)k0 * TILE] : 0x7fffffff;
      }
      if (b == f) {
        v |= DUO_B1;
        ++k1;
        b = (k1 < c1) ? r1[(size_t)k1 * TILE] : 0x7fffffff;
      }
      if (cnt < cap2) out[(size_t)cnt * TILE] = v;
      ++cnt;
    }
    duoCount[d] = min(cnt, cap2);
  }
  int mx = cnt;
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0 && mx > 0) {
    atomicMax(&flags[2], mx);
    if (mx > cap2) flags[3] = 1;
  }
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE>
__global__ void __launch_bounds__(TPB) k_pair_forces_duo(const __grid_constant__ ForceArgs a, int cap2,
                                                         const unsigned int* __restrict__ duoNbr,
                                                         const int* __restrict__ duoCount) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PairEntry* tab = a.tab;
  if (!SINGLE && a.nt <= MAX_SMEM_TYPES) {
    PairEntry* st = reinterpret_cast<PairEntry*>(smem_raw);
    const int words = a.nt * a.nt * (int)(sizeof(PairEntry) / sizeof(int));
    for (int w = threadIdx.x; w < words; w += blockDim.x)
      reinterpret_cast<int*>(st)[w] = reinterpret_cast<const int*>(a.tab)[w];
    __syncthreads();
    tab = st;
  }
  constexpr bool LJ_FAST = SINGLE && PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int e0 = 2 * d, e1 = 2 * d + 1;
  PairAcc s0, s1;
  double Wb = 0.0;
  if (e0 < a.Next) {
    const bool has1 = e1 < a.Next;
    const int cnt = duoCount[d];
    const double4 p0 = a.pos[e0];
    const double4 p1 = has1 ? a.pos[e1] : p0;
    const int t0 = SINGLE ? 0 : a.sType[e0];
    const int t1 = (SINGLE || !has1) ? 0 : a.sType[e1];
    const bool q0 = fabs(p0.w) > DEPS, q1 = fabs(p1.w) > DEPS;
    const unsigned int* row = duoNbr + ((size_t)(d >> 5) * cap2) * TILE + (d & 31);
    const double c1 = a.single.model.c * a.invL2;
    int k = 0;
    for (; k + 2 <= cnt; k += 2) {   // two gathers in flight before either is consumed
      const unsigned int va = row[(size_t)k * TILE];
      const unsigned int vb = row[(size_t)(k + 1) * TILE];
      const int fa = (int)(va & DUO_IDX), fb = (int)(vb & DUO_IDX);
      const double4 pa = ld_pos(a.pos + fa);
      const double4 pb = ld_pos(a.pos + fb);
      if (va & DUO_B0) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p0, t0, q0, c1, pa, fa, s0);
      if (va & DUO_B1) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p1, t1, q1, c1, pa, fa, s1);
      if (vb & DUO_B0) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p0, t0, q0, c1, pb, fb, s0);
      if (vb & DUO_B1) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p1, t1, q1, c1, pb, fb, s1);
    }
    if (k < cnt) {
      const unsigned int va = row[(size_t)k * TILE];
      const int fa = (int)(va & DUO_IDX);
      const double4 pa = ld_pos(a.pos + fa);
      if (va & DUO_B0) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p0, t0, q0, c1, pa, fa, s0);
      if (va & DUO_B1) pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, p1, t1, q1, c1, pa, fa, s1);
    }
    if (!a.sGhost[e0]) Wb += finish_atom<LJ_FAST>(a, a.sMeta[e0].x, s0);
    else s0 = PairAcc();
    if (has1 && !a.sGhost[e1]) Wb += finish_atom<LJ_FAST>(a, a.sMeta[e1].x, s1);
    else s1 = PairAcc();
  }
  reduce_scalars(a, s0.Ep + s1.Ep, s0.Ec + s1.Ec, s0.Wp + s1.Wp, s0.Wc + s1.Wc, Wb);
}

// ------------------------------------------------------------------------------------------------
// Cluster-2 path (opt-in, EMDEE_CLUSTER2=1; not yet measured on a GPU): one WARP owns the duo (2d, 2d+1): lanes
// 0-15 work for the first atom, lanes 16-31 for the second, and lane pair (s, s+16) reads the SAME entry of the
// duo's union row (row-major copy, k_transpose_rows), so a warp-gather touches 16 sectors (~7 lines) for up to 32
// pair terms, and unlike the duo kernel above no thread carries two atoms (no extra registers, no two-body
// divergence). A lane whose atom does not list the entry (mask bit clear) idles for that slot: ~68 % of the slots
// are useful (tools/gather_model.py). Force partials are folded over the 16 lanes of each atom by shuffles.
// ------------------------------------------------------------------------------------------------
template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE, int UNROLL>
__global__ void __launch_bounds__(256) k_pair_forces_cluster2(const __grid_constant__ ForceArgs a, int pitch,
                                                              const unsigned int* __restrict__ rows,
                                                              const int* __restrict__ duoCount) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PairEntry* tab = a.tab;
  if (!SINGLE && a.nt <= MAX_SMEM_TYPES) {
    PairEntry* st = reinterpret_cast<PairEntry*>(smem_raw);
    const int words = a.nt * a.nt * (int)(sizeof(PairEntry) / sizeof(int));
    for (int w = threadIdx.x; w < words; w += blockDim.x)
      reinterpret_cast<int*>(st)[w] = reinterpret_cast<const int*>(a.tab)[w];
    __syncthreads();
    tab = st;
  }
  constexpr bool LJ_FAST = SINGLE && PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  const int lane = threadIdx.x & 31;
  const int which = lane >> 4, sub = lane & 15;
  const unsigned int mybit = which ? DUO_B1 : DUO_B0;
  const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // duo of this warp
  const long long e = 2 * d + which;
  const bool valid = e < a.Next;
  PairAcc s;
  double Wb = 0.0;
  if (valid) {
    const int cnt = duoCount[d];   // 2d < Next whenever e is valid
    if (cnt > 0) {
      const double4 pi = a.pos[e];
      const int itype = SINGLE ? 0 : a.sType[e];
      const bool icharged = fabs(pi.w) > DEPS;
      const unsigned int* row = rows + (size_t)d * pitch;
      const double c1 = a.single.model.c * a.invL2;
      int k = sub;
      for (; k + 16 * (UNROLL - 1) < cnt; k += 16 * UNROLL) {
        unsigned int v[UNROLL];
        double4 p[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = row[k + 16 * u];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) p[u] = ld_pos(a.pos + (v[u] & DUO_IDX));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (v[u] & mybit)
            pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, p[u], (int)(v[u] & DUO_IDX), s);
      }
      for (; k < cnt; k += 16) {
        const unsigned int v0 = row[k];
        const double4 p0 = ld_pos(a.pos + (v0 & DUO_IDX));
        if (v0 & mybit)
          pair_term<PK, PM, CK, CM, SINGLE, NEED_INVR, COMPUTE>(a, tab, pi, itype, icharged, c1, p0, (int)(v0 & DUO_IDX), s);
      }
    }
  }
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) {   // fold the 16 partial forces of each atom (fixed butterfly order)
    s.fx += __shfl_xor_sync(0xffffffffu, s.fx, off);
    s.fy += __shfl_xor_sync(0xffffffffu, s.fy, off);
    s.fz += __shfl_xor_sync(0xffffffffu, s.fz, off);
  }
  if (LJ_FAST) {
    s.Ep *= a.single.model.a;
    s.Wp *= a.single.model.b;
  }
  if (valid && sub == 0 && !a.sGhost[e]) {
    const double fs = LJ_FAST ? a.single.model.b * a.invL2 * a.L : a.L;
    const size_t atom = (size_t)a.sMeta[e].x;
    const double fx = s.fx * fs, fy = s.fy * fs, fz = s.fz * fs;
    a.F[3 * atom] = fx;
    a.F[3 * atom + 1] = fy;
    a.F[3 * atom + 2] = fz;
    if (a.delta != nullptr) Wb = -(fx * a.delta[3 * atom] + fy * a.delta[3 * atom + 1] + fz * a.delta[3 * atom + 2]);
  }
  reduce_scalars(a, s.Ep, s.Ec, s.Wp, s.Wc, Wb);
}

// ================================================================================================
// Brick path (single-type systems whose cell occupancy fits): the real cells are tiled by bricks of
// about b^3 cells; one CTA owns a brick, stages the positions of the brick plus its 2-cell halo in
// shared memory with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier: the
// TMA path, one copy per contiguous x-run of cells), and gathers neighbors from shared memory through
// 16-bit brick-local indices. Versus the global path: a divergent gather costs shared-memory bank
// conflicts instead of one LSU wavefront per 32-byte sector, and the list is half the bytes.
// ================================================================================================
constexpr int BRICK_MAX_SEG = 144;    // (b+4)^2 staged x-runs, b <= 8
constexpr int BRICK_MAX_ROWS = 64;    // b^2 owned x-runs
constexpr int BRICK_TPB = 640;        // upper bound of the brick kernels' block size
constexpr int BRICK_SMAX = 3328;      // staged entries per brick (x 32 B = 104 KB of shared memory)

struct BrickGrid {
  int M, Mx;
  int nbx, nby, nbz;   // bricks per dimension; brick i covers real cells [floor(i*M/nb), floor((i+1)*M/nb))
};

struct BrickDesc {
  int nseg, nrows, S, B;           // staged runs, owned runs, staged entries, owned (real) entries
  int segG[BRICK_MAX_SEG];         // first global sorted entry of each staged run
  int segL[BRICK_MAX_SEG + 1];     // prefix sum of run lengths = local index of each run's first entry
  int rowG[BRICK_MAX_ROWS];        // first global entry of each owned run
  int rowL[BRICK_MAX_ROWS];        // its local (staged) index
  int rowT[BRICK_MAX_ROWS + 1];    // prefix sum of owned-run lengths = first owned-atom ordinal of the run
};

__device__ __forceinline__ void brick_range(int i, int nb, int M, int& c0, int& c1) {
  c0 = 2 + (int)(((long long)i * M) / nb);
  c1 = 2 + (int)(((long long)(i + 1) * M) / nb);
}

__global__ void __launch_bounds__(TPB) k_brick_setup(BrickGrid g, const int* __restrict__ cellStart,
                                                     BrickDesc* __restrict__ desc, int* __restrict__ flags) {
  __shared__ int len[BRICK_MAX_SEG];
  __shared__ int rlen[BRICK_MAX_ROWS];
  const int brick = blockIdx.x;
  const int ix = brick % g.nbx, iy = (brick / g.nbx) % g.nby, iz = brick / (g.nbx * g.nby);
  int x0, x1, y0, y1, z0, z1;
  brick_range(ix, g.nbx, g.M, x0, x1);
  brick_range(iy, g.nby, g.M, y0, y1);
  brick_range(iz, g.nbz, g.M, z0, z1);
  const int nys = (y1 - y0) + 4, nzs = (z1 - z0) + 4, nseg = nys * nzs;
  const int nyr = (y1 - y0), nzr = (z1 - z0), nrows = nyr * nzr;
  BrickDesc& d = desc[brick];
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    const int y = y0 - 2 + (s % nys), z = z0 - 2 + (s / nys);
    const int row = g.Mx * (y + g.Mx * z);
    const int a = cellStart[row + x0 - 2], b = cellStart[row + x1 + 2];
    d.segG[s] = a;
    len[s] = b - a;
  }
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
    const int y = y0 + (r % nyr), z = z0 + (r / nyr);
    const int row = g.Mx * (y + g.Mx * z);
    const int a = cellStart[row + x0], b = cellStart[row + x1];
    d.rowG[r] = a;
    rlen[r] = b - a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < nseg; ++s) {
      d.segL[s] = acc;
      acc += len[s];
    }
    d.segL[nseg] = acc;
    d.S = acc;
    d.nseg = nseg;
    int t = 0;
    for (int r = 0; r < nrows; ++r) {
      d.rowT[r] = t;
      t += rlen[r];
      // local index of the owned run = local start of its staged run + offset of x0 inside that run
      const int sy = (r % nyr) + 2, sz = (r / nyr) + 2, sidx = sy + nys * sz;
      d.rowL[r] = d.segL[sidx] + (d.rowG[r] - d.segG[sidx]);
    }
    d.rowT[nrows] = t;
    d.B = t;
    d.nrows = nrows;
    atomicMax(&flags[2], acc);
    atomicMax(&flags[3], t);
  }
}

// ---- mbarrier + bulk-copy helpers -------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned int phase) {
  unsigned int ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}

// stage `elem_bytes`-sized records of all runs of a brick into shared memory (one bulk copy per run)
__device__ __forceinline__ void brick_stage(const BrickDesc& d, const int* sSegG, const int* sSegL, const void* src,
                                            void* dst, int elem_bytes, unsigned long long* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) mbar_expect_tx(bar, (unsigned int)d.S * (unsigned int)elem_bytes);
  for (int s = threadIdx.x; s < d.nseg; s += blockDim.x) {
    const int n = sSegL[s + 1] - sSegL[s];
    if (n > 0)
      bulk_g2s(reinterpret_cast<char*>(dst) + (size_t)sSegL[s] * elem_bytes,
               reinterpret_cast<const char*>(src) + (size_t)sSegG[s] * elem_bytes, (unsigned int)n * elem_bytes, bar);
  }
  while (!mbar_try_wait(bar, 0u)) {
  }
}

// owned-atom ordinal b -> (global sorted entry, local staged index)
__device__ __forceinline__ void brick_locate(const int* sRowT, const int* sRowG, const int* sRowL, int nrows, int b,
                                             int& e, int& li) {
  int lo = 0, hi = nrows - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (sRowT[mid] <= b) lo = mid;
    else hi = mid - 1;
  }
  const int off = b - sRowT[lo];
  e = sRowG[lo] + off;
  li = sRowL[lo] + off;
}

struct BrickArgs {
  BrickGrid g;
  const BrickDesc* desc;
  unsigned short* nbr16;   // [brick][slot][Bmax]
  int cap, Bmax;
};

// ---- list build, brick version ----------------------------------------------------------------------
__global__ void __launch_bounds__(BRICK_TPB) k_build_list_brick(const __grid_constant__ BuildArgs a,
                                                                const __grid_constant__ BrickArgs k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sSegG[BRICK_MAX_SEG], sSegL[BRICK_MAX_SEG + 1];
  __shared__ int sRowG[BRICK_MAX_ROWS], sRowL[BRICK_MAX_ROWS], sRowT[BRICK_MAX_ROWS + 1];
  __shared__ __align__(8) unsigned long long bar;
  float4* sPos = reinterpret_cast<float4*>(smem_raw);
  const int brick = blockIdx.x;
  const BrickDesc& d = k.desc[brick];
  const int nseg = d.nseg, nrows = d.nrows, B = d.B;
  for (int s = threadIdx.x; s <= nseg; s += blockDim.x) {
    sSegL[s] = d.segL[s];
    if (s < nseg) sSegG[s] = d.segG[s];
  }
  for (int r = threadIdx.x; r <= nrows; r += blockDim.x) {
    sRowT[r] = d.rowT[r];
    if (r < nrows) {
      sRowG[r] = d.rowG[r];
      sRowL[r] = d.rowL[r];
    }
  }
  __syncthreads();
  brick_stage(d, sSegG, sSegL, a.sPosF, sPos, (int)sizeof(float4), &bar);

  const int ix = brick % k.g.nbx, iy = (brick / k.g.nbx) % k.g.nby, iz = brick / (k.g.nbx * k.g.nby);
  int bx0, bx1, by0, by1, bz0, bz1;
  brick_range(ix, k.g.nbx, k.g.M, bx0, bx1);
  brick_range(iy, k.g.nby, k.g.M, by0, by1);
  brick_range(iz, k.g.nbz, k.g.M, bz0, bz1);
  const int nys = (by1 - by0) + 4;
  const int Mx = a.g.Mx;
  const float w = 1.0f / (float)a.g.M;
  const float slack = 1.0e-5f * w + 4.0e-7f;
  const float rc = (float)a.xRcs + slack;
  const float rc2 = rc * rc;
  int mxcnt = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int e, li;
    brick_locate(sRowT, sRowG, sRowL, nrows, b, e, li);
    unsigned short* out = k.nbr16 + ((size_t)brick * k.cap) * k.Bmax + b;
    const int atom_i = a.sMeta[e].x;
    const int body_i = a.sBody[e];
    const double4 ri = a.sRs[e];
    const float4 pf = sPos[li];
    const int x0 = a.exFirst[atom_i], x1 = a.exFirst[atom_i + 1];
    const int cell = a.sCell[e];
    const int ez = cell / (Mx * Mx), ey = (cell - ez * Mx * Mx) / Mx, ex = cell - Mx * (ey + Mx * ez);
    int cnt = 0;
    for (int dz = -2; dz <= 2; ++dz) {
      const float zlo = (float)(ez + dz - 2 + a.g.z0) * w, zhi = zlo + w;   // local layer -> global coordinate
      const float gz = fmaxf(0.0f, fmaxf(zlo - pf.z, pf.z - zhi) - slack);
      for (int dy = -2; dy <= 2; ++dy) {
        const float ylo = (float)(ey + dy - 2) * w, yhi = ylo + w;
        const float gy = fmaxf(0.0f, fmaxf(ylo - pf.y, pf.y - yhi) - slack);
        const float rem = rc2 - gz * gz - gy * gy;
        if (rem <= 0.0f) continue;
        const float hx = sqrtf(rem) + slack;
        int cl = (int)floorf((pf.x - hx) * (float)a.g.M) + 2;
        int ch = (int)floorf((pf.x + hx) * (float)a.g.M) + 2;
        cl = max(cl, ex - 2);
        ch = min(ch, ex + 2);
        const int row = Mx * ((ey + dy) + Mx * (ez + dz));
        const int f0 = a.cellStart[row + cl], f1 = a.cellStart[row + ch + 1];
        const int sidx = (ey + dy - (by0 - 2)) + nys * (ez + dz - (bz0 - 2));
        const int toLocal = sSegL[sidx] - sSegG[sidx];
        for (int f = f0; f < f1; ++f) {
          const float4 qf = sPos[f + toLocal];
          const float dxf = pf.x - qf.x, dyf = pf.y - qf.y, dzf = pf.z - qf.z;
          const float r2f = fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf));
          if (r2f > a.r2_reject) continue;
          if (f == e) continue;
          if (r2f >= a.r2_accept) {
            const double4 rj = a.sRs[f];
            const double r2 = __dadd_rn(__dadd_rn(strict_pbc_sq(ri.x, rj.x), strict_pbc_sq(ri.y, rj.y)),
                                        strict_pbc_sq(ri.z, rj.z));
            if (!(r2 < a.xRc2s)) continue;
          }
          bool ok = (__float_as_int(qf.w) != body_i) && a.all_interact;   // brick path: single type
          if (ok && x0 < x1) {
            const int atom_j = a.sMeta[f].x;
            for (int q = x0; ok && q < x1; ++q) ok = (a.exItem[q] != atom_j);
          }
          if (ok) {
            if (cnt < k.cap) out[(size_t)cnt * k.Bmax] = (unsigned short)(f + toLocal);
            ++cnt;
          }
        }
      }
    }
    a.nbrCount[e] = min(cnt, k.cap);
    mxcnt = max(mxcnt, cnt);
  }
  for (int off = 16; off > 0; off >>= 1) mxcnt = max(mxcnt, __shfl_xor_sync(0xffffffffu, mxcnt, off));
  if ((threadIdx.x & 31) == 0 && mxcnt > 0) {
    atomicMax(&a.flags[0], mxcnt);
    if (mxcnt > k.cap) a.flags[1] = 1;
  }
}

// ---- pair forces, brick version -----------------------------------------------------------------------
template <int PK, int PM, int CK, int CM, bool NEED_INVR, bool COMPUTE>
__global__ void __launch_bounds__(BRICK_TPB, (PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE) ? 2 : 1)
    k_pair_forces_brick(const __grid_constant__ ForceArgs a,
                                                                 const __grid_constant__ BrickArgs k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sSegG[BRICK_MAX_SEG], sSegL[BRICK_MAX_SEG + 1];
  __shared__ int sRowG[BRICK_MAX_ROWS], sRowL[BRICK_MAX_ROWS], sRowT[BRICK_MAX_ROWS + 1];
  __shared__ __align__(8) unsigned long long bar;
  double4* sPos = reinterpret_cast<double4*>(smem_raw);
  constexpr bool LJ_FAST = PK == nb::K_PAIR_LJ_CUT && PM == nb::M_NONE && CK == nb::K_COUL_NONE;
  const int brick = blockIdx.x;
  const BrickDesc& d = k.desc[brick];
  const int nseg = d.nseg, nrows = d.nrows, B = d.B;
  for (int s = threadIdx.x; s <= nseg; s += blockDim.x) {
    sSegL[s] = d.segL[s];
    if (s < nseg) sSegG[s] = d.segG[s];
  }
  for (int r = threadIdx.x; r <= nrows; r += blockDim.x) {
    sRowT[r] = d.rowT[r];
    if (r < nrows) {
      sRowG[r] = d.rowG[r];
      sRowL[r] = d.rowL[r];
    }
  }
  __syncthreads();
  brick_stage(d, sSegG, sSegL, a.pos, sPos, (int)sizeof(double4), &bar);

  double Ep = 0.0, Ec = 0.0, Wp = 0.0, Wc = 0.0, Wb = 0.0;
  const double c1 = a.single.model.c * a.invL2;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int e, li;
    brick_locate(sRowT, sRowG, sRowL, nrows, b, e, li);
    const int cnt = a.nbrCount[e];
    const double4 pi = sPos[li];
    const bool icharged = fabs(pi.w) > DEPS;
    const unsigned short* nb_ptr = k.nbr16 + ((size_t)brick * k.cap) * k.Bmax + b;
    PairAcc s;
    int q = 0;
    for (; q + 2 <= cnt; q += 2) {
      const int l0 = nb_ptr[(size_t)q * k.Bmax];
      const int l1 = nb_ptr[(size_t)(q + 1) * k.Bmax];
      const double4 p0 = sPos[l0];
      const double4 p1 = sPos[l1];
      pair_term<PK, PM, CK, CM, true, NEED_INVR, COMPUTE>(a, nullptr, pi, 0, icharged, c1, p0, 0, s);
      pair_term<PK, PM, CK, CM, true, NEED_INVR, COMPUTE>(a, nullptr, pi, 0, icharged, c1, p1, 0, s);
    }
    if (q < cnt) {
      const double4 p0 = sPos[nb_ptr[(size_t)q * k.Bmax]];
      pair_term<PK, PM, CK, CM, true, NEED_INVR, COMPUTE>(a, nullptr, pi, 0, icharged, c1, p0, 0, s);
    }
    Wb += finish_atom<LJ_FAST>(a, a.sMeta[e].x, s);
    Ep += s.Ep;
    Ec += s.Ec;
    Wp += s.Wp;
    Wc += s.Wc;
  }
  reduce_scalars(a, Ep, Ec, Wp, Wc, Wb);
}

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
__device__ __forceinline__ void store12(double* __restrict__ p, long long a0, int n, const double (&v)[3 * APT]) {
  if (vec_ok(p, a0, n)) {
    double4* q = reinterpret_cast<double4*>(p + 3 * a0);
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k] = make_double4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else {
    for (int k = 0; k < 3 * n; ++k) p[3 * a0 + k] = v[k];
  }
}

__global__ void __launch_bounds__(TPB) k_boost(int N, double CP, double CF, double* __restrict__ P,
                                               const double* __restrict__ F, const double* __restrict__ invMass,
                                               const unsigned char* __restrict__ owned, int want_ke,
                                               double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                               double* __restrict__ out) {
  __shared__ double red[TPB / 32][3];
  const long long a0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * APT;
  const int n = (a0 >= N) ? 0 : (int)min((long long)APT, N - a0);
  double k[3] = {0.0, 0.0, 0.0};
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
          }
        }
      }
      store12(P, a0, n, p);   // non-owned slots are written back unchanged
    }
  }
  if (!want_ke) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    double v = k[x];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[threadIdx.x >> 5][x] = v;
  }
  __syncthreads();
  double mine[3] = {0.0, 0.0, 0.0};
  if (threadIdx.x == 0) {
#pragma unroll
    for (int x = 0; x < 3; ++x)
      for (int w = 0; w < TPB / 32; ++w) mine[x] += red[w][x];
  }
  grid_finish<3>(mine, partial, ticket, out, 1.0);
}

// R = CR*R + CP*P/m, fused with the rebuild criterion on the NEW coordinates (|R - R0|^2 ordered scan).
// Multi-GPU: only owned atoms move here and the criterion is evaluated by the distributed kernels below
// (partial == nullptr skips the fused scan).
__global__ void __launch_bounds__(TPB) k_displace(int N, double CR, double CP, double* __restrict__ R,
                                                  const double* __restrict__ P, const double* __restrict__ invMass,
                                                  const unsigned char* __restrict__ owned,
                                                  const double* __restrict__ R0, MaxNext* __restrict__ partial,
                                                  unsigned int* __restrict__ ticket, double* __restrict__ result) {
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
#pragma unroll
      for (int j = 0; j < APT; ++j) {
        if (own[j]) {
          const double im = invMass[a0 + j];
#pragma unroll
          for (int x = 0; x < 3; ++x)
            r[3 * j + x] = __dadd_rn(__dmul_rn(CR, r[3 * j + x]), __dmul_rn(__dmul_rn(CP, p[3 * j + x]), im));
        }
      }
      store12(R, a0, n, r);
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

// ------------------------------------------------------------------------------------------------
// Extension kernels: export the pair set; count interacting entries.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_export_pairs(int Next, int cap, const int* __restrict__ nbr,
                                                      const int* __restrict__ nbrCount, const int4* __restrict__ sMeta,
                                                      int* __restrict__ pairs, long long capacity,
                                                      unsigned long long* __restrict__ counter) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Next) return;
  int cnt = nbrCount[e];
  const int lane = threadIdx.x & 31;
  const int ai = sMeta[e].x;
  const int* p = nbr + ((size_t)(e >> 5) * cap) * TILE + lane;
  for (int k = 0; k < cnt; ++k) {
    int aj = sMeta[p[(size_t)k * TILE]].x;
    if (ai < aj) {
      unsigned long long slot = atomicAdd(counter, 1ull);
      if (pairs != nullptr && (long long)slot < capacity) {
        pairs[2 * slot] = ai;
        pairs[2 * slot + 1] = aj;
      }
    }
  }
}

__global__ void __launch_bounds__(TPB) k_count_interacting(int Next, int cap, double Rc2s, const double4* __restrict__ pos,
                                                           const int* __restrict__ nbr, const int* __restrict__ nbrCount,
                                                           unsigned long long* __restrict__ counter) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n = 0;
  if (e < Next) {
    int cnt = nbrCount[e];
    const int lane = threadIdx.x & 31;
    const double4 pi = pos[e];
    const int* p = nbr + ((size_t)(e >> 5) * cap) * TILE + lane;
    for (int k = 0; k < cnt; ++k) {
      double4 pj = pos[p[(size_t)k * TILE]];
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      if (dx * dx + dy * dy + dz * dz < Rc2s) ++n;
    }
  }
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

// brick-local staged index -> global sorted entry
__device__ __forceinline__ int brick_to_global(const BrickDesc& d, int lf) {
  int lo = 0, hi = d.nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (d.segL[mid] <= lf) lo = mid;
    else hi = mid - 1;
  }
  return d.segG[lo] + (lf - d.segL[lo]);
}

// mode 0: export pairs (ai < aj); mode 1: count entries with r^2 < Rc2s
__global__ void __launch_bounds__(TPB) k_brick_list_walk(BrickArgs k, int mode, const int* __restrict__ nbrCount,
                                                         const int4* __restrict__ sMeta, const double4* __restrict__ pos,
                                                         double Rc2s, int* __restrict__ pairs, long long capacity,
                                                         unsigned long long* __restrict__ counter) {
  const int brick = blockIdx.x;
  const BrickDesc& d = k.desc[brick];
  unsigned long long n = 0;
  for (int b = threadIdx.x; b < d.B; b += blockDim.x) {
    int lo = 0, hi = d.nrows - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (d.rowT[mid] <= b) lo = mid;
      else hi = mid - 1;
    }
    const int e = d.rowG[lo] + (b - d.rowT[lo]);
    const int cnt = nbrCount[e];
    const int ai = sMeta[e].x;
    const double4 pi = pos[e];
    const unsigned short* p = k.nbr16 + ((size_t)brick * k.cap) * k.Bmax + b;
    for (int q = 0; q < cnt; ++q) {
      const int f = brick_to_global(d, p[(size_t)q * k.Bmax]);
      if (mode == 0) {
        const int aj = sMeta[f].x;
        if (ai < aj) {
          unsigned long long slot = atomicAdd(counter, 1ull);
          if (pairs != nullptr && (long long)slot < capacity) {
            pairs[2 * slot] = ai;
            pairs[2 * slot + 1] = aj;
          }
        }
      } else {
        const double4 pj = pos[f];
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        if (dx * dx + dy * dy + dz * dz < Rc2s) ++n;
      }
    }
  }
  if (mode == 1 && n) atomicAdd(counter, n);
}

// ------------------------------------------------------------------------------------------------
// Pair-distance histogram over the resident list (reference count_pairs, src/EmDeeCode.f90:1346-1388).
// One thread per real entry walks its row of the FULL list, so every pair is met twice (host halves the
// integer counts). Block-private shared-memory histogram when it fits, flushed with 64-bit global atomics:
// integer sums, so the result does not depend on the order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_rdf(int Next, int cap, int nt, int bins, int nsym, double Rc2s, double binsByRcS,
                                             const double4* __restrict__ pos, const int* __restrict__ nbr,
                                             const int* __restrict__ nbrCount, const int* __restrict__ sType,
                                             const unsigned short* __restrict__ pairSym, int use_smem,
                                             unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned int* local = reinterpret_cast<unsigned int*>(smem_raw);
  const int nbin = bins * nsym;
  if (use_smem) {
    for (int q = threadIdx.x; q < nbin; q += blockDim.x) local[q] = 0u;
    __syncthreads();
  }
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < Next) {
    const int cnt = nbrCount[e];   // ghosts hold 0
    if (cnt > 0) {
      const int it = sType[e];
      const double4 pi = pos[e];
      const int* row = nbr + ((size_t)(e >> 5) * cap) * TILE + (e & 31);
      for (int k = 0; k < cnt; ++k) {
        const int f = row[(size_t)k * TILE];
        const int sym = pairSym[it * nt + sType[f]];
        if (sym == 0) continue;
        const double4 pj = ld_pos(pos + f);
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < Rc2s) {
          const int bin = (int)(sqrt(r2) * binsByRcS);
          if (bin < bins) {
            const int slot = (sym - 1) * bins + bin;
            if (use_smem) atomicAdd(&local[slot], 1u);
            else atomicAdd(&hist[slot], 1ull);
          }
        }
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int q = threadIdx.x; q < nbin; q += blockDim.x)
      if (local[q] != 0u) atomicAdd(&hist[q], (unsigned long long)local[q]);
  }
}

// ---- FP64 issue-rate microbenchmark (roofline denominator that MEASURED_PEAKS.json does not carry) ----
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;   // keep the chain alive
}

}  // namespace

// =================================================================================================
struct Engine::Impl {
  int N = 0, nt = 1, nlayers = 1, nbodies = 0;
  double Rc = 0, skin = 0, RcSq = 0, xRc = 0, xRcSq = 0, skinSq = 0, InRcSq = 0;
  int device = 0;
  cudaStream_t stream = nullptr;   // legacy default stream: visible to the caller's own CUDA events

  // original-order state
  DBuf<double> R, P, R0, F, q, invMass, delta;
  DBuf<int> type, body;
  bool has_delta = false, has_R = false, any_charged = false;
  DBuf<int> exFirst, exItem;
  DBuf<unsigned char> interact;
  bool all_interact = false;
  std::vector<LayerTable> layers;
  std::vector<DBuf<PairEntry>> tabs;

  // rebuild artifacts
  GridDesc grid{0, 0, 0, 0, 0};
  int Next = 0, cap = 0;
  DBuf<double> Rs;
  DBuf<double4> sRs;
  DBuf<float4> sPosF;
  DBuf<int> atomCell, atomFloor, cellCount, cellStart, cellFill, slotAtom, slotImg, slotCell;
  DBuf<int4> sMeta;
  DBuf<int> sCell, sType, sBody, nbr, nbrCount, flags;
  DBuf<unsigned char> sGhost;
  DBuf<double4> pos;
  DBuf<unsigned char> scanTmp;
  size_t scanTmpBytes = 0;
  bool list_valid = false;

  // duo path: union rows of consecutive entry pairs
  bool use_duos = false;
  int cap2 = 0;
  int rows_group = 0;   // EMDEE_ROWS: lanes per atom of the rows path (0 = off)
  bool use_cluster2 = false;   // EMDEE_CLUSTER2: warp-per-duo kernel over the row-major union rows
  int rows_pitch = 0;
  DBuf<int> rowsNbr;
  DBuf<unsigned int> duoNbr;
  DBuf<int> duoCount;

  // brick path (single-type systems): per-brick descriptors and the 16-bit brick-local list
  bool use_bricks = false;
  BrickGrid bgrid{0, 0, 0, 0, 0};
  int nbricks = 0, Bmax = 0, brick_threads = 0;
  DBuf<BrickDesc> bdesc;
  DBuf<unsigned short> nbr16;

  // multi-GPU (one rank per GPU, z-slabs): NCCL is loaded lazily, only when EmDeeX_comm_init is called
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  DBuf<unsigned char> owned;       // per atom: this rank integrates it and computes its force
  bool owned_valid = false;        // false until the first distributed rebuild (all ranks hold full arrays)
  bool halo_fresh = true;          // halo positions are current (false after displace)
  DBuf<unsigned char> known;       // per atom: this rank holds a current position (owned, halo or just received)
  bool all_known = true;           // every rank holds full, current arrays (after uploads / downloads)
  bool p_partial = false;          // momenta of non-owned atoms may be stale (coordinates were uploaded alone)
  DBuf<int> migList[2], migCounts;
  DBuf<double> migSend[2], migRecv[2];
  DBuf<double> scratch3;           // 3N doubles: masked copies for the all-reduce that rebuilds full arrays
  DBuf<unsigned char> haloFlags;   // 4N
  DBuf<int> haloList[4];           // send up, send down, receive from below, receive from above
  int haloCount[4] = {0, 0, 0, 0};
  DBuf<double> haloBuf[4];
  DBuf<int> selCount;
  DBuf<MaxIdx> miPartial, miResult;
  MaxIdx* h_mi = nullptr;          // pinned, world entries

  // reductions
  DBuf<MaxNext> chkPartial;
  DBuf<double> partial, scalars;
  DBuf<unsigned int> tickets;      // [0] force kernel, [1] boost, [2] displacement check
  DBuf<unsigned long long> counter;
  double* h_scalars = nullptr;     // pinned: [0..4] force scalars, [8] rebuild criterion, [10..12] kinetic
  bool check_cached = false;       // h_scalars[8] (after check_event) holds the criterion for the current R
  // host-side phase timers (printed by the destructor when EMDEE_PROFILE is set)
  double t_check = 0, t_rebuild = 0, t_force = 0, t_boost = 0, t_displace = 0;
  long long n_force = 0, n_boost = 0, n_displace = 0, n_rebuild = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, check_event = nullptr;
};

// ---- NCCL, loaded lazily (a single-GPU client never needs libnccl) ------------------------------------
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.handle == nullptr) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle == nullptr) fatal("multi-GPU setup", "libnccl.so.2 could not be loaded");
    auto sym = [&](const char* n) {
      void* p = dlsym(api.handle, n);
      if (p == nullptr) fatal("multi-GPU setup", "a required NCCL symbol is missing");
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  }
  return api;
}
#define NCCL_CHECK(call)                                                                              \
  do {                                                                                                \
    ncclResult_t r__ = (call);                                                                        \
    if (r__ != ncclSuccess) {                                                                         \
      std::fprintf(stderr, "Error in NCCL: %s (%s:%d).\n", nccl().GetErrorString(r__), __FILE__, __LINE__); \
      std::exit(1);                                                                                   \
    }                                                                                                 \
  } while (0)
}  // namespace

Engine::Engine(int natoms, int ntypes, int nlayers, double Rc, double skin, const int* atomType1,
               const double* mass, const double* invMass, const int* atomBody, int nbodies) {
  d_ = new Impl();
  Impl& s = *d_;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    fatal("system setup", "no CUDA device is available (this library has no CPU fallback)");
  const char* env = std::getenv("EMDEE_DEVICE");
  if (env == nullptr) env = std::getenv("LOCAL_RANK");
  s.device = env ? std::atoi(env) % ndev : 0;
  CUDA_CHECK(cudaSetDevice(s.device));
  s.N = natoms;
  s.nt = ntypes;
  s.nlayers = nlayers;
  s.nbodies = nbodies;
  s.Rc = Rc;
  s.skin = skin;
  s.RcSq = Rc * Rc;
  s.xRc = Rc + skin;
  s.xRcSq = s.xRc * s.xRc;
  s.skinSq = skin * skin;
  s.InRcSq = s.RcSq;
  const size_t n3 = 3 * (size_t)natoms;
  s.R.ensure(n3);
  s.P.ensure(n3);
  s.R0.ensure(n3);
  s.F.ensure(n3 * nlayers);
  s.q.ensure(natoms);
  s.invMass.ensure(natoms);
  s.type.ensure(natoms);
  s.body.ensure(natoms);
  CUDA_CHECK(cudaMemset(s.P.p, 0, n3 * sizeof(double)));
  CUDA_CHECK(cudaMemset(s.R0.p, 0, n3 * sizeof(double)));
  CUDA_CHECK(cudaMemset(s.F.p, 0, n3 * nlayers * sizeof(double)));
  CUDA_CHECK(cudaMemset(s.q.p, 0, natoms * sizeof(double)));
  std::vector<int> t0(natoms);
  for (int i = 0; i < natoms; ++i) t0[i] = atomType1[i] - 1;
  CUDA_CHECK(cudaMemcpy(s.type.p, t0.data(), natoms * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.body.p, atomBody, natoms * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.invMass.p, invMass, natoms * sizeof(double), cudaMemcpyHostToDevice));
  (void)mass;
  s.exFirst.ensure(natoms + 1);
  CUDA_CHECK(cudaMemset(s.exFirst.p, 0, (natoms + 1) * sizeof(int)));
  s.exItem.ensure(1);
  s.interact.ensure((size_t)ntypes * ntypes);
  CUDA_CHECK(cudaMemset(s.interact.p, 0, (size_t)ntypes * ntypes));
  s.layers.resize(nlayers);
  s.tabs.resize(nlayers);
  s.flags.ensure(4);
  s.scalars.ensure(16);
  s.counter.ensure(2);
  s.tickets.ensure(4);
  CUDA_CHECK(cudaMemset(s.tickets.p, 0, 4 * sizeof(unsigned int)));
  s.chkPartial.ensure(nblocks(natoms));
  s.partial.ensure((size_t)nblocks(natoms) * 5);
  CUDA_CHECK(cudaMallocHost(&s.h_scalars, 16 * sizeof(double)));
  CUDA_CHECK(cudaEventCreate(&s.ev0));
  CUDA_CHECK(cudaEventCreate(&s.ev1));
  CUDA_CHECK(cudaEventCreateWithFlags(&s.check_event, cudaEventDisableTiming));
  stats_.device = s.device;
}

static inline double wall_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Engine::~Engine() {
  Impl& s = *d_;
  cudaDeviceSynchronize();
  if (std::getenv("EMDEE_PROFILE") != nullptr)
    std::fprintf(stderr, "[emdee profile] host ms per call: check %.3f (n=%lld) rebuild %.3f (n=%lld) force %.3f boost %.3f (n=%lld) displace %.3f (n=%lld)\n",
                 1e3 * s.t_check / std::max(1LL, s.n_force), s.n_force, 1e3 * s.t_rebuild / std::max(1LL, s.n_rebuild), s.n_rebuild,
                 1e3 * s.t_force / std::max(1LL, s.n_force), 1e3 * s.t_boost / std::max(1LL, s.n_boost), s.n_boost,
                 1e3 * s.t_displace / std::max(1LL, s.n_displace), s.n_displace);
  s.R.release(); s.P.release(); s.R0.release(); s.F.release(); s.q.release(); s.invMass.release();
  s.delta.release(); s.type.release(); s.body.release(); s.exFirst.release(); s.exItem.release();
  s.interact.release();
  for (auto& t : s.tabs) t.release();
  s.Rs.release(); s.sRs.release(); s.sPosF.release(); s.atomCell.release(); s.atomFloor.release();
  s.cellCount.release(); s.cellStart.release(); s.cellFill.release(); s.slotAtom.release(); s.slotImg.release();
  s.slotCell.release(); s.sMeta.release(); s.sCell.release(); s.sType.release(); s.sBody.release(); s.nbr.release();
  s.nbrCount.release(); s.flags.release(); s.sGhost.release(); s.pos.release(); s.scanTmp.release();
  s.bdesc.release(); s.nbr16.release(); s.duoNbr.release(); s.duoCount.release(); s.rowsNbr.release();
  s.owned.release(); s.scratch3.release(); s.haloFlags.release(); s.selCount.release(); s.miPartial.release(); s.miResult.release();
  for (int k = 0; k < 4; ++k) { s.haloList[k].release(); s.haloBuf[k].release(); }
  s.known.release(); s.migCounts.release();
  for (int k = 0; k < 2; ++k) { s.migList[k].release(); s.migSend[k].release(); s.migRecv[k].release(); }
  if (s.h_mi) cudaFreeHost(s.h_mi);
  if (s.comm) nccl().CommDestroy(s.comm);
  s.chkPartial.release(); s.partial.release(); s.scalars.release(); s.counter.release(); s.tickets.release();
  if (s.h_scalars) cudaFreeHost(s.h_scalars);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  if (s.check_event) cudaEventDestroy(s.check_event);
  delete d_;
}

void slab_range(int M, int rank, int world, int& z0, int& z1) {
  z0 = (int)(((long long)rank * M) / world);
  z1 = (int)(((long long)(rank + 1) * M) / world);
}

void comm_unique_id(void* out128) {
  ncclUniqueId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
}

void Engine::comm_init(int rank, int world, const void* unique_id) {
  Impl& s = *d_;
  if (world <= 1) return;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  CUDA_CHECK(cudaSetDevice(s.device));
  NCCL_CHECK(nccl().CommInitRank(&s.comm, world, id, rank));
  s.rank = rank;
  s.world = world;
  CUDA_CHECK(cudaMallocHost(&s.h_mi, (size_t)world * sizeof(MaxIdx)));
  s.miResult.ensure(world + 1);
  s.miPartial.ensure(nblocks(s.N));
  s.scratch3.ensure(3 * (size_t)s.N);
  s.check_cached = false;
}

namespace {

// X (3N doubles, valid for the atoms each rank owns) -> full array on every rank:
// all-reduce(sum) of the owned parts (every atom is owned by exactly one rank, so the sum is exact)
void gather_full(Engine::Impl& s, double* X) {
  if (s.world <= 1 || !s.owned_valid) return;
  k_mask_owned<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, s.owned.p, X, s.scratch3.p);
  NCCL_CHECK(nccl().AllReduce(s.scratch3.p, X, 3 * (size_t)s.N, ncclDouble, ncclSum, s.comm, s.stream));
}

// per-rebuild: compact the four halo lists (ascending atom index) from the cell layers
void build_halo_lists(Engine::Impl& s) {
  const int N = s.N;
  s.haloFlags.ensure(4 * (size_t)N);
  s.selCount.ensure(4);
  k_halo_flags<<<nblocks(N), TPB, 0, s.stream>>>(N, s.grid, s.atomCell.p, s.haloFlags.p);
  cub::CountingInputIterator<int> ids(0);
  size_t need = 0;
  cub::DeviceSelect::Flagged(nullptr, need, ids, s.haloFlags.p, (int*)nullptr, s.selCount.p, N, s.stream);
  if (need > s.scanTmpBytes) {
    s.scanTmp.ensure(need);
    s.scanTmpBytes = need;
  }
  for (int k = 0; k < 4; ++k) {
    s.haloList[k].ensure(N / 2 + 64);   // a halo is two layers out of >= 4 per rank; grown below if ever exceeded
    if (s.haloList[k].n < (size_t)N) s.haloList[k].ensure(N);
    cub::DeviceSelect::Flagged(s.scanTmp.p, need, ids, s.haloFlags.p + (size_t)k * N, s.haloList[k].p, s.selCount.p + k, N,
                               s.stream);
  }
  int h[4];
  CUDA_CHECK(cudaMemcpyAsync(h, s.selCount.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  if (std::getenv("EMDEE_DEBUG")) std::fprintf(stderr, "[emdee r%d] halo lists: send up %d dn %d, recv below %d above %d\n", s.rank, h[0], h[1], h[2], h[3]);
  for (int k = 0; k < 4; ++k) {
    s.haloCount[k] = h[k];
    s.haloBuf[k].ensure(3 * (size_t)h[k] + 8, 1.2);
  }
}

// per step: ghost positions of the two layers above and below the slab (no reverse exchange: full list)
void halo_exchange(Engine::Impl& s) {
  if (s.world <= 1 || !s.owned_valid || s.halo_fresh) return;
  const int up = (s.rank + 1) % s.world, dn = (s.rank + s.world - 1) % s.world;
  for (int k = 0; k < 2; ++k)
    if (s.haloCount[k] > 0)
      k_pack3<<<nblocks(s.haloCount[k]), TPB, 0, s.stream>>>(s.haloCount[k], s.haloList[k].p, s.R.p, s.haloBuf[k].p);
  NCCL_CHECK(nccl().GroupStart());
  NCCL_CHECK(nccl().Send(s.haloBuf[0].p, 3 * (size_t)s.haloCount[0], ncclDouble, up, s.comm, s.stream));
  NCCL_CHECK(nccl().Send(s.haloBuf[1].p, 3 * (size_t)s.haloCount[1], ncclDouble, dn, s.comm, s.stream));
  NCCL_CHECK(nccl().Recv(s.haloBuf[2].p, 3 * (size_t)s.haloCount[2], ncclDouble, dn, s.comm, s.stream));
  NCCL_CHECK(nccl().Recv(s.haloBuf[3].p, 3 * (size_t)s.haloCount[3], ncclDouble, up, s.comm, s.stream));
  NCCL_CHECK(nccl().GroupEnd());
  for (int k = 2; k < 4; ++k)
    if (s.haloCount[k] > 0)
      k_unpack3<<<nblocks(s.haloCount[k]), TPB, 0, s.stream>>>(s.haloCount[k], s.haloList[k].p, s.haloBuf[k].p, s.R.p);
  s.halo_fresh = true;
}

// rebuild-time migration: see k_mig_flags. Returns with R, P current for every atom this rank may need.
void migrate(Engine::Impl& s, double Lbox) {
  const int N = s.N;
  s.known.ensure(N);
  if (s.all_known) {   // arrays are full on every rank (first build, or after an upload/download)
    if (s.p_partial) {
      gather_full(s, s.P.p);
      s.p_partial = false;
    }
    CUDA_CHECK(cudaMemsetAsync(s.known.p, 1, N, s.stream));
    return;
  }
  const int up = (s.rank + 1) % s.world, dn = (s.rank + s.world - 1) % s.world;
  s.haloFlags.ensure(4 * (size_t)N);
  s.selCount.ensure(4);
  k_mig_flags<<<nblocks(N), TPB, 0, s.stream>>>(N, Lbox, s.grid, s.R.p, s.owned.p, s.haloFlags.p);
  cub::CountingInputIterator<int> ids(0);
  size_t need = 0;
  cub::DeviceSelect::Flagged(nullptr, need, ids, s.haloFlags.p, (int*)nullptr, s.selCount.p, N, s.stream);
  if (need > s.scanTmpBytes) {
    s.scanTmp.ensure(need);
    s.scanTmpBytes = need;
  }
  for (int k = 0; k < 2; ++k) {
    s.migList[k].ensure(N);
    cub::DeviceSelect::Flagged(s.scanTmp.p, need, ids, s.haloFlags.p + (size_t)k * N, s.migList[k].p, s.selCount.p + k, N,
                               s.stream);
  }
  // counts: every rank learns every rank's (up, down) record counts with one small all-gather
  s.migCounts.ensure(2 * (size_t)s.world);
  NCCL_CHECK(nccl().AllGather(s.selCount.p, s.migCounts.p, 2, ncclInt, s.comm, s.stream));
  std::vector<int> all(2 * (size_t)s.world);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), s.migCounts.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  int c[4];
  c[0] = all[2 * s.rank];       // what I send up
  c[1] = all[2 * s.rank + 1];   // what I send down
  c[2] = all[2 * dn];           // what the rank below sends up = what arrives from below
  c[3] = all[2 * up + 1];       // what the rank above sends down = what arrives from above
  if (std::getenv("EMDEE_DEBUG")) std::fprintf(stderr, "[emdee r%d] migrate: send up %d dn %d, recv from-below %d from-above %d\n", s.rank, c[0], c[1], c[2], c[3]);
  for (int k = 0; k < 2; ++k) {
    s.migSend[k].ensure(7 * (size_t)c[k] + 8, 1.2);
    s.migRecv[k].ensure(7 * (size_t)c[2 + k] + 8, 1.2);
    if (c[k] > 0) k_pack7<<<nblocks(c[k]), TPB, 0, s.stream>>>(c[k], s.migList[k].p, s.R.p, s.P.p, s.migSend[k].p);
  }
  NCCL_CHECK(nccl().GroupStart());   // zero-length messages are skipped on both sides (both know the counts)
  if (c[0] > 0) NCCL_CHECK(nccl().Send(s.migSend[0].p, 7 * (size_t)c[0], ncclDouble, up, s.comm, s.stream));
  if (c[1] > 0) NCCL_CHECK(nccl().Send(s.migSend[1].p, 7 * (size_t)c[1], ncclDouble, dn, s.comm, s.stream));
  if (c[2] > 0) NCCL_CHECK(nccl().Recv(s.migRecv[0].p, 7 * (size_t)c[2], ncclDouble, dn, s.comm, s.stream));
  if (c[3] > 0) NCCL_CHECK(nccl().Recv(s.migRecv[1].p, 7 * (size_t)c[3], ncclDouble, up, s.comm, s.stream));
  NCCL_CHECK(nccl().GroupEnd());
  // known = previously owned + received
  CUDA_CHECK(cudaMemcpyAsync(s.known.p, s.owned.p, N, cudaMemcpyDeviceToDevice, s.stream));
  for (int k = 0; k < 2; ++k)
    if (c[2 + k] > 0)
      k_unpack7<<<nblocks(c[2 + k]), TPB, 0, s.stream>>>(c[2 + k], s.migRecv[k].p, s.R.p, s.P.p, s.known.p);
}

// distributed rebuild decision: identical on every rank (all inputs are all-gathered / all-reduced)
bool rebuild_needed_dist(Engine::Impl& s) {
  const int N = s.N;
  const long long all = 0x7fffffffffffffffLL;
  k_check_dist<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, s.R0.p, s.owned.p, N, all, s.miPartial.p, s.tickets.p + 2,
                                                 s.miResult.p + s.world);
  NCCL_CHECK(nccl().AllGather(s.miResult.p + s.world, s.miResult.p, sizeof(MaxIdx), ncclChar, s.comm, s.stream));
  CUDA_CHECK(cudaMemcpyAsync(s.h_mi, s.miResult.p, (size_t)s.world * sizeof(MaxIdx), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  MaxIdx g = s.h_mi[0];
  for (int r = 1; r < s.world; ++r)
    if (s.h_mi[r].m > g.m || (s.h_mi[r].m == g.m && s.h_mi[r].i < g.i)) g = s.h_mi[r];
  if (g.m > s.skinSq) return true;            // value >= maximum
  if (4.0 * g.m <= s.skinSq) return false;    // value <= 4*maximum
  double next = g.m;                          // i* == 0: `next` still holds the first atom's value
  if (g.i > 0) {
    k_check_dist<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, s.R0.p, s.owned.p, N, g.i, s.miPartial.p, s.tickets.p + 2,
                                                   s.miResult.p + s.world);
    NCCL_CHECK(nccl().AllGather(s.miResult.p + s.world, s.miResult.p, sizeof(MaxIdx), ncclChar, s.comm, s.stream));
    CUDA_CHECK(cudaMemcpyAsync(s.h_mi, s.miResult.p, (size_t)s.world * sizeof(MaxIdx), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    next = s.h_mi[0].m;
    for (int r = 1; r < s.world; ++r) next = std::max(next, s.h_mi[r].m);
  }
  const double value = g.m + 2 * std::sqrt(g.m * next) + next;   // reference neighbor_lists.f90:57
  return value > s.skinSq;
}

}  // namespace

void Engine::set_inner_cutoff(double InRc) { d_->InRcSq = InRc * InRc; }

void Engine::set_exclusions(const std::vector<int>& first, const std::vector<int>& last,
                            const std::vector<int>& item) {
  Impl& s = *d_;
  std::vector<int> f0(s.N + 1, 0), it;
  for (int i = 0; i < s.N; ++i) {
    f0[i] = (int)it.size();
    for (int q = first[i]; q <= last[i]; ++q) it.push_back(item[q - 1] - 1);
  }
  f0[s.N] = (int)it.size();
  CUDA_CHECK(cudaMemcpy(s.exFirst.p, f0.data(), (s.N + 1) * sizeof(int), cudaMemcpyHostToDevice));
  s.exItem.ensure(it.size() + 1);
  if (!it.empty()) CUDA_CHECK(cudaMemcpy(s.exItem.p, it.data(), it.size() * sizeof(int), cudaMemcpyHostToDevice));
}

void Engine::set_charges(const double* q) {
  Impl& s = *d_;
  CUDA_CHECK(cudaMemcpy(s.q.p, q, s.N * sizeof(double), cudaMemcpyHostToDevice));
  s.any_charged = false;
  for (int i = 0; i < s.N; ++i)
    if (std::fabs(q[i]) > DEPS) { s.any_charged = true; break; }
}

void Engine::set_interact(const std::vector<char>& interact) {
  d_->all_interact = true;
  for (char c : interact) d_->all_interact = d_->all_interact && (c != 0);
  CUDA_CHECK(cudaMemcpy(d_->interact.p, interact.data(), interact.size(), cudaMemcpyHostToDevice));
}

void Engine::set_layer(int layer0, const LayerTable& t) {
  Impl& s = *d_;
  s.layers[layer0] = t;
  s.tabs[layer0].ensure(t.pair.size());
  CUDA_CHECK(cudaMemcpy(s.tabs[layer0].p, t.pair.data(), t.pair.size() * sizeof(PairEntry), cudaMemcpyHostToDevice));
}

void Engine::upload_coordinates(const double* R) {
  Impl& s = *d_;
  CUDA_CHECK(cudaMemcpyAsync(s.R.p, R, 3 * (size_t)s.N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  s.has_R = true;
  s.check_cached = false;
  s.halo_fresh = true;   // every rank uploads the full array
  if (s.world > 1 && s.owned_valid && !s.all_known) {
    s.all_known = true;    // coordinates are full everywhere again ...
    s.p_partial = true;    // ... but the momenta of atoms owned elsewhere are stale here until the next rebuild
  }
}
void Engine::upload_body_delta(const double* delta) {
  Impl& s = *d_;
  s.delta.ensure(3 * (size_t)s.N);
  CUDA_CHECK(cudaMemcpy(s.delta.p, delta, 3 * (size_t)s.N * sizeof(double), cudaMemcpyHostToDevice));
  s.has_delta = true;
}
void Engine::upload_momenta(const double* P) {
  CUDA_CHECK(cudaMemcpy(d_->P.p, P, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
  d_->p_partial = false;
}
void Engine::upload_forces(int layer0, const double* F) {
  CUDA_CHECK(cudaMemcpy(d_->F.p + (size_t)layer0 * 3 * d_->N, F, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
}
void Engine::download_coordinates(double* R) {
  gather_full(*d_, d_->R.p);   // multi-GPU: collective, every rank must call
  d_->halo_fresh = true;
  CUDA_CHECK(cudaMemcpy(R, d_->R.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_momenta(double* P) {
  gather_full(*d_, d_->P.p);
  CUDA_CHECK(cudaMemcpy(P, d_->P.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_forces(int layer0, double* F) {
  gather_full(*d_, d_->F.p + (size_t)layer0 * 3 * d_->N);
  CUDA_CHECK(cudaMemcpy(F, d_->F.p + (size_t)layer0 * 3 * d_->N, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::synchronize() { CUDA_CHECK(cudaStreamSynchronize(d_->stream)); }
void* Engine::stream_handle() { return (void*)d_->stream; }

// ---- force-kernel dispatch ---------------------------------------------------------------------
namespace {

// UNROLL gathers in flight per thread, THREADS per block, MINBLOCKS resident blocks (register cap); the plain
// LJ instantiation runs 6 / 512 / 2 (tools/tune_force.py sweep, profiles/r1h_force_tuning.txt)
template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, int UNROLL = 2, int THREADS = TPB, int MINBLOCKS = 1>
void launch_force(ForceArgs& a, DBuf<double>& partial, bool compute, size_t smem, cudaStream_t st) {
  const int grid = nblocks(a.Next, THREADS);
  partial.ensure((size_t)grid * 5);
  a.partial = partial.p;
  if (compute) k_pair_forces<PK, PM, CK, CM, SINGLE, NEED_INVR, true, UNROLL, THREADS, MINBLOCKS><<<grid, THREADS, smem, st>>>(a);
  else k_pair_forces<PK, PM, CK, CM, SINGLE, NEED_INVR, false, UNROLL, THREADS, MINBLOCKS><<<grid, THREADS, smem, st>>>(a);
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR>
void launch_force_duo(const ForceArgs& a, int cap2, const unsigned int* duoNbr, const int* duoCount, bool compute, int grid,
                      size_t smem, cudaStream_t st) {
  if (compute) k_pair_forces_duo<PK, PM, CK, CM, SINGLE, NEED_INVR, true><<<grid, TPB, smem, st>>>(a, cap2, duoNbr, duoCount);
  else k_pair_forces_duo<PK, PM, CK, CM, SINGLE, NEED_INVR, false><<<grid, TPB, smem, st>>>(a, cap2, duoNbr, duoCount);
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, int UNROLL>
void launch_force_rows(ForceArgs& a, DBuf<double>& partial, int group, int pitch, const int* rows, bool compute, size_t smem,
                       cudaStream_t st) {
  const long long warps = ((long long)a.Next * group + 31) / 32;   // 32/group atoms per warp
  const int grid = (int)((warps + 7) / 8);                         // 8 warps per block
  partial.ensure((size_t)grid * 5);
  a.partial = partial.p;
#define EMDEE_ROWS_CASE(GG)                                                                                               \
  if (group == GG) {                                                                                                      \
    if (compute) k_pair_forces_rows<PK, PM, CK, CM, SINGLE, NEED_INVR, true, GG, UNROLL><<<grid, 256, smem, st>>>(a, pitch, rows); \
    else k_pair_forces_rows<PK, PM, CK, CM, SINGLE, NEED_INVR, false, GG, UNROLL><<<grid, 256, smem, st>>>(a, pitch, rows);       \
  }
  EMDEE_ROWS_CASE(8)
  EMDEE_ROWS_CASE(16)
  EMDEE_ROWS_CASE(32)
#undef EMDEE_ROWS_CASE
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, int UNROLL>
void launch_force_cluster2(ForceArgs& a, DBuf<double>& partial, int pitch, const unsigned int* rows, const int* duoCount,
                           bool compute, size_t smem, cudaStream_t st) {
  const long long nduo = ((long long)a.Next + 1) / 2;   // one warp per duo, 8 warps per block
  const int grid = (int)((nduo + 7) / 8);
  partial.ensure((size_t)grid * 5);
  a.partial = partial.p;
  if (compute) k_pair_forces_cluster2<PK, PM, CK, CM, SINGLE, NEED_INVR, true, UNROLL><<<grid, 256, smem, st>>>(a, pitch, rows, duoCount);
  else k_pair_forces_cluster2<PK, PM, CK, CM, SINGLE, NEED_INVR, false, UNROLL><<<grid, 256, smem, st>>>(a, pitch, rows, duoCount);
}

template <int PK, int PM, int CK, int CM, bool NEED_INVR>
void launch_force_brick(const ForceArgs& a, const BrickArgs& k, bool compute, int grid, int threads, size_t smem,
                        cudaStream_t st) {
  if (compute) k_pair_forces_brick<PK, PM, CK, CM, NEED_INVR, true><<<grid, threads, smem, st>>>(a, k);
  else k_pair_forces_brick<PK, PM, CK, CM, NEED_INVR, false><<<grid, threads, smem, st>>>(a, k);
}

template <class K>
void allow_big_smem(K kernel, size_t bytes) {
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void configure_brick_kernels() {
  static bool done = false;
  if (done) return;
  done = true;
  using namespace nb;
  const size_t fb = (size_t)BRICK_SMAX * sizeof(double4), bb = (size_t)BRICK_SMAX * sizeof(float4);
  allow_big_smem(k_build_list_brick, bb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, false, true>, fb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, false, false>, fb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true>, fb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, false>, fb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true>, fb);
  allow_big_smem(k_pair_forces_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, false>, fb);
  allow_big_smem(k_pair_forces_brick<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, true, true>, fb);
  allow_big_smem(k_pair_forces_brick<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, true, false>, fb);
}

// FP32 pre-test band (see k_build_list). Positions are ghost-shifted scaled coordinates, |p| <= pmax.
// Rounding p to FP32 moves each coordinate by <= 2^-24*pmax; the FP32 difference adds <= 2^-24*|d|.
// So every component of the FP32 separation is within delta of the true one, and
//   |r2_f - r2| <= 2*sqrt(3)*r*delta + 3*delta^2 + (FP32 rounding of 3 products and 2 sums) ~ 6*2^-24*r2.
// Evaluated at r = sqrt(xRc2s) (+ band) and doubled. The reference's own r2 differs from the
// ghost-shifted one by O(1e-16), far inside the band.
void build_band(double xRc2s, int M, float& accept, float& reject) {
  const double u = std::ldexp(1.0, -24);
  const double pmax = 1.0 + 2.0 / M + 1e-6;
  const double dmax = 3.0 / M;
  const double delta = 2.0 * u * pmax + u * dmax;
  const double r = std::sqrt(xRc2s) * 1.01;
  const double band = 2.0 * (2.0 * std::sqrt(3.0) * r * delta + 3.0 * delta * delta + 8.0 * u * xRc2s * 1.03);
  accept = std::nextafterf((float)(xRc2s - band), -1.0f);
  reject = std::nextafterf((float)(xRc2s + band), 2.0f);
}

}  // namespace

bool Engine::compute_forces(int layer0, bool compute, double Lbox, ForceScalars& out, double& neighbor_seconds) {
  Impl& s = *d_;
  const int N = s.N;
  const LayerTable& lt = s.layers[layer0];
  auto t_start = std::chrono::steady_clock::now();

  const double tp0 = wall_now();
  // ---- K0: rebuild trigger (reference handle_neighbor_lists) -----------------------------------
  bool rebuild;
  if (s.world > 1 && s.owned_valid) {
    halo_exchange(s);
    rebuild = rebuild_needed_dist(s);
    if (std::getenv("EMDEE_DEBUG")) std::fprintf(stderr, "[emdee r%d] compute_forces: rebuild=%d all_known=%d\n", s.rank, (int)rebuild, (int)s.all_known);
    stats_.launches += 1;
  } else {
    if (s.check_cached) {
      CUDA_CHECK(cudaEventSynchronize(s.check_event));   // evaluated by k_displace when the atoms moved
    } else {
      k_displacement_check<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, s.R0.p, N, s.chkPartial.p, s.tickets.p + 2,
                                                             s.scalars.p + 8);
      stats_.launches += 1;
      CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 8, s.scalars.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaEventRecord(s.check_event, s.stream));
      CUDA_CHECK(cudaEventSynchronize(s.check_event));
      s.check_cached = true;   // stays valid until the coordinates or R0 change
    }
    rebuild = s.h_scalars[8] > s.skinSq;
  }
  const double tp1 = wall_now();
  s.t_check += tp1 - tp0;

  const double invL2 = 1.0 / (Lbox * Lbox);
  if (rebuild) {
    int M = (int)std::floor(2 * Lbox / s.xRc);
    M = std::max(M, 5);
    if (2 * Lbox / s.xRc < 5.0)
      fatal("neighbor list handling", "box length is smaller than 2.5*(Rc + skin): the reference's 5x5x5 cell stencil cannot cover the cutoff sphere");
    if (M > 1019) fatal("neighbor list handling", "more than 1019 cells per dimension are not supported");
    const int prevM = s.grid.M;   // the grid of the previous build (0 before the first one)
    s.grid.M = M;
    s.grid.Mx = M + 4;
    s.grid.z0 = 0;
    s.grid.nzl = M;
    if (s.world > 1) {
      int z0, z1;
      slab_range(M, s.rank, s.world, z0, z1);
      if (z1 - z0 < 3 || M / s.world < 3)
        fatal("neighbor list handling", "fewer than three cell layers per GPU: use fewer GPUs for this box");
      if (s.owned_valid && prevM != M) {   // same decision on every rank (M is global)
        // the cell grid itself changed (box rescaled): fall back to re-assembling the full arrays
        gather_full(s, s.R.p);
        gather_full(s, s.P.p);
        s.all_known = true;
        s.p_partial = false;
      }
      if (s.owned_valid && !s.all_known && std::getenv("EMDEE_NO_MIGRATE") != nullptr) {
        gather_full(s, s.R.p);   // simpler scheme: re-assemble the full arrays on every rank
        gather_full(s, s.P.p);
        s.all_known = true;
        s.p_partial = false;
      }
      s.owned.ensure(N);
      if (!s.owned_valid) CUDA_CHECK(cudaMemsetAsync(s.owned.p, 0, N, s.stream));
      s.grid.z0 = z0;       // migrate() classifies against the slab faces of the grid being built
      s.grid.nzl = z1 - z0;
      s.grid.Mz = s.grid.nzl + 4;
      migrate(s, Lbox);
    }
    s.grid.Mz = s.grid.nzl + 4;
    s.owned.ensure(N);
    const long long ncell = (long long)s.grid.Mx * s.grid.Mx * s.grid.Mz;
    s.Rs.ensure(3 * (size_t)N);
    s.atomCell.ensure(N);
    s.atomFloor.ensure(3 * (size_t)N);
    s.cellCount.ensure(ncell + 1);
    s.cellStart.ensure(ncell + 1);
    s.cellFill.ensure(ncell + 1);
    CUDA_CHECK(cudaMemsetAsync(s.cellCount.p, 0, (ncell + 1) * sizeof(int), s.stream));
    CUDA_CHECK(cudaMemsetAsync(s.cellFill.p, 0, (ncell + 1) * sizeof(int), s.stream));
    k_bin<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, N, Lbox, s.grid, s.Rs.p, s.atomCell.p, s.atomFloor.p, s.owned.p,
                                            s.world > 1 ? s.known.p : nullptr, s.cellCount.p);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, s.cellCount.p, s.cellStart.p, (int)(ncell + 1), s.stream);
    if (need > s.scanTmpBytes) {
      s.scanTmp.ensure(need);
      s.scanTmpBytes = need;
    }
    cub::DeviceScan::ExclusiveSum(s.scanTmp.p, need, s.cellCount.p, s.cellStart.p, (int)(ncell + 1), s.stream);
    int Next = 0;
    CUDA_CHECK(cudaMemcpyAsync(&Next, s.cellStart.p + ncell, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    s.Next = Next;
    s.slotAtom.ensure(Next, 1.1);
    s.slotImg.ensure(Next, 1.1);
    s.slotCell.ensure(Next, 1.1);
    s.sMeta.ensure(Next, 1.1);
    s.sCell.ensure(Next, 1.1);
    s.sGhost.ensure(Next, 1.1);
    s.sType.ensure(Next, 1.1);
    s.sBody.ensure(Next, 1.1);
    s.sRs.ensure(Next, 1.1);
    s.sPosF.ensure(Next, 1.1);
    s.nbrCount.ensure(Next, 1.1);
    s.pos.ensure(Next, 1.1);
    k_fill<<<nblocks(N), TPB, 0, s.stream>>>(N, s.grid, s.atomCell.p, s.cellStart.p, s.cellFill.p, s.slotAtom.p,
                                             s.slotImg.p, s.slotCell.p);
    PlaceArgs pa;
    pa.Next = Next; pa.slotAtom = s.slotAtom.p; pa.slotImg = s.slotImg.p; pa.slotCell = s.slotCell.p;
    pa.cellStart = s.cellStart.p; pa.atomFloor = s.atomFloor.p; pa.Rs = s.Rs.p; pa.atomType = s.type.p;
    pa.atomBody = s.body.p; pa.owned = s.owned.p; pa.sMeta = s.sMeta.p; pa.sCell = s.sCell.p; pa.sGhost = s.sGhost.p; pa.sType = s.sType.p;
    pa.sBody = s.sBody.p; pa.sRs = s.sRs.p; pa.sPosF = s.sPosF.p; pa.nbrCount = s.nbrCount.p;
    k_place<<<nblocks(Next), TPB, 0, s.stream>>>(pa);
    stats_.launches += 4;
    // capacity guess from the mean density; grown on overflow
    if (s.cap == 0) {
      double nbar = (double)N / (Lbox * Lbox * Lbox) * (4.0 / 3.0) * 3.14159265358979323846 * s.xRc * s.xRcSq;
      s.cap = std::max(16, (int)(1.35 * nbar) + 16);
    }
    // ---- brick decomposition (single-type systems): largest brick whose staged halo fits shared memory ----
    s.use_bricks = false;
    if (s.nt == 1 && s.world == 1 && std::getenv("EMDEE_BRICKS") != nullptr) {   // opt-in: measured slower than the global path
      configure_brick_kernels();
      const double per_cell = (double)Next / (double)ncell;
      for (int b = 6; b >= 2 && !s.use_bricks; --b) {
        if ((b + 4.0) * (b + 4.0) * (b + 4.0) * per_cell > 1.02 * BRICK_SMAX) continue;
        BrickGrid bg;
        bg.M = M;
        bg.Mx = M + 4;
        bg.nbx = bg.nby = bg.nbz = (M + b - 1) / b;
        const int nbricks = bg.nbx * bg.nby * bg.nbz;
        s.bdesc.ensure(nbricks, 1.0);
        CUDA_CHECK(cudaMemsetAsync(s.flags.p, 0, 4 * sizeof(int), s.stream));
        k_brick_setup<<<nbricks, TPB, 0, s.stream>>>(bg, s.cellStart.p, s.bdesc.p, s.flags.p);
        stats_.launches += 1;
        int hf[4];
        CUDA_CHECK(cudaMemcpyAsync(hf, s.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (hf[2] <= BRICK_SMAX && hf[2] < 65536) {
          s.use_bricks = true;
          s.bgrid = bg;
          s.nbricks = nbricks;
          s.Bmax = ((hf[3] + 31) / 32) * 32;
          s.brick_threads = std::max(TPB, std::min(BRICK_TPB, s.Bmax));   // >= TPB: grid_finish folds with TPB threads
        }
      }
    }
    BuildArgs b;
    b.Next = Next; b.nt = s.nt; b.g = s.grid; b.all_interact = s.all_interact ? 1 : 0;
    b.xRc2s = s.xRcSq * invL2;
    b.xRcs = std::sqrt(b.xRc2s) * (1.0 + 1e-9);
    build_band(b.xRc2s, M, b.r2_accept, b.r2_reject);
    b.cellStart = s.cellStart.p; b.sMeta = s.sMeta.p; b.sCell = s.sCell.p; b.sGhost = s.sGhost.p;
    b.sType = s.sType.p; b.sBody = s.sBody.p; b.sRs = s.sRs.p; b.sPosF = s.sPosF.p; b.exFirst = s.exFirst.p;
    b.exItem = s.exItem.p; b.interact = s.interact.p; b.nbrCount = s.nbrCount.p; b.flags = s.flags.p;
    const long long ntiles = ((long long)Next + TILE - 1) / TILE;
    for (;;) {
      CUDA_CHECK(cudaMemsetAsync(s.flags.p, 0, 4 * sizeof(int), s.stream));
      b.cap = s.cap;
      if (timing_) CUDA_CHECK(cudaEventRecord(s.ev0, s.stream));
      if (s.use_bricks) {
        s.nbr16.ensure((size_t)s.nbricks * s.cap * s.Bmax, 1.1);
        b.nbr = nullptr;
        BrickArgs k;
        k.g = s.bgrid; k.desc = s.bdesc.p; k.nbr16 = s.nbr16.p; k.cap = s.cap; k.Bmax = s.Bmax;
        k_build_list_brick<<<s.nbricks, s.brick_threads, (size_t)BRICK_SMAX * sizeof(float4), s.stream>>>(b, k);
      } else {
        s.nbr.ensure((size_t)ntiles * s.cap * TILE, 1.1);
        b.nbr = s.nbr.p;
        // (a warp-per-tile variant with __ballot_sync compaction into shared-memory rows and a transposed
        // write-out was measured 6x slower at LJ-1M -- serial per-atom dependency chains at ~12 warps/SM --
        // and removed)
        k_build_list<<<nblocks(Next), TPB, 0, s.stream>>>(b);
      }
      if (timing_) CUDA_CHECK(cudaEventRecord(s.ev1, s.stream));
      stats_.launches += 1;
      stats_.build_launches += 1;
      int hflags[4];
      CUDA_CHECK(cudaMemcpyAsync(hflags, s.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      CUDA_CHECK(cudaGetLastError());
      if (timing_) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        stats_.build_ms += ms;
      }
      if (!hflags[1]) break;
      s.cap = (int)(hflags[0] * 1.15) + 8;   // overflow: regrow to the observed maximum and redo
    }
    // ---- duo rows: sorted merge of the rows of entries (2d, 2d+1) ------------------------------------------
    s.use_duos = !s.use_bricks && std::getenv("EMDEE_DUOS") != nullptr;   // opt-in: fewer LSU wavefronts but more FP64 issue + divergence; measured slower
    s.use_cluster2 = !s.use_bricks && !s.use_duos && std::getenv("EMDEE_CLUSTER2") != nullptr;   // opt-in, unmeasured
    if (s.use_duos || s.use_cluster2) {
      const int nduo = (Next + 1) / 2;
      const long long dtiles = ((long long)nduo + TILE - 1) / TILE;
      if (s.cap2 == 0) s.cap2 = (int)(1.45 * s.cap) + 8;
      s.duoCount.ensure(nduo, 1.1);
      for (;;) {
        s.duoNbr.ensure((size_t)dtiles * s.cap2 * TILE, 1.1);
        CUDA_CHECK(cudaMemsetAsync(s.flags.p, 0, 4 * sizeof(int), s.stream));
        k_merge_duos<<<nblocks(nduo), TPB, 0, s.stream>>>(Next, s.cap, s.cap2, s.nbr.p, s.nbrCount.p, s.duoNbr.p,
                                                          s.duoCount.p, s.flags.p);
        stats_.launches += 1;
        int hf[4];
        CUDA_CHECK(cudaMemcpyAsync(hf, s.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (!hf[3]) break;
        s.cap2 = (int)(hf[2] * 1.1) + 8;
      }
      if (s.use_cluster2) {   // row-major copy of the union rows (one row per duo)
        s.rows_pitch = (s.cap2 + 7) & ~7;
        s.rowsNbr.ensure((size_t)nduo * s.rows_pitch, 1.1);
        const int tgrid = (int)((dtiles + ROWS_TILES_PER_BLOCK - 1) / ROWS_TILES_PER_BLOCK);
        k_transpose_rows<<<tgrid, 32 * ROWS_TILES_PER_BLOCK, 0, s.stream>>>(nduo, s.cap2, s.rows_pitch,
                                                                             reinterpret_cast<const int*>(s.duoNbr.p), s.duoCount.p,
                                                                             s.rowsNbr.p);
        stats_.launches += 1;
      }
    }
    // ---- rows path: row-major copy of the list -------------------------------------------------------------
    s.rows_group = 0;
    if (!s.use_bricks && !s.use_duos && !s.use_cluster2 && std::getenv("EMDEE_ROWS") != nullptr) {   // opt-in experiment (see k_pair_forces_rows)
      const int g = std::atoi(std::getenv("EMDEE_ROWS"));
      s.rows_group = (g == 8 || g == 16 || g == 32) ? g : 8;
      s.rows_pitch = (s.cap + 7) & ~7;   // rows start on 32-byte boundaries
      s.rowsNbr.ensure((size_t)Next * s.rows_pitch, 1.1);
      const int tgrid = (int)((ntiles + ROWS_TILES_PER_BLOCK - 1) / ROWS_TILES_PER_BLOCK);
      k_transpose_rows<<<tgrid, 32 * ROWS_TILES_PER_BLOCK, 0, s.stream>>>(Next, s.cap, s.rows_pitch, s.nbr.p, s.nbrCount.p,
                                                                           s.rowsNbr.p);
      stats_.launches += 1;
    }
    if (s.world > 1) {
      build_halo_lists(s);
      s.owned_valid = true;
      s.all_known = false;
      s.halo_fresh = true;   // migrate() just made every needed position current
    }
    CUDA_CHECK(cudaMemcpyAsync(s.R0.p, s.R.p, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
    s.h_scalars[8] = 0.0;   // R0 = R: the criterion for the current coordinates is now zero displacement
    s.list_valid = true;
    stats_.cells_per_dim = M;
  }
  neighbor_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  const double tp2 = wall_now();
  if (rebuild) { s.t_rebuild += tp2 - tp1; s.n_rebuild += 1; }

  // ---- pair loop -------------------------------------------------------------------------------
  double* Fl = s.F.p + (size_t)layer0 * 3 * N;
  if (!lt.pairs_exist) {
    CUDA_CHECK(cudaMemsetAsync(Fl, 0, 3 * (size_t)N * sizeof(double), s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    out = ForceScalars();
    return rebuild;
  }
  const int Next = s.Next;
  k_refresh_positions<<<nblocks(Next), TPB, 0, s.stream>>>(Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p);
  ForceArgs a;
  a.Next = Next; a.cap = s.cap; a.nt = s.nt;
  a.Rc2s = (lt.useInRc ? s.InRcSq : s.RcSq) * invL2;
  a.L = Lbox; a.L2 = Lbox * Lbox; a.invL = 1.0 / Lbox; a.invL2 = invL2;
  a.pos = s.pos.p; a.nbr = s.nbr.p; a.nbrCount = s.nbrCount.p; a.sMeta = s.sMeta.p; a.sGhost = s.sGhost.p;
  a.sType = s.sType.p;
  a.delta = (s.has_delta && s.nbodies != 0) ? s.delta.p : nullptr;
  a.tab = s.tabs[layer0].p;
  a.single = lt.pair[0];
  a.coul = lt.coul;
  a.F = Fl; a.partial = s.partial.p; a.ticket = s.tickets.p; a.out = s.scalars.p;

  // classify the layer for kernel selection
  bool uniform = true;
  for (auto& pe : lt.pair)
    uniform = uniform && pe.model.kind == lt.pair[0].model.kind && pe.model.modifier == lt.pair[0].model.modifier;
  bool any_pair_coulomb = false;
  for (auto& pe : lt.pair) any_pair_coulomb = any_pair_coulomb || pe.coulomb;
  const bool coul_active = s.any_charged && any_pair_coulomb;
  a.q4_quirk = (!compute && lt.coul.kind == nb::K_COUL_NONE && coul_active) ? 1 : 0;
  const int pk = lt.pair[0].model.kind, pm = lt.pair[0].model.modifier;
  const int ck = coul_active ? lt.coul.kind : (int)nb::K_COUL_NONE, cm = lt.coul.modifier;
  const size_t smem_dyn = (s.nt <= MAX_SMEM_TYPES) ? (size_t)s.nt * s.nt * sizeof(PairEntry) : 0;

  if (timing_) CUDA_CHECK(cudaEventRecord(s.ev0, s.stream));
  using namespace nb;
  const bool lj_plain = uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_NONE && !a.q4_quirk;
  const bool lj_sf = uniform && pk == K_PAIR_LJ_CUT && pm == M_SHIFTED_FORCE && ck == K_COUL_NONE && !a.q4_quirk;
  const bool lj_coul_sf = uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_SF && cm == M_NONE;
  if (s.use_bricks) {
    BrickArgs k;
    k.g = s.bgrid; k.desc = s.bdesc.p; k.nbr16 = s.nbr16.p; k.cap = s.cap; k.Bmax = s.Bmax;
    const int bgrid = s.nbricks, bt = s.brick_threads;
    s.partial.ensure((size_t)bgrid * 5);
    a.partial = s.partial.p;
    const size_t sm = (size_t)BRICK_SMAX * sizeof(double4);
    if (lj_plain) launch_force_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, false>(a, k, compute, bgrid, bt, sm, s.stream);
    else if (lj_sf) launch_force_brick<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true>(a, k, compute, bgrid, bt, sm, s.stream);
    else if (lj_coul_sf) launch_force_brick<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true>(a, k, compute, bgrid, bt, sm, s.stream);
    else launch_force_brick<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, true>(a, k, compute, bgrid, bt, sm, s.stream);
  } else if (s.use_duos) {
    const int dgrid = nblocks((Next + 1) / 2);
    s.partial.ensure((size_t)dgrid * 5);
    a.partial = s.partial.p;
    if (s.nt == 1 && lj_plain)
      launch_force_duo<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false>(a, s.cap2, s.duoNbr.p, s.duoCount.p, compute, dgrid, 0, s.stream);
    else if (s.nt == 1 && lj_sf)
      launch_force_duo<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true>(a, s.cap2, s.duoNbr.p, s.duoCount.p, compute, dgrid, 0, s.stream);
    else if (s.nt == 1 && lj_coul_sf)
      launch_force_duo<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true>(a, s.cap2, s.duoNbr.p, s.duoCount.p, compute, dgrid, 0, s.stream);
    else
      launch_force_duo<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true>(a, s.cap2, s.duoNbr.p, s.duoCount.p, compute, dgrid, smem_dyn, s.stream);
  } else if (s.use_cluster2) {
    const unsigned int* rows = reinterpret_cast<const unsigned int*>(s.rowsNbr.p);
    if (s.nt == 1 && lj_plain)
      launch_force_cluster2<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, 2>(a, s.partial, s.rows_pitch, rows, s.duoCount.p, compute, 0, s.stream);
    else if (s.nt == 1 && lj_sf)
      launch_force_cluster2<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true, 2>(a, s.partial, s.rows_pitch, rows, s.duoCount.p, compute, 0, s.stream);
    else if (s.nt == 1 && lj_coul_sf)
      launch_force_cluster2<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true, 2>(a, s.partial, s.rows_pitch, rows, s.duoCount.p, compute, 0, s.stream);
    else
      launch_force_cluster2<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true, 2>(a, s.partial, s.rows_pitch, rows, s.duoCount.p, compute, smem_dyn, s.stream);
  } else if (s.rows_group != 0) {
    const int* rows = s.rowsNbr.p;
    const int g = s.rows_group, pitch = s.rows_pitch;
    if (s.nt == 1 && lj_plain)
      launch_force_rows<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, 3>(a, s.partial, g, pitch, rows, compute, 0, s.stream);
    else if (s.nt == 1 && lj_sf)
      launch_force_rows<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true, 2>(a, s.partial, g, pitch, rows, compute, 0, s.stream);
    else if (s.nt == 1 && lj_coul_sf)
      launch_force_rows<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true, 2>(a, s.partial, g, pitch, rows, compute, 0, s.stream);
    else
      launch_force_rows<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true, 2>(a, s.partial, g, pitch, rows, compute, smem_dyn, s.stream);
  } else if (s.nt == 1 && lj_plain && compute && std::getenv("EMDEE_FORCE_TUNE") != nullptr) {
    // tuning hook (bench experiments only): EMDEE_FORCE_TUNE="<variant>"
    const int v = std::atoi(std::getenv("EMDEE_FORCE_TUNE"));
#define EMDEE_TUNE_CASE(ID, UN, TH, MB) \
    if (v == ID) launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, UN, TH, MB>(a, s.partial, true, 0, s.stream);
    EMDEE_TUNE_CASE(0, 2, 128, 1)
    EMDEE_TUNE_CASE(1, 4, 256, 4)
    EMDEE_TUNE_CASE(2, 4, 256, 5)
    EMDEE_TUNE_CASE(4, 3, 256, 5)
    EMDEE_TUNE_CASE(5, 6, 256, 4)
    EMDEE_TUNE_CASE(7, 4, 512, 2)
    EMDEE_TUNE_CASE(10, 6, 512, 2)
    EMDEE_TUNE_CASE(12, 8, 512, 2)
    EMDEE_TUNE_CASE(13, 6, 1024, 1)
#undef EMDEE_TUNE_CASE
  } else if (s.nt == 1 && lj_plain)
    launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, 6, 512, 2>(a, s.partial, compute, 0, s.stream);
  else if (s.nt == 1 && lj_sf)
    launch_force<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true, 4, 256, 3>(a, s.partial, compute, 0, s.stream);
  else if (s.nt == 1 && lj_coul_sf)
    launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true, 4, 256, 3>(a, s.partial, compute, 0, s.stream);
  else
    launch_force<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true, 4, 256, 2>(a, s.partial, compute, smem_dyn, s.stream);
  if (timing_) CUDA_CHECK(cudaEventRecord(s.ev1, s.stream));
  stats_.launches += 2;
  stats_.force_launches += 1;
  if (s.world > 1) NCCL_CHECK(nccl().AllReduce(s.scalars.p, s.scalars.p, 5, ncclDouble, ncclSum, s.comm, s.stream));
  CUDA_CHECK(cudaMemcpyAsync(s.h_scalars, s.scalars.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  CUDA_CHECK(cudaGetLastError());
  if (timing_) {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    stats_.force_ms += ms;
  }
  s.t_force += wall_now() - tp2;
  s.n_force += 1;
  out.Epair = s.h_scalars[0];
  out.Ecoul = s.h_scalars[1];
  out.Wpair = s.h_scalars[2];
  out.Wcoul = s.h_scalars[3];
  out.Wbody = s.h_scalars[4];
  return rebuild;
}

void Engine::boost(int layer0, double CP, double CF, bool want_kinetic, KineticScalars& ke) {
  Impl& s = *d_;
  const double tp0 = wall_now();
  const int grid = nblocks((s.N + APT - 1) / APT);
  const unsigned char* owned = (s.world > 1 && s.owned_valid) ? s.owned.p : nullptr;
  k_boost<<<grid, TPB, 0, s.stream>>>(s.N, CP, CF, s.P.p, s.F.p + (size_t)layer0 * 3 * s.N, s.invMass.p, owned,
                                      want_kinetic ? 1 : 0, s.partial.p, s.tickets.p + 1, s.scalars.p + 10);
  stats_.launches += 1;
  if (want_kinetic) {
    if (owned != nullptr)
      NCCL_CHECK(nccl().AllReduce(s.scalars.p + 10, s.scalars.p + 10, 3, ncclDouble, ncclSum, s.comm, s.stream));
    CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 10, s.scalars.p + 10, 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    for (int x = 0; x < 3; ++x) ke.twoKE[x] = s.h_scalars[10 + x];
  }
  s.t_boost += wall_now() - tp0;
  s.n_boost += 1;
}

void Engine::displace(double CR, double CP) {
  Impl& s = *d_;
  const double tp0 = wall_now();
  if (s.world > 1 && s.owned_valid) {
    k_displace<<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p, s.owned.p, s.R0.p, nullptr,
                                                   s.tickets.p + 2, s.scalars.p + 8);
    stats_.launches += 1;
    s.halo_fresh = false;
    s.check_cached = false;
    s.all_known = false;   // from now on only owned + halo positions are current on this rank
  } else {
    k_displace<<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p, nullptr, s.R0.p, s.chkPartial.p,
                                                   s.tickets.p + 2, s.scalars.p + 8);
    stats_.launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 8, s.scalars.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaEventRecord(s.check_event, s.stream));
    s.check_cached = true;
  }
  s.t_displace += wall_now() - tp0;
  s.n_displace += 1;
}

long long Engine::pair_count() { return download_pairs(nullptr, 0); }

void Engine::update_list_stats(int layer0, double Lbox) {
  Impl& s = *d_;
  if (!s.list_valid) return;
  pair_count();
  const double invL2 = 1.0 / (Lbox * Lbox);
  const double Rc2s = (s.layers[layer0].useInRc ? s.InRcSq : s.RcSq) * invL2;
  CUDA_CHECK(cudaMemsetAsync(s.counter.p, 0, sizeof(unsigned long long), s.stream));
  k_refresh_positions<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p);
  if (s.use_bricks) {
    BrickArgs k;
    k.g = s.bgrid; k.desc = s.bdesc.p; k.nbr16 = s.nbr16.p; k.cap = s.cap; k.Bmax = s.Bmax;
    k_brick_list_walk<<<s.nbricks, TPB, 0, s.stream>>>(k, 1, s.nbrCount.p, s.sMeta.p, s.pos.p, Rc2s, nullptr, 0, s.counter.p);
  } else {
    k_count_interacting<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, s.cap, Rc2s, s.pos.p, s.nbr.p, s.nbrCount.p, s.counter.p);
  }
  unsigned long long n = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n, s.counter.p, sizeof(n), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  stats_.interacting = (long long)n;
}

long long Engine::download_pairs(int* pairs, long long capacity) {
  Impl& s = *d_;
  if (!s.list_valid) return 0;
  DBuf<int> dp;
  if (pairs != nullptr && capacity > 0) dp.ensure(2 * (size_t)capacity);
  CUDA_CHECK(cudaMemsetAsync(s.counter.p, 0, sizeof(unsigned long long), s.stream));
  if (s.use_bricks) {
    BrickArgs k;
    k.g = s.bgrid; k.desc = s.bdesc.p; k.nbr16 = s.nbr16.p; k.cap = s.cap; k.Bmax = s.Bmax;
    k_brick_list_walk<<<s.nbricks, TPB, 0, s.stream>>>(k, 0, s.nbrCount.p, s.sMeta.p, s.pos.p, 0.0, dp.p, capacity, s.counter.p);
  } else {
    k_export_pairs<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, s.cap, s.nbr.p, s.nbrCount.p, s.sMeta.p, dp.p, capacity, s.counter.p);
  }
  unsigned long long n = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n, s.counter.p, sizeof(n), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  stats_.list_entries = 2 * (long long)n;
  long long got = std::min<long long>((long long)n, capacity);
  if (pairs != nullptr && got > 0)
    CUDA_CHECK(cudaMemcpy(pairs, dp.p, 2 * (size_t)got * sizeof(int), cudaMemcpyDeviceToHost));
  dp.release();
  return pairs == nullptr ? (long long)n : got;
}

void Engine::rdf(double Lbox, int bins, double Rc2_scaled, double bins_by_Rc_scaled,
                 const std::vector<unsigned short>& pairSym, int nsym, std::vector<long long>& counts) {
  Impl& s = *d_;
  if (!s.list_valid) fatal("radial distribution calculation", "no neighbor list has been built yet");
  if (s.world > 1) fatal("radial distribution calculation", "not available on a multi-GPU system yet");
  if (s.use_bricks) fatal("radial distribution calculation", "not available with EMDEE_BRICKS");
  const size_t nbin = (size_t)bins * nsym;
  DBuf<unsigned long long> hist;
  DBuf<unsigned short> sym;
  hist.ensure(nbin);
  sym.ensure(pairSym.size());
  CUDA_CHECK(cudaMemsetAsync(hist.p, 0, nbin * sizeof(unsigned long long), s.stream));
  CUDA_CHECK(cudaMemcpyAsync(sym.p, pairSym.data(), pairSym.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, s.stream));
  // the list is walked with the CURRENT coordinates (the reference rescales me%R on entry, EmDeeCode.f90:1324)
  k_refresh_positions<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p);
  const int use_smem = nbin * sizeof(unsigned int) <= 40 * 1024 ? 1 : 0;
  k_rdf<<<nblocks(s.Next), TPB, use_smem ? nbin * sizeof(unsigned int) : 0, s.stream>>>(
      s.Next, s.cap, s.nt, bins, nsym, Rc2_scaled, bins_by_Rc_scaled, s.pos.p, s.nbr.p, s.nbrCount.p, s.sType.p, sym.p,
      use_smem, hist.p);
  stats_.launches += 2;
  std::vector<unsigned long long> h(nbin);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), hist.p, nbin * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  CUDA_CHECK(cudaGetLastError());
  counts.resize(nbin);
  for (size_t q = 0; q < nbin; ++q) counts[q] = (long long)(h[q] / 2ull);   // full list: each pair met twice
  hist.release();
  sym.release();
}

double measure_fp64_fma_tflops() {
  int dev = 0, sms = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  const int iters = 1 << 14, grid = sms * 8, tpb = 256;
  k_dfma_peak<<<grid, tpb>>>(d, iters, 1.0);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_CHECK(cudaEventRecord(e0));
    k_dfma_peak<<<grid, tpb>>>(d, iters, 1.0 + rep);
    CUDA_CHECK(cudaEventRecord(e1));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 8.0 * (double)iters * grid * tpb / (ms * 1e-3) / 1e12;
    best = tf > best ? tf : best;
  }
  CUDA_CHECK(cudaFree(d));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace emdee
