// engine.cu -- CUDA hot path (sm_100a): cell binning, cell sort with periodic ghost images, Verlet
// list build, pair force/energy/virial kernels, device-resident velocity-Verlet pieces.
//
// Reference loops replaced (paths relative to the reference tree):
//   k_displacement_check  <- src/neighbor_lists.f90:41-59,184      (maximum_approach_sq + trigger)
//   k_bin / k_fill / k_place <- src/neighbor_lists.f90:63-171       (distribute_atoms)
//   k_build_list          <- src/neighbor_lists.f90:199-300         (build_neighbor_lists)
//   k_refresh_positions   <- src/EmDeeCode.f90:1228                 (Rs = R/L)
//   k_pair_forces         <- src/compute.f90:20-100 + src/apply_modifier.f90 + model bodies,
//                            src/EmDeeData.f90:644-685, src/EmDeeCode.f90:1247 (sum of thread forces),
//                            src/EmDeeData.f90:926-953 (rigid_body_virial, fused in the epilogue)
//   k_boost / k_displace  <- src/EmDeeData.f90:823-922 (free atoms)
//
// Design: atoms are sorted by cell of an EXTENDED grid (M+4)^3 that carries explicit
// periodic ghost images in a 2-cell shell, so the force kernel needs no minimum-image arithmetic; the
// Verlet list is a FULL list (both directions of every pair) stored as 32-lane tiles (ELL-in-tile,
// coalesced 128-byte rows), so every atom's force is finished inside one thread: no atomics, no
// scatter, deterministic summation order. List MEMBERSHIP is decided with the reference's exact
// arithmetic (un-fused IEEE operations on unwrapped scaled coordinates), so pair sets are bit-identical.
#include "engine.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
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

inline int nblocks(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

// ------------------------------------------------------------------------------------------------
// K0: rebuild trigger. Ordered reduction reproducing the sequential scan of maximum_approach_sq:
// state (m, n): m = running maximum, n = value `next` holds. combine(A then B) =
//   B.m > A.m ? (B.m, max(A.m, B.n)) : A.     Atom 0 contributes (d0, d0), atom i>0 (d_i, -inf).
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
constexpr int CHK_ITEMS = 8;   // consecutive atoms per thread

__device__ __forceinline__ MaxNext block_ordered_reduce(MaxNext s) {
  __shared__ MaxNext warp_state[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ordered tree inside the warp: lane L absorbs lane L+off (its right-hand neighbour segment)
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
  return r;   // valid in thread 0
}

__global__ void __launch_bounds__(TPB) k_displacement_check(const double* __restrict__ R,
                                                            const double* __restrict__ R0, int N,
                                                            MaxNext* __restrict__ partial) {
  const double NEG_INF = -1.0 / 0.0;
  long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * CHK_ITEMS;
  MaxNext s;
  s.m = NEG_INF;
  s.n = NEG_INF;
  for (int q = 0; q < CHK_ITEMS; ++q) {
    long long i = first + q;
    if (i < N) {
      double dx = __dsub_rn(R[3 * i], R0[3 * i]);
      double dy = __dsub_rn(R[3 * i + 1], R0[3 * i + 1]);
      double dz = __dsub_rn(R[3 * i + 2], R0[3 * i + 2]);
      double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      MaxNext e;
      e.m = d;
      e.n = (i == 0) ? d : NEG_INF;
      s = mn_combine(s, e);
    }
  }
  s = block_ordered_reduce(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_displacement_final(const MaxNext* __restrict__ partial, int nparts,
                                                            double* __restrict__ result) {
  const double NEG_INF = -1.0 / 0.0;
  // each thread folds a contiguous range of block partials, then an ordered block reduction
  int per = (nparts + blockDim.x - 1) / blockDim.x;
  MaxNext s;
  s.m = NEG_INF;
  s.n = NEG_INF;
  for (int q = 0; q < per; ++q) {
    int i = threadIdx.x * per + q;
    if (i < nparts) s = mn_combine(s, partial[i]);
  }
  s = block_ordered_reduce(s);
  if (threadIdx.x == 0)
    result[0] = __dadd_rn(__dadd_rn(s.m, __dmul_rn(2.0, __dsqrt_rn(__dmul_rn(s.m, s.n)))), s.n);
}

// ------------------------------------------------------------------------------------------------
// K1-K3: binning into the extended grid (real cells [2, M+2) per dimension + 2-cell ghost shell).
// ------------------------------------------------------------------------------------------------
struct GridDesc {
  int M;    // real cells per dimension (reference: max(floor(2L/xRc), 5))
  int Mx;   // extended cells per dimension = M + 4
};

// enumerate the images of an atom in real cell coordinate c (one dimension): s in {0} U {+1 if c<=1}
// U {-1 if c>=M-2}; at most two because M >= 5.
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

__global__ void __launch_bounds__(TPB) k_bin(const double* __restrict__ R, int N, double L, GridDesc g,
                                             double* __restrict__ Rs, int* __restrict__ atomCell,
                                             int* __restrict__ atomFloor, int* __restrict__ cellCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
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
  int sx[2], sy[2], sz[2];
  int nx = image_shifts(c[0], g.M, sx), ny = image_shifts(c[1], g.M, sy), nz = image_shifts(c[2], g.M, sz);
  for (int a = 0; a < nz; ++a)
    for (int b = 0; b < ny; ++b)
      for (int d = 0; d < nx; ++d) {
        int ex = c[0] + 2 + sx[d] * g.M, ey = c[1] + 2 + sy[b] * g.M, ez = c[2] + 2 + sz[a] * g.M;
        atomicAdd(&cellCount[ex + g.Mx * (ey + g.Mx * ez)], 1);
      }
}

__global__ void __launch_bounds__(TPB) k_fill(int N, GridDesc g, const int* __restrict__ atomCell,
                                              const int* __restrict__ cellStart, int* __restrict__ cellFill,
                                              int* __restrict__ slotAtom, int* __restrict__ slotImg,
                                              int* __restrict__ slotCell) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int pc = atomCell[i];
  int c[3] = {pc & 1023, (pc >> 10) & 1023, (pc >> 20) & 1023};
  int sx[2], sy[2], sz[2];
  int nx = image_shifts(c[0], g.M, sx), ny = image_shifts(c[1], g.M, sy), nz = image_shifts(c[2], g.M, sz);
  for (int a = 0; a < nz; ++a)
    for (int b = 0; b < ny; ++b)
      for (int d = 0; d < nx; ++d) {
        int ex = c[0] + 2 + sx[d] * g.M, ey = c[1] + 2 + sy[b] * g.M, ez = c[2] + 2 + sz[a] * g.M;
        int cell = ex + g.Mx * (ey + g.Mx * ez);
        int slot = cellStart[cell] + atomicAdd(&cellFill[cell], 1);
        slotAtom[slot] = i;
        slotImg[slot] = (sx[d] + 1) | ((sy[b] + 1) << 2) | ((sz[a] + 1) << 4);
        slotCell[slot] = cell;
      }
}

// Deterministic order inside a cell: ascending atom index (rank by counting), then materialise the
// per-entry arrays the list build and the force kernel read.
__global__ void __launch_bounds__(TPB) k_place(int Next, const int* __restrict__ slotAtom,
                                               const int* __restrict__ slotImg, const int* __restrict__ slotCell,
                                               const int* __restrict__ cellStart, const int* __restrict__ atomFloor,
                                               const double* __restrict__ Rs, const int* __restrict__ atomType,
                                               const int* __restrict__ atomBody, int4* __restrict__ sMeta,
                                               int* __restrict__ sCell, unsigned char* __restrict__ sGhost,
                                               int* __restrict__ sType, int* __restrict__ sBody,
                                               double* __restrict__ sRs, int* __restrict__ nbrCount) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Next) return;
  int cell = slotCell[t];
  int a = slotAtom[t];
  int lo = cellStart[cell], hi = cellStart[cell + 1];
  int rank = 0;
  for (int u = lo; u < hi; ++u) rank += (slotAtom[u] < a);
  int e = lo + rank;
  int img = slotImg[t];
  int sx = (img & 3) - 1, sy = ((img >> 2) & 3) - 1, sz = ((img >> 4) & 3) - 1;
  sMeta[e] = make_int4(a, sx - atomFloor[3 * (size_t)a], sy - atomFloor[3 * (size_t)a + 1],
                       sz - atomFloor[3 * (size_t)a + 2]);
  sCell[e] = cell;
  sGhost[e] = (img != (1 | (1 << 2) | (1 << 4)));
  sType[e] = atomType[a];
  sBody[e] = atomBody[a];
  sRs[3 * (size_t)e] = Rs[3 * (size_t)a];
  sRs[3 * (size_t)e + 1] = Rs[3 * (size_t)a + 1];
  sRs[3 * (size_t)e + 2] = Rs[3 * (size_t)a + 2];
  nbrCount[e] = 0;
}

// ------------------------------------------------------------------------------------------------
// K4: Verlet list build. One thread per real entry; candidates = the 5x5x5 block of extended cells
// around the entry's cell (25 contiguous x-runs). Membership test is the reference's, bit for bit:
//   d = Rs_i - Rs_j (unwrapped scaled), d -= anint(d), r2 = (dx^2 + dy^2) + dz^2, r2 < xRc^2/L^2
// with every operation individually rounded (no FMA contraction). rint replaces anint: they differ
// only at |d| = k + 1/2 exactly, where (d - round(d))^2 is the same number.
// ------------------------------------------------------------------------------------------------
struct BuildArgs {
  int Next, cap, nt;
  GridDesc g;
  double xRc2s;   // xRcSq * invL2
  const int* cellStart;
  const int4* sMeta;
  const int* sCell;
  const unsigned char* sGhost;
  const int* sType;
  const int* sBody;
  const double* sRs;
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

__global__ void __launch_bounds__(TPB) k_build_list(BuildArgs a) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  const int lane = threadIdx.x & 31;
  if (e < a.Next && !a.sGhost[e]) {
    const size_t base = ((size_t)(e >> 5) * a.cap) * TILE + lane;
    const int atom_i = a.sMeta[e].x;
    const int type_i = a.sType[e], body_i = a.sBody[e];
    const double xi = a.sRs[3 * (size_t)e], yi = a.sRs[3 * (size_t)e + 1], zi = a.sRs[3 * (size_t)e + 2];
    const int x0 = a.exFirst[atom_i], x1 = a.exFirst[atom_i + 1];
    const int cell = a.sCell[e];
    const int Mx = a.g.Mx;
    const int ez = cell / (Mx * Mx), ey = (cell - ez * Mx * Mx) / Mx, ex = cell - Mx * (ey + Mx * ez);
    for (int dz = -2; dz <= 2; ++dz)
      for (int dy = -2; dy <= 2; ++dy) {
        const int row = Mx * ((ey + dy) + Mx * (ez + dz));
        const int f0 = a.cellStart[row + ex - 2], f1 = a.cellStart[row + ex + 3];
        for (int f = f0; f < f1; ++f) {
          double r2 = __dadd_rn(__dadd_rn(strict_pbc_sq(xi, a.sRs[3 * (size_t)f]),
                                          strict_pbc_sq(yi, a.sRs[3 * (size_t)f + 1])),
                                strict_pbc_sq(zi, a.sRs[3 * (size_t)f + 2]));
          if (r2 < a.xRc2s && f != e) {
            const int atom_j = a.sMeta[f].x;
            bool ok = (a.sBody[f] != body_i) && a.interact[type_i * a.nt + a.sType[f]];
            for (int q = x0; ok && q < x1; ++q) ok = (a.exItem[q] != atom_j);
            if (ok) {
              if (cnt < a.cap) a.nbr[base + (size_t)cnt * TILE] = f;
              ++cnt;
            }
          }
        }
      }
    a.nbrCount[e] = min(cnt, a.cap);
  }
  // warp max of counts -> one atomicMax per warp
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
  double L, invL, invL2;
  const double4* pos;
  const int* nbr;
  const int* nbrCount;
  const int4* sMeta;
  const unsigned char* sGhost;
  const int* sType;
  const double* delta;     // (3,N) body-frame offsets for the rigid-body virial, or nullptr
  const PairEntry* tab;    // nt*nt (device)
  PairEntry single;        // the only entry when nt == 1
  nb::DevModel coul;
  int q4_quirk;            // virial-only + coul_none: Wij keeps the pair value (reference make_virial_compute.sh:24-29)
  double* F;               // (3,N) output, original atom order
  double* partial;         // gridDim.x * 5
};

__device__ __forceinline__ double fast_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  return x;
}

__device__ __forceinline__ double4 ld_pos(const double4* p) {
  // 2 x 16-byte read-only loads (one 32-byte sector)
  const double2* q = reinterpret_cast<const double2*>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR, bool COMPUTE>
__global__ void __launch_bounds__(TPB) k_pair_forces(const __grid_constant__ ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[TPB / 32][5];
  const PairEntry* tab = a.tab;
  if (!SINGLE && a.nt <= MAX_SMEM_TYPES) {
    PairEntry* st = reinterpret_cast<PairEntry*>(smem_raw);
    const int words = a.nt * a.nt * (int)(sizeof(PairEntry) / sizeof(int));
    for (int w = threadIdx.x; w < words; w += blockDim.x)
      reinterpret_cast<int*>(st)[w] = reinterpret_cast<const int*>(a.tab)[w];
    __syncthreads();
    tab = st;
  }
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double Ep = 0.0, Ec = 0.0, Wp = 0.0, Wc = 0.0, Wb = 0.0;
  const int cnt = (e < a.Next) ? a.nbrCount[e] : 0;   // ghosts hold 0
  const bool has_coul = (CK == nb::K_DYNAMIC) ? true : (CK != nb::K_COUL_NONE);
  if (e < a.Next) {
    const double4 pi = a.pos[e];
    const int itype = SINGLE ? 0 : a.sType[e];
    const bool icharged = fabs(pi.w) > 2.220446049250313e-16;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const int* nb_ptr = a.nbr + ((size_t)(e >> 5) * a.cap) * TILE + lane;
#pragma unroll 2
    for (int k = 0; k < cnt; ++k) {
      const int f = nb_ptr[(size_t)k * TILE];
      const double4 pj = ld_pos(a.pos + f);
      const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < a.Rc2s) {
        double invR, invR2;
        if (NEED_INVR) {
          invR = rsqrt(r2) * a.invL;
          invR2 = invR * invR;
        } else {
          invR2 = fast_rcp(r2) * a.invL2;
          invR = 0.0;
        }
        const PairEntry& pe = SINGLE ? a.single : tab[itype * a.nt + a.sType[f]];
        double E, W;
        nb::eval_kind<PK>(pe.model, invR, invR2, E, W);
        nb::eval_modifier<PM>(pe.model, invR, invR2, E, W);
        if (COMPUTE) Ep += E;
        Wp += W;
        double Wsum = W;
        if (has_coul) {
          if (icharged && fabs(pj.w) > 2.220446049250313e-16 && pe.coulomb) {
            double Eq, Wq;
            if (!COMPUTE && a.q4_quirk) {
              Eq = 0.0;
              Wq = W;
            } else {
              nb::eval_kind<CK>(a.coul, invR, invR2, Eq, Wq);
              nb::eval_modifier<CM>(a.coul, invR, invR2, Eq, Wq);
            }
            const double QiQj = pe.kCoul * pi.w * pj.w;
            if (COMPUTE) Ec += QiQj * Eq;
            Wq = QiQj * Wq;
            Wc += Wq;
            Wsum += Wq;
          }
        }
        const double s = Wsum * invR2;
        fx += s * dx;
        fy += s * dy;
        fz += s * dz;
      }
    }
    if (!a.sGhost[e]) {   // ghost images hold no list (count 0) and own no force slot
      const int atom = a.sMeta[e].x;
      fx *= a.L;
      fy *= a.L;
      fz *= a.L;
      a.F[3 * (size_t)atom] = fx;
      a.F[3 * (size_t)atom + 1] = fy;
      a.F[3 * (size_t)atom + 2] = fz;
      if (a.delta != nullptr)
        Wb = -(fx * a.delta[3 * (size_t)atom] + fy * a.delta[3 * (size_t)atom + 1] + fz * a.delta[3 * (size_t)atom + 2]);
    }
  }
  // block reduction of the five scalars (deterministic: fixed shuffle tree + fixed warp order)
  double v[5] = {Ep, Ec, Wp, Wc, Wb};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double x = v[q];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) red[threadIdx.x >> 5][q] = x;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int w = 0; w < TPB / 32; ++w) s += red[w][threadIdx.x];
    a.partial[(size_t)blockIdx.x * 5 + threadIdx.x] = s;
  }
}

// final fixed-order reduction of per-block partials; pair sums are halved (full list counts i-j and j-i)
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partial, int nparts, int width,
                                                         double half_mask_scale, double* __restrict__ out) {
  __shared__ double sm[256];
  for (int q = 0; q < width; ++q) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[(size_t)i * width + q];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = sm[0] * ((q < 4) ? half_mask_scale : 1.0);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Device-resident dynamics for free atoms. Un-fused arithmetic so trajectories track the reference.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_boost(int N, double CP, double CF, double* __restrict__ P,
                                               const double* __restrict__ F, const double* __restrict__ invMass,
                                               int want_ke, double* __restrict__ partial) {
  __shared__ double red[TPB / 32][3];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double k[3] = {0.0, 0.0, 0.0};
  if (i < N) {
    double im = invMass[i];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      double p = __dadd_rn(__dmul_rn(CP, P[3 * (size_t)i + x]), __dmul_rn(CF, F[3 * (size_t)i + x]));
      P[3 * (size_t)i + x] = p;
      k[x] = __dmul_rn(__dmul_rn(im, p), p);
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
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < TPB / 32; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(TPB) k_displace(int N, double CR, double CP, double* __restrict__ R,
                                                  const double* __restrict__ P, const double* __restrict__ invMass) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double im = invMass[i];
#pragma unroll
  for (int x = 0; x < 3; ++x)
    R[3 * (size_t)i + x] = __dadd_rn(__dmul_rn(CR, R[3 * (size_t)i + x]), __dmul_rn(__dmul_rn(CP, P[3 * (size_t)i + x]), im));
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
  std::vector<LayerTable> layers;
  std::vector<DBuf<PairEntry>> tabs;

  // rebuild artifacts
  GridDesc grid{0, 0};
  int Next = 0, cap = 0;
  double Lbuild = 0;
  DBuf<double> Rs, sRs;
  DBuf<int> atomCell, atomFloor, cellCount, cellStart, cellFill, slotAtom, slotImg, slotCell;
  DBuf<int4> sMeta;
  DBuf<int> sCell, sType, sBody, nbr, nbrCount, flags;
  DBuf<unsigned char> sGhost;
  DBuf<double4> pos;
  DBuf<unsigned char> scanTmp;
  size_t scanTmpBytes = 0;
  bool list_valid = false;

  // reductions
  DBuf<MaxNext> chkPartial;
  DBuf<double> partial, scalars;
  DBuf<unsigned long long> counter;
  double* h_scalars = nullptr;   // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

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
  CUDA_CHECK(cudaMallocHost(&s.h_scalars, 16 * sizeof(double)));
  CUDA_CHECK(cudaEventCreate(&s.ev0));
  CUDA_CHECK(cudaEventCreate(&s.ev1));
  stats_.device = s.device;
}

Engine::~Engine() {
  Impl& s = *d_;
  cudaDeviceSynchronize();
  s.R.release(); s.P.release(); s.R0.release(); s.F.release(); s.q.release(); s.invMass.release();
  s.delta.release(); s.type.release(); s.body.release(); s.exFirst.release(); s.exItem.release();
  s.interact.release();
  for (auto& t : s.tabs) t.release();
  s.Rs.release(); s.sRs.release(); s.atomCell.release(); s.atomFloor.release(); s.cellCount.release();
  s.cellStart.release(); s.cellFill.release(); s.slotAtom.release(); s.slotImg.release(); s.slotCell.release();
  s.sMeta.release(); s.sCell.release(); s.sType.release(); s.sBody.release(); s.nbr.release();
  s.nbrCount.release(); s.flags.release(); s.sGhost.release(); s.pos.release(); s.scanTmp.release();
  s.chkPartial.release(); s.partial.release(); s.scalars.release(); s.counter.release();
  if (s.h_scalars) cudaFreeHost(s.h_scalars);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  delete d_;
}

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
  s.list_valid = false;
}

void Engine::set_charges(const double* q) {
  Impl& s = *d_;
  CUDA_CHECK(cudaMemcpy(s.q.p, q, s.N * sizeof(double), cudaMemcpyHostToDevice));
  s.any_charged = false;
  for (int i = 0; i < s.N; ++i)
    if (std::fabs(q[i]) > 2.220446049250313e-16) { s.any_charged = true; break; }
}

void Engine::set_interact(const std::vector<char>& interact) {
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
}
void Engine::upload_body_delta(const double* delta) {
  Impl& s = *d_;
  s.delta.ensure(3 * (size_t)s.N);
  CUDA_CHECK(cudaMemcpy(s.delta.p, delta, 3 * (size_t)s.N * sizeof(double), cudaMemcpyHostToDevice));
  s.has_delta = true;
}
void Engine::upload_momenta(const double* P) {
  CUDA_CHECK(cudaMemcpy(d_->P.p, P, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
}
void Engine::upload_forces(int layer0, const double* F) {
  CUDA_CHECK(cudaMemcpy(d_->F.p + (size_t)layer0 * 3 * d_->N, F, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
}
void Engine::download_coordinates(double* R) {
  CUDA_CHECK(cudaMemcpy(R, d_->R.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_momenta(double* P) {
  CUDA_CHECK(cudaMemcpy(P, d_->P.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_forces(int layer0, double* F) {
  CUDA_CHECK(cudaMemcpy(F, d_->F.p + (size_t)layer0 * 3 * d_->N, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::synchronize() { CUDA_CHECK(cudaStreamSynchronize(d_->stream)); }

// ---- force-kernel dispatch ---------------------------------------------------------------------
namespace {

template <int PK, int PM, int CK, int CM, bool SINGLE, bool NEED_INVR>
void launch_force(const ForceArgs& a, bool compute, int grid, size_t smem, cudaStream_t st) {
  if (compute) k_pair_forces<PK, PM, CK, CM, SINGLE, NEED_INVR, true><<<grid, TPB, smem, st>>>(a);
  else k_pair_forces<PK, PM, CK, CM, SINGLE, NEED_INVR, false><<<grid, TPB, smem, st>>>(a);
}

}  // namespace

bool Engine::compute_forces(int layer0, bool compute, double Lbox, ForceScalars& out, double& neighbor_seconds) {
  Impl& s = *d_;
  const int N = s.N;
  const LayerTable& lt = s.layers[layer0];
  auto t_start = std::chrono::steady_clock::now();

  // ---- K0: rebuild trigger (reference handle_neighbor_lists) -----------------------------------
  const int chkBlocks = nblocks(((long long)N + CHK_ITEMS - 1) / CHK_ITEMS);
  s.chkPartial.ensure(chkBlocks);
  k_displacement_check<<<chkBlocks, TPB, 0, s.stream>>>(s.R.p, s.R0.p, N, s.chkPartial.p);
  k_displacement_final<<<1, 256, 0, s.stream>>>(s.chkPartial.p, chkBlocks, s.scalars.p + 8);
  stats_.launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 8, s.scalars.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  const bool rebuild = s.h_scalars[8] > s.skinSq;

  const double invL2 = 1.0 / (Lbox * Lbox);
  if (rebuild) {
    int M = (int)std::floor(2 * Lbox / s.xRc);
    M = std::max(M, 5);
    if (2 * Lbox / s.xRc < 5.0)
      fatal("neighbor list handling", "box length is smaller than 2.5*(Rc + skin): the reference's 5x5x5 cell stencil cannot cover the cutoff sphere");
    if (M > 1019) fatal("neighbor list handling", "more than 1019 cells per dimension are not supported");
    s.grid.M = M;
    s.grid.Mx = M + 4;
    const long long ncell = (long long)s.grid.Mx * s.grid.Mx * s.grid.Mx;
    s.Rs.ensure(3 * (size_t)N);
    s.atomCell.ensure(N);
    s.atomFloor.ensure(3 * (size_t)N);
    s.cellCount.ensure(ncell + 1);
    s.cellStart.ensure(ncell + 1);
    s.cellFill.ensure(ncell + 1);
    CUDA_CHECK(cudaMemsetAsync(s.cellCount.p, 0, (ncell + 1) * sizeof(int), s.stream));
    CUDA_CHECK(cudaMemsetAsync(s.cellFill.p, 0, (ncell + 1) * sizeof(int), s.stream));
    k_bin<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, N, Lbox, s.grid, s.Rs.p, s.atomCell.p, s.atomFloor.p, s.cellCount.p);
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
    s.sRs.ensure(3 * (size_t)Next, 1.1);
    s.nbrCount.ensure(Next, 1.1);
    s.pos.ensure(Next, 1.1);
    k_fill<<<nblocks(N), TPB, 0, s.stream>>>(N, s.grid, s.atomCell.p, s.cellStart.p, s.cellFill.p, s.slotAtom.p,
                                             s.slotImg.p, s.slotCell.p);
    k_place<<<nblocks(Next), TPB, 0, s.stream>>>(Next, s.slotAtom.p, s.slotImg.p, s.slotCell.p, s.cellStart.p,
                                                 s.atomFloor.p, s.Rs.p, s.type.p, s.body.p, s.sMeta.p, s.sCell.p,
                                                 s.sGhost.p, s.sType.p, s.sBody.p, s.sRs.p, s.nbrCount.p);
    stats_.launches += 4;
    // capacity guess from the mean density; grown on overflow
    if (s.cap == 0) {
      double nbar = (double)N / (Lbox * Lbox * Lbox) * (4.0 / 3.0) * 3.14159265358979323846 * s.xRc * s.xRcSq;
      s.cap = std::max(16, (int)(1.35 * nbar) + 16);
    }
    const long long ntiles = ((long long)Next + TILE - 1) / TILE;
    for (;;) {
      s.nbr.ensure((size_t)ntiles * s.cap * TILE);
      CUDA_CHECK(cudaMemsetAsync(s.flags.p, 0, 4 * sizeof(int), s.stream));
      BuildArgs b;
      b.Next = Next; b.cap = s.cap; b.nt = s.nt; b.g = s.grid;
      b.xRc2s = s.xRcSq * invL2;
      b.cellStart = s.cellStart.p; b.sMeta = s.sMeta.p; b.sCell = s.sCell.p; b.sGhost = s.sGhost.p;
      b.sType = s.sType.p; b.sBody = s.sBody.p; b.sRs = s.sRs.p; b.exFirst = s.exFirst.p; b.exItem = s.exItem.p;
      b.interact = s.interact.p; b.nbr = s.nbr.p; b.nbrCount = s.nbrCount.p; b.flags = s.flags.p;
      if (timing_) CUDA_CHECK(cudaEventRecord(s.ev0, s.stream));
      k_build_list<<<nblocks(Next), TPB, 0, s.stream>>>(b);
      if (timing_) CUDA_CHECK(cudaEventRecord(s.ev1, s.stream));
      stats_.launches += 1;
      stats_.build_launches += 1;
      int hflags[4];
      CUDA_CHECK(cudaMemcpyAsync(hflags, s.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      if (timing_) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        stats_.build_ms += ms;
      }
      if (!hflags[1]) break;
      s.cap = (int)(hflags[0] * 1.15) + 8;   // overflow: regrow to the observed maximum and redo
    }
    CUDA_CHECK(cudaMemcpyAsync(s.R0.p, s.R.p, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
    s.Lbuild = Lbox;
    s.list_valid = true;
    stats_.cells_per_dim = M;
  }
  neighbor_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();

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
  const int grid = nblocks(Next);
  s.partial.ensure((size_t)grid * 5);
  ForceArgs a;
  a.Next = Next; a.cap = s.cap; a.nt = s.nt;
  a.Rc2s = (lt.useInRc ? s.InRcSq : s.RcSq) * invL2;
  a.L = Lbox; a.invL = 1.0 / Lbox; a.invL2 = invL2;
  a.pos = s.pos.p; a.nbr = s.nbr.p; a.nbrCount = s.nbrCount.p; a.sMeta = s.sMeta.p; a.sGhost = s.sGhost.p; a.sType = s.sType.p;
  a.delta = (s.has_delta && s.nbodies != 0) ? s.delta.p : nullptr;
  a.tab = s.tabs[layer0].p;
  a.single = lt.pair[0];
  a.coul = lt.coul;
  a.F = Fl; a.partial = s.partial.p;

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
  if (s.nt == 1 && uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_NONE && !a.q4_quirk)
    launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false>(a, compute, grid, 0, s.stream);
  else if (s.nt == 1 && uniform && pk == K_PAIR_LJ_CUT && pm == M_SHIFTED_FORCE && ck == K_COUL_NONE && !a.q4_quirk)
    launch_force<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true>(a, compute, grid, 0, s.stream);
  else if (s.nt == 1 && uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_SF && cm == M_NONE)
    launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true>(a, compute, grid, 0, s.stream);
  else
    launch_force<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true>(a, compute, grid, smem_dyn, s.stream);
  if (timing_) CUDA_CHECK(cudaEventRecord(s.ev1, s.stream));
  k_reduce_partials<<<1, 256, 0, s.stream>>>(s.partial.p, grid, 5, 0.5, s.scalars.p);
  stats_.launches += 3;
  stats_.force_launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(s.h_scalars, s.scalars.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  CUDA_CHECK(cudaGetLastError());
  if (timing_) {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    stats_.force_ms += ms;
  }
  out.Epair = s.h_scalars[0];
  out.Ecoul = s.h_scalars[1];
  out.Wpair = s.h_scalars[2];
  out.Wcoul = s.h_scalars[3];
  out.Wbody = s.h_scalars[4];
  return rebuild;
}

void Engine::boost(int layer0, double CP, double CF, bool want_kinetic, KineticScalars& ke) {
  Impl& s = *d_;
  const int grid = nblocks(s.N);
  s.partial.ensure((size_t)grid * 5);
  k_boost<<<grid, TPB, 0, s.stream>>>(s.N, CP, CF, s.P.p, s.F.p + (size_t)layer0 * 3 * s.N, s.invMass.p,
                                      want_kinetic ? 1 : 0, s.partial.p);
  stats_.launches += 1;
  if (want_kinetic) {
    k_reduce_partials<<<1, 256, 0, s.stream>>>(s.partial.p, grid, 3, 1.0, s.scalars.p + 10);
    stats_.launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 10, s.scalars.p + 10, 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    for (int x = 0; x < 3; ++x) ke.twoKE[x] = s.h_scalars[10 + x];
  }
}

void Engine::displace(double CR, double CP) {
  Impl& s = *d_;
  k_displace<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p);
  stats_.launches += 1;
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
  k_count_interacting<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, s.cap, Rc2s, s.pos.p, s.nbr.p, s.nbrCount.p, s.counter.p);
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
  k_export_pairs<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, s.cap, s.nbr.p, s.nbrCount.p, s.sMeta.p, dp.p, capacity, s.counter.p);
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

}  // namespace emdee

// ---- FP64 issue-rate microbenchmark (roofline denominator that MEASURED_PEAKS.json does not carry) ----
namespace emdee {
namespace {
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
