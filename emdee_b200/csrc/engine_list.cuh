// engine_list.cuh -- neighbor-list side of the hot path: rebuild criterion, cell binning / sort with ghost images, and the
// Verlet-list build (reference src/neighbor_lists.f90:41-59, 63-171, 199-300).
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
#pragma once

namespace emdee {
namespace {

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
                                             unsigned int* __restrict__ ticket, double* __restrict__ result,
                                             HostSlot* hs = nullptr, unsigned long long seq = 0ull) {
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
    const double value = __dadd_rn(__dadd_rn(s.m, __dmul_rn(2.0, __dsqrt_rn(__dmul_rn(s.m, s.n)))), s.n);
    result[0] = value;
    *ticket = 0u;
    if (hs != nullptr) {
      hs->v[0] = value;
      slot_publish(hs, seq);
    }
  }
}

__global__ void __launch_bounds__(TPB) k_displacement_check(const double* __restrict__ R,
                                                            const double* __restrict__ R0, int N,
                                                            MaxNext* __restrict__ partial,
                                                            unsigned int* __restrict__ ticket,
                                                            double* __restrict__ result, HostSlot* hs,
                                                            unsigned long long seq) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  MaxNext s = (i < N) ? mn_atom(R, R0, i) : mn_identity();
  s = block_ordered_reduce(s);
  check_finish(s, partial, ticket, result, hs, seq);
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

// `list` (several GPUs between full uploads): only the n listed atoms are binned -- the atoms this rank owned at the last
// build plus what the neighbors just sent -- so that the work follows the slab, not the whole system
__global__ void __launch_bounds__(TPB) k_bin(const double* __restrict__ R, int N, double L, GridDesc g,
                                             double* __restrict__ Rs, int* __restrict__ atomCell,
                                             int* __restrict__ atomFloor, unsigned char* __restrict__ owned,
                                             const unsigned char* __restrict__ known, int* __restrict__ cellCount,
                                             const int* __restrict__ list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (list != nullptr) i = list[i];   // N is the list length
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
                                              int* __restrict__ slotCell, const int* __restrict__ list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (list != nullptr) i = list[i];
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

// row append without a branch: if (on) { if (cnt < cap) *p = v; ++cnt; }   (p = the row's slot `cnt`)
__device__ __forceinline__ void append_if(int* p, int v, bool on, int& cnt, int cap) {
#if defined(__CUDACC__)
  asm volatile(
      "{\n\t.reg .pred q, s;\n\tsetp.ne.s32 q, %3, 0;\n\tsetp.lt.and.s32 s, %0, %4, q;\n\t@s st.global.b32 [%1], %2;\n\t@q add.s32 %0, %0, 1;\n\t}"
      : "+r"(cnt)
      : "l"(p), "r"(v), "r"((int)on), "r"(cap)
      : "memory");
#else   // tests/cusim emulation build
  if (on) {
    if (cnt < cap) *p = v;
    ++cnt;
  }
#endif
}

// (A flat form -- runs pre-clipped in lock step into shared memory, then ONE predicated loop with a per-lane trip count
// -- was measured in round 2: 3.4x fewer warp-iterations but ~40 SASS instructions per candidate against ~16 here, and
// 29 % slower overall (0.847 vs 0.658 ms at LJ-1M, profiles/r2c_force_build_variants.txt); removed.)
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
    const float r2_accept = a.r2_accept, r2_reject = a.r2_reject;
    const int cap = a.cap;
    const bool slow_masks = !a.all_interact || x0 < x1;
    // bounds [f0, f1) of the entries of stencil row r = 5*(dz+2) + (dy+2) that can hold a neighbor (empty: f0 = f1)
    auto run_bounds = [&](int r, int& f0, int& f1) {
      const int dz = r / 5 - 2, dy = r - 5 * (r / 5) - 2;
      const float zlo = (float)(ez + dz - 2 + a.g.z0) * w, zhi = zlo + w;   // local layer -> global coordinate
      const float gz = fmaxf(0.0f, fmaxf(zlo - pf.z, pf.z - zhi) - slack);
      const float ylo = (float)(ey + dy - 2) * w, yhi = ylo + w;
      const float gy = fmaxf(0.0f, fmaxf(ylo - pf.y, pf.y - yhi) - slack);
      const float rem = rc2 - gz * gz - gy * gy;
      f0 = f1 = 0;
      if (rem <= 0.0f) return;
      const float hx = sqrtf(rem) + slack;
      int cl = (int)floorf((pf.x - hx) * (float)a.g.M) + 2;   // `slack` (inside hx) exceeds the FP32 rounding here
      int ch = (int)floorf((pf.x + hx) * (float)a.g.M) + 2;
      cl = max(cl, ex - 2);
      ch = min(ch, ex + 2);
      const int row = Mx * ((ey + dy) + Mx * (ez + dz));
      f0 = a.cellStart[row + cl];
      f1 = a.cellStart[row + ch + 1];
    };
    // Software pipeline (the kernel waits on loads, not on issue slots): the bounds of the NEXT row and the NEXT candidate's
    // position are requested before the current one is examined.
    int f0, f1;
    run_bounds(0, f0, f1);
    for (int r = 0; r < 25; ++r) {
      int nf0 = 0, nf1 = 0;
      if (r + 1 < 25) run_bounds(r + 1, nf0, nf1);
      if (f0 < f1) {
        // Straight-line candidate test: in a warp some lane accepts at almost every iteration, so a branch around the
        // accept path costs every lane both paths. The common case (FP32 decides, no exclusion row, every type pair
        // interacts) runs without a branch -- predicated store, count += ok -- and one rare branch covers the rest.
        auto examine = [&](const float4& qf, int f) {
          const float dxf = pf.x - qf.x, dyf = pf.y - qf.y, dzf = pf.z - qf.z;
          const float r2f = fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf));
          const bool in32 = r2f < r2_accept;
          const bool other = (f != e) & (__float_as_int(qf.w) != body_i);
          if ((!in32 & !(r2f > r2_reject)) | (in32 & other & slow_masks)) {   // rare
            bool ok = other;
            if (!in32) {   // inside the FP32 uncertainty band: decide exactly
              const double4 rj = a.sRs[f];
              const double r2 = __dadd_rn(__dadd_rn(strict_pbc_sq(ri.x, rj.x), strict_pbc_sq(ri.y, rj.y)),
                                          strict_pbc_sq(ri.z, rj.z));
              ok = ok && (r2 < a.xRc2s);
            }
            if (ok && slow_masks) {   // some type pair does not interact, or this atom has an exclusion row
              ok = a.all_interact || a.interact[type_i * a.nt + a.sType[f]];
              if (ok && x0 < x1) {
                const int atom_j = a.sMeta[f].x;
                for (int q = x0; ok && q < x1; ++q) ok = (a.exItem[q] != atom_j);
              }
            }
            if (ok) {
              if (cnt < cap) out[(size_t)cnt * TILE] = f;
              ++cnt;
            }
            return;
          }
          append_if(out + (size_t)cnt * TILE, f, in32 & other, cnt, cap);
        };
        // (two candidates per iteration with both positions requested ahead measured the same 0.56 ms at 62 registers
        // instead of 48 -- GPU call 24 -- and was not kept)
        const float4* qp = a.sPosF + f0;
        float4 qn = __ldg(qp);
        for (int f = f0; f < f1; ++f) {
          const float4 qf = qn;
          ++qp;
          if (f + 1 < f1) qn = __ldg(qp);
          examine(qf, f);
        }
      }
      f0 = nf0;
      f1 = nf1;
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

}  // namespace
}  // namespace emdee
