// engine_brick.cuh -- opt-in brick path (EMDEE_BRICKS=1): shared-memory staging with cp.async.bulk + mbarrier, 16-bit local lists.
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
#pragma once

namespace emdee {
namespace {

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
#if defined(__CUDACC__)
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
#else   // tests/cusim emulation build: the copy is synchronous; the barrier word counts outstanding bytes
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int) { *bar = 0ull; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  *bar = ((*bar + bytes) & 0xffffffffull) | (1ull << 32);   // low word: bytes still expected; bit 32: arrived
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  std::memcpy(dst, src, bytes);
  *bar = (*bar & ~0xffffffffull) | ((*bar - bytes) & 0xffffffffull);
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned int) {
  if (*bar == (1ull << 32)) return true;
  cusim::yield();
  return false;
}
#endif

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

}  // namespace
}  // namespace emdee
