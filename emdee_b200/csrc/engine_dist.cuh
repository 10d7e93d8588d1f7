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

// Block-wide then grid-wide (last block, fixed order) fold of MaxIdx states; thread 0 of the last block writes the result.
__device__ __forceinline__ void maxidx_finish(MaxIdx v, MaxIdx* __restrict__ partial, unsigned int* __restrict__ ticket,
                                              MaxIdx* __restrict__ result) {
  __shared__ MaxIdx sm[TPB];
  __shared__ bool last;
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

// phase 1 / phase 2 of the criterion over the rank's OWNED atoms (compact ascending list): max of d_i over owned i < below
__global__ void __launch_bounds__(TPB) k_check_owned(int n, const int* __restrict__ list, const double* __restrict__ R,
                                                     const double* __restrict__ R0, long long below,
                                                     MaxIdx* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                     MaxIdx* __restrict__ result) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  MaxIdx v;
  v.m = -1.0 / 0.0;
  v.i = 0x7fffffffffffffffLL;
  if (k < n) {
    const long long i = list[k];
    if (i < below) {
      v.m = mn_atom(R, R0, i).m;
      v.i = i;
    }
  }
  maxidx_finish(v, partial, ticket, result);
}

// Free-atom kick of the owned atoms (k_boost through the owned list: the work is O(atoms of this rank), not O(N)); the
// two sets of kinetic sums and the speculative `crit` as in k_boost
constexpr int BOOST_EPT = 4;
__global__ void __launch_bounds__(TPB) k_boost_owned(int n, const int* __restrict__ list, double CP, double CF,
                                                     double* __restrict__ P, const double* __restrict__ F,
                                                     const double* __restrict__ invMass, int want_ke,
                                                     double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                     double* __restrict__ out, const double* __restrict__ crit, double skinSq) {
  __shared__ double red[TPB / 32][6];
  if (crit != nullptr && __ldcg(crit) > skinSq) return;
  // BOOST_EPT list entries per thread, a block-stride apart (coalesced): the grid-wide finish folds 4x fewer block partials
  double ke[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < BOOST_EPT; ++j) {
    const int k = blockIdx.x * (blockDim.x * BOOST_EPT) + j * blockDim.x + threadIdx.x;
    if (k < n) {
      const size_t a = (size_t)list[k];
      const double im = invMass[a];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const double f = F[3 * a + x];
        const double q = __dadd_rn(__dmul_rn(CP, P[3 * a + x]), __dmul_rn(CF, f));
        P[3 * a + x] = q;
        ke[x] += __dmul_rn(__dmul_rn(im, q), q);
        const double q2 = __dadd_rn(__dmul_rn(CP, q), __dmul_rn(CF, f));
        ke[3 + x] += __dmul_rn(__dmul_rn(im, q2), q2);
      }
    }
  }
  if (!want_ke) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int x = 0; x < 6; ++x) {
    double v = ke[x];
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
  grid_finish<6>(mine, partial, ticket, out, 1.0);
}

// Drift of the owned atoms fused with phase 1 of the criterion on the NEW coordinates: (max d_i, first index) over the list
// (no deferred kick here, unlike k_displace: over the owned list the momenta and forces are gathered, and the fused form
// measured 1 % slower than the two kernels on 2 GPUs, profiles/r2f_bench_2gpu.txt)
__global__ void __launch_bounds__(TPB) k_displace_owned(int n, const int* __restrict__ list, double CR, double CP,
                                                        double* __restrict__ R, const double* __restrict__ P,
                                                        const double* __restrict__ invMass, const double* __restrict__ R0,
                                                        MaxIdx* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                        MaxIdx* __restrict__ result) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  MaxIdx v;
  v.m = -1.0 / 0.0;
  v.i = 0x7fffffffffffffffLL;
  if (k < n) {
    const size_t a = (size_t)list[k];
    const double im = invMass[a];
    double r[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      r[x] = __dadd_rn(__dmul_rn(CR, R[3 * a + x]), __dmul_rn(__dmul_rn(CP, P[3 * a + x]), im));
      R[3 * a + x] = r[x];
    }
    const double dx = __dsub_rn(r[0], R0[3 * a]), dy = __dsub_rn(r[1], R0[3 * a + 1]), dz = __dsub_rn(r[2], R0[3 * a + 2]);
    v.m = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    v.i = (long long)a;
  }
  maxidx_finish(v, partial, ticket, result);
}

// Global rebuild decision from every rank's phase-1 state (all[r] for the peers, all[world] for this rank), identical on
// every rank:  crit[0] = what the speculative pair kernel compares with skinSq (+inf = do not compute),
//              crit[1] = code: 0 no rebuild, 1 rebuild, 2 undecided (maximum <= skin^2 < 4*maximum with i* > 0: phase 2 needed),
//              crit[2] = i*, crit[3] = maximum   (reference neighbor_lists.f90:41-59)
constexpr int CRIT_DIST = 16;   // offset of this block in Engine::Impl::scalars
__global__ void k_decide(int world, int rank, const MaxIdx* __restrict__ all, double skinSq, double* __restrict__ crit) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  MaxIdx g = all[world];
  for (int r = 0; r < world; ++r)
    if (r != rank) g = mi_better(g, all[r]);
  int code;
  if (g.m > skinSq) code = 1;
  else if (__dmul_rn(4.0, g.m) <= skinSq) code = 0;
  else if (g.i == 0)   // `next` still holds the first atom's value
    code = (__dadd_rn(__dadd_rn(g.m, __dmul_rn(2.0, __dsqrt_rn(__dmul_rn(g.m, g.m)))), g.m) > skinSq) ? 1 : 0;
  else code = 2;
  crit[0] = (code == 0) ? 0.0 : 1.0 / 0.0;
  crit[1] = (double)code;
  crit[2] = (double)g.i;
  crit[3] = g.m;
}

// force scalars (already summed over the ranks) + the decision block -> pinned host slot:
// v[0..4] scalars, v[5] code, v[6] i*; when the pair kernel did not run (code != 0) v[0] carries `maximum` instead
__global__ void k_publish_force(const double* __restrict__ scalars, const double* __restrict__ crit, HostSlot* hs,
                                unsigned long long seq) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int q = 0; q < 5; ++q) hs->v[q] = scalars[q];
  hs->v[5] = 0.0;
  hs->v[6] = 0.0;
  if (crit != nullptr) {
    hs->v[5] = crit[1];
    hs->v[6] = crit[2];
    if (crit[1] != 0.0) hs->v[0] = crit[3];
  }
  slot_publish(hs, seq);
}

// n kinetic sums (already summed over the ranks) -> pinned host slot
__global__ void k_publish_n(int n, const double* __restrict__ src, HostSlot* hs, unsigned long long seq) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int q = 0; q < n; ++q) hs->v[q] = src[q];
  slot_publish(hs, seq);
}

// ================================================================================================
// NVLink peer path (one process per GPU, every rank's mailbox and coordinate array mapped into every peer with cudaIpc):
// the per-step communication is done by these kernels with plain peer stores -- no NCCL call between two rebuilds.
//   k_push_step    halo coordinates written straight into the neighbors' R arrays (same atom index on every rank: the
//                  arrays are full-length), this rank's criterion state into every peer's mailbox, then the flags
//   k_wait_decide  waits for the neighbors' flags and the peers' criterion states, takes the rebuild decision (k_decide)
//   k_reduce_small all-reduce of <= 6 doubles: contribution stored into every peer's mailbox, sum in RANK ORDER (the same
//                  bits on every rank), result + decision published to the pinned host slot
// Flow control: a rank pushes the halo of step n+1 only after its host has the reduced scalars of step n, and a peer
// contributes to that reduction after its own force kernel of step n -- so nobody's R is overwritten while it is being read.
// Mail slots are double-buffered by sequence parity (a rank can be at most one collective ahead of a peer).
// ================================================================================================
constexpr int PEER_MAX = 16;
constexpr int MAIL_DOUBLES = 14;
struct alignas(128) PeerMail {
  double v[MAIL_DOUBLES];
  unsigned long long seq;   // written last (release.sys)
  unsigned long long pad;
};
struct PeerBox {
  PeerMail red[2][PEER_MAX];
  PeerMail crit[2][PEER_MAX];
  unsigned long long haloFrom[2];   // [0] written by the rank below, [1] by the rank above: exchange sequence number
  unsigned long long pad[6];
};
struct PeerPtrs {
  PeerBox* box[PEER_MAX];   // box[rank] is this rank's own
};

#if defined(__CUDACC__)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
#else   // tests/cusim emulation build: single address space, no peers (the path is never taken there)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { *p = v; }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { return *p; }
__device__ __forceinline__ void st_sys(double* p, double v) { *p = v; }
__device__ __forceinline__ double ld_sys(const double* p) { return *p; }
#endif

__global__ void __launch_bounds__(TPB) k_push_step(int nUp, const int* __restrict__ listUp, int nDn, const int* __restrict__ listDn,
                                                   const double* __restrict__ R, double* __restrict__ Rup, double* __restrict__ Rdn,
                                                   int with_halo, int with_crit, const MaxIdx* __restrict__ myCrit, PeerPtrs peers,
                                                   int world, int rank, int up, int dn, unsigned long long seq,
                                                   unsigned int* __restrict__ ticket) {
  __shared__ bool last;
  // one thread per DOUBLE (3 per halo atom): consecutive lanes write consecutive words of the neighbor's array wherever the
  // halo atoms are consecutive in the atom index, so the peer stores leave the SM as full sectors instead of 8-byte
  // fragments 24 bytes apart; one system fence per block (by the thread that then takes the ticket) instead of one per thread
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (with_halo && j < 3LL * (nUp + nDn)) {
    const int k = (int)(j / 3);
    const int x = (int)(j - 3LL * k);
    const bool isUp = k < nUp;
    const size_t a = (size_t)(isUp ? listUp[k] : listDn[k - nUp]);
    double* dst = isUp ? Rup : Rdn;
    dst[3 * a + x] = R[3 * a + x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (take_ticket(ticket) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x == 0) *ticket = 0u;
  const int t = threadIdx.x;
  if (with_crit && t < world && t != rank) {
    PeerMail* m = &peers.box[t]->crit[seq & 1ull][rank];
    st_sys(&m->v[0], myCrit->m);
    st_sys(&m->v[1], __longlong_as_double(myCrit->i));
    st_release_sys(&m->seq, seq);
  }
  if (with_halo && t == PEER_MAX) st_release_sys(&peers.box[up]->haloFrom[0], seq);       // I am the rank below `up`
  if (with_halo && t == PEER_MAX + 1) st_release_sys(&peers.box[dn]->haloFrom[1], seq);   // and the rank above `dn`
}

__global__ void k_wait_decide(PeerBox* mine, int world, int rank, int with_halo, int with_crit, unsigned long long seq,
                              const MaxIdx* __restrict__ myCrit, double skinSq, double* __restrict__ crit) {
  __shared__ MaxIdx got[PEER_MAX];
  const int t = threadIdx.x;
  if (with_crit && t < world) {
    if (t == rank) {
      got[t] = *myCrit;
    } else {
      const PeerMail* m = &mine->crit[seq & 1ull][t];
      while (ld_acquire_sys(&m->seq) != seq) {
      }
      got[t].m = ld_sys(&m->v[0]);
      got[t].i = __double_as_longlong(ld_sys(&m->v[1]));
    }
  }
  if (with_halo && (t == PEER_MAX || t == PEER_MAX + 1)) {
    while (ld_acquire_sys(&mine->haloFrom[t - PEER_MAX]) != seq) {
    }
  }
  __syncthreads();
  if (t != 0 || !with_crit) return;
  MaxIdx g = got[0];
  for (int r = 1; r < world; ++r) g = mi_better(g, got[r]);
  int code;
  if (g.m > skinSq) code = 1;
  else if (__dmul_rn(4.0, g.m) <= skinSq) code = 0;
  else if (g.i == 0)
    code = (__dadd_rn(__dadd_rn(g.m, __dmul_rn(2.0, __dsqrt_rn(__dmul_rn(g.m, g.m)))), g.m) > skinSq) ? 1 : 0;
  else code = 2;
  crit[0] = (code == 0) ? 0.0 : 1.0 / 0.0;
  crit[1] = (double)code;
  crit[2] = (double)g.i;
  crit[3] = g.m;
}

// All-reduce of n1 + n2 <= MAIL_DOUBLES doubles (contiguous in `src`): the first n1 sums go to host slot hs1 -- when
// n1 == 5 they are the force scalars and the decision block rides along as in k_publish_force --, the next n2 to hs2
// (the kinetic sums of a kick that was launched right behind the pair kernel). Either slot may be absent (n = 0).
__global__ void k_reduce_small(int n1, int n2, const double* __restrict__ src, PeerPtrs peers, PeerBox* mine, int world, int rank,
                               unsigned long long seq, double* __restrict__ out, const double* __restrict__ crit, HostSlot* hs1,
                               unsigned long long hseq1, HostSlot* hs2, unsigned long long hseq2) {
  __shared__ double got[PEER_MAX][MAIL_DOUBLES];
  const int t = threadIdx.x;
  const int n = n1 + n2;
  if (t < world) {
    PeerMail* m = &peers.box[t]->red[seq & 1ull][rank];
    for (int q = 0; q < n; ++q) st_sys(&m->v[q], src[q]);
    st_release_sys(&m->seq, seq);
  }
  if (t < world) {
    const PeerMail* m = &mine->red[seq & 1ull][t];
    while (ld_acquire_sys(&m->seq) != seq) {
    }
    for (int q = 0; q < n; ++q) got[t][q] = ld_sys(&m->v[q]);
  }
  __syncthreads();
  if (t != 0) return;
  for (int q = 0; q < n; ++q) {
    double sum = got[0][q];
    for (int r = 1; r < world; ++r) sum += got[r][q];
    out[q] = sum;
    if (q < n1) hs1->v[q] = sum;
    else hs2->v[q - n1] = sum;
  }
  if (n1 == 5) {
    hs1->v[5] = 0.0;
    hs1->v[6] = 0.0;
    if (crit != nullptr) {
      hs1->v[5] = crit[1];
      hs1->v[6] = crit[2];
      if (crit[1] != 0.0) hs1->v[0] = crit[3];
    }
  }
  if (n1 > 0) slot_publish(hs1, hseq1);
  if (n2 > 0) slot_publish(hs2, hseq2);
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
// than skin/2 < one layer between rebuilds -- and re-bins only those (k_mig_flags_listed / k_unpack7_listed below).
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

// ---- compact forms: the same decisions over a LIST of atoms (what this rank owned at the last build, plus what it just
// received) instead of all N, so that a rebuild costs O(atoms of the slab) on every rank ----
__global__ void __launch_bounds__(TPB) k_mig_flags_listed(int n, const int* __restrict__ list, double L, GridDesc g,
                                                          const double* __restrict__ R, unsigned char* __restrict__ fl) {   // 2 flag arrays of n
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t i = (size_t)list[k];
  const double rs = __ddiv_rn(R[3 * i + 2], L);
  int cz = (int)__dmul_rn((double)g.M, __dsub_rn(rs, floor(rs)));
  if (cz >= g.M) cz = g.M - 1;
  const int z1 = g.z0 + g.nzl;
  fl[k] = ((cz - (z1 - 3)) % g.M + g.M) % g.M < 5;                 // layers z1-3 .. z1+1 (periodic)
  fl[(size_t)n + k] = (((g.z0 + 2) - cz) % g.M + g.M) % g.M < 5;   // layers z0-2 .. z0+2
}
__global__ void __launch_bounds__(TPB) k_stamp_listed(int n, const int* __restrict__ list, int* __restrict__ stamp, int epoch) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) stamp[list[k]] = epoch;
}
// received (id, R, P) records: coordinates and momenta stored; mode 0: every record is new to this rank -> id appended at
// cand[k] and stamped; mode 1: ids[k] = id and fresh[k] = "not seen in this rebuild yet" (two ranks only: the same atom can
// arrive in both messages), the caller compacts the fresh ones behind the others
__global__ void __launch_bounds__(TPB) k_unpack7_listed(int n, const double* __restrict__ buf, double* __restrict__ R,
                                                        double* __restrict__ P, int* __restrict__ stamp, int epoch, int mode,
                                                        int* __restrict__ ids, unsigned char* __restrict__ fresh) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* b = buf + 7 * (size_t)k;
  const int a = (int)b[0];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    R[3 * (size_t)a + x] = b[1 + x];
    P[3 * (size_t)a + x] = b[4 + x];
  }
  ids[k] = a;
  if (mode == 0) stamp[a] = epoch;
  else fresh[k] = stamp[a] != epoch;
}
// five flag arrays of n over the listed atoms: send up, send down, receive from below, receive from above, owned
__global__ void __launch_bounds__(TPB) k_halo_flags_listed(int n, const int* __restrict__ list, GridDesc g,
                                                           const int* __restrict__ atomCell, unsigned char* __restrict__ fl) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cz = (atomCell[list[k]] >> 20) & 1023;
  const int z1 = g.z0 + g.nzl;
  const bool own = cz >= g.z0 && cz < z1;
  const int up0 = z1 % g.M, up1 = (z1 + 1) % g.M;
  const int dn0 = (g.z0 - 2 + g.M) % g.M, dn1 = (g.z0 - 1 + g.M) % g.M;
  fl[k] = own && cz >= z1 - 2;
  fl[(size_t)n + k] = own && cz < g.z0 + 2;
  fl[2 * (size_t)n + k] = !own && (cz == dn0 || cz == dn1);
  fl[3 * (size_t)n + k] = !own && (cz == up0 || cz == up1);
  fl[4 * (size_t)n + k] = own;
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

// local I/O (EmDeeX_tune "local_io"): three doubles per listed atom between the caller's pinned, device-mapped host array
// and the device array, both indexed by the atom; dst[a] = src[a] for a in list (zero-copy over PCIe, no staging buffer).
// A warp moves the 96 doubles of its 32 listed atoms element-wise (lane l takes elements l, l+32, l+64), so a run of
// consecutive atom indices -- what a slab of a lattice-ordered system consists of -- becomes full, contiguous sectors on
// the PCIe side instead of 8-byte pieces at a 24-byte stride.
__global__ void __launch_bounds__(TPB) k_copy3_listed(int n, const int* __restrict__ list, const double* __restrict__ src,
                                                      double* __restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int mine = (k < n) ? list[k] : -1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int e = c * 32 + lane;          // element of the warp's 96
    const int from = e / 3;               // lane that holds the atom index
    const int comp = e - 3 * from;
    const int a = __shfl_sync(0xffffffffu, mine, from);
    if (a >= 0) dst[3 * (size_t)a + comp] = src[3 * (size_t)a + comp];
  }
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
