// engine.cu -- CUDA hot path (sm_100a): cell binning, cell sort with periodic ghost images, Verlet
// list build, pair force/energy/virial kernels, device-resident velocity-Verlet pieces.
// This file holds the device-resident state (Engine::Impl) and the host side (launch logic, multi-GPU
// collectives); the kernels live in the engine_*.cuh parts included below, same translation unit:
//   engine_common.cuh  error macros, DBuf, grid_finish      engine_brick.cuh     opt-in shared-memory brick path
//   engine_list.cuh    criterion, binning/sort, list build  engine_dynamics.cuh  k_boost, k_displace
//   engine_force.cuh   pair term + force kernels             engine_dist.cuh      multi-GPU kernels
//   engine_extra.cuh   pair export, rdf histogram, FP64 probe   engine_bodies.cuh    rigid-body integrator, verlet_step bookkeeping
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
//   k_body_* / k_shadow_*                           <- src/ArBee.f90, src/EmDeeData.f90:157-189,823-922, src/EmDeeCode.f90:1107-1211
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
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine_common.cuh"
#include "engine_list.cuh"
#include "engine_force.cuh"
#include "engine_dynamics.cuh"
#include "engine_bodies.cuh"
#include "engine_bonded.cuh"
#include "engine_ewald.cuh"
#include "engine_dist.cuh"
#include "engine_extra.cuh"

namespace emdee {

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

  // typed path (EMDEE_TYPED, opt-in experiment): compact per-layer tables, built on first use
  std::vector<DBuf<TypedEntry>> ttabs;
  std::vector<int> typedState;     // per layer: 0 = not examined, 1 = eligible, -1 = not eligible
  std::vector<int> typedPM;

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
  // compact rebuild path: the atoms this rank bins at a rebuild (previously owned + just received), a per-atom stamp that
  // tells which atoms are in that list, and scratch for sorting the halo lists of the NCCL-per-step mode
  DBuf<int> candList, stamp, sortTmp, recvIds;
  int nCand = 0, epoch = 0;
  bool compact = false;            // the current rebuild runs over candList
  DBuf<int> ownedList;             // compact ascending list of the atoms this rank owns (per-step kernels run over it)
  int nOwn = 0;
  bool mi_fresh = false;           // miResult[world] holds phase 1 of the criterion for the current coordinates (k_displace_owned)
  // local I/O (EmDeeX_tune "local_io", several GPUs): coordinate uploads read only the atoms this rank owns or keeps as
  // halo, force downloads write only the atoms it owns, both straight from / to the caller's pinned host array
  bool local_io = false;
  // NVLink peer path (engine_dist.cuh): every rank's mailbox and R array mapped into every peer (cudaIpc)
  bool peer_ok = false;
  PeerBox* box = nullptr;          // this rank's mailbox (device memory, 2 MB allocation of its own)
  PeerPtrs peers{};                // every rank's mailbox as seen from this device
  double* peerR[PEER_MAX] = {};    // every rank's coordinate array as seen from this device
  unsigned long long xseq = 0, rseq = 0;   // exchange / small-reduction sequence numbers (advance in lock step on all ranks)
  long long io_h2d = 0, io_d2h = 0;   // bytes moved by coordinate uploads / force downloads (EmDeeX_io_bytes)

  // rigid bodies (engine_bodies.cuh): CSR of members + SoA state, 27 doubles per body
  int nitems = 0;
  DBuf<int> bFirst, bAtom;
  DBuf<double> bMItem, bD, bState, bPartial, bScalars, shR0, shQ0, shS0;
  DBuf<unsigned char> freeMask;    // per atom: 1 = free atom (integrated by k_boost / k_displace), 0 = body member
  DBuf<unsigned char> ownedFree;   // several GPUs: free atoms this rank owns (refreshed after every distributed rebuild)
  double* h_bscalars = nullptr;    // pinned, 16 doubles
  bool frames_valid = false;
  // bonded terms (engine_bonded.cuh): per-atom CSR of (term, role) references
  int nterms = 0;
  DBuf<BondedTerm> terms;
  DBuf<int> termFirst, termRef;
  // reciprocal-space Ewald (engine_ewald.cuh)
  bool ewald_on = false;
  double ew_alpha = 0, ew_beta = 0;
  int ew_nvecs = 0, ew_ntk = 0;
  DBuf<int> ewN, ewKType;
  DBuf<double> ewPrefac, ewLambda, ewSigma, ewPartial;
  // EmDee_memory_address / EmDee_share_phase_space: coordinates may change behind the engine's back
  bool exposed = false;            // a raw pointer to R, P or F was handed out: every call ends with a stream sync
  bool foreign_R = false;          // R is written by the client or by another system: never trust the cached criterion

  // environment switches, read once at construction (nothing on the per-step path calls getenv)
  bool env_debug = false, env_profile = false, env_no_migrate = false;
  // developer knobs (EmDeeX_tune; tools/force_lab.py): force-kernel variant and L1/shared carveout of the plain-LJ kernel
  int tune_variant = 0, tune_carveout = -1;

  // kick bookkeeping (Engine::boost): the sums of the NEXT identical kick predicted by the last one, and a kick that
  // compute_forces launches itself right behind the pair kernel (Engine::plan_kick)
  bool ke_valid = false;
  double ke_CP = 0, ke_CF = 0, ke_next[3] = {0, 0, 0};
  int ke_layer = -1;
  bool kick_planned = false, kick_done = false, kick_want_ke = false;
  double kick_CP = 0, kick_CF = 0, kick_ke[3] = {0, 0, 0};
  int kick_layer = -1;
  // deferred kick (Engine::boost): a kick whose kinetic sums need no reduction (predicted by the previous kick, or not
  // wanted) is not launched but applied by the drift kernel that follows (k_displace<true>); everything else that reads or
  // writes momenta or forces executes it first (Engine::flush_kick)
  bool defer_on = false, env_no_defer = false;
  double defer_CP = 0, defer_CF = 0;
  int defer_layer = -1;

  // host-visible results: pinned slots the last block of a reducing kernel writes; the host spins on the sequence number
  HostSlot* slots = nullptr;
  unsigned long long slot_seq[NSLOTS] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long next_seq(int slot) { return ++slot_seq[slot]; }
  void wait_slot(int slot) {
    const volatile unsigned long long* flag = &slots[slot].seq;
    const unsigned long long want = slot_seq[slot];
    for (unsigned int spins = 1; *flag != want; ++spins) {
      if ((spins & 0xffffu) == 0u) {   // a kernel that died would leave us spinning: ask the runtime now and then
        cudaError_t err = cudaStreamQuery(stream);
        if (err != cudaSuccess && err != cudaErrorNotReady) {
          std::fprintf(stderr, "Error in CUDA runtime: %s (while waiting for a kernel result).\n", cudaGetErrorString(err));
          std::exit(1);
        }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }

  // non-intrusive kernel timing (EmDeeX_set_kernel_timing): event pairs in a ring, harvested when the ring wraps or when
  // the statistics are read, never by a synchronisation on the step path
  static constexpr int NTIMERS = 256, NKINDS = 8;
  struct Timer { cudaEvent_t a = nullptr, b = nullptr; int kind = -1; };
  Timer timers[NTIMERS];
  int timer_head = 0, last_force_timer = -1;
  double timed_ms[NKINDS] = {0, 0, 0, 0, 0, 0, 0, 0};   // [0] force kernel, [1] list-build kernel, then engine.h TIMER_* kinds
  long long timed_n[NKINDS] = {0, 0, 0, 0, 0, 0, 0, 0};

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
    // EMDEE_NCCL_LIB: explicit path (a site's own NCCL build; the test-suite's shared-memory stand-in)
    const char* names[] = {std::getenv("EMDEE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (n == nullptr || *n == '\0') continue;
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
  s.ttabs.resize(nlayers);
  s.typedState.assign(nlayers, 0);
  s.typedPM.assign(nlayers, 0);
  s.flags.ensure(8);
  s.scalars.ensure(32);
  CUDA_CHECK(cudaMemset(s.scalars.p, 0, 32 * sizeof(double)));
  s.counter.ensure(2);
  s.tickets.ensure(4);
  CUDA_CHECK(cudaMemset(s.tickets.p, 0, 4 * sizeof(unsigned int)));
  s.chkPartial.ensure(nblocks(natoms));
  s.partial.ensure((size_t)nblocks(natoms) * 6);
  CUDA_CHECK(cudaMallocHost(&s.h_scalars, 16 * sizeof(double)));
  CUDA_CHECK(cudaHostAlloc(&s.slots, NSLOTS * sizeof(HostSlot), cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(s.slots, 0, NSLOTS * sizeof(HostSlot));
  s.env_debug = std::getenv("EMDEE_DEBUG") != nullptr;
  s.env_profile = std::getenv("EMDEE_PROFILE") != nullptr;
  s.env_no_migrate = std::getenv("EMDEE_NO_MIGRATE") != nullptr;
  s.env_no_defer = std::getenv("EMDEE_NO_DEFER_KICK") != nullptr;
  if (const char* fv = std::getenv("EMDEE_FORCE_VARIANT")) s.tune_variant = std::atoi(fv);   // developer knob (profiling a lab variant)
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
  if (s.env_profile)
    std::fprintf(stderr, "[emdee profile] host ms per call: check %.3f (n=%lld) rebuild %.3f (n=%lld) force %.3f boost %.3f (n=%lld) displace %.3f (n=%lld)\n",
                 1e3 * s.t_check / std::max(1LL, s.n_force), s.n_force, 1e3 * s.t_rebuild / std::max(1LL, s.n_rebuild), s.n_rebuild,
                 1e3 * s.t_force / std::max(1LL, s.n_force), 1e3 * s.t_boost / std::max(1LL, s.n_boost), s.n_boost,
                 1e3 * s.t_displace / std::max(1LL, s.n_displace), s.n_displace);
  s.R.release(); s.P.release(); s.R0.release(); s.F.release(); s.q.release(); s.invMass.release();
  s.delta.release(); s.type.release(); s.body.release(); s.exFirst.release(); s.exItem.release();
  s.interact.release();
  for (auto& t : s.tabs) t.release();
  for (auto& t : s.ttabs) t.release();
  s.Rs.release(); s.sRs.release(); s.sPosF.release(); s.atomCell.release(); s.atomFloor.release();
  s.cellCount.release(); s.cellStart.release(); s.cellFill.release(); s.slotAtom.release(); s.slotImg.release();
  s.slotCell.release(); s.sMeta.release(); s.sCell.release(); s.sType.release(); s.sBody.release(); s.nbr.release();
  s.nbrCount.release(); s.flags.release(); s.sGhost.release(); s.pos.release(); s.scanTmp.release();
  s.owned.release(); s.scratch3.release(); s.haloFlags.release(); s.selCount.release(); s.miPartial.release(); s.miResult.release();
  for (int k = 0; k < 4; ++k) { s.haloList[k].release(); s.haloBuf[k].release(); }
  s.known.release(); s.migCounts.release(); s.ownedList.release(); s.candList.release(); s.stamp.release(); s.sortTmp.release(); s.recvIds.release();
  s.terms.release(); s.termFirst.release(); s.termRef.release();
  s.ewN.release(); s.ewKType.release(); s.ewPrefac.release(); s.ewLambda.release(); s.ewSigma.release(); s.ewPartial.release();
  s.bFirst.release(); s.bAtom.release(); s.bMItem.release(); s.bD.release(); s.bState.release(); s.bPartial.release();
  s.bScalars.release(); s.shR0.release(); s.shQ0.release(); s.shS0.release(); s.freeMask.release(); s.ownedFree.release();
  if (s.h_bscalars) cudaFreeHost(s.h_bscalars);
  for (int k = 0; k < 2; ++k) { s.migList[k].release(); s.migSend[k].release(); s.migRecv[k].release(); }
  if (s.h_mi) cudaFreeHost(s.h_mi);
#if defined(__CUDACC__)
  if (s.peer_ok)
    for (int r = 0; r < s.world; ++r)
      if (r != s.rank) {
        cudaIpcCloseMemHandle(s.peers.box[r]);
        cudaIpcCloseMemHandle(s.peerR[r]);
      }
  if (s.box) cudaFree(s.box);
#endif
  if (s.comm) nccl().CommDestroy(s.comm);
  s.chkPartial.release(); s.partial.release(); s.scalars.release(); s.counter.release(); s.tickets.release();
  if (s.h_scalars) cudaFreeHost(s.h_scalars);
  if (s.slots) cudaFreeHost(s.slots);
  for (auto& t : s.timers) {
    if (t.a) cudaEventDestroy(t.a);
    if (t.b) cudaEventDestroy(t.b);
  }
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

namespace {
// Maps every rank's mailbox and coordinate array into this process (cudaIpc over NVLink). All-or-nothing across the
// ranks: the flags are all-reduced, so either every rank uses the peer kernels or every rank uses NCCL per step.
void setup_peer_links(Engine::Impl& s) {
#if defined(__CUDACC__)
  if (std::getenv("EMDEE_NO_PEER") != nullptr || s.world > PEER_MAX) return;
  int ok = 1;
  // allocations of their own (>= 2 MB), so that an IPC handle opens at the array's first byte
  const size_t n3 = std::max(3 * (size_t)s.N, (size_t)(2u << 20) / sizeof(double) + 16);
  s.R.release();
  s.R.ensure(n3);
  CUDA_CHECK(cudaMemset(s.R.p, 0, n3 * sizeof(double)));
  void* raw = nullptr;
  CUDA_CHECK(cudaMalloc(&raw, std::max(sizeof(PeerBox), (size_t)(2u << 20))));
  s.box = static_cast<PeerBox*>(raw);
  CUDA_CHECK(cudaMemset(s.box, 0, sizeof(PeerBox)));
  struct Handles { cudaIpcMemHandle_t box, R; };
  Handles mine;
  if (cudaIpcGetMemHandle(&mine.box, s.box) != cudaSuccess || cudaIpcGetMemHandle(&mine.R, s.R.p) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
    std::memset(&mine, 0, sizeof(mine));
  }
  DBuf<unsigned char> all;
  all.ensure((size_t)s.world * sizeof(Handles));
  CUDA_CHECK(cudaMemcpy(all.p + (size_t)s.rank * sizeof(Handles), &mine, sizeof(Handles), cudaMemcpyHostToDevice));
  NCCL_CHECK(nccl().AllGather(all.p + (size_t)s.rank * sizeof(Handles), all.p, sizeof(Handles), ncclChar, s.comm, s.stream));
  std::vector<Handles> h(s.world);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), all.p, (size_t)s.world * sizeof(Handles), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  all.release();
  for (int r = 0; r < s.world && ok; ++r) {
    if (r == s.rank) {
      s.peers.box[r] = s.box;
      s.peerR[r] = s.R.p;
      continue;
    }
    void *pb = nullptr, *pr = nullptr;
    if (cudaIpcOpenMemHandle(&pb, h[r].box, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pr, h[r].R, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    s.peers.box[r] = static_cast<PeerBox*>(pb);
    s.peerR[r] = static_cast<double*>(pr);
  }
  // agree: all ranks or none
  DBuf<int> flag;
  flag.ensure(1);
  CUDA_CHECK(cudaMemcpy(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice));
  NCCL_CHECK(nccl().AllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, s.comm, s.stream));
  CUDA_CHECK(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  flag.release();
  s.peer_ok = ok != 0;
  if (s.env_debug) std::fprintf(stderr, "[emdee r%d] NVLink peer path %s\n", s.rank, s.peer_ok ? "on" : "off (NCCL per step)");
#else
  (void)s;
#endif
}
}  // namespace

void Engine::comm_init(int rank, int world, const void* unique_id) {
  flush_kick();
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
  setup_peer_links(s);
}

namespace {

// X (3N doubles, valid for the atoms each rank owns) -> full array on every rank:
// all-reduce(sum) of the owned parts (every atom is owned by exactly one rank, so the sum is exact)
void gather_full(Engine::Impl& s, double* X) {
  if (s.world <= 1 || !s.owned_valid) return;
  k_mask_owned<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, s.owned.p, X, s.scratch3.p);
  NCCL_CHECK(nccl().AllReduce(s.scratch3.p, X, 3 * (size_t)s.N, ncclDouble, ncclSum, s.comm, s.stream));
}

// per-rebuild: compact the four halo lists and the owned list (all ascending in the atom index) from the cell layers
void build_halo_lists(Engine::Impl& s) {
  const int N = s.N;
  if (s.compact) {
    // over the atoms this rank just binned (candList) instead of all N; stable compactions keep candList's order, which
    // is the same from run to run. The NCCL-per-step mode needs a sender's list and the matching receiver's list in the
    // SAME order (no indices travel): there the four small halo lists are sorted by atom index afterwards.
    const int n = s.nCand;
    s.haloFlags.ensure(5 * (size_t)n + 16, 1.25);   // (slack: the candidate count wobbles from rebuild to rebuild; no re-allocation in steady state)
    s.selCount.ensure(8);
    k_halo_flags_listed<<<std::max(1, nblocks(n)), TPB, 0, s.stream>>>(n, s.candList.p, s.grid, s.atomCell.p, s.haloFlags.p);
    size_t need = 0;
    cub::DeviceSelect::Flagged(nullptr, need, s.candList.p, s.haloFlags.p, (int*)nullptr, s.selCount.p, n, s.stream);
    if (need > s.scanTmpBytes) {
      s.scanTmp.ensure(need);
      s.scanTmpBytes = need;
    }
    s.ownedList.ensure((size_t)n + 16, 1.25);
    for (int k = 0; k < 4; ++k) {
      s.haloList[k].ensure((size_t)n + 16, 1.25);
      cub::DeviceSelect::Flagged(s.scanTmp.p, need, s.candList.p, s.haloFlags.p + (size_t)k * n, s.haloList[k].p, s.selCount.p + k, n,
                                 s.stream);
    }
    cub::DeviceSelect::Flagged(s.scanTmp.p, need, s.candList.p, s.haloFlags.p + 4 * (size_t)n, s.ownedList.p, s.selCount.p + 4, n,
                               s.stream);
    int h[5];
    CUDA_CHECK(cudaMemcpyAsync(h, s.selCount.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    if (s.env_debug) std::fprintf(stderr, "[emdee r%d] halo lists (compact, %d candidates): send up %d dn %d, recv below %d above %d, owned %d\n", s.rank, n, h[0], h[1], h[2], h[3], h[4]);
    for (int k = 0; k < 4; ++k) {
      s.haloCount[k] = h[k];
      s.haloBuf[k].ensure(3 * (size_t)h[k] + 8, 1.2);
      if (!s.peer_ok && h[k] > 1) {   // NCCL-per-step mode: ascending atom index on both sides of every message
        s.sortTmp.ensure((size_t)h[k] + 16, 1.25);
        size_t sb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, sb, s.haloList[k].p, s.sortTmp.p, h[k], 0, 32, s.stream);
        if (sb > s.scanTmpBytes) {
          s.scanTmp.ensure(sb);
          s.scanTmpBytes = sb;
        }
        cub::DeviceRadixSort::SortKeys(s.scanTmp.p, sb, s.haloList[k].p, s.sortTmp.p, h[k], 0, 32, s.stream);
        CUDA_CHECK(cudaMemcpyAsync(s.haloList[k].p, s.sortTmp.p, (size_t)h[k] * sizeof(int), cudaMemcpyDeviceToDevice, s.stream));
      }
    }
    s.nOwn = h[4];
    return;
  }
  s.haloFlags.ensure(4 * (size_t)N);
  s.selCount.ensure(8);
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
  s.ownedList.ensure(N);
  cub::DeviceSelect::Flagged(s.scanTmp.p, need, ids, s.owned.p, s.ownedList.p, s.selCount.p + 4, N, s.stream);
  int h[5];
  CUDA_CHECK(cudaMemcpyAsync(h, s.selCount.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  if (s.env_debug) std::fprintf(stderr, "[emdee r%d] halo lists: send up %d dn %d, recv below %d above %d, owned %d\n", s.rank, h[0], h[1], h[2], h[3], h[4]);
  for (int k = 0; k < 4; ++k) {
    s.haloCount[k] = h[k];
    s.haloBuf[k].ensure(3 * (size_t)h[k] + 8, 1.2);
  }
  s.nOwn = h[4];
}

// Per step, ONE NCCL group (one launch): the ghost positions of the two layers above and below the slab (no reverse
// exchange: full list) and -- with_criterion -- every rank's phase-1 state of the rebuild criterion (16 bytes to and from
// every peer); the decision is then taken on the device by k_decide (same inputs, same result on every rank) and read by
// the speculatively launched pair kernel, so the host does not wait here.
void exchange_step(Engine::Impl& s, bool with_criterion) {
  if (s.world <= 1 || !s.owned_valid) return;
  const bool halo = !s.halo_fresh;
  if (!halo && !with_criterion) return;
  const int up = (s.rank + 1) % s.world, dn = (s.rank + s.world - 1) % s.world;
  if (with_criterion && !s.mi_fresh) {   // coordinates did not come from k_displace_owned (upload, first step after a rebuild)
    const long long all = 0x7fffffffffffffffLL;
    k_check_owned<<<std::max(1, nblocks(s.nOwn)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, s.R.p, s.R0.p, all, s.miPartial.p,
                                                                      s.tickets.p + 2, s.miResult.p + s.world);
    s.mi_fresh = true;
  }
  if (s.peer_ok) {
    // NVLink peer path: halo coordinates go straight into the neighbors' arrays, the criterion state into every mailbox
    const unsigned long long seq = ++s.xseq;
    const int nUp = halo ? s.haloCount[0] : 0, nDn = halo ? s.haloCount[1] : 0;
    k_push_step<<<std::max(1, nblocks(3LL * (nUp + nDn))), TPB, 0, s.stream>>>(nUp, s.haloList[0].p, nDn, s.haloList[1].p, s.R.p, s.peerR[up],
                                                                      s.peerR[dn], halo ? 1 : 0, with_criterion ? 1 : 0,
                                                                      s.miResult.p + s.world, s.peers, s.world, s.rank, up, dn, seq,
                                                                      s.tickets.p + 3);
    k_wait_decide<<<1, 32, 0, s.stream>>>(s.box, s.world, s.rank, halo ? 1 : 0, with_criterion ? 1 : 0, seq, s.miResult.p + s.world,
                                          s.skinSq, s.scalars.p + CRIT_DIST);
    s.halo_fresh = true;
    return;
  }
  if (halo)
    for (int k = 0; k < 2; ++k)
      if (s.haloCount[k] > 0)
        k_pack3<<<nblocks(s.haloCount[k]), TPB, 0, s.stream>>>(s.haloCount[k], s.haloList[k].p, s.R.p, s.haloBuf[k].p);
  NCCL_CHECK(nccl().GroupStart());
  if (halo) {
    NCCL_CHECK(nccl().Send(s.haloBuf[0].p, 3 * (size_t)s.haloCount[0], ncclDouble, up, s.comm, s.stream));
    NCCL_CHECK(nccl().Send(s.haloBuf[1].p, 3 * (size_t)s.haloCount[1], ncclDouble, dn, s.comm, s.stream));
    NCCL_CHECK(nccl().Recv(s.haloBuf[2].p, 3 * (size_t)s.haloCount[2], ncclDouble, dn, s.comm, s.stream));
    NCCL_CHECK(nccl().Recv(s.haloBuf[3].p, 3 * (size_t)s.haloCount[3], ncclDouble, up, s.comm, s.stream));
  }
  if (with_criterion)
    for (int r = 0; r < s.world; ++r) {
      if (r == s.rank) continue;
      NCCL_CHECK(nccl().Send(s.miResult.p + s.world, sizeof(MaxIdx), ncclChar, r, s.comm, s.stream));
      NCCL_CHECK(nccl().Recv(s.miResult.p + r, sizeof(MaxIdx), ncclChar, r, s.comm, s.stream));
    }
  NCCL_CHECK(nccl().GroupEnd());
  if (halo)
    for (int k = 2; k < 4; ++k)
      if (s.haloCount[k] > 0)
        k_unpack3<<<nblocks(s.haloCount[k]), TPB, 0, s.stream>>>(s.haloCount[k], s.haloList[k].p, s.haloBuf[k].p, s.R.p);
  s.halo_fresh = true;
  if (with_criterion) k_decide<<<1, 32, 0, s.stream>>>(s.world, s.rank, s.miResult.p, s.skinSq, s.scalars.p + CRIT_DIST);
}
void halo_exchange(Engine::Impl& s) { exchange_step(s, false); }

// rebuild-time migration: see k_mig_flags_listed. Returns with R, P current for every atom this rank may need.
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
  // Compact path: everything below runs over the atoms this rank owned at the last build (ownedList) and the records it
  // receives -- O(atoms of the slab), not O(N). The atoms to bin afterwards are collected in candList:
  //   [previously owned atoms, in ownedList's order | records from below | records from above (minus duplicates)]
  const int n = s.nOwn;
  s.haloFlags.ensure(5 * (size_t)std::max(n, 1) + 16, 1.25);
  s.selCount.ensure(8);
  k_mig_flags_listed<<<std::max(1, nblocks(n)), TPB, 0, s.stream>>>(n, s.ownedList.p, Lbox, s.grid, s.R.p, s.haloFlags.p);
  size_t need = 0;
  cub::DeviceSelect::Flagged(nullptr, need, s.ownedList.p, s.haloFlags.p, (int*)nullptr, s.selCount.p, std::max(n, 1), s.stream);
  if (need > s.scanTmpBytes) {
    s.scanTmp.ensure(need);
    s.scanTmpBytes = need;
  }
  for (int k = 0; k < 2; ++k) {
    s.migList[k].ensure((size_t)n + 16, 1.25);
    cub::DeviceSelect::Flagged(s.scanTmp.p, need, s.ownedList.p, s.haloFlags.p + (size_t)k * n, s.migList[k].p, s.selCount.p + k, n,
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
  if (s.env_debug) std::fprintf(stderr, "[emdee r%d] migrate: send up %d dn %d, recv from-below %d from-above %d\n", s.rank, c[0], c[1], c[2], c[3]);
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
  // candList = previously owned + received
  s.epoch += 1;
  if (s.stamp.n < (size_t)N) {
    s.stamp.ensure(N);
    CUDA_CHECK(cudaMemsetAsync(s.stamp.p, 0, (size_t)N * sizeof(int), s.stream));
  }
  s.candList.ensure((size_t)n + c[2] + c[3] + 16, 1.25);
  s.recvIds.ensure((size_t)c[3] + 16, 1.25);
  if (n > 0) {
    CUDA_CHECK(cudaMemcpyAsync(s.candList.p, s.ownedList.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s.stream));
    k_stamp_listed<<<nblocks(n), TPB, 0, s.stream>>>(n, s.ownedList.p, s.stamp.p, s.epoch);
  }
  if (c[2] > 0)
    k_unpack7_listed<<<nblocks(c[2]), TPB, 0, s.stream>>>(c[2], s.migRecv[0].p, s.R.p, s.P.p, s.stamp.p, s.epoch, 0, s.candList.p + n,
                                                          nullptr);
  int fresh = c[3];
  if (c[3] > 0) {
    if (s.world > 2) {   // the two messages come from different ranks: no atom can be in both
      k_unpack7_listed<<<nblocks(c[3]), TPB, 0, s.stream>>>(c[3], s.migRecv[1].p, s.R.p, s.P.p, s.stamp.p, s.epoch, 0,
                                                            s.candList.p + n + c[2], nullptr);
    } else {             // two ranks: both messages come from the same peer and may name the same atom
      unsigned char* fl = s.haloFlags.p;   // (the migration flags are no longer needed)
      if (s.haloFlags.n < (size_t)c[3] + 16) {
        s.haloFlags.ensure((size_t)c[3] + 16, 1.25);
        fl = s.haloFlags.p;
      }
      k_unpack7_listed<<<nblocks(c[3]), TPB, 0, s.stream>>>(c[3], s.migRecv[1].p, s.R.p, s.P.p, s.stamp.p, s.epoch, 1, s.recvIds.p, fl);
      size_t nb2 = 0;
      cub::DeviceSelect::Flagged(nullptr, nb2, s.recvIds.p, fl, (int*)nullptr, s.selCount.p + 5, c[3], s.stream);
      if (nb2 > s.scanTmpBytes) {
        s.scanTmp.ensure(nb2);
        s.scanTmpBytes = nb2;
      }
      cub::DeviceSelect::Flagged(s.scanTmp.p, nb2, s.recvIds.p, fl, s.candList.p + n + c[2], s.selCount.p + 5, c[3], s.stream);
      CUDA_CHECK(cudaMemcpyAsync(&fresh, s.selCount.p + 5, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
    }
  }
  s.nCand = n + c[2] + fresh;
  s.compact = true;
}

// Distributed rebuild decision, phase 2 (rare: maximum <= skin^2 < 4*maximum and i* > 0): `next` = max of d_i over the
// atoms with index below i*, then the reference's value. Identical on every rank (all inputs are all-gathered).
bool rebuild_needed_phase2(Engine::Impl& s, double maximum, long long istar) {
  k_check_owned<<<std::max(1, nblocks(s.nOwn)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, s.R.p, s.R0.p, istar, s.miPartial.p,
                                                                    s.tickets.p + 2, s.miResult.p + s.world);
  NCCL_CHECK(nccl().AllGather(s.miResult.p + s.world, s.miResult.p, sizeof(MaxIdx), ncclChar, s.comm, s.stream));
  CUDA_CHECK(cudaMemcpyAsync(s.h_mi, s.miResult.p, (size_t)s.world * sizeof(MaxIdx), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  double next = s.h_mi[0].m;
  for (int r = 1; r < s.world; ++r) next = std::max(next, s.h_mi[r].m);
  s.mi_fresh = false;   // miResult[world] now holds the phase-2 state
  const double value = maximum + 2 * std::sqrt(maximum * next) + next;   // reference neighbor_lists.f90:57
  return value > s.skinSq;
}

// energies / virial of all ranks summed and delivered to the host slot together with the device-side rebuild decision:
// one collective, one host wait per force evaluation
void finish_pair_dist(Engine::Impl& s, bool with_decision) {
  if (s.peer_ok) {
    // a planned kick (launched behind the pair kernel, sums at scalars[5..10]) shares the reduction and the host wait
    const bool kick = s.kick_planned && s.kick_want_ke;
    const unsigned long long kseq = kick ? s.next_seq(SLOT_KINETIC) : 0ull;
    k_reduce_small<<<1, 32, 0, s.stream>>>(5, kick ? 6 : 0, s.scalars.p, s.peers, s.box, s.world, s.rank, ++s.rseq, s.scalars.p,
                                           with_decision ? s.scalars.p + CRIT_DIST : nullptr, s.slots + SLOT_FORCE,
                                           s.slot_seq[SLOT_FORCE], s.slots + SLOT_KINETIC, kseq);
    s.wait_slot(SLOT_FORCE);
    if (kick) s.wait_slot(SLOT_KINETIC);
    return;
  }
  NCCL_CHECK(nccl().AllReduce(s.scalars.p, s.scalars.p, 5, ncclDouble, ncclSum, s.comm, s.stream));
  k_publish_force<<<1, 32, 0, s.stream>>>(s.scalars.p, with_decision ? s.scalars.p + CRIT_DIST : nullptr, s.slots + SLOT_FORCE,
                                          s.slot_seq[SLOT_FORCE]);
  s.wait_slot(SLOT_FORCE);
}

// the decision alone (layers without a pair kernel)
void publish_decision(Engine::Impl& s) {
  s.next_seq(SLOT_FORCE);
  k_publish_force<<<1, 32, 0, s.stream>>>(s.scalars.p, s.scalars.p + CRIT_DIST, s.slots + SLOT_FORCE, s.slot_seq[SLOT_FORCE]);
  s.wait_slot(SLOT_FORCE);
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
  s.typedState[layer0] = 0;
  s.tabs[layer0].ensure(t.pair.size());
  CUDA_CHECK(cudaMemcpy(s.tabs[layer0].p, t.pair.data(), t.pair.size() * sizeof(PairEntry), cudaMemcpyHostToDevice));
}

namespace {
// device-side alias of a pinned (page-locked, mapped) host pointer, or nullptr when the memory is pageable
double* mapped_alias(const double* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? static_cast<double*>(at.devicePointer) : nullptr;
}
}  // namespace

void Engine::upload_coordinates(const double* R) {
  flush_kick();
  Impl& s = *d_;
  if (s.local_io && s.world > 1 && s.owned_valid && !s.all_known) {
    // local I/O: this rank reads only the atoms it owns and its halo atoms, straight from the caller's pinned array.
    // Contract (as for the resident dynamics): between two uploads no atom moves further than one cell layer.
    if (const double* src = mapped_alias(R)) {
      if (s.nOwn > 0) k_copy3_listed<<<nblocks(s.nOwn), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, src, s.R.p);
      for (int k = 2; k < 4; ++k)
        if (s.haloCount[k] > 0)
          k_copy3_listed<<<nblocks(s.haloCount[k]), TPB, 0, s.stream>>>(s.haloCount[k], s.haloList[k].p, src, s.R.p);
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      stats_.launches += 3;
      s.io_h2d += 24LL * ((long long)s.nOwn + s.haloCount[2] + s.haloCount[3]);
      s.check_cached = false;
      s.mi_fresh = false;
      s.halo_fresh = true;
      return;
    }
  }
  CUDA_CHECK(cudaMemcpyAsync(s.R.p, R, 3 * (size_t)s.N * sizeof(double), cudaMemcpyHostToDevice, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  s.io_h2d += 24LL * s.N;
  s.has_R = true;
  s.check_cached = false;
  s.mi_fresh = false;
  s.halo_fresh = true;   // every rank uploads the full array
  if (s.world > 1 && s.owned_valid && !s.all_known) {
    s.all_known = true;    // coordinates are full everywhere again ...
    s.p_partial = true;    // ... but the momenta of atoms owned elsewhere are stale here until the next rebuild
  }
}
void Engine::upload_momenta(const double* P) {
  flush_kick();
  d_->ke_valid = false;
  CUDA_CHECK(cudaMemcpy(d_->P.p, P, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
  d_->p_partial = false;
}
void Engine::upload_forces(int layer0, const double* F) {
  flush_kick();
  d_->ke_valid = false;
  CUDA_CHECK(cudaMemcpy(d_->F.p + (size_t)layer0 * 3 * d_->N, F, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyHostToDevice));
}
void Engine::download_coordinates(double* R) {
  flush_kick();
  gather_full(*d_, d_->R.p);   // multi-GPU: collective, every rank must call
  d_->halo_fresh = true;
  CUDA_CHECK(cudaMemcpy(R, d_->R.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_momenta(double* P) {
  flush_kick();
  gather_full(*d_, d_->P.p);
  CUDA_CHECK(cudaMemcpy(P, d_->P.p, 3 * (size_t)d_->N * sizeof(double), cudaMemcpyDeviceToHost));
}
void Engine::download_forces(int layer0, double* F) {
  flush_kick();
  Impl& s = *d_;
  double* Fl = s.F.p + (size_t)layer0 * 3 * s.N;
  if (s.local_io && s.world > 1 && s.owned_valid) {
    // local I/O: only the forces of the atoms this rank owns are written (the caller's array keeps its other entries);
    // not a collective
    if (double* dst = mapped_alias(F)) {
      if (s.nOwn > 0) k_copy3_listed<<<nblocks(s.nOwn), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, Fl, dst);
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      stats_.launches += 1;
      s.io_d2h += 24LL * s.nOwn;
      return;
    }
  }
  gather_full(s, Fl);
  CUDA_CHECK(cudaMemcpy(F, Fl, 3 * (size_t)s.N * sizeof(double), cudaMemcpyDeviceToHost));
  s.io_d2h += 24LL * s.N;
}
void Engine::io_bytes(long long& h2d, long long& d2h) { h2d = d_->io_h2d; d2h = d_->io_d2h; }
int Engine::comm_mode() { return d_->world <= 1 ? 0 : (d_->peer_ok ? 2 : 1); }
void Engine::synchronize() {
  flush_kick(); CUDA_CHECK(cudaStreamSynchronize(d_->stream)); }
void Engine::tune(const char* knob, int value) {
  if (std::strcmp(knob, "force_variant") == 0) d_->tune_variant = value;
  else if (std::strcmp(knob, "carveout") == 0) d_->tune_carveout = value;
  else if (std::strcmp(knob, "local_io") == 0) d_->local_io = value != 0;
  else fatal("tuning", "unknown knob");
}
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

// Plain single-type Lennard-Jones: the benchmark kernel. Variant 0 is what ships; the others exist for tools/force_lab.py
// (EmDeeX_tune "force_variant" / "carveout"), which times them back to back on one resident system: 1 = the form with a
// branch per pair, 3 / 4 = the "gathers only" / "arithmetic only" probes. (Launch shapes, unroll depths and cache policies
// were swept in round 2 -- profiles/r2c_force_build_variants.txt -- and their instantiations removed.)
//   columns: UNROLL, THREADS, MINBLOCKS, index-stream load, position-gather load, PROBE, FORM
#define EMDEE_LJ_VARIANTS(X)                                                \
  X(0, 6, 512, 2, LD_PLAIN, LD_PLAIN, 0, FORM_BRANCHLESS)                   \
  X(1, 6, 512, 2, LD_PLAIN, LD_PLAIN, 0, FORM_DEFAULT)                      \
  X(3, 6, 512, 2, LD_PLAIN, LD_PLAIN, 1, FORM_DEFAULT)                      \
  X(4, 6, 512, 2, LD_PLAIN, LD_PLAIN, 2, FORM_DEFAULT)

void launch_lj_plain(Engine::Impl& s, ForceArgs& a, bool compute) {
  using namespace nb;
  const int v = s.tune_variant;
  switch (v) {
#define EMDEE_LJ_CASE(ID, UN, TH, MB, LL, PL, PR, FO)                                                                       \
    case ID: {                                                                                                             \
      const int grid = nblocks(a.Next, TH);                                                                                \
      s.partial.ensure((size_t)grid * 5);                                                                                  \
      a.partial = s.partial.p;                                                                                             \
      auto kt = k_pair_forces<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, true, UN, TH, MB, LL, PL, PR, FO>;   \
      auto kf = k_pair_forces<K_PAIR_LJ_CUT, M_NONE, K_COUL_NONE, M_NONE, true, false, false, UN, TH, MB, LL, PL, PR, FO>;  \
      if (s.tune_carveout >= 0) {                                                                                          \
        CUDA_CHECK(cudaFuncSetAttribute(kt, cudaFuncAttributePreferredSharedMemoryCarveout, s.tune_carveout));             \
        CUDA_CHECK(cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, s.tune_carveout));             \
      }                                                                                                                    \
      if (compute) kt<<<grid, TH, 0, s.stream>>>(a);                                                                       \
      else kf<<<grid, TH, 0, s.stream>>>(a);                                                                               \
      break;                                                                                                               \
    }
    EMDEE_LJ_VARIANTS(EMDEE_LJ_CASE)
#undef EMDEE_LJ_CASE
    default: fatal("force kernel selection", "unknown force_variant");
  }
}

// ---- typed path (k_pair_forces_typed): eligibility + compact table of one layer ----------------------------------------
// eligible: every pair entry is pair_none or pair_lj_cut, all LJ entries carry the same modifier (none or shifted_force)
bool build_typed_table(const LayerTable& lt, std::vector<TypedEntry>& out, int& pm) {
  pm = -1;
  for (const PairEntry& pe : lt.pair) {
    if (pe.model.kind == nb::K_PAIR_NONE) continue;
    if (pe.model.kind != nb::K_PAIR_LJ_CUT) return false;
    if (pe.model.modifier != nb::M_NONE && pe.model.modifier != nb::M_SHIFTED_FORCE) return false;
    if (pm >= 0 && pm != pe.model.modifier) return false;
    pm = pe.model.modifier;
  }
  if (pm < 0) pm = nb::M_NONE;
  out.clear();
  for (const PairEntry& pe : lt.pair) {
    TypedEntry te;
    const bool lj = pe.model.kind == nb::K_PAIR_LJ_CUT;
    te.a = lj ? pe.model.a : 0.0;        // pair_none == Lennard-Jones of zero strength
    te.b = lj ? pe.model.b : 0.0;
    te.c = lj ? pe.model.c : 1.0;
    te.eshift = lj ? pe.model.eshift : 0.0;
    te.fshift = lj ? pe.model.fshift : 0.0;
    te.kCoul = pe.kCoul;
    te.coulomb = pe.coulomb;
    te.pad = 0;
    out.push_back(te);
  }
  return true;
}

template <int PM, int CK>
void launch_typed(ForceArgs& a, DBuf<double>& partial, const TypedEntry* ttab, bool compute, cudaStream_t st) {
  constexpr int THREADS = 128;
  const int grid = nblocks(a.Next, THREADS);
  partial.ensure((size_t)grid * 5);
  a.partial = partial.p;
  const size_t smem = (size_t)a.nt * a.nt * sizeof(TypedEntry);
  if (a.nt == 2) {
    if (compute) k_pair_forces_typed<PM, CK, true, 2><<<grid, THREADS, 0, st>>>(a, ttab);
    else k_pair_forces_typed<PM, CK, false, 2><<<grid, THREADS, 0, st>>>(a, ttab);
  } else {
    if (compute) k_pair_forces_typed<PM, CK, true, 0><<<grid, THREADS, smem, st>>>(a, ttab);
    else k_pair_forces_typed<PM, CK, false, 0><<<grid, THREADS, smem, st>>>(a, ttab);
  }
}

template <int PM>
bool launch_typed_ck(int ck, ForceArgs& a, DBuf<double>& partial, const TypedEntry* ttab, bool compute, cudaStream_t st) {
  using namespace nb;
  switch (ck) {
    case K_COUL_NONE: launch_typed<PM, K_COUL_NONE>(a, partial, ttab, compute, st); return true;
    case K_COUL_CUT: launch_typed<PM, K_COUL_CUT>(a, partial, ttab, compute, st); return true;
    case K_COUL_SF: launch_typed<PM, K_COUL_SF>(a, partial, ttab, compute, st); return true;
    case K_COUL_DAMPED: launch_typed<PM, K_COUL_DAMPED>(a, partial, ttab, compute, st); return true;
    case K_COUL_DAMPED_SMOOTHED: launch_typed<PM, K_COUL_DAMPED_SMOOTHED>(a, partial, ttab, compute, st); return true;
    case K_COUL_DAMPED_SQUARE_SMOOTHED: launch_typed<PM, K_COUL_DAMPED_SQUARE_SMOOTHED>(a, partial, ttab, compute, st); return true;
    default: return false;
  }
}

// typed path: builds the layer's compact table on first use and launches k_pair_forces_typed.
// Returns false (nothing launched) when the layer is not eligible.
bool try_typed_path(Engine::Impl& s, int layer0, const LayerTable& lt, int ck, ForceArgs& a, bool compute) {
  if (s.typedState[layer0] == 0) {
    std::vector<TypedEntry> tt;
    int pmod = 0;
    if (build_typed_table(lt, tt, pmod)) {
      s.ttabs[layer0].ensure(tt.size());
      CUDA_CHECK(cudaMemcpyAsync(s.ttabs[layer0].p, tt.data(), tt.size() * sizeof(TypedEntry), cudaMemcpyHostToDevice, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));   // `tt` is a local
      s.typedState[layer0] = 1;
      s.typedPM[layer0] = pmod;
      if (s.env_debug)
        std::fprintf(stderr, "[emdee] typed path: layer %d eligible (modifier %d, coulomb kind %d, %d types)\n", layer0, pmod, ck, s.nt);
    } else {
      s.typedState[layer0] = -1;
    }
  }
  if (s.typedState[layer0] != 1) return false;
  return s.typedPM[layer0] == nb::M_NONE ? launch_typed_ck<nb::M_NONE>(ck, a, s.partial, s.ttabs[layer0].p, compute, s.stream)
                                         : launch_typed_ck<nb::M_SHIFTED_FORCE>(ck, a, s.partial, s.ttabs[layer0].p, compute, s.stream);
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

// Kernel timing without a synchronisation on the step path: an event pair from the ring brackets the launch; a pair is
// read back (it finished long ago) when its slot is reused, and all pending pairs when the statistics are requested.
void Engine::timer_harvest(int idx) {
  Impl& s = *d_;
  Impl::Timer& t = s.timers[idx];
  if (t.kind < 0) return;
  CUDA_CHECK(cudaEventSynchronize(t.b));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, t.a, t.b));
  s.timed_ms[t.kind] += ms;
  s.timed_n[t.kind] += 1;
  t.kind = -1;
}
int Engine::timer_begin(int kind) {
  Impl& s = *d_;
  if (!timing_) return -1;
  const int idx = s.timer_head;
  s.timer_head = (s.timer_head + 1) % Impl::NTIMERS;
  timer_harvest(idx);
  Impl::Timer& t = s.timers[idx];
  if (t.a == nullptr) {
    CUDA_CHECK(cudaEventCreate(&t.a));
    CUDA_CHECK(cudaEventCreate(&t.b));
  }
  t.kind = kind;
  CUDA_CHECK(cudaEventRecord(t.a, s.stream));
  return idx;
}
void Engine::timer_end(int idx) {
  if (idx >= 0) CUDA_CHECK(cudaEventRecord(d_->timers[idx].b, d_->stream));
}
EngineStats Engine::stats() {
  for (int k = 0; k < Impl::NTIMERS; ++k) timer_harvest(k);
  stats_.force_ms = d_->timed_ms[0];
  stats_.build_ms = d_->timed_ms[1];
  return stats_;
}
void Engine::kernel_times(double* ms8, long long* n8) {
  for (int k = 0; k < Impl::NTIMERS; ++k) timer_harvest(k);
  for (int k = 0; k < Impl::NKINDS; ++k) {
    ms8[k] = d_->timed_ms[k];
    n8[k] = d_->timed_n[k];
  }
}

bool Engine::compute_forces(int layer0, bool compute, double Lbox, ForceScalars& out, double& neighbor_seconds) {
  flush_kick();
  Impl& s = *d_;
  const int N = s.N;
  const LayerTable& lt = s.layers[layer0];
  auto t_start = std::chrono::steady_clock::now();

  const double tp0 = wall_now();
  s.ke_valid = false;   // new forces: the sums predicted by the last kick no longer describe the next one
  if (s.foreign_R) s.check_cached = false;
  // ---- K0: rebuild trigger (reference handle_neighbor_lists) -----------------------------------
  // Single GPU, coordinates moved by k_displace: the criterion of the current coordinates is (or will be) in
  // scalars[8] on the device and the host does not wait for it -- the pair kernel is launched SPECULATIVELY, checks the
  // criterion itself and reports "rebuild needed" instead of forces when it fires (about one step in seven at LJ-1M).
  bool rebuild = false, speculative = false;
  const bool dist = s.world > 1 && s.owned_valid;
  if (dist) {
    // several GPUs: halo positions + every rank's criterion state travel in one NCCL group, the decision is taken on the
    // device (scalars[CRIT_DIST..]) and the pair kernel is launched speculatively against it, as on one GPU
    const int tmr_x = timer_begin(TIMER_EXCHANGE);
    exchange_step(s, true);
    timer_end(tmr_x);
    speculative = true;
    stats_.launches += 2;
  } else if (s.check_cached && s.list_valid && s.world == 1) {
    speculative = true;
  } else if (s.list_valid && s.world == 1) {
    // the coordinates came from an upload (or may have been written through a raw pointer): the criterion is evaluated on the
    // device and stays there -- the pair kernel is launched speculatively against it, as after a drift; no host wait here
    k_displacement_check<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, s.R0.p, N, s.chkPartial.p, s.tickets.p + 2, s.scalars.p + 8,
                                                           nullptr, 0ull);
    stats_.launches += 1;
    s.check_cached = true;   // scalars[8] stays valid until the coordinates or R0 change
    speculative = true;
  } else {
    const unsigned long long seq = s.next_seq(SLOT_CHECK);
    k_displacement_check<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, s.R0.p, N, s.chkPartial.p, s.tickets.p + 2,
                                                           s.scalars.p + 8, s.slots + SLOT_CHECK, seq);
    stats_.launches += 1;
    s.wait_slot(SLOT_CHECK);
    s.check_cached = true;   // scalars[8] stays valid until the coordinates or R0 change
    rebuild = s.slots[SLOT_CHECK].v[0] > s.skinSq;
  }
  const double tp1 = wall_now();
  s.t_check += tp1 - tp0;
  if (rebuild) rebuild_list(Lbox);
  neighbor_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  const double tp2 = wall_now();
  if (rebuild) { s.t_rebuild += tp2 - tp1; s.n_rebuild += 1; }

  // ---- pair loop -------------------------------------------------------------------------------
  // several GPUs: what to do with the decision that came back with the scalars (code 1 = rebuild, 2 = phase 2 needed)
  auto settle_dist = [&](int code) -> bool {
    const double maximum = s.slots[SLOT_FORCE].v[0];
    const long long istar = (long long)s.slots[SLOT_FORCE].v[6];
    const bool rb = code == 1 || rebuild_needed_phase2(s, maximum, istar);
    if (s.env_debug) std::fprintf(stderr, "[emdee r%d] compute_forces: code=%d rebuild=%d\n", s.rank, code, (int)rb);
    if (rb) {
      const double tr0 = wall_now();
      auto t_rb = std::chrono::steady_clock::now();
      rebuild_list(Lbox);
      neighbor_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_rb).count();
      s.t_rebuild += wall_now() - tr0;
      s.n_rebuild += 1;
    }
    return rb;
  };
  if (!lt.pairs_exist) {
    if (speculative && dist) {   // no pair kernel to carry the speculation: ask for the decision now
      publish_decision(s);
      const int code = (int)s.slots[SLOT_FORCE].v[5];
      if (code != 0) rebuild = settle_dist(code);
    } else if (speculative) {
      CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 8, s.scalars.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      if (s.h_scalars[8] > s.skinSq) {
        rebuild = true;
        rebuild_list(Lbox);
      }
    }
    CUDA_CHECK(cudaMemsetAsync(s.F.p + (size_t)layer0 * 3 * N, 0, 3 * (size_t)N * sizeof(double), s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    out = ForceScalars();
    s.kick_planned = false;   // no pair kernel to follow: EmDee_boost issues the kick itself
    return rebuild;
  }
  launch_pair_kernel(layer0, compute, Lbox, speculative);
  if (s.world > 1) {
    finish_pair_dist(s, dist);
    const int code = dist ? (int)s.slots[SLOT_FORCE].v[5] : 0;
    if (code != 0) {   // the speculative launch did nothing: settle the decision, then the real launch
      if (s.last_force_timer >= 0) s.timers[s.last_force_timer].kind = -1;   // the aborted launch is not a force evaluation
      stats_.force_launches -= 1;
      rebuild = settle_dist(code);
      launch_pair_kernel(layer0, compute, Lbox, false);
      finish_pair_dist(s, false);
    }
  } else {
    s.wait_slot(SLOT_FORCE);
    if (speculative && s.slots[SLOT_FORCE].v[SLOT_STATUS] != 0.0) {   // the criterion fired: rebuild, then the real launch
      const double tr0 = wall_now();
      auto t_rb = std::chrono::steady_clock::now();
      rebuild = true;
      if (s.last_force_timer >= 0) s.timers[s.last_force_timer].kind = -1;   // the aborted launch is not a force evaluation
      stats_.force_launches -= 1;
      rebuild_list(Lbox);
      neighbor_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_rb).count();
      s.t_rebuild += wall_now() - tr0;
      s.n_rebuild += 1;
      launch_pair_kernel(layer0, compute, Lbox, false);
      s.wait_slot(SLOT_FORCE);
    }
  }
  collect_planned_kick();
  s.t_force += wall_now() - tp2;
  s.n_force += 1;
  const double* v = s.slots[SLOT_FORCE].v;
  out.Epair = v[0];
  out.Ecoul = v[1];
  out.Wpair = v[2];
  out.Wcoul = v[3];
  out.Wbody = v[4];
  return rebuild;
}

// List rebuild (reference distribute_atoms + build_neighbor_lists); ends with R0 = R and a zero criterion on the device.
void Engine::rebuild_list(double Lbox) {
  Impl& s = *d_;
  const int N = s.N;
  const double invL2 = 1.0 / (Lbox * Lbox);
  {

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
      if (s.owned_valid && !s.all_known && s.env_no_migrate) {
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
    const int tmr_b = timer_begin(TIMER_BINNING);
    if (s.compact)
      k_bin<<<std::max(1, nblocks(s.nCand)), TPB, 0, s.stream>>>(s.R.p, s.nCand, Lbox, s.grid, s.Rs.p, s.atomCell.p, s.atomFloor.p,
                                                                 s.owned.p, nullptr, s.cellCount.p, s.candList.p);
    else
      k_bin<<<nblocks(N), TPB, 0, s.stream>>>(s.R.p, N, Lbox, s.grid, s.Rs.p, s.atomCell.p, s.atomFloor.p, s.owned.p,
                                              s.world > 1 ? s.known.p : nullptr, s.cellCount.p, nullptr);
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
    if (s.compact)
      k_fill<<<std::max(1, nblocks(s.nCand)), TPB, 0, s.stream>>>(s.nCand, s.grid, s.atomCell.p, s.cellStart.p, s.cellFill.p, s.slotAtom.p,
                                                                  s.slotImg.p, s.slotCell.p, s.candList.p);
    else
      k_fill<<<nblocks(N), TPB, 0, s.stream>>>(N, s.grid, s.atomCell.p, s.cellStart.p, s.cellFill.p, s.slotAtom.p,
                                               s.slotImg.p, s.slotCell.p, nullptr);
    PlaceArgs pa;
    pa.Next = Next; pa.slotAtom = s.slotAtom.p; pa.slotImg = s.slotImg.p; pa.slotCell = s.slotCell.p;
    pa.cellStart = s.cellStart.p; pa.atomFloor = s.atomFloor.p; pa.Rs = s.Rs.p; pa.atomType = s.type.p;
    pa.atomBody = s.body.p; pa.owned = s.owned.p; pa.sMeta = s.sMeta.p; pa.sCell = s.sCell.p; pa.sGhost = s.sGhost.p; pa.sType = s.sType.p;
    pa.sBody = s.sBody.p; pa.sRs = s.sRs.p; pa.sPosF = s.sPosF.p; pa.nbrCount = s.nbrCount.p;
    k_place<<<nblocks(Next), TPB, 0, s.stream>>>(pa);
    timer_end(tmr_b);
    stats_.launches += 4;
    // capacity guess from the mean density; grown on overflow
    if (s.cap == 0) {
      double nbar = (double)N / (Lbox * Lbox * Lbox) * (4.0 / 3.0) * 3.14159265358979323846 * s.xRc * s.xRcSq;
      s.cap = std::max(16, (int)(1.35 * nbar) + 16);
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
      s.nbr.ensure((size_t)ntiles * s.cap * TILE, 1.1);
      b.nbr = s.nbr.p;
      // (a warp-per-tile variant with __ballot_sync compaction into shared-memory rows and a transposed
      // write-out was measured 6x slower at LJ-1M -- serial per-atom dependency chains at ~12 warps/SM --
      // and removed)
      const int tmr = timer_begin(1);
      k_build_list<<<nblocks(Next), TPB, 0, s.stream>>>(b);
      timer_end(tmr);
      stats_.launches += 1;
      stats_.build_launches += 1;
      int hflags[4];
      CUDA_CHECK(cudaMemcpyAsync(hflags, s.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CUDA_CHECK(cudaStreamSynchronize(s.stream));
      CUDA_CHECK(cudaGetLastError());
      if (!hflags[1]) break;
      s.cap = (int)(hflags[0] * 1.15) + 8;   // overflow: regrow to the observed maximum and redo
    }
    if (s.world > 1) {
      build_halo_lists(s);
      if (s.nbodies != 0) {   // free atoms this rank integrates (body state is replicated, see boost_all / move_all)
        s.ownedFree.ensure(N);
        k_mask_and<<<nblocks(N), TPB, 0, s.stream>>>(N, s.owned.p, s.freeMask.p, s.ownedFree.p);
        stats_.launches += 1;
      }
      s.owned_valid = true;
      s.all_known = false;
      s.halo_fresh = true;   // migrate() just made every needed position current
    }
    if (s.compact) {   // the criterion of a rank only looks at the atoms it owns
      if (s.nOwn > 0) k_copy3_listed<<<nblocks(s.nOwn), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, s.R.p, s.R0.p);
    } else {
      CUDA_CHECK(cudaMemcpyAsync(s.R0.p, s.R.p, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
    }
    s.compact = false;
    CUDA_CHECK(cudaMemsetAsync(s.scalars.p + 8, 0, sizeof(double), s.stream));   // R0 = R: zero displacement
    s.check_cached = true;
    s.mi_fresh = false;   // R0 changed: the distributed criterion state is re-evaluated on the next force call
    s.list_valid = true;
    stats_.cells_per_dim = M;
  }
}

// Position refresh + pair kernel of one layer. `speculative`: the kernel reads the rebuild criterion itself (see above).
void Engine::launch_pair_kernel(int layer0, bool compute, double Lbox, bool speculative) {
  Impl& s = *d_;
  const int N = s.N;
  const LayerTable& lt = s.layers[layer0];
  const double invL2 = 1.0 / (Lbox * Lbox);
  double* Fl = s.F.p + (size_t)layer0 * 3 * N;
  const int Next = s.Next;
  const int tmr_r = timer_begin(TIMER_REFRESH);
  k_refresh_positions<<<nblocks(Next), TPB, 0, s.stream>>>(Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p,
                                                           !speculative ? nullptr : (s.world > 1 ? s.scalars.p + CRIT_DIST : s.scalars.p + 8),
                                                           s.skinSq);
  timer_end(tmr_r);
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
  a.crit = !speculative ? nullptr : (s.world > 1 ? s.scalars.p + CRIT_DIST : s.scalars.p + 8);
  a.skinSq = s.skinSq;
  a.hs = (s.world > 1) ? nullptr : s.slots + SLOT_FORCE;   // several GPUs: the scalars are all-reduced first
  a.seq = s.next_seq(SLOT_FORCE);

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

  const int tmr = timer_begin(0);
  using namespace nb;
  const bool lj_plain = uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_NONE && !a.q4_quirk;
  const bool lj_sf = uniform && pk == K_PAIR_LJ_CUT && pm == M_SHIFTED_FORCE && ck == K_COUL_NONE && !a.q4_quirk;
  const bool lj_coul_sf = uniform && pk == K_PAIR_LJ_CUT && pm == M_NONE && ck == K_COUL_SF && cm == M_NONE;
  // several types, every pair model pair_lj_cut (one modifier: none or shifted_force) or pair_none, Coulomb kind cut / sf /
  // damped family: k_pair_forces_typed (branch-free pair term), launched by try_typed_path. (Single-type LJ + coul_sf through
  // the same kernel, without type lookups, measured 0.443 ms against 0.385 ms for the generic kernel below at 1M atoms --
  // 128 registers against 85 -- and stays with the generic kernel; profiles/r2g_coul_sf_typed_vs_generic.txt.)
  if (s.nt > 1 && s.nt <= MAX_SMEM_TYPES && !a.q4_quirk && cm == M_NONE && try_typed_path(s, layer0, lt, ck, a, compute)) {
  } else if (s.nt == 1 && lj_plain)
    launch_lj_plain(s, a, compute);
  else if (s.nt == 1 && lj_sf)
    launch_force<K_PAIR_LJ_CUT, M_SHIFTED_FORCE, K_COUL_NONE, M_NONE, true, true, 4, 256, 3>(a, s.partial, compute, 0, s.stream);
  else if (s.nt == 1 && lj_coul_sf)   // unroll 3 / 256 threads / 4 blocks: fastest of six launch shapes (profiles/r2g_coul_sf_typed_vs_generic.txt)
    launch_force<K_PAIR_LJ_CUT, M_NONE, K_COUL_SF, M_NONE, true, true, 3, 256, 4>(a, s.partial, compute, 0, s.stream);
  else
    launch_force<K_DYNAMIC, M_DYNAMIC, K_DYNAMIC, M_DYNAMIC, false, true, 4, 256, 2>(a, s.partial, compute, smem_dyn, s.stream);
  timer_end(tmr);
  s.last_force_timer = tmr;
  stats_.launches += 2;
  stats_.force_launches += 1;
  if (s.kick_planned && s.kick_layer == layer0) launch_planned_kick(speculative);
}

// A kick that EmDee_boost is about to issue right after the force evaluation it triggers: compute_forces launches it
// itself behind the pair kernel (one host wait, and on several GPUs one reduction, for both). Only where the kick can
// follow the pair kernel directly: free atoms, no bonded / reciprocal-space terms added afterwards (abi.cpp decides).
void Engine::plan_kick(int layer0, double CP, double CF, bool want_kinetic) {
  Impl& s = *d_;
  if (s.world > 1 && !(s.owned_valid && s.peer_ok)) return;   // NCCL-per-step mode: the kick stays a call of its own
  if (s.exposed || s.foreign_R) return;
  s.kick_planned = true;
  s.kick_done = false;
  s.kick_layer = layer0;
  s.kick_CP = CP;
  s.kick_CF = CF;
  s.kick_want_ke = want_kinetic;
}

// launches the planned kick behind the pair kernel of the same layer (`speculative`: it checks the criterion like the pair kernel)
void Engine::launch_planned_kick(bool speculative) {
  Impl& s = *d_;
  const double* Fl = s.F.p + (size_t)s.kick_layer * 3 * s.N;
  const int ke = s.kick_want_ke ? 1 : 0;
  const int tmr = timer_begin(TIMER_BOOST);
  if (s.world > 1) {
    k_boost_owned<<<std::max(1, nblocks(s.nOwn, TPB * BOOST_EPT)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, s.kick_CP, s.kick_CF, s.P.p, Fl,
                                                                      s.invMass.p, ke, s.partial.p, s.tickets.p + 1, s.scalars.p + 5,
                                                                      speculative ? s.scalars.p + CRIT_DIST : nullptr, s.skinSq);
  } else {
    const unsigned long long seq = ke ? s.next_seq(SLOT_KINETIC) : 0ull;
    k_boost<<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, s.kick_CP, s.kick_CF, s.P.p, Fl, s.invMass.p, nullptr, ke,
                                                                 s.partial.p, s.tickets.p + 1, s.scalars.p + 10,
                                                                 ke ? s.slots + SLOT_KINETIC : nullptr, seq,
                                                                 speculative ? s.scalars.p + 8 : nullptr, s.skinSq);
  }
  timer_end(tmr);
  stats_.launches += 1;
}

// the planned kick has run (its sums, if any, are in the kinetic slot): remember them for Engine::boost
void Engine::collect_planned_kick() {
  Impl& s = *d_;
  if (!s.kick_planned) return;
  if (s.kick_want_ke) {
    if (s.world == 1) s.wait_slot(SLOT_KINETIC);   // several GPUs: finish_pair_dist has waited for both slots
    for (int x = 0; x < 3; ++x) {
      s.kick_ke[x] = s.slots[SLOT_KINETIC].v[x];
      s.ke_next[x] = s.slots[SLOT_KINETIC].v[3 + x];
    }
    s.ke_valid = true;
    s.ke_CP = s.kick_CP;
    s.ke_CF = s.kick_CF;
    s.ke_layer = s.kick_layer;
  }
  s.kick_planned = false;
  s.kick_done = true;
}

// executes a deferred kick now (its kinetic sums were already answered): every entry point that reads or writes momenta
// or forces, other than the drift that absorbs it, starts with this
void Engine::flush_kick() {
  Impl& s = *d_;
  if (!s.defer_on) return;
  s.defer_on = false;
  const double* Fl = s.F.p + (size_t)s.defer_layer * 3 * s.N;
  const int tmr = timer_begin(TIMER_BOOST);
  if (s.world > 1 && s.owned_valid)
    k_boost_owned<<<std::max(1, nblocks(s.nOwn, TPB * BOOST_EPT)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, s.defer_CP, s.defer_CF, s.P.p, Fl,
                                                                      s.invMass.p, 0, s.partial.p, s.tickets.p + 1, s.scalars.p + 10,
                                                                      nullptr, 0.0);
  else
    k_boost<<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, s.defer_CP, s.defer_CF, s.P.p, Fl, s.invMass.p, nullptr, 0,
                                                                 s.partial.p, s.tickets.p + 1, s.scalars.p + 10, nullptr, 0ull,
                                                                 nullptr, 0.0);
  timer_end(tmr);
  stats_.launches += 1;
}

void Engine::boost(int layer0, double CP, double CF, bool want_kinetic, KineticScalars& ke) {
  Impl& s = *d_;
  const double tp0 = wall_now();
  if (s.kick_done) {   // compute_forces has already executed this very kick (plan_kick)
    s.kick_done = false;
    if (s.kick_layer == layer0 && s.kick_CP == CP && s.kick_CF == CF && s.kick_want_ke == want_kinetic) {
      for (int x = 0; x < 3; ++x) ke.twoKE[x] = s.kick_ke[x];
      s.t_boost += wall_now() - tp0;
      s.n_boost += 1;
      return;
    }
    fatal("momentum update", "internal: a planned kick does not match the kick requested");
  }
  s.kick_planned = false;
  flush_kick();
  const double* Fl = s.F.p + (size_t)layer0 * 3 * s.N;
  const bool dist = s.world > 1 && s.owned_valid;
  // the previous kick predicted this one's sums (same coefficients, same forces, momenta untouched in between)
  const bool predicted = want_kinetic && s.ke_valid && s.ke_layer == layer0 && s.ke_CP == CP && s.ke_CF == CF && !s.exposed;
  const int want = (want_kinetic && !predicted) ? 1 : 0;
  s.ke_valid = false;
  if (!want && !s.exposed && !s.foreign_R && !s.env_no_defer && s.world == 1) {   // (several GPUs: measured 1 % slower, see k_displace_owned)
    // nothing to reduce: the drift that follows applies this kick (k_displace<true>); flush_kick otherwise
    s.defer_on = true;
    s.defer_layer = layer0;
    s.defer_CP = CP;
    s.defer_CF = CF;
    if (predicted)
      for (int x = 0; x < 3; ++x) ke.twoKE[x] = s.ke_next[x];
    s.t_boost += wall_now() - tp0;
    s.n_boost += 1;
    return;
  }
  const int tmr = timer_begin(TIMER_BOOST);
  if (dist) {
    // several GPUs: the kick runs over the compact list of owned atoms (work ~ atoms of this rank, not N)
    k_boost_owned<<<std::max(1, nblocks(s.nOwn, TPB * BOOST_EPT)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, CP, CF, s.P.p, Fl, s.invMass.p, want,
                                                                      s.partial.p, s.tickets.p + 1, s.scalars.p + 10, nullptr, 0.0);
    timer_end(tmr);
    stats_.launches += 1;
    if (want) {
      const unsigned long long seq = s.next_seq(SLOT_KINETIC);
      if (s.peer_ok) {
        k_reduce_small<<<1, 32, 0, s.stream>>>(0, 6, s.scalars.p + 10, s.peers, s.box, s.world, s.rank, ++s.rseq, s.scalars.p + 10,
                                               nullptr, nullptr, 0ull, s.slots + SLOT_KINETIC, seq);
      } else {
        NCCL_CHECK(nccl().AllReduce(s.scalars.p + 10, s.scalars.p + 10, 6, ncclDouble, ncclSum, s.comm, s.stream));
        k_publish_n<<<1, 32, 0, s.stream>>>(6, s.scalars.p + 10, s.slots + SLOT_KINETIC, seq);
      }
      s.wait_slot(SLOT_KINETIC);
    }
  } else {
    const int grid = nblocks((s.N + APT - 1) / APT);
    HostSlot* hs = want ? s.slots + SLOT_KINETIC : nullptr;   // the last block writes the sums to the host slot
    const unsigned long long seq = want ? s.next_seq(SLOT_KINETIC) : 0ull;
    k_boost<<<grid, TPB, 0, s.stream>>>(s.N, CP, CF, s.P.p, Fl, s.invMass.p, nullptr, want, s.partial.p, s.tickets.p + 1,
                                        s.scalars.p + 10, hs, seq, nullptr, 0.0);
    timer_end(tmr);
    stats_.launches += 1;
    if (want) s.wait_slot(SLOT_KINETIC);
    else if (s.exposed) CUDA_CHECK(cudaStreamSynchronize(s.stream));
  }
  if (want) {
    for (int x = 0; x < 3; ++x) {
      ke.twoKE[x] = s.slots[SLOT_KINETIC].v[x];
      s.ke_next[x] = s.slots[SLOT_KINETIC].v[3 + x];
    }
    s.ke_valid = !s.exposed;
    s.ke_CP = CP;
    s.ke_CF = CF;
    s.ke_layer = layer0;
  } else if (predicted) {
    for (int x = 0; x < 3; ++x) ke.twoKE[x] = s.ke_next[x];
  }
  s.t_boost += wall_now() - tp0;
  s.n_boost += 1;
}

void Engine::displace(double CR, double CP) {
  Impl& s = *d_;
  const double tp0 = wall_now();
  if (s.world > 1) flush_kick();   // (a kick is only ever deferred on one GPU)
  const bool kick = s.defer_on;    // the deferred kick rides in the drift kernel
  const double* Fk = kick ? s.F.p + (size_t)s.defer_layer * 3 * s.N : nullptr;
  s.defer_on = false;
  if (s.world > 1 && s.owned_valid) {
    // owned atoms only (compact list); phase 1 of the rebuild criterion on the new coordinates lands in miResult[world]
    const int tmr = timer_begin(TIMER_DISPLACE);
    k_displace_owned<<<std::max(1, nblocks(s.nOwn)), TPB, 0, s.stream>>>(s.nOwn, s.ownedList.p, CR, CP, s.R.p, s.P.p, s.invMass.p,
                                                                         s.R0.p, s.miPartial.p, s.tickets.p + 2,
                                                                         s.miResult.p + s.world);
    timer_end(tmr);
    stats_.launches += 1;
    s.mi_fresh = true;
    s.halo_fresh = false;
    s.check_cached = false;
    s.all_known = false;   // from now on only owned + halo positions are current on this rank
  } else {
    // the rebuild criterion of the new coordinates lands in scalars[8] on the device; compute_forces launches the pair
    // kernel speculatively against it instead of waiting for it here
    const int tmr = timer_begin(TIMER_DISPLACE);
    if (kick)
      k_displace<true><<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p, nullptr, s.R0.p,
                                                                             s.chkPartial.p, s.tickets.p + 2, s.scalars.p + 8,
                                                                             s.defer_CP, s.defer_CF, Fk);
    else
      k_displace<false><<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p, nullptr, s.R0.p,
                                                                              s.chkPartial.p, s.tickets.p + 2, s.scalars.p + 8,
                                                                              0.0, 0.0, nullptr);
    timer_end(tmr);
    stats_.launches += 1;
    s.check_cached = true;
  }
  if (s.exposed) CUDA_CHECK(cudaStreamSynchronize(s.stream));
  s.t_displace += wall_now() - tp0;
  s.n_displace += 1;
}

// ---- rigid bodies ----------------------------------------------------------------------------------------------
namespace {
enum { B_MASS = 0, B_MOI = 1, B_RCM = 4, B_PCM = 7, B_Q = 10, B_PI = 14, B_OMEGA = 18, B_F = 21, B_TAU = 24, B_WIDTH = 27 };

BodyView body_view(Engine::Impl& s) {
  BodyView v;
  const size_t nb = (size_t)s.nbodies;
  v.nb = s.nbodies;
  v.first = s.bFirst.p; v.atom = s.bAtom.p; v.mItem = s.bMItem.p; v.d = s.bD.p;
  double* base = s.bState.p;
  v.mass = base + B_MASS * nb; v.MoI = base + B_MOI * nb; v.rcm = base + B_RCM * nb; v.pcm = base + B_PCM * nb;
  v.q = base + B_Q * nb; v.pi = base + B_PI * nb; v.omega = base + B_OMEGA * nb; v.Fb = base + B_F * nb;
  v.tau = base + B_TAU * nb;
  return v;
}

}  // namespace

void Engine::set_bodies(const std::vector<int>& first, const std::vector<int>& atoms, const std::vector<double>& memberMass) {
  Impl& s = *d_;
  const size_t nb = (size_t)s.nbodies;
  if (nb == 0) return;
  s.nitems = (int)atoms.size();
  s.bFirst.ensure(nb + 1);
  s.bAtom.ensure(atoms.size());
  s.bMItem.ensure(atoms.size());
  s.bD.ensure(3 * atoms.size());
  s.bState.ensure(B_WIDTH * nb);
  s.bPartial.ensure((size_t)std::max(nblocks(s.N), nblocks((long long)nb)) * 6);
  s.bScalars.ensure(16);
  s.freeMask.ensure(s.N);
  CUDA_CHECK(cudaMallocHost(&s.h_bscalars, 16 * sizeof(double)));
  CUDA_CHECK(cudaMemcpy(s.bFirst.p, first.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.bAtom.p, atoms.data(), atoms.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.bMItem.p, memberMass.data(), atoms.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemset(s.bState.p, 0, B_WIDTH * nb * sizeof(double)));
  CUDA_CHECK(cudaMemset(s.bD.p, 0, 3 * atoms.size() * sizeof(double)));
  std::vector<unsigned char> mask(s.N, 1);
  for (int a : atoms) mask[a] = 0;
  CUDA_CHECK(cudaMemcpy(s.freeMask.p, mask.data(), s.N, cudaMemcpyHostToDevice));
}

void Engine::update_body_frames(double Lbox) {
  flush_kick();
  Impl& s = *d_;
  if (s.nbodies == 0) return;
  if (!s.has_R) fatal("rigid-body update", "coordinates have not been uploaded");
  if (!s.has_delta) {
    s.delta.ensure(3 * (size_t)s.N);
    CUDA_CHECK(cudaMemsetAsync(s.delta.p, 0, 3 * (size_t)s.N * sizeof(double), s.stream));   // free atoms keep a zero offset
    s.has_delta = true;
  }
  k_body_frame<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), s.R.p, s.delta.p, Lbox);
  stats_.launches += 1;
  s.frames_valid = true;
  s.check_cached = false;   // member coordinates may have been shifted by whole box lengths
  s.mi_fresh = false;
}

void Engine::boost_all(int layer0, double CP, double CF, bool translate, bool rotate, bool want_kinetic, KineticAll& ke) {
  flush_kick();
  Impl& s = *d_;
  const double* Fl = s.F.p + (size_t)layer0 * 3 * s.N;
  const bool bodies = s.nbodies != 0;
  const bool dist = s.world > 1 && s.owned_valid;   // several GPUs: forces exist only for the atoms this rank owns
  const bool have_free = translate && s.nitems < s.N;
  s.ke_valid = false;
  if (have_free) {
    const unsigned char* mask = !bodies ? (dist ? s.owned.p : nullptr) : (dist ? s.ownedFree.p : s.freeMask.p);
    const int grid = nblocks((s.N + APT - 1) / APT);
    k_boost<<<grid, TPB, 0, s.stream>>>(s.N, CP, CF, s.P.p, Fl, s.invMass.p, mask, want_kinetic ? 1 : 0,
                                        s.partial.p, s.tickets.p + 1, s.scalars.p + 10, nullptr, 0ull, nullptr, 0.0);
    stats_.launches += 1;
    if (dist && want_kinetic)
      NCCL_CHECK(nccl().AllReduce(s.scalars.p + 10, s.scalars.p + 10, 3, ncclDouble, ncclSum, s.comm, s.stream));
  }
  if (bodies) {
    const BodyView v = body_view(s);
    const int grid = nblocks(s.nbodies);
    if (dist) {
      // the body state is replicated: sum the owned members' forces and torques, all-reduce them (F and tau are
      // contiguous, 6 doubles per body), then every rank applies the same kick to every body
      k_body_boost<<<grid, TPB, 0, s.stream>>>(v, Fl, s.delta.p, s.owned.p, 1, CP, CF, 0, 0, 0, s.bPartial.p, s.tickets.p + 3,
                                               s.bScalars.p);
      NCCL_CHECK(nccl().AllReduce(v.Fb, v.Fb, 6 * (size_t)s.nbodies, ncclDouble, ncclSum, s.comm, s.stream));
      k_body_boost<<<grid, TPB, 0, s.stream>>>(v, Fl, s.delta.p, nullptr, 2, CP, CF, translate ? 1 : 0, rotate ? 1 : 0,
                                               want_kinetic ? 1 : 0, s.bPartial.p, s.tickets.p + 3, s.bScalars.p);
      stats_.launches += 2;
    } else {
      k_body_boost<<<grid, TPB, 0, s.stream>>>(v, Fl, s.delta.p, nullptr, 0, CP, CF, translate ? 1 : 0, rotate ? 1 : 0,
                                               want_kinetic ? 1 : 0, s.bPartial.p, s.tickets.p + 3, s.bScalars.p);
      stats_.launches += 1;
    }
  }
  if (!want_kinetic) {
    if (s.exposed) CUDA_CHECK(cudaStreamSynchronize(s.stream));
    return;
  }
  double free3[3] = {0, 0, 0}, body6[6] = {0, 0, 0, 0, 0, 0};
  if (have_free)
    CUDA_CHECK(cudaMemcpyAsync(s.h_scalars + 10, s.scalars.p + 10, 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  if (bodies) CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars, s.bScalars.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  if (have_free)
    for (int x = 0; x < 3; ++x) free3[x] = s.h_scalars[10 + x];
  if (bodies)
    for (int x = 0; x < 6; ++x) body6[x] = s.h_bscalars[x];
  for (int x = 0; x < 3; ++x) {
    ke.twoKEt[x] = free3[x] + body6[x];
    ke.twoKEr[x] = body6[3 + x];
  }
}

void Engine::move_all(double CR, double CP, double dt, bool translate, bool rotate, int mode) {
  flush_kick();
  Impl& s = *d_;
  const bool bodies = s.nbodies != 0;
  const bool dist = s.world > 1 && s.owned_valid;
  if (translate && s.nitems < s.N) {
    // the fused rebuild criterion is not used here: body members move in k_body_move below
    const unsigned char* mask = !bodies ? (dist ? s.owned.p : nullptr) : (dist ? s.ownedFree.p : s.freeMask.p);
    k_displace<false><<<nblocks((s.N + APT - 1) / APT), TPB, 0, s.stream>>>(s.N, CR, CP, s.R.p, s.P.p, s.invMass.p, mask, s.R0.p, nullptr,
                                                                            s.tickets.p + 2, s.scalars.p + 8, 0.0, 0.0, nullptr);
    stats_.launches += 1;
  }
  if (bodies) {   // several GPUs: every rank moves every body (replicated state), so member coordinates stay current everywhere
    k_body_move<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), s.R.p, s.delta.p, CR, CP, dt, translate ? 1 : 0,
                                                          rotate ? 1 : 0, mode);
    stats_.launches += 1;
  }
  s.check_cached = false;   // compute_forces evaluates the rebuild criterion on the new coordinates
  s.mi_fresh = false;
  if (dist) {
    s.halo_fresh = false;
    s.all_known = false;    // free atoms: only owned + halo positions are current on this rank
  }
  if (s.exposed) CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

// ---- bonded terms ---------------------------------------------------------------------------------------------------
void Engine::set_bonded(const std::vector<BondedTerm>& terms) {
  Impl& s = *d_;
  s.nterms = (int)terms.size();
  if (s.nterms == 0) return;
  std::vector<int> first(s.N + 1, 0), ref;
  auto members = [](const BondedTerm& t) { return t.kind <= T_BOND_HARMONIC ? 2 : 3; };
  for (const BondedTerm& t : terms) {
    const int m[3] = {t.a0, t.a1, t.a2};
    for (int r = 0; r < members(t); ++r) first[m[r] + 1] += 1;
  }
  for (int a = 0; a < s.N; ++a) first[a + 1] += first[a];
  ref.resize(first[s.N]);
  std::vector<int> fill(first.begin(), first.end() - 1);
  for (size_t q = 0; q < terms.size(); ++q) {   // terms in the order they were added: fixed summation order per atom
    const int m[3] = {terms[q].a0, terms[q].a1, terms[q].a2};
    for (int r = 0; r < members(terms[q]); ++r) ref[fill[m[r]]++] = (int)(q << 2) | r;
  }
  s.terms.ensure(terms.size());
  s.termFirst.ensure(s.N + 1);
  s.termRef.ensure(ref.size() + 1);
  CUDA_CHECK(cudaMemcpy(s.terms.p, terms.data(), terms.size() * sizeof(BondedTerm), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.termFirst.p, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.termRef.p, ref.data(), ref.size() * sizeof(int), cudaMemcpyHostToDevice));
  s.bPartial.ensure((size_t)nblocks(s.N) * 6);
  s.bScalars.ensure(16);
  if (s.h_bscalars == nullptr) CUDA_CHECK(cudaMallocHost(&s.h_bscalars, 16 * sizeof(double)));
}

void Engine::add_bonded(int layer0, double Lbox, bool bonded, bool kspace, BondedScalars& out) {
  flush_kick();
  Impl& s = *d_;
  out = BondedScalars();
  if (s.nterms == 0) return;
  const bool dist = s.world > 1 && s.owned_valid;   // partners of an owned atom lie inside the halo (bonds are shorter than the cutoff)
  const double* delta = (s.has_delta && s.nbodies != 0) ? s.delta.p : nullptr;
  BondedEwald ks;
  ks.on = (kspace && s.ewald_on) ? 1 : 0;
  ks.alpha = s.ew_alpha; ks.beta = s.ew_beta; ks.q = s.q.p; ks.type = s.type.p; ks.tab = s.tabs[layer0].p; ks.nt = s.nt;
  if (!bonded && !ks.on) return;
  if (dist) CUDA_CHECK(cudaMemsetAsync(s.flags.p + 4, 0, sizeof(int), s.stream));
  k_bonded<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, s.termFirst.p, s.termRef.p, s.terms.p, s.R.p, Lbox, bonded ? 1 : 0, ks,
                                               dist ? s.owned.p : nullptr, s.F.p + (size_t)layer0 * 3 * s.N, delta, s.bPartial.p,
                                               s.tickets.p + 3, s.bScalars.p, s.xRcSq, s.flags.p + 4);
  stats_.launches += 1;
  int toolong = 0;
  if (dist) {
    NCCL_CHECK(nccl().AllReduce(s.bScalars.p, s.bScalars.p, 6, ncclDouble, ncclSum, s.comm, s.stream));
    NCCL_CHECK(nccl().AllReduce(s.flags.p + 4, s.flags.p + 4, 1, ncclInt, ncclMax, s.comm, s.stream));   // every rank stops, or none
    CUDA_CHECK(cudaMemcpyAsync(&toolong, s.flags.p + 4, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  }
  CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars, s.bScalars.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  if (toolong)
    fatal("bonded force computation", "a bond or angle arm is longer than Rc + skin: on several GPUs its other end lies outside the halo of the atom's rank");
  out.Ebond = s.h_bscalars[0];
  out.Wbond = s.h_bscalars[1];
  out.Eangle = s.h_bscalars[2];
  out.Wangle = s.h_bscalars[3];
  out.Wbody = s.h_bscalars[4];
  out.Ecoul = s.h_bscalars[5];
}

// ---- reciprocal-space Ewald ---------------------------------------------------------------------------------------
void Engine::set_ewald(const EwaldSetup& e) {
  Impl& s = *d_;
  if (e.ntk > EWALD_MAX_TYPES) fatal("kspace model initialization", "more than 8 distinct types of charged atoms are not supported");
  s.ewald_on = true;
  s.ew_alpha = e.alpha;
  s.ew_beta = e.beta;
  s.ew_ntk = e.ntk;
  s.ew_nvecs = (int)e.prefac.size();
  s.ewN.ensure(e.n.size() + 1);
  s.ewKType.ensure(s.N);
  s.ewPrefac.ensure(e.prefac.size() + 1);
  s.ewLambda.ensure(e.lambda.size() + 1);
  s.ewSigma.ensure(2 * (size_t)e.ntk * s.ew_nvecs + 1);
  s.ewPartial.ensure((size_t)std::max(s.ew_nvecs, nblocks(s.N)) * 2 + 2);
  CUDA_CHECK(cudaMemcpy(s.ewN.p, e.n.data(), e.n.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.ewKType.p, e.atomKType.data(), s.N * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.ewPrefac.p, e.prefac.data(), e.prefac.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(s.ewLambda.p, e.lambda.data(), e.lambda.size() * sizeof(double), cudaMemcpyHostToDevice));
  s.bScalars.ensure(16);
  if (s.h_bscalars == nullptr) CUDA_CHECK(cudaMallocHost(&s.h_bscalars, 16 * sizeof(double)));
}

void Engine::add_ewald(int layer0, double Lbox, double& Elong, double& Wbody) {
  flush_kick();
  Impl& s = *d_;
  Elong = Wbody = 0.0;
  if (!s.ewald_on || s.ew_nvecs == 0) return;
  const bool dist = s.world > 1 && s.owned_valid;
  EwaldView v;
  v.nvecs = s.ew_nvecs; v.ntk = s.ew_ntk; v.N = s.N; v.n = s.ewN.p; v.prefac = s.ewPrefac.p; v.ktype = s.ewKType.p;
  v.q = s.q.p; v.owned = dist ? s.owned.p : nullptr; v.R = s.R.p; v.sigma = s.ewSigma.p;
  const double* lambda = s.ewLambda.p + (size_t)layer0 * s.ew_ntk * s.ew_ntk;
  const double* delta = (s.has_delta && s.nbodies != 0) ? s.delta.p : nullptr;
  if (dist) {
    // every rank sums the structure factors of the atoms it owns; one all-reduce makes them global, after which sigma
    // and the energy are the same numbers on every rank and each rank finishes the forces of its own atoms
    k_ewald_structure<<<s.ew_nvecs, TPB, 0, s.stream>>>(v, Lbox, lambda, 1, s.ewPartial.p, s.tickets.p + 3, s.bScalars.p + 12);
    NCCL_CHECK(nccl().AllReduce(s.ewSigma.p, s.ewSigma.p, 2 * (size_t)s.ew_ntk * s.ew_nvecs, ncclDouble, ncclSum, s.comm, s.stream));
    k_ewald_sigma<<<nblocks(s.ew_nvecs), TPB, 0, s.stream>>>(v, lambda, s.ewPartial.p, s.tickets.p + 3, s.bScalars.p + 12);
    stats_.launches += 1;
  } else {
    k_ewald_structure<<<s.ew_nvecs, TPB, 0, s.stream>>>(v, Lbox, lambda, 0, s.ewPartial.p, s.tickets.p + 3, s.bScalars.p + 12);
  }
  k_ewald_forces<<<nblocks(s.N), TPB, 0, s.stream>>>(v, Lbox, s.F.p + (size_t)layer0 * 3 * s.N, delta, s.ewPartial.p,
                                                     s.tickets.p + 3, s.bScalars.p + 14);
  stats_.launches += 2;
  if (dist) NCCL_CHECK(nccl().AllReduce(s.bScalars.p + 14, s.bScalars.p + 14, 1, ncclDouble, ncclSum, s.comm, s.stream));
  CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars + 12, s.bScalars.p + 12, 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  Elong = s.h_bscalars[12];
  Wbody = s.h_bscalars[14];
}

// ---- EmDee_memory_address / EmDee_share_phase_space ---------------------------------------------------------------
// The reference hands out pointers into its own host arrays (src/EmDeeCode.f90:212-235). Here the arrays live in HBM,
// so the requested array is moved into CUDA managed memory once: the SAME allocation is then the kernels' array and
// the client's. Contract: the client touches it only between library calls (every call now ends with a stream
// synchronisation), and -- as with the reference -- writing coordinates through the pointer does not invalidate
// anything by itself; the next EmDee_compute_forces re-evaluates the rebuild criterion on whatever it finds.
void* Engine::expose(int what, int layer0) {
  flush_kick();
  Impl& s = *d_;
  if (s.world > 1) fatal("memory address retrieving", "not available on several GPUs (every rank holds a slab)");
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  s.exposed = true;
  switch (what) {
    case EXPOSE_R: s.R.to_managed(); s.foreign_R = true; return s.R.p;
    case EXPOSE_P: s.P.to_managed(); return s.P.p;
    case EXPOSE_F: s.F.to_managed(); return s.F.p + (size_t)layer0 * 3 * s.N;
    default: s.F.to_managed(); return s.F.p;
  }
}

// `this` gives up its own R, P and rigid-body state for `keep`'s (src/EmDeeCode.f90:259-263)
void Engine::share_phase_space(Engine& keep) {
  flush_kick();
  keep.flush_kick();
  Impl& s = *d_;
  Impl& k = *keep.d_;
  if (s.world > 1 || k.world > 1) fatal("phase space sharing", "not available on several GPUs");
  if (s.device != k.device) fatal("phase space sharing", "the two systems live on different GPUs");
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  s.R.alias(k.R);
  s.P.alias(k.P);
  if (s.nbodies != 0) {
    s.delta.alias(k.delta);
    s.bState.alias(k.bState);
    s.bD.alias(k.bD);
  }
  s.foreign_R = k.foreign_R = true;   // either system may now move the atoms
  s.check_cached = k.check_cached = false;
  s.exposed = s.exposed || k.exposed;
}

void Engine::refresh_member_momenta() {
  flush_kick();
  Impl& s = *d_;
  if (s.nbodies == 0) return;
  k_body_momenta<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), s.delta.p, s.P.p);
  stats_.launches += 1;
}

// kinetic sums of the momenta just uploaded: free atoms (k_free_kinetic) + bodies (k_body_take_momenta = assign_momenta)
void Engine::take_member_momenta(KineticAll& ke) {
  flush_kick();
  Impl& s = *d_;
  for (int x = 0; x < 3; ++x) ke.twoKEt[x] = ke.twoKEr[x] = 0.0;
  if (s.nitems < s.N) {
    s.bPartial.ensure((size_t)nblocks(s.N) * 6);
    s.bScalars.ensure(16);
    if (s.h_bscalars == nullptr) CUDA_CHECK(cudaMallocHost(&s.h_bscalars, 16 * sizeof(double)));
    k_free_kinetic<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, s.nbodies != 0 ? s.freeMask.p : nullptr, s.P.p, s.invMass.p,
                                                       s.bPartial.p, s.tickets.p + 3, s.bScalars.p + 8);
    stats_.launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars + 8, s.bScalars.p + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    for (int x = 0; x < 3; ++x) ke.twoKEt[x] = s.h_bscalars[8 + x];
  }
  if (s.nbodies == 0) return;
  k_body_take_momenta<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), s.delta.p, s.P.p, s.bPartial.p, s.tickets.p + 3,
                                                                 s.bScalars.p);
  stats_.launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars, s.bScalars.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  for (int x = 0; x < 3; ++x) {
    ke.twoKEt[x] += s.h_bscalars[x];
    ke.twoKEr[x] = s.h_bscalars[3 + x];
  }
}

namespace {
int body_item_offset(int what, int& width) {
  switch (what) {
    case Engine::BODY_QUATERNION: width = 4; return B_Q;
    case Engine::BODY_QUATMOM: width = 4; return B_PI;
    case Engine::BODY_OMEGA: width = 3; return B_OMEGA;
    case Engine::BODY_RCM: width = 3; return B_RCM;
    case Engine::BODY_PCM: width = 3; return B_PCM;
    case Engine::BODY_FORCE: width = 3; return B_F;
    case Engine::BODY_TORQUE: width = 3; return B_TAU;
    case Engine::BODY_INERTIA: width = 3; return B_MOI;
    default: width = 1; return B_MASS;
  }
}
}  // namespace

// out is body-major (width, nbodies) in the reference's Fortran sense: out[b*width + c]
void Engine::download_body(int what, double* out) {
  flush_kick();
  Impl& s = *d_;
  const size_t nb = (size_t)s.nbodies;
  if (nb == 0) return;
  int width = 0;
  const int off = body_item_offset(what, width);
  std::vector<double> soa(width * nb);
  CUDA_CHECK(cudaMemcpyAsync(soa.data(), s.bState.p + (size_t)off * nb, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  for (size_t b = 0; b < nb; ++b)
    for (int c = 0; c < width; ++c) out[b * width + c] = soa[(size_t)c * nb + b];
}

void Engine::upload_body(int what, const double* in) {
  flush_kick();
  Impl& s = *d_;
  const size_t nb = (size_t)s.nbodies;
  if (nb == 0) return;
  int width = 0;
  const int off = body_item_offset(what, width);
  std::vector<double> soa(width * nb);
  for (size_t b = 0; b < nb; ++b)
    for (int c = 0; c < width; ++c) soa[(size_t)c * nb + b] = in[b * width + c];
  CUDA_CHECK(cudaMemcpy(s.bState.p + (size_t)off * nb, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
}

void Engine::derive_quaternion_momenta() {
  flush_kick();
  Impl& s = *d_;
  if (s.nbodies == 0) return;
  k_body_set_omega<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s));
  stats_.launches += 1;
}

void Engine::shadow_pre(int layer0, double dt, int mode) {
  flush_kick();
  Impl& s = *d_;
  const double* Fl = s.F.p + (size_t)layer0 * 3 * s.N;
  const bool bodies = s.nbodies != 0;
  const bool dist = s.world > 1 && s.owned_valid;
  if (bodies) {   // several GPUs: body state (incl. the all-reduced F and tau of the last kick) is replicated
    s.shR0.ensure(3 * (size_t)s.nbodies);
    s.shQ0.ensure(4 * (size_t)s.nbodies);
    k_shadow_pre_bodies<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), dt, mode, s.shR0.p, s.shQ0.p);
    stats_.launches += 1;
  }
  if (s.nitems < s.N) {
    const unsigned char* mask = !bodies ? (dist ? s.owned.p : nullptr) : (dist ? s.ownedFree.p : s.freeMask.p);
    s.shS0.ensure(3 * (size_t)s.N);
    if (dist) CUDA_CHECK(cudaMemsetAsync(s.shS0.p, 0, 3 * (size_t)s.N * sizeof(double), s.stream));
    k_shadow_pre_atoms<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, mask, dt, s.R.p, s.P.p, Fl, s.invMass.p, s.shS0.p);
    stats_.launches += 1;
    // several GPUs: the force call in the middle of the step may rebuild the list and hand atoms to another rank, whose
    // post_force sum then needs the s0 the old owner stored: make s0 full on every rank (all-reduce of the owned parts)
    if (dist) gather_full(s, s.shS0.p);
  }
}

void Engine::shadow_post(int layer0, double dt, int mode, double& Us, double& Ks_t, double& Ks_r) {
  flush_kick();
  Impl& s = *d_;
  const double* Fl = s.F.p + (size_t)layer0 * 3 * s.N;
  const bool bodies = s.nbodies != 0;
  const bool dist = s.world > 1 && s.owned_valid;
  if (!bodies) {   // the reduction buffers of the body path are created by set_bodies
    s.bPartial.ensure((size_t)nblocks(s.N) * 6);
    s.bScalars.ensure(16);
    if (s.h_bscalars == nullptr) CUDA_CHECK(cudaMallocHost(&s.h_bscalars, 16 * sizeof(double)));
  }
  Us = Ks_t = Ks_r = 0.0;
  if (bodies) {
    k_shadow_post_bodies<<<nblocks(s.nbodies), TPB, 0, s.stream>>>(body_view(s), dt, mode, s.shR0.p, s.shQ0.p, s.bPartial.p,
                                                                   s.tickets.p + 3, s.bScalars.p);
    stats_.launches += 1;
  }
  if (s.nitems < s.N) {
    // (an atom that migrated during this step is summed by its NEW owner: s0 was made full on every rank in shadow_pre)
    const unsigned char* mask = !bodies ? (dist ? s.owned.p : nullptr) : (dist ? s.ownedFree.p : s.freeMask.p);
    k_shadow_post_atoms<<<nblocks(s.N), TPB, 0, s.stream>>>(s.N, mask, dt, s.R.p, s.P.p, Fl, s.invMass.p, s.shS0.p, s.bPartial.p,
                                                            s.tickets.p + 3, s.bScalars.p + 8);
    stats_.launches += 1;
    if (dist) NCCL_CHECK(nccl().AllReduce(s.bScalars.p + 8, s.bScalars.p + 8, 2, ncclDouble, ncclSum, s.comm, s.stream));
  }
  CUDA_CHECK(cudaMemcpyAsync(s.h_bscalars, s.bScalars.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  if (bodies) { Us += s.h_bscalars[0]; Ks_t += s.h_bscalars[1]; Ks_r += s.h_bscalars[2]; }
  if (s.nitems < s.N) { Us += s.h_bscalars[8]; Ks_t += s.h_bscalars[9]; }
}

long long Engine::pair_count() { return download_pairs(nullptr, 0); }

void Engine::update_list_stats(int layer0, double Lbox) {
  Impl& s = *d_;
  if (!s.list_valid) return;
  pair_count();
  const double invL2 = 1.0 / (Lbox * Lbox);
  const double Rc2s = (s.layers[layer0].useInRc ? s.InRcSq : s.RcSq) * invL2;
  CUDA_CHECK(cudaMemsetAsync(s.counter.p, 0, sizeof(unsigned long long), s.stream));
  k_refresh_positions<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p, nullptr, 0.0);
  {
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
  {
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
  const size_t nbin = (size_t)bins * nsym;
  DBuf<unsigned long long> hist;
  DBuf<unsigned short> sym;
  hist.ensure(nbin);
  sym.ensure(pairSym.size());
  CUDA_CHECK(cudaMemsetAsync(hist.p, 0, nbin * sizeof(unsigned long long), s.stream));
  CUDA_CHECK(cudaMemcpyAsync(sym.p, pairSym.data(), pairSym.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, s.stream));
  // the list is walked with the CURRENT coordinates (the reference rescales me%R on entry, EmDeeCode.f90:1324);
  // on several GPUs that includes the neighbors' halo atoms
  halo_exchange(s);
  k_refresh_positions<<<nblocks(s.Next), TPB, 0, s.stream>>>(s.Next, Lbox, s.R.p, s.q.p, s.sMeta.p, s.pos.p, nullptr, 0.0);
  const int use_smem = nbin * sizeof(unsigned int) <= 40 * 1024 ? 1 : 0;
  k_rdf<<<nblocks(s.Next), TPB, use_smem ? nbin * sizeof(unsigned int) : 0, s.stream>>>(
      s.Next, s.cap, s.nt, bins, nsym, Rc2_scaled, bins_by_Rc_scaled, s.pos.p, s.nbr.p, s.nbrCount.p, s.sType.p, sym.p,
      use_smem, hist.p);
  stats_.launches += 2;
  // slab decomposition: every rank walks the rows of the atoms it owns, so a pair is met twice over all ranks
  if (s.world > 1) NCCL_CHECK(nccl().AllReduce(hist.p, hist.p, nbin, ncclUint64, ncclSum, s.comm, s.stream));
  std::vector<unsigned long long> h(nbin);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), hist.p, nbin * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
  CUDA_CHECK(cudaStreamSynchronize(s.stream));
  CUDA_CHECK(cudaGetLastError());
  counts.resize(nbin);
  for (size_t q = 0; q < nbin; ++q) counts[q] = (long long)(h[q] / 2ull);   // full list: each pair met twice
  hist.release();
  sym.release();
}

namespace {
// device arithmetic helpers evaluated on an array of inputs (EmDeeX_math_probe; tests/test_gpu_device_math.py)
__global__ void k_math_probe(int what, int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = in[i];
  double y;
  switch (what) {
    case 0: y = fast_rcp(x); break;                        // plain-LJ pair term
    case 1: y = nb::rcp(x); break;                         // model bodies
    case 2: y = exp_nonpos(x); break;                      // typed kernel: exp of a non-positive argument
    case 3: y = uerfc_c(x, exp_nonpos(-x * x)); break;     // typed kernel: the reference's erfc
    case 4: y = nb::uerfc(x, exp(-x * x)); break;          // generic kernel: the same formula with the library exp
    default: y = 0.0;
  }
  out[i] = y;
}
}  // namespace

void math_probe(int what, int n, const double* in, double* out) {
  if (n <= 0) return;
  double *din = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&din, (size_t)n * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&dout, (size_t)n * sizeof(double)));
  CUDA_CHECK(cudaMemcpy(din, in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  k_math_probe<<<nblocks(n), TPB>>>(what, n, din, dout);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaFree(din));
  CUDA_CHECK(cudaFree(dout));
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
