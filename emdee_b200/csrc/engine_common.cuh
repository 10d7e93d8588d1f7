// engine_common.cuh -- shared device/host helpers of the engine: error macros, device buffers, constants, and the
// last-block grid-wide finish used by every reducing kernel.
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
#pragma once

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
  int* refs = nullptr;    // non-null once two systems share this allocation (EmDee_share_phase_space)
  bool managed = false;   // cudaMallocManaged: the client holds a raw pointer to it (EmDee_memory_address)
  void ensure(size_t m, double slack = 1.0) {
    if (m > n) {
      if (refs != nullptr) fatal("device buffer", "a shared phase-space array cannot be resized");
      if (p) CUDA_CHECK(cudaFree(p));
      n = (size_t)(m * slack) + 16;
      if (managed) CUDA_CHECK(cudaMallocManaged(&p, n * sizeof(T)));
      else CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
    }
  }
  void release() {
    if (refs != nullptr && --*refs > 0) {   // another system still uses the allocation
      p = nullptr;
      refs = nullptr;
      n = 0;
      return;
    }
    delete refs;
    refs = nullptr;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  // this buffer becomes a second name for `o`'s allocation (reference: `lose%R => keep%R`)
  void alias(DBuf& o) {
    if (p == o.p) return;
    release();
    if (o.refs == nullptr) o.refs = new int(1);
    ++*o.refs;
    p = o.p;
    n = o.n;
    refs = o.refs;
    managed = o.managed;
  }
  // move the contents into managed memory so that a host pointer to them can be handed out
  void to_managed() {
    if (managed || p == nullptr) return;
    if (refs != nullptr) fatal("memory address retrieving", "request the address before EmDee_share_phase_space");
    T* q = nullptr;
    CUDA_CHECK(cudaMallocManaged(&q, n * sizeof(T)));
    CUDA_CHECK(cudaMemcpy(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice));
    CUDA_CHECK(cudaFree(p));
    p = q;
    managed = true;
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
#if defined(__CUDACC__)
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ticket) : "memory");
  return old;
#else   // host-side kernel emulation used by tests/cusim (never part of the product build)
  return atomicAdd(ticket, 1u);
#endif
}

// ------------------------------------------------------------------------------------------------
// Host-visible result slot in pinned (mapped) host memory. The last block of a reducing kernel stores its scalars
// here, fences system-wide and then publishes a sequence number; the host spins on `seq` (Engine::Impl::wait_slot)
// instead of paying a D2H copy + stream synchronisation per result. One 64-byte line per slot.
// ------------------------------------------------------------------------------------------------
struct HostSlot {
  double v[7];
  unsigned long long seq;   // written last
};
constexpr int SLOT_FORCE = 0, SLOT_KINETIC = 1, SLOT_CHECK = 2, SLOT_BODY = 3, SLOT_AUX = 4, NSLOTS = 8;
constexpr int SLOT_STATUS = 5;   // v[5] of the force slot: 0 = forces computed, 1 = the kernel found "rebuild needed" and did nothing

__device__ __forceinline__ void slot_publish(HostSlot* hs, unsigned long long seq) {
#if defined(__CUDACC__)
  __threadfence_system();
  asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(&hs->seq), "l"(seq) : "memory");
#else   // tests/cusim emulation build
  hs->seq = seq;
#endif
}

// ------------------------------------------------------------------------------------------------
// Grid-wide finish without a second launch: every block publishes WIDTH partial sums, takes a ticket,
// and the block drawing the last ticket folds all partials in a FIXED order (thread t sums blocks
// t, t+T, ...; then a fixed shared-memory tree), so the result never depends on which block is last.
// ------------------------------------------------------------------------------------------------
template <int WIDTH>
__device__ __forceinline__ void grid_finish(const double (&mine)[WIDTH], double* __restrict__ partial,
                                            unsigned int* __restrict__ ticket, double* __restrict__ out,
                                            double scale_first4, HostSlot* hs = nullptr, unsigned long long seq = 0ull) {
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
#pragma unroll 4   // the loads of four blocks in flight; the additions keep their order
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
    for (int q = 0; q < WIDTH; ++q) {
      const double r = fin[0][q] * ((q < 4) ? scale_first4 : 1.0);
      out[q] = r;
      if (hs != nullptr) hs->v[q] = r;
    }
    *ticket = 0u;
    if (hs != nullptr) {
      if (WIDTH <= SLOT_STATUS) hs->v[SLOT_STATUS] = 0.0;
      slot_publish(hs, seq);
    }
  }
}

}  // namespace
}  // namespace emdee
