// cusim.h -- TEST INFRASTRUCTURE: a minimal functional emulator of the CUDA execution model, just large enough
// to run emdee_b200/csrc/engine.cu and its engine_*.cuh kernels on the host, unchanged, so that kernel LOGIC
// (indexing, loop bounds, reductions, masks) can be checked against the oracle on a machine without a GPU.
//
// It is not a product path and not a CPU fallback: the emulated library is built only by the test-suite
// (tests/cusim/build.py -> tests/_build/libemdee_cusim.so), nothing under emdee_b200/ refers to it, and the
// product library keeps failing loudly without a device. It says nothing about performance or about
// memory-model races; it executes one block at a time, the threads of a block as cooperatively scheduled fibers
// (a small x86-64 context switch), with __syncthreads / __syncwarp / warp shuffles / ballots implemented as fiber barriers.
//
// The sources are compiled by g++ after tests/cusim/build.py has rewritten the two constructs C++ cannot parse:
// `kernel<<<grid, block, smem, stream>>>(args)` -> `cusim::launch([&] { kernel(args); }, grid, block, smem, stream)`
// and `extern __shared__ T name[];` -> a pointer to the launch's dynamic shared-memory buffer.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- keywords ------------------------------------------------------------------------------------------------
#define __global__
#define __constant__ static const
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ----------------------------------------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct uint3 { unsigned int x, y, z; };
struct dim3 {
  unsigned int x = 1, y = 1, z = 1;
  dim3() {}
  template <class T>
  dim3(T x_) : x((unsigned int)x_) {}
  dim3(unsigned int x_, unsigned int y_, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

// ---- runtime API (memory is host memory; streams and events are inert; everything is synchronous) -----------------
typedef int cudaError_t;
typedef struct cusimStream* cudaStream_t;
typedef struct cusimEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaHostAllocMapped = 2, cudaHostAllocPortable = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
// texture objects over linear memory: the "object" is the base pointer
typedef unsigned long long cudaTextureObject_t;
enum cudaResourceType { cudaResourceTypeLinear = 2 };
enum cudaTextureReadMode { cudaReadModeElementType = 0 };
struct cudaChannelFormatDesc { int x, y, z, w, f; };
template <class T>
inline cudaChannelFormatDesc cudaCreateChannelDesc() { return cudaChannelFormatDesc{32, 32, 32, 32, 0}; }
struct cudaResourceDesc {
  cudaResourceType resType;
  struct { struct { void* devPtr; cudaChannelFormatDesc desc; size_t sizeInBytes; } linear; } res;
};
struct cudaTextureDesc { cudaTextureReadMode readMode; };

inline const char* cudaGetErrorString(cudaError_t) { return "cusim error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// every host pointer counts as pinned + mapped (the emulated "device" is the host)
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  *p = reinterpret_cast<T*>(std::aligned_alloc(256, ((bytes + 255) / 256 + 1) * 256));
  return *p ? cudaSuccess : 2;
}
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
template <class T>
inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
template <class T>
inline cudaError_t cudaMallocManaged(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- the fiber scheduler -------------------------------------------------------------------------------------------
// Context switch: callee-saved registers + stack pointer (x86-64 SysV). swapcontext() would cost two signal-mask
// system calls per switch, and a warp shuffle is two barriers per lane.
extern "C" void cusim_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cusim_switch,.-cusim_switch
)");

namespace cusim {

struct Fiber {
  void* sp = nullptr;
  uint3 tid;
  int linear = 0;
  bool done = false;
  const volatile unsigned* wait = nullptr;   // blocked until *wait != wait_val (nullptr: runnable)
  unsigned wait_val = 0;
  unsigned exchanges = 0;                    // warp exchanges done so far (selects the slot buffer)
};

struct Warp {
  int live = 0, count = 0;
  unsigned gen = 0;
  unsigned long long slot[2][32];   // double-buffered: exchange k uses buffer k&1, so one barrier per exchange suffices
  unsigned snap[2] = {0u, 0u};      // lanes that took part in the exchange (a lane may have RETURNED by the time a slower lane reads)
  bool alive[32];
};

struct BlockState {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int live = 0, count = 0;
  unsigned gen = 0;
  Fiber* cur = nullptr;
  void* main_sp = nullptr;
  const std::function<void()>* body = nullptr;
  unsigned char* dyn = nullptr;
};

inline BlockState& st() {
  static BlockState s;
  return s;
}
inline dim3& block_idx() { static dim3 v; return v; }
inline dim3& block_dim() { static dim3 v; return v; }
inline dim3& grid_dim() { static dim3 v; return v; }

inline void yield() {   // give the other threads of the block a turn (used directly by spin-waits)
  BlockState& b = st();
  cusim_switch(&b.cur->sp, b.main_sp);
}

inline void release_if_complete_block(BlockState& b) {
  if (b.count > 0 && b.count >= b.live) {
    b.count = 0;
    b.gen++;
  }
}
inline void release_if_complete_warp(Warp& w) {
  if (w.count > 0 && w.count >= w.live) {
    w.count = 0;
    w.gen++;
  }
}

inline void wait_for(const volatile unsigned* gen, unsigned seen) {
  BlockState& b = st();
  while (*gen == seen) {
    b.cur->wait = gen;
    b.cur->wait_val = seen;
    yield();
  }
  b.cur->wait = nullptr;
}
inline void sync_block() {
  BlockState& b = st();
  const unsigned gen = b.gen;
  b.count++;
  release_if_complete_block(b);
  wait_for(&b.gen, gen);
}
inline void sync_warp() {
  BlockState& b = st();
  Warp& w = b.warps[b.cur->linear >> 5];
  const unsigned gen = w.gen;
  w.count++;
  release_if_complete_warp(w);
  wait_for(&w.gen, gen);
}

inline void fiber_main() {
  BlockState& b = st();
  (*b.body)();
  Fiber* f = b.cur;
  f->done = true;
  b.live--;
  Warp& w = b.warps[f->linear >> 5];
  w.live--;
  w.alive[f->linear & 31] = false;
  release_if_complete_block(b);   // a thread that returned early must not hold a barrier the others wait on
  release_if_complete_warp(w);
  cusim_switch(&f->sp, b.main_sp);
  std::abort();                   // a finished fiber is never resumed
}

constexpr size_t STACK_BYTES = 256 * 1024;

// CUSIM_ORDER: unset/"forward" = threads of a block take turns in index order; "reverse"; "random[:seed]" = a new
// random order every scheduling round (results must not depend on it: anything else is a missing barrier)
inline unsigned long long& rng_state() { static unsigned long long s = 88172645463325252ull; return s; }
inline unsigned long long next_random() {
  unsigned long long& x = rng_state();
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  return x;
}
inline int schedule_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("CUSIM_ORDER");
    mode = 0;
    if (e != nullptr && std::strncmp(e, "reverse", 7) == 0) mode = 1;
    if (e != nullptr && std::strncmp(e, "random", 6) == 0) {
      mode = 2;
      if (e[6] == ':') rng_state() = 0x9e3779b97f4a7c15ull * (std::strtoull(e + 7, nullptr, 10) + 1);
    }
  }
  return mode;
}

inline unsigned char* stack_pool(size_t nthreads) {
  static std::vector<unsigned char> pool;
  if (pool.size() < nthreads * STACK_BYTES + 64) pool.resize(nthreads * STACK_BYTES + 64);
  return pool.data();
}

template <class F>
inline void launch(F fn, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr) {
  BlockState& b = st();
  const std::function<void()> body = fn;
  const size_t nthreads = (size_t)block.x * block.y * block.z;
  if (nthreads == 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  std::vector<unsigned char> dyn(smem + 256);
  unsigned char* stacks = stack_pool(nthreads);
  grid_dim() = grid;
  block_dim() = block;
  b.body = &body;
  b.dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~(uintptr_t)127);
  // blocks run one after the other; CUSIM_ORDER also permutes WHICH block goes when (grid-wide finishes must not
  // care which block happens to be the last one)
  const size_t nblocks_total = (size_t)grid.x * grid.y * grid.z;
  std::vector<size_t> block_order(nblocks_total);
  for (size_t q = 0; q < nblocks_total; ++q) block_order[q] = q;
  if (schedule_mode() == 1) std::reverse(block_order.begin(), block_order.end());
  if (schedule_mode() >= 2)
    for (size_t q = nblocks_total - 1; q > 0; --q) std::swap(block_order[q], block_order[next_random() % (q + 1)]);
  for (size_t q = 0; q < nblocks_total; ++q) {
        const size_t lin = block_order[q];
        const unsigned bx = (unsigned)(lin % grid.x), by = (unsigned)((lin / grid.x) % grid.y), bz = (unsigned)(lin / ((size_t)grid.x * grid.y));
        block_idx() = dim3(bx, by, bz);
        b.fibers.assign(nthreads, Fiber());
        b.warps.assign((nthreads + 31) / 32, Warp());
        b.live = (int)nthreads;
        b.count = 0;
        b.gen = 0;
        for (size_t t = 0; t < nthreads; ++t) {
          Fiber& f = b.fibers[t];
          f.linear = (int)t;
          f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
          Warp& w = b.warps[t >> 5];
          w.live++;
          w.alive[t & 31] = true;
          // initial frame: six callee-saved registers, then the entry point as the return address; after the `ret`
          // the stack pointer is 8 mod 16, as at any function entry
          uintptr_t top = (reinterpret_cast<uintptr_t>(stacks + (t + 1) * STACK_BYTES)) & ~(uintptr_t)15;
          void** sp = reinterpret_cast<void**>(top - 8);
          *--sp = reinterpret_cast<void*>(&fiber_main);
          for (int r = 0; r < 6; ++r) *--sp = nullptr;
          f.sp = sp;
        }
        for (size_t w = 0; w < b.warps.size(); ++w)
          for (int l = 0; l < 32; ++l)
            if (w * 32 + l >= nthreads) b.warps[w].alive[l] = false;
        long long idle_rounds = 0;
        std::vector<size_t> order(nthreads);
        for (size_t t = 0; t < nthreads; ++t) order[t] = t;
        const int mode = schedule_mode();
        if (mode == 1) std::reverse(order.begin(), order.end());
        while (b.live > 0) {
          bool ran = false;
          if (mode >= 2)   // a fresh random interleaving every round: shakes out missing barriers
            for (size_t t = nthreads - 1; t > 0; --t) std::swap(order[t], order[next_random() % (t + 1)]);
          for (size_t k = 0; k < nthreads; ++k) {
            const size_t t = order[k];
            Fiber& f = b.fibers[t];
            if (f.done || (f.wait != nullptr && *f.wait == f.wait_val)) continue;
            b.cur = &f;
            cusim_switch(&b.main_sp, f.sp);
            ran = true;
          }
          idle_rounds = ran ? 0 : idle_rounds + 1;
          if (idle_rounds > 2) {
            std::fprintf(stderr, "cusim: deadlock -- every live thread of block (%u,%u,%u) waits on a barrier\n", bx, by, bz);
            std::abort();
          }
        }
      }
  b.body = nullptr;
}

inline unsigned char* dyn_smem() { return st().dyn; }
inline const uint3& thread_idx() { return st().cur->tid; }

// One warp-wide exchange: every live lane publishes 8 bytes, one barrier, every lane may read any lane's value.
// A lane can only reach exchange k+2 after all lanes passed the barrier of exchange k+1, i.e. after they finished
// reading the buffer of exchange k, so two alternating buffers are enough.
inline const unsigned long long* exchange(unsigned long long mine) {
  BlockState& b = st();
  Warp& w = b.warps[b.cur->linear >> 5];
  const unsigned buf = b.cur->exchanges++ & 1u;
  w.slot[buf][b.cur->linear & 31] = mine;
  if (w.count + 1 >= w.live) {   // last lane to arrive: record who is taking part
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
      if (w.alive[l]) m |= 1u << l;
    w.snap[buf] = m;
  }
  sync_warp();
  return w.slot[buf];
}

template <class T>
inline T shuffle(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  const unsigned buf = st().cur->exchanges & 1u;
  const unsigned long long* all = exchange(raw);
  BlockState& b = st();
  const Warp& w = b.warps[b.cur->linear >> 5];
  T r = v;
  if (src_lane >= 0 && src_lane < 32 && ((w.snap[buf] >> src_lane) & 1u)) std::memcpy(&r, &all[src_lane], sizeof(T));
  return r;
}

}  // namespace cusim

#define threadIdx (cusim::thread_idx())
#define blockIdx (cusim::block_idx())
#define blockDim (cusim::block_dim())
#define gridDim (cusim::grid_dim())

// ---- synchronisation and warp primitives ---------------------------------------------------------------------------
inline void __syncthreads() { cusim::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cusim::sync_warp(); }
inline void __threadfence() {}
inline void __threadfence_system() {}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cusim::shuffle(v, src); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int mask, int = 32) { return cusim::shuffle(v, (cusim::st().cur->linear & 31) ^ mask); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32) {
  const int lane = cusim::st().cur->linear & 31;
  return cusim::shuffle(v, lane + (int)delta < 32 ? lane + (int)delta : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned buf = cusim::st().cur->exchanges & 1u;
  const unsigned long long* all = cusim::exchange(pred ? 1ull : 0ull);
  cusim::BlockState& b = cusim::st();
  const cusim::Warp& w = b.warps[b.cur->linear >> 5];
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l)
    if (((w.snap[buf] >> l) & 1u) && all[l]) bits |= 1u << l;
  return bits;
}

// ---- atomics (one OS thread, cooperative fibers: plain read-modify-write is atomic) ---------------------------------
template <class T>
inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <class T>
inline T atomicMax(T* p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <class T>
inline T atomicMin(T* p, T v) { T old = *p; if (v < old) *p = v; return old; }

// ---- arithmetic intrinsics (compile with -ffp-contract=off: the _rn forms must not fuse) ----------------------------
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int cudaCreateTextureObject(cudaTextureObject_t* t, const cudaResourceDesc* r, const cudaTextureDesc*, const void*) {
  *t = reinterpret_cast<cudaTextureObject_t>(r->res.linear.devPtr);
  return 0;
}
inline int cudaDestroyTextureObject(cudaTextureObject_t) { return 0; }
template <class T>
inline T tex1Dfetch(cudaTextureObject_t t, int i) { return reinterpret_cast<const T*>(t)[i]; }
inline double __hiloint2double(int hi, int lo) {
  const unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
}
inline long long __double2ll_rn(double a) { return std::llrint(a); }
inline int __double2hiint(double a) { unsigned long long u; std::memcpy(&u, &a, 8); return (int)(u >> 32); }
inline int __double2loint(double a) { unsigned long long u; std::memcpy(&u, &a, 8); return (int)(u & 0xffffffffull); }
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
inline void sincospi(double x, double* s, double* c) { sincos(3.14159265358979323846 * x, s, c); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline void __stcg(T* p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
inline long long clock64() { return 0; }
using std::max;
using std::min;
