// lsu_probe.cu -- microbenchmark: what does one warp-wide GATHER of 32-byte position records cost on sm_100a,
// as a function of how the 32 lane addresses fall on 32-byte sectors and 128-byte lines?
//
// Why: k_pair_forces (emdee_b200/csrc/engine.cu) is bound by l1tex data-pipe wavefronts (DESIGN.md section 5:
// ~26 wavefronts per warp-iteration). Whether a wavefront is spent per distinct SECTOR or per distinct LINE,
// and what the same gather costs from shared memory in AoS (LDS.128 x2) or SoA (LDS.64 x3) form, decides which
// data layout is worth building next. This tool measures it; it is not part of the product.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lsu_probe tools/lsu_probe.cu && ./lsu_probe
//
// Also measured: the same gather through the texture front-end (TLD), LDG/TEX alternation (do the two front-ends add
// up or share one data stage?) and a 16-byte record (what a compact position format would buy).
//
// Output: one line per (path, pattern): SM cycles per warp-gather at full occupancy (throughput, not latency).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

enum Pattern { SAME = 0, COALESCED, GROUP4, GROUP2, WINDOW42, WINDOW128, RANDOM, NPATTERN };
static const char* pattern_name[NPATTERN] = {
    "all lanes one record (1 sector, 1 line)",
    "lane l -> record b+l (32 sectors, 8 lines)",
    "4-lane groups contiguous, groups scattered (32 sectors, 8 lines)",
    "2-lane groups contiguous, groups scattered (32 sectors, 16 lines)",
    "random inside a 42-record window (~22 sectors, <=11 lines)",
    "random inside a 128-record window (~28 sectors, <=32 lines)",
    "fully scattered (32 sectors, 32 lines)"};

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// record index for (warp-iteration key, lane) under a pattern; nrec is a power of two
__device__ __forceinline__ unsigned pick(int pattern, unsigned key, unsigned lane, unsigned mask) {
  const unsigned base = hash32(key) & mask;
  switch (pattern) {
    case SAME: return base;
    case COALESCED: return (base + lane) & mask;
    case GROUP4: return ((hash32(key * 8u + (lane >> 2)) & mask & ~3u) + (lane & 3u)) & mask;
    case GROUP2: return ((hash32(key * 16u + (lane >> 1)) & mask & ~1u) + (lane & 1u)) & mask;
    case WINDOW42: return (base + hash32(key * 32u + lane) % 42u) & mask;
    case WINDOW128: return (base + (hash32(key * 32u + lane) & 127u)) & mask;
    default: return hash32(key * 32u + lane) & mask;
  }
}

__device__ __forceinline__ double4 ldg256(const double4* p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

constexpr int ITER = 2048;
constexpr int UNROLL = 4;

// path 0: LDG.E.256 of an AoS record   path 1: 2 x LDG.E.128   path 2: 3 x LDG.E.64 from SoA arrays
// path 3: the same 32-byte record through the TEXTURE front-end (2 x tex1Dfetch<int4>)
// path 4: alternate gathers between LDG.E.256 and the texture front-end (do the two front-ends overlap?)
// path 5: a 16-byte record (what a cell-relative fixed-point position format would gather), one LDG.E.128
template <int PATH>
__global__ void __launch_bounds__(256) k_global(const double4* __restrict__ rec, const double* __restrict__ soa,
                                                cudaTextureObject_t tex, unsigned mask, int pattern, double* sink,
                                                long long* cycles) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double acc = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < ITER; it += UNROLL) {
    double4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned j = pick(pattern, warp * ITER + it + u, lane, mask);
      if (PATH == 0) {
        v[u] = ldg256(rec + j);
      } else if (PATH == 1) {
        const double2* p = reinterpret_cast<const double2*>(rec + j);
        const double2 a = __ldg(p), b = __ldg(p + 1);
        v[u] = make_double4(a.x, a.y, b.x, b.y);
      } else if (PATH == 2) {
        const size_t n = (size_t)mask + 1;
        v[u] = make_double4(__ldg(soa + j), __ldg(soa + n + j), __ldg(soa + 2 * n + j), 0.0);
      } else if (PATH == 3 || (PATH == 4 && (u & 1))) {
        const int4 a = tex1Dfetch<int4>(tex, 2 * (int)j), b = tex1Dfetch<int4>(tex, 2 * (int)j + 1);
        v[u] = make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), 0.0);
      } else if (PATH == 4) {
        v[u] = ldg256(rec + j);
      } else {
        const double2 a = __ldg(reinterpret_cast<const double2*>(rec) + j);   // 16-byte records, same index pattern
        v[u] = make_double4(a.x, a.y, 0.0, 0.0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z;
  }
  const long long t1 = clock64();
  if (acc == 123.456) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory paths: path 0 = AoS double4 records (LDS.128 x2), path 1 = SoA (LDS.64 x3),
// path 2 = AoS with a 40-byte record pitch (LDS.64 x3 at odd 8-byte strides)
template <int PATH>
__global__ void __launch_bounds__(256) k_shared(int nrec, int pattern, double* sink, long long* cycles) {
  extern __shared__ __align__(16) double sm[];
  const int pitch = (PATH == 2) ? 5 : 4;
  for (int i = threadIdx.x; i < nrec * pitch; i += blockDim.x) sm[i] = (double)i;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned mask = (unsigned)nrec - 1u;
  double acc = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < ITER; it += UNROLL) {
    double x[UNROLL], y[UNROLL], z[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned j = pick(pattern, warp * ITER + it + u, lane, mask);
      if (PATH == 0) {
        const double4 v = reinterpret_cast<const double4*>(sm)[j];
        x[u] = v.x; y[u] = v.y; z[u] = v.z;
      } else if (PATH == 1) {
        x[u] = sm[j]; y[u] = sm[nrec + j]; z[u] = sm[2 * nrec + j];
      } else {
        x[u] = sm[5 * j]; y[u] = sm[5 * j + 1]; z[u] = sm[5 * j + 2];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += x[u] + y[u] + z[u];
  }
  const long long t1 = clock64();
  if (acc == 123.456) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  std::printf("# %s, %d SMs; cycles per warp-gather per SM at 8 warps/block x 4 blocks/SM (global) or x 2 (shared)\n",
              prop.name, sms);

  double* sink;
  long long* cycles;
  CK(cudaMalloc(&sink, 8));
  const int maxblocks = sms * 8;
  CK(cudaMalloc(&cycles, maxblocks * sizeof(long long)));
  std::vector<long long> h(maxblocks);

  auto report = [&](const char* path, const char* foot, int pattern, int blocks, int blocks_per_sm) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), cycles, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int b = 0; b < blocks; ++b) mean += (double)h[b];
    mean /= blocks;
    // every block runs ITER gathers in each of its 8 warps; blocks_per_sm blocks share an SM concurrently
    const double per = mean / ((double)ITER * 8 * blocks_per_sm);
    std::printf("%-26s %-10s %-66s %7.2f cyc\n", path, foot, pattern_name[pattern], per);
  };

  // global-memory paths at two footprints: 2^11 records (64 KB, L1-resident) and 2^20 records (32 MB, L2-resident)
  for (int logn : {11, 20}) {
    const size_t n = (size_t)1 << logn;
    double4* rec;
    double* soa;
    CK(cudaMalloc(&rec, n * sizeof(double4)));
    CK(cudaMalloc(&soa, 3 * n * sizeof(double)));
    CK(cudaMemset(rec, 0, n * sizeof(double4)));
    CK(cudaMemset(soa, 0, 3 * n * sizeof(double)));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = rec;
    rd.res.linear.desc = cudaCreateChannelDesc<int4>();
    rd.res.linear.sizeInBytes = n * sizeof(double4);
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const char* foot = logn == 11 ? "L1 64KB" : "L2 32MB";
    const int blocks = sms * 4;
    for (int p = 0; p < NPATTERN; ++p) {
      for (int rep = 0; rep < 2; ++rep) k_global<0><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("LDG.E.256 AoS", foot, p, blocks, 4);
      for (int rep = 0; rep < 2; ++rep) k_global<1><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("2 x LDG.E.128 AoS", foot, p, blocks, 4);
      for (int rep = 0; rep < 2; ++rep) k_global<2><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("3 x LDG.E.64 SoA", foot, p, blocks, 4);
      for (int rep = 0; rep < 2; ++rep) k_global<3><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("2 x TEX int4 AoS", foot, p, blocks, 4);
      for (int rep = 0; rep < 2; ++rep) k_global<4><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("LDG.256 / TEX alternating", foot, p, blocks, 4);
      for (int rep = 0; rep < 2; ++rep) k_global<5><<<blocks, 256>>>(rec, soa, tex, (unsigned)n - 1, p, sink, cycles);
      report("LDG.E.128 16-byte record", foot, p, blocks, 4);
    }
    CK(cudaDestroyTextureObject(tex));
    CK(cudaFree(rec));
    CK(cudaFree(soa));
  }

  // shared-memory paths: 2048 records per block (64 KB AoS / 48 KB SoA / 80 KB pitch-5), 2 blocks per SM
  {
    const int nrec = 2048;
    const int blocks = sms * 2;
    CK(cudaFuncSetAttribute(k_shared<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrec * 32));
    CK(cudaFuncSetAttribute(k_shared<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrec * 24));
    CK(cudaFuncSetAttribute(k_shared<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrec * 40));
    for (int p = 0; p < NPATTERN; ++p) {
      for (int rep = 0; rep < 2; ++rep) k_shared<0><<<blocks, 256, nrec * 32>>>(nrec, p, sink, cycles);
      report("LDS.128 x2 AoS(32B)", "smem", p, blocks, 2);
      for (int rep = 0; rep < 2; ++rep) k_shared<1><<<blocks, 256, nrec * 24>>>(nrec, p, sink, cycles);
      report("LDS.64 x3 SoA", "smem", p, blocks, 2);
      for (int rep = 0; rep < 2; ++rep) k_shared<2><<<blocks, 256, nrec * 40>>>(nrec, p, sink, cycles);
      report("LDS.64 x3 AoS(40B pitch)", "smem", p, blocks, 2);
    }
  }
  return 0;
}
