// stand-in for cub::DeviceRadixSort in the host-side kernel emulation build (tests/cusim)
#pragma once
#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int = 0,
                               int = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
    for (int i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return cudaSuccess;
  }
  template <class K>
  static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* kin, K* kout, int n, int = 0, int = sizeof(K) * 8,
                              cudaStream_t = nullptr) {
    if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
    std::vector<K> v(kin, kin + n);
    std::sort(v.begin(), v.end());
    for (int i = 0; i < n; ++i) kout[i] = v[i];
    return cudaSuccess;
  }
};
}  // namespace cub
