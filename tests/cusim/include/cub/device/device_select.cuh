// stand-in for cub::DeviceSelect in the host-side kernel emulation build (tests/cusim)
#pragma once
#include <cstddef>
namespace cub {
struct DeviceSelect {
  template <class In, class Flag, class Out, class Num>
  static cudaError_t Flagged(void* tmp, size_t& bytes, In in, Flag flags, Out out, Num num_out, int n, cudaStream_t = nullptr) {
    if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
    int k = 0;
    for (int i = 0; i < n; ++i)
      if (flags[i]) out[k++] = in[i];
    *num_out = k;
    return cudaSuccess;
  }
};
}  // namespace cub
