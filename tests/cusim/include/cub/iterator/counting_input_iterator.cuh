// stand-in for cub::CountingInputIterator in the host-side kernel emulation build (tests/cusim)
#pragma once
#include <cstddef>
namespace cub {
template <class T>
struct CountingInputIterator {
  T base;
  explicit CountingInputIterator(T b) : base(b) {}
  T operator[](size_t i) const { return base + (T)i; }
};
}  // namespace cub
