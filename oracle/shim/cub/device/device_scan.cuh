// Build shim (test infrastructure only): stands in for CUB while compiling the
// reference prototype's *CPU* loops out of /root/reference/seg_ops_cuda/seg_ops.cu.
// The prototype passes NULL offset iterators to CUB size queries, which CUB 2.8
// rejects at compile time; none of the CUB-backed GPU code is ever called through
// oracle/_ref, so every entry point is an inert stub.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace cub {
struct Max {}; struct Min {}; struct Sum {};
struct DeviceScan {
  template <typename... E, typename... A> static cudaError_t InclusiveScan(A...) { return cudaSuccess; }
};
struct DeviceSegmentedReduce {
  template <typename... E, typename... A> static cudaError_t Reduce(A...) { return cudaSuccess; }
};
struct DeviceRadixSort {
  template <typename... E, typename... A> static cudaError_t SortPairs(A...) { return cudaSuccess; }
};
}  // namespace cub
