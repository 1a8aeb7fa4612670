// Shared host-side helpers of the popnet_b200 CUDA library (error capture, launch accounting).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "../../include/popnet_b200.h"

namespace popnet {

// The only process-wide state of the library: a sticky copy of the last CUDA error code and a
// launch counter (both diagnostics; no results are ever kept between calls).
extern std::atomic<int> g_last_cuda_error;
extern std::atomic<long long> g_launch_count;

inline int record_cuda_error(cudaError_t e) {
  g_last_cuda_error.store(static_cast<int>(e));
  return POPNET_ERR_CUDA;
}

}  // namespace popnet

#define POPNET_CUDA_TRY(expr)                                        \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) return popnet::record_cuda_error(_e);     \
  } while (0)

// call right after every <<< >>> launch
#define POPNET_AFTER_LAUNCH()                                        \
  do {                                                               \
    popnet::g_launch_count.fetch_add(1);                             \
    cudaError_t _e = cudaGetLastError();                             \
    if (_e != cudaSuccess) return popnet::record_cuda_error(_e);     \
  } while (0)
