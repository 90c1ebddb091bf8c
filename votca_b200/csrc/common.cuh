// Shared helpers for the gwbse_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>

namespace gwbse {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
                    std::to_string(line) + ")");
  }
}
#define GW_CUDA(x) ::gwbse::check_cuda((x), #x, __FILE__, __LINE__)
#define GW_REQUIRE(cond, msg)                                                  \
  do {                                                                         \
    if (!(cond)) throw std::runtime_error(std::string(msg) + " [" #cond "]"); \
  } while (0)

// cudaMemcpy2DAsync that collapses to one linear copy when both sides are contiguous (pageable 2-D copies
// are issued row by row by the driver and are an order of magnitude slower)
inline cudaError_t copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                                cudaMemcpyKind kind, cudaStream_t stream) {
  if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * height, kind, stream);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, stream);
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

}  // namespace gwbse
