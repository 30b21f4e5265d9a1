// Shared helpers for libgppvae_b200 (error plumbing, argument checks, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gppvae_b200.h"

namespace gpp {

void set_error(const char* fmt, ...);
void count_launch();  // bumps the process-wide kernel-launch counter read by gpp_launch_count()

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define GPP_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      gpp::set_error(__VA_ARGS__);             \
      return GPP_ERR_INVALID_ARGUMENT;         \
    }                                          \
  } while (0)

#define GPP_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      gpp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return GPP_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define GPP_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    gpp::count_launch();                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      gpp::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return GPP_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define GPP_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != GPP_OK) return rc__; \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

int sm_count();  // SMs of the current device (cached)

// softmax of the two log-variances (gp.py:48-50), in double.
__device__ __forceinline__ void softmax2(const float* __restrict__ lvs, double& v0, double& vn) {
  const double a = (double)lvs[0], b = (double)lvs[1];
  const double m = a > b ? a : b;
  const double ea = exp(a - m), eb = exp(b - m);
  v0 = ea / (ea + eb);
  vn = eb / (ea + eb);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace gpp
