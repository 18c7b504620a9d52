// Shared helpers for libcsn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <algorithm>

#include "../../include/csn_b200.h"

namespace csn {

void set_error(const char* fmt, ...);
int sm_count();
void count_launches(unsigned long long n);  // bookkeeping for csn_launch_count()

#define CSN_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      csn::set_error(__VA_ARGS__);            \
      return CSN_EINVAL;                      \
    }                                         \
  } while (0)

#define CSN_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      csn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CSN_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

#define CSN_LAUNCH_CHECK()          \
  do {                              \
    csn::count_launches(1);         \
    CSN_CUDA(cudaGetLastError());   \
  } while (0)

#define CSN_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != CSN_OK) return _r; \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace csn
