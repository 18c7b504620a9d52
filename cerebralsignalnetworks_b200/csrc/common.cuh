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

// ---- deterministic cross-CTA sum ("the last CTA folds") -----------------------------------------------------------
// Float atomics make a result depend on the order in which CTAs happen to retire; the reference's CPU reductions do
// not.  Every CTA stores its partial into a caller-provided scratch block, takes a ticket, and the LAST CTA to arrive
// adds the partials in CTA order (lane-strided, then a fixed shuffle tree): same inputs -> same bits, every run.
// Scratch layout: 64-byte header (word 0 = ticket, zeroed by the host wrapper before the launch) + one float per CTA.
constexpr size_t kDetHeader = 64;
inline size_t det_scratch_bytes(size_t n_partials) { return kDetHeader + ((n_partials * 4 + 63) & ~size_t(63)); }

// Call from ALL threads of a participating CTA (it synchronises the CTA).  `v` is the CTA's partial, read from the
// first thread only.  `cta` / `n_cta` index the participants.  *out = sum (overwritten) by the last CTA.
__device__ __forceinline__ void det_cta_sum(float v, void* scratch, unsigned cta, unsigned n_cta, float* out) {
  __shared__ int s_det_last;
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch);
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + kDetHeader);
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid == 0) {
    __stcg(part + cta, v);
    __threadfence();
    s_det_last = (atomicAdd(ticket, 1u) == n_cta - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_det_last || tid >= 32) return;
  __threadfence();
  float a = 0.f;
  for (unsigned i = tid; i < n_cta; i += 32) a += __ldcg(part + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (tid == 0) {
    *out = a;
    *ticket = 0u;
  }
}

// out[n] (+)= sum_m x[m, n], fixed summation order (defined in optim_elementwise.cu)
int colsum_det(const float* x, float* out, int M, int N, int ldx, int accumulate, cudaStream_t s);

}  // namespace csn
