// Error plumbing, device info, casts and layout changes.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace csn {

static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(y + i) = o;
  }
  // tail (n % 4) handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    size_t j = (n & ~size_t(3)) + threadIdx.x;
    y[j] = __float2bfloat16_rn(x[j]);
  }
}

__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}

// [B,T,C] fp32 -> [T,B,C] (f32 or bf16).  C is contiguous on both sides, so this is a row gather:
// one warp moves one (b,t) row of C floats with float4 loads.
template <typename OutT>
__global__ void btc_to_tbc_kernel(const float* __restrict__ x, OutT* __restrict__ y, int B, int T, int C) {
  const int rows = B * T;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const int t = r / B, b = r - t * B;  // destination-major order keeps the writes streaming
    const float* src = x + (size_t(b) * T + t) * C;
    OutT* dst = y + size_t(r) * C;
    for (int c = lane; c < C; c += 32) {
      float v = src[c];
      if constexpr (sizeof(OutT) == 4) dst[c] = v; else dst[c] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace csn

using namespace csn;

extern "C" int csn_version(void) { return CSN_VERSION; }
extern "C" const char* csn_last_error(void) { return csn::g_err; }
extern "C" unsigned long long csn_launch_count(void) { return csn::g_launches.load(std::memory_order_relaxed); }

extern "C" int csn_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  CSN_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  CSN_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  CSN_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  CSN_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  if (b != 10) {
    set_error("libcsn_b200 is built for sm_100a only; device reports sm_%d%d", b, c);
    return CSN_EARCH;
  }
  return CSN_OK;
}

extern "C" int csn_cast(const void* x, int x_dtype, void* y, int y_dtype, size_t n, void* stream) {
  CSN_REQUIRE(x && y, "csn_cast: null pointer");
  if (n == 0) return CSN_OK;
  cudaStream_t s = as_stream(stream);
  if (x_dtype == CSN_F32 && y_dtype == CSN_BF16) {
    int blocks = (int)std::min<size_t>(ceil_div<size_t>(n, 256 * 4), size_t(sm_count()) * 8);
    if (blocks < 1) blocks = 1;
    cast_f32_to_bf16_kernel<<<blocks, 256, 0, s>>>((const float*)x, (__nv_bfloat16*)y, n);
  } else if (x_dtype == CSN_BF16 && y_dtype == CSN_F32) {
    int blocks = (int)std::min<size_t>(ceil_div<size_t>(n, 256), size_t(sm_count()) * 8);
    cast_bf16_to_f32_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, (float*)y, n);
  } else if (x_dtype == y_dtype) {
    CSN_CUDA(cudaMemcpyAsync(y, x, n * (x_dtype == CSN_F32 ? 4 : 2), cudaMemcpyDeviceToDevice, s));
    return CSN_OK;
  } else {
    CSN_REQUIRE(false, "csn_cast: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
  }
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_btc_to_tbc(const float* x, void* y, int B, int T, int C, int out_dtype, void* stream) {
  CSN_REQUIRE(x && y && B > 0 && T > 0 && C > 0, "csn_btc_to_tbc: bad arguments");
  int blocks = min(ceil_div(B * T, 8), sm_count() * 16);
  if (out_dtype == CSN_F32)
    btc_to_tbc_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(x, (float*)y, B, T, C);
  else if (out_dtype == CSN_BF16)
    btc_to_tbc_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)y, B, T, C);
  else
    CSN_REQUIRE(false, "csn_btc_to_tbc: bad out_dtype %d", out_dtype);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
