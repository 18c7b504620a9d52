// Bring-up / self-test hooks (include/csn_b200_debug.h): NOT part of the product ABI, nothing on the train step calls them.
// One tcgen05.mma tile through the no-swizzle canonical operand layouts the recurrence kernels use (tests/test_gpu_gemm.py),
// and the tcgen05.mma issue / completion cost microbenchmark behind the cost model in DESIGN.md (scripts/umma_bench*.py).
#include "tc.cuh"
#include "../../include/csn_b200_debug.h"

namespace csn {

using namespace tc;

// ---------------------------------------------------------------------------------------- bring-up hook
// D[128, N] = A[128, K] . B[N, K]^T with both operands staged by threads into the no-swizzle canonical layouts
// (K-major or MN-major), one tcgen05.mma chain, tcgen05.ld epilogue.  Exercises exactly the descriptor
// conventions the recurrence kernels rely on.
__global__ void __launch_bounds__(128, 1) dbg_umma_tile_kernel(const __nv_bfloat16* __restrict__ A,
                                                              const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                                                              int N, int K, int a_mode, int b_mn) {
  // a_mode: 0 = K-major shared memory, 1 = MN-major shared memory, 2 = tensor memory (TS form of tcgen05.mma)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
  uint8_t* sa = base;
  uint8_t* sb = base + a_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + b_bytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lbo_a = 16 * 128, lbo_b = (N / 8) * 128, sbo = 128;
  if (a_mode != 2) {
    for (int e = tid; e < 128 * K; e += 128) {
      int r = e / K, k = e % K;
      uint32_t off = a_mode ? canon_mn_off(r, k, lbo_a, sbo) : canon_k_off(r, k, lbo_a, sbo);
      *reinterpret_cast<__nv_bfloat16*>(sa + off) = A[e];
    }
  }
  for (int e = tid; e < N * K; e += 128) {
    int r = e / K, k = e % K;
    uint32_t off = b_mn ? canon_mn_off(r, k, lbo_b, sbo) : canon_k_off(r, k, lbo_b, sbo);
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = Bm[e];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  const uint32_t a_col0 = 256;  // A operand columns when it lives in tensor memory (K/2 <= 128 columns)
  const uint32_t ncols = (a_mode == 2) ? 512 : (N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)));
  if (warp == 0) tmem_alloc(slot, ncols);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *slot;
  if (a_mode == 2) {
    const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = (c0 + c) * 2;
        __nv_bfloat162 bb;
        bb.x = (k < K) ? A[size_t(tid) * K + k] : __float2bfloat16(0.f);
        bb.y = (k + 1 < K) ? A[size_t(tid) * K + k + 1] : __float2bfloat16(0.f);
        r[c] = *reinterpret_cast<uint32_t*>(&bb);
      }
      tmem_st32(lane_addr + a_col0 + c0, r);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mode == 1 ? 1 : 0, b_mn);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t db = make_smem_desc(smem_u32(sb) + kk * 2 * lbo_b, lbo_b, sbo, kLayoutNone);
      if (a_mode == 2) {
        umma_f16_ts(tmem_base, tmem_base + a_col0 + kk * 8, db, idesc, kk != 0);
      } else {
        uint64_t da = make_smem_desc(smem_u32(sa) + kk * 2 * lbo_a, lbo_a, sbo, kLayoutNone);
        umma_f16(tmem_base, da, db, idesc, kk != 0);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld<16>(tmem_base + (uint32_t(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[size_t(tid) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, ncols); }
}


// ---- tcgen05.mma issue / completion cost microbenchmark (bring-up): one CTA, 32 unrolled MMAs per repetition ----
// Addresses are immediates and the issue is elect-guarded, like the recurrence kernels.  NACC accumulators are used
// round-robin (NACC = 1: every MMA accumulates into the same D).
template <int NACC, bool TS, int BREUSE = 1>
__device__ __forceinline__ void bench_issue32(uint64_t da0, uint64_t db0, uint32_t idesc, uint32_t lbo_a16, uint32_t lbo_b16) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int kk = (i / BREUSE) & 7;  // BREUSE consecutive MMAs share one B operand
    const uint64_t db = db0 + uint64_t(kk) * (2 * lbo_b16);
    if (TS) umma_f16_ts((i % NACC) * 16, 256 + kk * 8 + (i >> 3) * 64, db, idesc, 1u);
    else umma_f16((i % NACC) * 16, da0 + uint64_t(kk) * (2 * lbo_a16), db, idesc, 1u);
  }
}

__global__ void __launch_bounds__(128, 1) dbg_umma_bench_kernel(long long* __restrict__ out, int M, int N, int n_acc, int a_mode,
                                                               int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sa = base;                    // A: 128 rows x 128 K bf16 (smem mode)
  uint8_t* sb = base + 128 * 128 * 2;    // B: up to 256 rows x 128 K bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 256 * 128 * 2);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 128 * 2 + 256 * 128 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *slot;
  {
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) r[c] = 0x3c003c00u;
    for (int c0 = 0; c0 < 512; c0 += 32) tmem_st32(tmem_base + (uint32_t(warp * 32) << 16) + c0, r);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && tmem_base == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
    const uint32_t lbo_a = (M / 8) * 128, lbo_b = (N / 8) * 128;
    const uint64_t da0 = make_smem_desc(smem_u32(sa), lbo_a, 128, kLayoutNone);
    const uint64_t db0 = make_smem_desc(smem_u32(sb), lbo_b, 128, kLayoutNone);
    uint32_t phase = 0;
    for (int rep = 0; rep < reps; ++rep) {
      long long t0 = 0, t1 = 0;
      if (elect_one()) {
        t0 = clock64();
        if (a_mode == 3) {  // TMEM A, four consecutive MMAs share one B descriptor (the forward kernel's pattern)
          bench_issue32<4, true, 4>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
        } else if (a_mode == 2) {
          if (n_acc == 1) bench_issue32<1, true>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
          else if (n_acc == 4) bench_issue32<4, true>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
          else bench_issue32<8, true>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
        } else {
          if (n_acc == 1) bench_issue32<1, false>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
          else if (n_acc == 4) bench_issue32<4, false>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
          else bench_issue32<8, false>(da0, db0, idesc, lbo_a >> 4, lbo_b >> 4);
        }
        t1 = clock64();
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      t0 = __shfl_sync(0xffffffffu, t0, 0) | 0;  // elected lane is lane 0 of a converged warp
      if (tid == 0 && blockIdx.x == 0) {
        out[rep * 2 + 0] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
  } else if (tid == 0 && tmem_base != 0 && blockIdx.x == 0) {
    out[0] = -1;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace csn

using namespace csn;

extern "C" int csn_dbg_umma_tile(const void* A, const void* B, float* D, int N, int K, int a_mn_major, int b_mn_major,
                                 void* stream) {
  CSN_REQUIRE(A && B && D, "csn_dbg_umma_tile: null pointer");
  CSN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "csn_dbg_umma_tile: bad N/K");
  CSN_REQUIRE(a_mn_major >= 0 && a_mn_major <= 2, "csn_dbg_umma_tile: a_mn_major must be 0 (K-major smem), 1 (MN-major smem) or 2 (tensor memory)");
  const size_t smem = size_t(128) * K * 2 + size_t(N) * K * 2 + 256;
  CSN_CUDA(cudaFuncSetAttribute(dbg_umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dbg_umma_tile_kernel<<<1, 128, smem, as_stream(stream)>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, N, K,
                                                           a_mn_major, b_mn_major);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_dbg_umma_bench(long long* out, int M, int N, int n_acc, int a_mode, int reps, void* stream) {
  static const int grid = [] { const char* e = getenv("CSN_UMMA_BENCH_GRID"); return e ? atoi(e) : 1; }();
  CSN_REQUIRE(out && (M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && (n_acc == 1 || n_acc == 4 || n_acc == 8) &&
                  reps >= 1 && reps <= 16, "csn_dbg_umma_bench: bad arguments");
  const size_t smem = 128 * 128 * 2 + 256 * 128 * 2 + 256;
  CSN_CUDA(cudaFuncSetAttribute(dbg_umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dbg_umma_bench_kernel<<<grid, 128, smem, as_stream(stream)>>>(out, M, N, n_acc, a_mode, reps);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

// ---- per-SM global store bandwidth microbenchmark (bring-up): n_cta CTAs of 256 threads each write `bytes` of their own
// region `reps` times.  mode 0: STG.128, a warp writes 512 contiguous bytes per instruction; mode 1: STG.128, a warp
// writes 4 rows x 128 B at a 2 KB row pitch (the GEMM-epilogue pattern); mode 2: TMA bulk stores of 16 KB from shared
// memory.  out[cta] = cycles of the CTA's loop.
namespace csn {
__global__ void __launch_bounds__(256, 1) dbg_store_bw_kernel(float* __restrict__ dst, long long* __restrict__ out, size_t bytes,
                                                             int mode, int reps) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* base = dst + size_t(blockIdx.x) * (bytes / 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384 / 4; i += 256) reinterpret_cast<float*>(smem_raw)[i] = 1.f;
  fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
      for (size_t o = size_t(tid) * 4; o < bytes / 4; o += 256 * 4) *reinterpret_cast<float4*>(base + o) = v;
    } else if (mode == 1) {
      // rows of 512 floats (2 KB); warp w handles row blocks: per instruction 4 rows x 32 floats
      const size_t n_rows = bytes / 2048;
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
      for (size_t rb = size_t(warp) * 4; rb < n_rows; rb += 8 * 4)
        for (int c = 0; c < 16; ++c)
          *reinterpret_cast<float4*>(base + (rb + (lane >> 3)) * 512 + c * 32 + (lane & 7) * 4) = v;
    } else {
      if (tid == 0) {
        for (size_t o = 0; o < bytes; o += 16384) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(base) + o),
                       "r"(smem_u32(smem_raw)), "r"(16384) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) out[blockIdx.x] = clock64() - t0;
}
}  // namespace csn

extern "C" int csn_dbg_store_bw(float* dst, long long* out, size_t bytes_per_cta, int n_cta, int mode, int reps, void* stream) {
  CSN_REQUIRE(dst && out && bytes_per_cta % 16384 == 0 && n_cta >= 1 && mode >= 0 && mode <= 2 && reps >= 1, "csn_dbg_store_bw: bad arguments");
  csn::dbg_store_bw_kernel<<<n_cta, 256, 16384, as_stream(stream)>>>(dst, out, bytes_per_cta, mode, reps);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
