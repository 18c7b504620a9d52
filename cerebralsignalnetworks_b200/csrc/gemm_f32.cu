// fp32 SIMT GEMM with optional transposes, bias and activation epilogue.  This is the tight-tolerance parity
// path (compute_dtype f32) for nn.Linear / input projections / dW; the throughput path is gemm_tc.cu (tcgen05).
#include "common.cuh"

namespace csn {

constexpr int BK = 16;  // tiles are (16*TM) x (16*TM): TM=4 -> 64x64 (large problems), TM=2 -> 32x32 (small, to fill the SMs)

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == CSN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CSN_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  return v;
}

// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta*C + bias
template <bool TA, bool TB, int TM>
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A,
                                                      int lda, const float* __restrict__ B, int ldb, float beta,
                                                      float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                                      int act) {
  constexpr int BM = 16 * TM, BN = 16 * TM;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TM] = {};

  // global -> registers for tile k0 (issued one tile ahead of its use: the loads of tile k+1 fly while tile k is
  // multiplied out of shared memory)
  constexpr int NA = (BM * BK) / 256, NB = (BN * BK) / 256;
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
      int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < K) ? (TA ? A[size_t(gk) * lda + gm] : A[size_t(gm) * lda + gk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int e = tid + i * 256;
      int nn, kk;
      if (TB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int gn = n0 + nn, gk = k0 + kk;
      rb[i] = (gn < N && gk < K) ? (TB ? B[size_t(gn) * ldb + gk] : B[size_t(gk) * ldb + gn]) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int e = tid + i * 256;
      int nn, kk;
      if (TB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      Bs[kk][nn] = rb[i];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) { av[i] = As[kk][ty * TM + i]; bv[i] = Bs[kk][tx * TM + i]; }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      int gn = n0 + tx * TM + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v = fmaf(beta, C[size_t(gm) * ldc + gn], v);
      if (bias) v += bias[gn];
      C[size_t(gm) * ldc + gn] = act_apply(v, act);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Small problems with a short contraction (the projection head of the distill step: 256 x 384 x 128 and its two
// gradients): the tiled kernel above is latency bound there -- K/16 dependent global-load round trips per CTA.  Here a
// CTA owns a 16 x 64 output tile and pulls its WHOLE K strip of both operands into shared memory with one burst of
// 16-byte cp.async (one HBM/L2 round trip in total), multiplies out of shared memory, and can emit the row sums of
// op(A) on the side (bias gradient of the dW product).  Supported operand forms: NT, NN, TN.
constexpr int RBM = 16, RBN = 64;

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_f32_resident_kernel(int M, int N, int K, float alpha,
                                                               const float* __restrict__ A, int lda,
                                                               const float* __restrict__ B, int ldb, float beta,
                                                               float* __restrict__ C, int ldc,
                                                               const float* __restrict__ bias, int act,
                                                               float* __restrict__ rowsum, int rowsum_accumulate) {
  static_assert(!(TA && TB), "TT form is not built");
  extern __shared__ __align__(16) float smem_g[];
  // As: TA ? [K][RBM] : [RBM][K+4];   Bs: TB ? [RBN][K+4] : [K][RBN]
  const int a_stride = TA ? RBM : K + 4, b_stride = TB ? K + 4 : RBN;
  float* As = smem_g;
  float* Bs = smem_g + (TA ? K * RBM : RBM * (K + 4));
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * RBM, n0 = blockIdx.x * RBN;

  // ---- one burst of loads: every 16-byte chunk of both strips ----
  if (TA) {  // K rows of RBM floats
    for (int ch = tid; ch < K * (RBM / 4); ch += 256) {
      const int k = ch / (RBM / 4), c = ch % (RBM / 4);
      const int gm = m0 + c * 4;
      const int nb = max(0, min(4, M - gm)) * 4;
      cp_async16_zfill(As + k * RBM + c * 4, nb ? A + size_t(k) * lda + gm : A, nb);
    }
  } else {   // RBM rows of K floats
    const int kc = K / 4;
    for (int ch = tid; ch < RBM * kc; ch += 256) {
      const int r = ch / kc, c = ch % kc;
      const bool ok = m0 + r < M;
      cp_async16_zfill(As + r * a_stride + c * 4, ok ? A + size_t(m0 + r) * lda + c * 4 : A, ok ? 16 : 0);
    }
  }
  if (TB) {  // RBN rows of K floats
    const int kc = K / 4;
    for (int ch = tid; ch < RBN * kc; ch += 256) {
      const int r = ch / kc, c = ch % kc;
      const bool ok = n0 + r < N;
      cp_async16_zfill(Bs + r * b_stride + c * 4, ok ? B + size_t(n0 + r) * ldb + c * 4 : B, ok ? 16 : 0);
    }
  } else {   // K rows of RBN floats
    for (int ch = tid; ch < K * (RBN / 4); ch += 256) {
      const int k = ch / (RBN / 4), c = ch % (RBN / 4);
      const int gn = n0 + c * 4;
      const int nb = max(0, min(4, N - gn)) * 4;
      cp_async16_zfill(Bs + k * RBN + c * 4, nb ? B + size_t(k) * ldb + gn : B, nb);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  const int r = tid >> 4, tx = tid & 15;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (TB) {
    // thread owns row r, columns tx + 16 j: both strips are k-contiguous -> float4 over k (conflict-free: the
    // row stride K+4 floats steps one 16-byte bank group per row)
    const float* ap = As + r * a_stride;
    const float* bp = Bs + tx * b_stride;
    for (int k = 0; k < K; k += 4) {
      const float4 a = *reinterpret_cast<const float4*>(ap + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(bp + j * 16 * b_stride + k);
        acc[j] = fmaf(a.x, b.x, acc[j]);
        acc[j] = fmaf(a.y, b.y, acc[j]);
        acc[j] = fmaf(a.z, b.z, acc[j]);
        acc[j] = fmaf(a.w, b.w, acc[j]);
      }
    }
  } else {
    // thread owns row r, columns 4 tx .. 4 tx + 3: B strip is k-major -> one float4 of B per k
    const float* bp = Bs + tx * 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float a = TA ? As[k * RBM + r] : As[r * a_stride + k];
      const float4 b = *reinterpret_cast<const float4*>(bp + k * RBN);
      acc[0] = fmaf(a, b.x, acc[0]);
      acc[1] = fmaf(a, b.y, acc[1]);
      acc[2] = fmaf(a, b.z, acc[2]);
      acc[3] = fmaf(a, b.w, acc[3]);
    }
  }
  const int gm = m0 + r;
  if (gm < M) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + (TB ? tx + 16 * j : tx * 4 + j);
      if (gn >= N) continue;
      float v = alpha * acc[j];
      if (beta != 0.f) v = fmaf(beta, C[size_t(gm) * ldc + gn], v);
      if (bias) v += bias[gn];
      C[size_t(gm) * ldc + gn] = act_apply(v, act);
    }
  }
  // row sums of op(A) (one column of CTAs): 16 partial sums per row, folded with shuffles inside a half warp
  if (rowsum && blockIdx.x == 0) {
    float sum = 0.f;
    for (int k = tx; k < K; k += 16) sum += TA ? As[k * RBM + r] : As[r * a_stride + k];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (tx == 0 && gm < M) rowsum[gm] = rowsum_accumulate ? rowsum[gm] + sum : sum;
  }
}

template <bool TA, bool TB>
static int launch_resident(int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb, float beta,
                           float* C, int ldc, const float* bias, int act, float* rowsum, int rowsum_accumulate,
                           cudaStream_t s) {
  const size_t smem = (size_t(TA ? K * RBM : RBM * (K + 4)) + size_t(TB ? RBN * (K + 4) : K * RBN)) * 4;
  auto kern = gemm_f32_resident_kernel<TA, TB>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid(ceil_div(N, RBN), ceil_div(M, RBM));
  kern<<<grid, 256, smem, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, rowsum, rowsum_accumulate);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

// the resident kernel applies when the whole K strip fits and every 16-byte chunk is aligned
static bool resident_ok(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb) {
  static const bool off = [] { const char* e = getenv("CSN_GEMM_NO_RESIDENT"); return e && e[0] == '1'; }();
  if (off || (transA && transB) || K < 4 || K > 512 || K % 4 != 0) return false;
  if (lda % 4 != 0 || ldb % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return false;
  if (transA && M % 4 != 0) return false;    // partial chunks would straddle a row end that is not 16-byte aligned
  if (!transB && N % 4 != 0) return false;
  const long long tiles = (long long)ceil_div(M, RBM) * ceil_div(N, RBN);
  return tiles <= 8LL * sm_count() && ceil_div(M, RBM) <= 65535;
}

}  // namespace csn

using namespace csn;

extern "C" int csn_gemm_f32_rowsum(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                                   const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int act,
                                   float* rowsum, int rowsum_accumulate, void* stream) {
  CSN_REQUIRE(A && B && C, "csn_gemm_f32_rowsum: null pointer");
  CSN_REQUIRE(M >= 0 && N >= 0 && K >= 0, "csn_gemm_f32_rowsum: negative dimension");
  if (M == 0 || N == 0) return CSN_OK;
  cudaStream_t s = as_stream(stream);
  if (resident_ok(transA, transB, M, N, K, A, lda, B, ldb)) {
    if (!transA && transB) return launch_resident<false, true>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, rowsum, rowsum_accumulate, s);
    if (!transA && !transB) return launch_resident<false, false>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, rowsum, rowsum_accumulate, s);
    return launch_resident<true, false>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, rowsum, rowsum_accumulate, s);
  }
  CSN_TRY(csn_gemm_f32(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, stream));
  if (rowsum) {
    // row sums of op(A): column sums of the stored matrix when transA, else row sums (= column sums of A^T)
    if (transA) return csn_colsum_f32(A, rowsum, K, M, lda, rowsum_accumulate, stream);
    set_error("csn_gemm_f32_rowsum: row sums of a non-transposed A need the resident path (K <= 512, aligned operands)");
    return CSN_EUNSUPPORTED;
  }
  return CSN_OK;
}

extern "C" int csn_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                            const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int act,
                            void* stream) {
  CSN_REQUIRE(A && B && C, "csn_gemm_f32: null pointer");
  CSN_REQUIRE(M >= 0 && N >= 0 && K >= 0, "csn_gemm_f32: negative dimension");
  if (M == 0 || N == 0) return CSN_OK;
  cudaStream_t s = as_stream(stream);
  if (resident_ok(transA, transB, M, N, K, A, lda, B, ldb)) {
    if (!transA && transB) return launch_resident<false, true>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, nullptr, 0, s);
    if (!transA && !transB) return launch_resident<false, false>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, nullptr, 0, s);
    return launch_resident<true, false>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, nullptr, 0, s);
  }
  // small outputs: 32x32 tiles so that e.g. the 256x384 projection runs on 96 CTAs instead of 24
  const bool small = ceil_div(N, 64) * ceil_div(M, 64) < 2 * sm_count();
  const int tile = small ? 32 : 64;
  dim3 grid(ceil_div(N, tile), ceil_div(M, tile));
  CSN_REQUIRE(grid.y <= 65535, "csn_gemm_f32: M too large (%d)", M);
#define CSN_GEMM_LAUNCH(TA_, TB_)                                                                                        \
  do {                                                                                                                   \
    if (small) gemm_f32_kernel<TA_, TB_, 2><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act); \
    else gemm_f32_kernel<TA_, TB_, 4><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);       \
  } while (0)
  if (!transA && !transB) CSN_GEMM_LAUNCH(false, false);
  else if (!transA && transB) CSN_GEMM_LAUNCH(false, true);
  else if (transA && !transB) CSN_GEMM_LAUNCH(true, false);
  else CSN_GEMM_LAUNCH(true, true);
#undef CSN_GEMM_LAUNCH
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
