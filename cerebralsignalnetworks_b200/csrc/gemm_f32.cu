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

}  // namespace csn

using namespace csn;

extern "C" int csn_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                            const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int act,
                            void* stream) {
  CSN_REQUIRE(A && B && C, "csn_gemm_f32: null pointer");
  CSN_REQUIRE(M >= 0 && N >= 0 && K >= 0, "csn_gemm_f32: negative dimension");
  if (M == 0 || N == 0) return CSN_OK;
  cudaStream_t s = as_stream(stream);
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
