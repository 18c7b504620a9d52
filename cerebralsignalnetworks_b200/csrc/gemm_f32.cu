// fp32 SIMT GEMM with optional transposes, bias and activation epilogue.  This is the tight-tolerance parity
// path (compute_dtype f32) for nn.Linear / input projections / dW; the throughput path is gemm_tc.cu (tcgen05).
#include "common.cuh"

namespace csn {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == CSN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CSN_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  return v;
}

// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta*C + bias
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A,
                                                      int lda, const float* __restrict__ B, int ldb, float beta,
                                                      float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                                      int act) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < K) v = TA ? A[size_t(gk) * lda + gm] : A[size_t(gm) * lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      int e = tid + i * 256;
      int nn, kk;
      if (TB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < K) v = TB ? B[size_t(gn) * ldb + gk] : B[size_t(gk) * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
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
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  CSN_REQUIRE(grid.y <= 65535, "csn_gemm_f32: M too large (%d)", M);
  cudaStream_t s = as_stream(stream);
  if (!transA && !transB) gemm_f32_kernel<false, false><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
  else if (!transA && transB) gemm_f32_kernel<false, true><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
  else if (transA && !transB) gemm_f32_kernel<true, false><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
  else gemm_f32_kernel<true, true><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
