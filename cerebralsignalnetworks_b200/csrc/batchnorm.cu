// BatchNorm1d over a [M, N] activation matrix: the optional normalisation layers of DINOHead (use_bn=True,
// LstmDistillation.py:72-80).  Training mode normalises with the batch statistics of the M rows (all crops x trials of the
// step) and moves the running statistics (momentum 0.1, unbiased variance, as torch.nn.BatchNorm1d); evaluation mode uses the
// running statistics.  A CTA owns 32 feature columns; its 8 row groups walk the rows with coalesced 128-byte loads, their
// partial sums are folded in row-group order (fixed order: same inputs, same bits).  Mean first, then the variance around
// it (two passes over an activation matrix that sits in L2: no catastrophic cancellation).
#include "common.cuh"

namespace csn {

constexpr int kBnCols = 32, kBnRowGroups = 8;

__device__ __forceinline__ float bn_fold(float v, float (*red)[kBnCols + 1], int rg, int col) {
  red[rg][col] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < kBnRowGroups; ++k) a += red[k][col];
  __syncthreads();
  return a;
}

__global__ void __launch_bounds__(kBnCols* kBnRowGroups) batchnorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
    float* __restrict__ running_var, float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd, int M, int N,
    float eps, float momentum, int training) {
  __shared__ float red[kBnRowGroups][kBnCols + 1];
  const int col = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * kBnCols + col;
  const bool ok = n < N;
  float mean, rstd;
  if (training) {
    float s = 0.f;
    if (ok)
      for (int m = rg; m < M; m += kBnRowGroups) s += x[size_t(m) * N + n];
    mean = bn_fold(s, red, rg, col) / float(M);
    float q = 0.f;
    if (ok)
      for (int m = rg; m < M; m += kBnRowGroups) {
        const float d = x[size_t(m) * N + n] - mean;
        q = fmaf(d, d, q);
      }
    const float var = bn_fold(q, red, rg, col) / float(M);
    rstd = rsqrtf(var + eps);
    if (ok && rg == 0) {
      save_mean[n] = mean;
      save_rstd[n] = rstd;
      if (running_mean) running_mean[n] = (1.f - momentum) * running_mean[n] + momentum * mean;
      if (running_var) running_var[n] = (1.f - momentum) * running_var[n] + momentum * var * (M > 1 ? float(M) / float(M - 1) : 1.f);
    }
  } else {
    mean = ok ? running_mean[n] : 0.f;
    rstd = ok ? rsqrtf(running_var[n] + eps) : 0.f;
    if (ok && rg == 0 && save_mean) {
      save_mean[n] = mean;
      save_rstd[n] = rstd;
    }
  }
  if (!ok) return;
  const float g = gamma ? gamma[n] : 1.f, b = beta ? beta[n] : 0.f;
  const float a = g * rstd, c = b - mean * a;
  for (int m = rg; m < M; m += kBnRowGroups) y[size_t(m) * N + n] = fmaf(x[size_t(m) * N + n], a, c);
}

// dgamma = sum dy xhat, dbeta = sum dy, dx = gamma rstd / M (M dy - dbeta - xhat dgamma)   (training statistics);
// evaluation mode (training == 0): the statistics are constants, dx = gamma rstd dy.
__global__ void __launch_bounds__(kBnCols* kBnRowGroups) batchnorm_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int N,
    int training) {
  __shared__ float red[kBnRowGroups][kBnCols + 1];
  const int col = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * kBnCols + col;
  const bool ok = n < N;
  const float mean = ok ? save_mean[n] : 0.f, rstd = ok ? save_rstd[n] : 0.f;
  float sb = 0.f, sg = 0.f;
  if (ok)
    for (int m = rg; m < M; m += kBnRowGroups) {
      const float d = dy[size_t(m) * N + n];
      sb += d;
      sg = fmaf(d, (x[size_t(m) * N + n] - mean) * rstd, sg);
    }
  const float db = bn_fold(sb, red, rg, col), dg = bn_fold(sg, red, rg, col);
  if (!ok) return;
  if (rg == 0) {
    if (dgamma) dgamma[n] = dg;
    if (dbeta) dbeta[n] = db;
  }
  if (!dx) return;
  const float g = gamma ? gamma[n] : 1.f;
  if (training) {
    const float k = g * rstd / float(M);
    for (int m = rg; m < M; m += kBnRowGroups) {
      const float xh = (x[size_t(m) * N + n] - mean) * rstd;
      dx[size_t(m) * N + n] = k * (float(M) * dy[size_t(m) * N + n] - db - xh * dg);
    }
  } else {
    const float k = g * rstd;
    for (int m = rg; m < M; m += kBnRowGroups) dx[size_t(m) * N + n] = k * dy[size_t(m) * N + n];
  }
}

}  // namespace csn

using namespace csn;

extern "C" int csn_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 float* y, float* save_mean, float* save_rstd, int M, int N, float eps, float momentum,
                                 int training, void* stream) {
  CSN_REQUIRE(x && y, "csn_batchnorm_fwd: null pointer");
  CSN_REQUIRE(M >= 1 && N >= 1, "csn_batchnorm_fwd: bad shape (M=%d N=%d)", M, N);
  CSN_REQUIRE(training ? (save_mean && save_rstd) : (running_mean && running_var),
              "csn_batchnorm_fwd: training needs save_mean / save_rstd, evaluation needs the running statistics");
  CSN_REQUIRE(!training || M > 1, "csn_batchnorm_fwd: training statistics need more than one row (torch raises too)");
  batchnorm_fwd_kernel<<<ceil_div(N, kBnCols), kBnCols * kBnRowGroups, 0, as_stream(stream)>>>(
      x, gamma, beta, running_mean, running_var, y, save_mean, save_rstd, M, N, eps, momentum, training);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_batchnorm_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean,
                                 const float* save_rstd, float* dx, float* dgamma, float* dbeta, int M, int N, int training,
                                 void* stream) {
  CSN_REQUIRE(x && dy && save_mean && save_rstd, "csn_batchnorm_bwd: null pointer");
  CSN_REQUIRE(M >= 1 && N >= 1, "csn_batchnorm_bwd: bad shape (M=%d N=%d)", M, N);
  batchnorm_bwd_kernel<<<ceil_div(N, kBnCols), kBnCols * kBnRowGroups, 0, as_stream(stream)>>>(x, dy, gamma, save_mean, save_rstd, dx,
                                                                                            dgamma, dbeta, M, N, training);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
