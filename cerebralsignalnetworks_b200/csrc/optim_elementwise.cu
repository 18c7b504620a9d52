// Fused Adam/AdamW over a flat fp32 buffer, activations, row L2-normalise, weight-norm, column sums.
// All HBM-streaming kernels: float4 where alignment allows, grid sized in multiples of the SM count.
#include "common.cuh"

namespace csn {

// torch.optim.Adam / AdamW single-tensor semantics (LSTMDistill.py:322, LstmDistillation.py:469-471):
//   g = grad*grad_scale (+ wd*p if coupled);  p *= 1 - lr*wd if decoupled
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            int decoupled, float step_size, float inv_sqrt_bc2, float grad_scale,
                            const float* __restrict__ dev_consts) {
  if (dev_consts) {  // CUDA-graph path: bias corrections computed on the device from the device-side step counter
    step_size = dev_consts[0];
    inv_sqrt_bc2 = dev_consts[1];
  }
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t n4 = n >> 2;
  for (; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = gg[j] * grad_scale;
      if (decoupled) pp[j] *= (1.f - lr * wd); else gr = fmaf(wd, pp[j], gr);
      mm[j] = b1 * mm[j] + (1.f - b1) * gr;
      vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
      float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  // tail
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    size_t j = (n4 << 2) + threadIdx.x;
    float gr = g[j] * grad_scale;
    float pj = p[j];
    if (decoupled) pj *= (1.f - lr * wd); else gr = fmaf(wd, pj, gr);
    float mj = b1 * m[j] + (1.f - b1) * gr;
    float vj = b2 * v[j] + (1.f - b2) * gr * gr;
    pj -= step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
    p[j] = pj; m[j] = mj; v[j] = vj;
  }
}

// ++step; consts = { lr / (1 - b1^step), 1 / sqrt(1 - b2^step) }
__global__ void adam_advance_kernel(int* __restrict__ step_counter, float* __restrict__ consts, float lr, float b1, float b2) {
  const int s = ++(*step_counter);
  const double bc1 = 1.0 - pow((double)b1, (double)s), bc2 = 1.0 - pow((double)b2, (double)s);
  consts[0] = (float)((double)lr / bc1);
  consts[1] = (float)(1.0 / sqrt(bc2));
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int act) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i];
    y[i] = act == CSN_ACT_RELU ? fmaxf(v, 0.f) : (act == CSN_ACT_GELU ? gelu_f(v) : v);
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                               size_t n, int act) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i];
    float d = act == CSN_ACT_RELU ? (v > 0.f ? 1.f : 0.f) : (act == CSN_ACT_GELU ? gelu_grad_f(v) : 1.f);
    dx[i] = dy[i] * d;
  }
}

// out[n] (+)= sum_m x[m,n] in a FIXED order (no atomics: the same input gives the same bits every run).  Block = 32
// columns x RY row-lanes; row-lane y adds rows y, y + RY, ... into four rotating accumulators (independent loads in
// flight), the lanes are folded in shared memory in lane order.  One CTA owns its 32 columns for all rows.
template <int RY>
__global__ void __launch_bounds__(32 * RY) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N,
                                                        int ldx, int accumulate) {
  __shared__ float red[RY][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (col < N) {
    const float* src = x + col;
    int m = threadIdx.y;
    for (; m + 3 * RY < M; m += 4 * RY) {
      a0 += src[size_t(m) * ldx];
      a1 += src[size_t(m + RY) * ldx];
      a2 += src[size_t(m + 2 * RY) * ldx];
      a3 += src[size_t(m + 3 * RY) * ldx];
    }
    for (; m < M; m += RY) a0 += src[size_t(m) * ldx];
  }
  red[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < RY; ++j) s += red[j][threadIdx.x];
    out[col] = accumulate ? out[col] + s : s;
  }
}

// one warp per row
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int M, int N) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + size_t(row) * N;
  float ss = 0.f;
  for (int k = lane; k < N; k += 32) ss = fmaf(xr[k], xr[k], ss);
  ss = warp_sum(ss);
  float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int k = lane; k < N; k += 32) y[size_t(row) * N + k] = xr[k] * inv;
  if (lane == 0) inv_norm[row] = inv;
}
// dx = inv * (dy - y * <dy, y>)
__global__ void l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm,
                                  const float* __restrict__ dy, float* __restrict__ dx, int M, int N) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* yr = y + size_t(row) * N;
  const float* gr = dy + size_t(row) * N;
  float dot = 0.f;
  for (int k = lane; k < N; k += 32) dot = fmaf(gr[k], yr[k], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int k = lane; k < N; k += 32) dx[size_t(row) * N + k] = inv * (gr[k] - yr[k] * dot);
}

__global__ void weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w,
                                       float* __restrict__ inv_norm, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* vr = v + size_t(row) * K;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) ss = fmaf(vr[k], vr[k], ss);
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss);
  const float sc = g[row] * inv;
  for (int k = lane; k < K; k += 32) w[size_t(row) * K + k] = vr[k] * sc;
  if (lane == 0) inv_norm[row] = inv;
}
__global__ void weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                       const float* __restrict__ inv_norm, const float* __restrict__ dw,
                                       float* __restrict__ dv, float* __restrict__ dg, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* vr = v + size_t(row) * K;
  const float* dr = dw + size_t(row) * K;
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) dot = fmaf(dr[k], vr[k], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row], gg = g[row];
  const float c = dot * inv * inv;
  for (int k = lane; k < K; k += 32) dv[size_t(row) * K + k] = gg * inv * (dr[k] - vr[k] * c);
  if (dg && lane == 0) dg[row] = dot * inv;
}

__global__ void scale_kernel(float* __restrict__ x, size_t n, const float* __restrict__ scale_dev, float scale_host) {
  const float sc = (scale_dev ? *scale_dev : 1.f) * scale_host;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) x[i] *= sc;
}

// teacher <- m * teacher + (1 - m) * student   (EMA teacher, LstmDistillation.py:616-619)
__global__ void ema_update_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n, float m) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = dst[i] * m + src[i] * (1.f - m);
}

// Per-parameter gradient clipping (utils/utils.py:132-141: clip_coef = clip / (||g_p|| + 1e-6), applied when < 1) over a
// flat gradient buffer cut into segments [seg_off[i], seg_off[i+1]); no host synchronisation.
__device__ __forceinline__ int find_segment(const long long* __restrict__ seg_off, int n_seg, long long idx) {
  int lo = 0, hi = n_seg;  // invariant: seg_off[lo] <= idx < seg_off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_off[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}
constexpr int kClipChunk = 2048;
__global__ void seg_sumsq_kernel(const float* __restrict__ g, const long long* __restrict__ seg_off, int n_seg,
                                 float* __restrict__ chunk_sumsq) {
  // each block handles one chunk of kClipChunk elements that never straddles a segment (chunks are per segment)
  __shared__ float red[8];
  long long chunk = blockIdx.x;
  // locate the segment of this chunk: chunk_off[i] = number of chunks before segment i, computed on the fly
  int seg = 0;
  long long first = 0;
  for (; seg < n_seg; ++seg) {
    long long len = seg_off[seg + 1] - seg_off[seg];
    long long nc = (len + kClipChunk - 1) / kClipChunk;
    if (chunk < first + nc) break;
    first += nc;
  }
  if (seg >= n_seg) {
    if (threadIdx.x == 0) chunk_sumsq[chunk] = 0.f;
    return;
  }
  const long long beg = seg_off[seg] + (chunk - first) * kClipChunk;
  const long long end = min(beg + (long long)kClipChunk, seg_off[seg + 1]);
  float acc = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    chunk_sumsq[chunk] = s;  // folded per segment, in chunk order, by seg_fold_kernel (no atomics)
  }
}
// one warp per segment: sumsq[seg] = sum of its chunks' partials, lane-strided then a fixed shuffle tree
__global__ void seg_fold_kernel(const float* __restrict__ chunk_sumsq, const long long* __restrict__ seg_off, int n_seg,
                                float* __restrict__ sumsq) {
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (seg >= n_seg) return;
  long long first = 0;
  for (int i = 0; i < seg; ++i) first += (seg_off[i + 1] - seg_off[i] + kClipChunk - 1) / kClipChunk;
  const long long nc = (seg_off[seg + 1] - seg_off[seg] + kClipChunk - 1) / kClipChunk;
  float a = 0.f;
  for (long long c = lane; c < nc; c += 32) a += chunk_sumsq[first + c];
  a = warp_sum(a);
  if (lane == 0) sumsq[seg] = a;
}
__global__ void seg_clip_kernel(float* __restrict__ g, const long long* __restrict__ seg_off, int n_seg,
                                const float* __restrict__ sumsq, float clip) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = seg_off[n_seg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int seg = find_segment(seg_off, n_seg, i);
    const float coef = clip / (sqrtf(sumsq[seg]) + 1e-6f);
    if (coef < 1.f) g[i] *= coef;
  }
}

static inline int stream_blocks(size_t n, int per_block) {
  size_t b = ceil_div<size_t>(n, per_block);
  size_t cap = size_t(sm_count()) * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace csn

using namespace csn;

extern "C" int csn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int decoupled, int step,
                             float grad_scale, void* stream) {
  CSN_REQUIRE(params && grads && exp_avg && exp_avg_sq, "csn_adam_step: null pointer");
  CSN_REQUIRE(step >= 1, "csn_adam_step: step is 1-based");
  CSN_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
              "csn_adam_step: buffers must be 16-byte aligned");
  if (n == 0) return CSN_OK;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  adam_kernel<<<stream_blocks(n, 256 * 4), 256, 0, as_stream(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, decoupled, (float)(lr / bc1),
      (float)(1.0 / sqrt(bc2)), grad_scale, nullptr);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_adam_step_graph(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int decoupled,
                                   int* step_counter, float* consts2, float grad_scale, void* stream) {
  CSN_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_counter && consts2, "csn_adam_step_graph: null pointer");
  CSN_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
              "csn_adam_step_graph: buffers must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  adam_advance_kernel<<<1, 1, 0, s>>>(step_counter, consts2, lr, beta1, beta2);
  CSN_LAUNCH_CHECK();
  if (n == 0) return CSN_OK;
  adam_kernel<<<stream_blocks(n, 256 * 4), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                        weight_decay, decoupled, 0.f, 0.f, grad_scale, consts2);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_scale_f32(float* x, size_t n, const float* scale_dev, float scale_host, void* stream) {
  CSN_REQUIRE(x, "csn_scale_f32: null pointer");
  if (n == 0) return CSN_OK;
  scale_kernel<<<stream_blocks(n, 256), 256, 0, as_stream(stream)>>>(x, n, scale_dev, scale_host);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_ema_update(float* dst, const float* src, size_t n, float momentum, void* stream) {
  CSN_REQUIRE(dst && src, "csn_ema_update: null pointer");
  if (n == 0) return CSN_OK;
  ema_update_kernel<<<stream_blocks(n, 256), 256, 0, as_stream(stream)>>>(dst, src, n, momentum);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_clip_grad_segments(float* grads, const long long* seg_off_dev, int n_seg, long long n_chunks,
                                      float* sumsq_dev, float clip, void* stream) {
  CSN_REQUIRE(grads && seg_off_dev && sumsq_dev && n_seg >= 1 && n_chunks >= 1 && clip > 0.f, "csn_clip_grad_segments: bad arguments");
  cudaStream_t s = as_stream(stream);
  seg_sumsq_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(grads, seg_off_dev, n_seg, sumsq_dev + n_seg);
  CSN_LAUNCH_CHECK();
  seg_fold_kernel<<<ceil_div(n_seg, 8), 256, 0, s>>>(sumsq_dev + n_seg, seg_off_dev, n_seg, sumsq_dev);
  CSN_LAUNCH_CHECK();
  seg_clip_kernel<<<sm_count() * 8, 256, 0, s>>>(grads, seg_off_dev, n_seg, sumsq_dev, clip);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_act_fwd(const float* x, float* y, size_t n, int act, void* stream) {
  CSN_REQUIRE(x && y, "csn_act_fwd: null pointer");
  if (n == 0) return CSN_OK;
  act_fwd_kernel<<<stream_blocks(n, 256), 256, 0, as_stream(stream)>>>(x, y, n, act);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
extern "C" int csn_act_bwd(const float* x, const float* dy, float* dx, size_t n, int act, void* stream) {
  CSN_REQUIRE(x && dy && dx, "csn_act_bwd: null pointer");
  if (n == 0) return CSN_OK;
  act_bwd_kernel<<<stream_blocks(n, 256), 256, 0, as_stream(stream)>>>(x, dy, dx, n, act);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

int csn::colsum_det(const float* x, float* out, int M, int N, int ldx, int accumulate, cudaStream_t s) {
  if (M == 0) {
    if (!accumulate) CSN_CUDA(cudaMemsetAsync(out, 0, size_t(N) * 4, s));
    return CSN_OK;
  }
  const int col_blocks = ceil_div(N, 32);
  // the row-lane count depends on the shape only, so the summation order is a function of (M, N) alone
  if (M >= 512 && col_blocks < 4 * sm_count()) colsum_kernel<32><<<col_blocks, dim3(32, 32), 0, s>>>(x, out, M, N, ldx, accumulate);
  else colsum_kernel<8><<<col_blocks, dim3(32, 8), 0, s>>>(x, out, M, N, ldx, accumulate);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_colsum_f32(const float* x, float* out, int M, int N, int ldx, int accumulate, void* stream) {
  CSN_REQUIRE(x && out && M >= 0 && N >= 1 && ldx >= N, "csn_colsum_f32: bad arguments");
  return colsum_det(x, out, M, N, ldx, accumulate, as_stream(stream));
}

extern "C" int csn_l2norm_fwd(const float* x, float* y, float* inv_norm, int M, int N, void* stream) {
  CSN_REQUIRE(x && y && inv_norm && M >= 1 && N >= 1, "csn_l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<ceil_div(M, 8), 256, 0, as_stream(stream)>>>(x, y, inv_norm, M, N);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
extern "C" int csn_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int M, int N, void* stream) {
  CSN_REQUIRE(y && inv_norm && dy && dx && M >= 1 && N >= 1, "csn_l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<ceil_div(M, 8), 256, 0, as_stream(stream)>>>(y, inv_norm, dy, dx, M, N);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
extern "C" int csn_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm, int N, int K, void* stream) {
  CSN_REQUIRE(v && g && w && inv_norm && N >= 1 && K >= 1, "csn_weight_norm_fwd: bad arguments");
  weight_norm_fwd_kernel<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(v, g, w, inv_norm, N, K);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
extern "C" int csn_weight_norm_bwd(const float* v, const float* g, const float* inv_norm, const float* dw, float* dv,
                                   float* dg, int N, int K, void* stream) {
  CSN_REQUIRE(v && g && inv_norm && dw && dv && N >= 1 && K >= 1, "csn_weight_norm_bwd: bad arguments");
  weight_norm_bwd_kernel<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(v, g, inv_norm, dw, dv, dg, N, K);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
