// The distillation losses LstmDistillFromDinoV2Train.py actually calls today, forward +
// backward in one pass, one WARP per batch row (K = 384 / 768 feature targets, 40 classes):
//   * FeatureDistributionLoss (LstmDistillFromDinoV2Train.py:107-140, live at :371):
//       term1 = alpha * cross_entropy(pred_label, label)
//       term2 = beta  * F.cross_entropy(softmax(teacher / T), softmax(student / T))
//     i.e. the teacher PROBABILITIES are used as logits (log-softmaxed once more) and the student probabilities are the
//     soft target: term2 = -mean_b sum_k p_s[b,k] * log_softmax(q_t[b,:])[k].  Reproduced as written.
//   * CosineSimilarityLoss (LstmDistillFromDinoV2Train.py:36-43): 1 - mean_b cos(student_b, teacher_b)
//     (nn.CosineSimilarity: dim = 1, eps = 1e-8, each norm clamped).
//   * loss_fn_kd, the knowledge-distillation family of the older training scripts:
//       LSTMDistillRetreival.py:40-70            w_soft * T^2/B * sum q (log q - log p) + w_ce * smooth_l1(student, teacher)
//       LstmDistillFromDinoV2TrainSpampinato.py:107-121   nn.KLDivLoss()(log p, q) * alpha T^2 + CE(student, label) * (1 - alpha)
//     with q = softmax(teacher / T), p = softmax(student / T): one kernel, three coefficients chosen by the host.
#include "common.cuh"

namespace csn {

constexpr int kLossWarps = 8;

__device__ __forceinline__ float ex2f_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kL2e = 1.4426950408889634f;

// block-level fold of one float per warp (warp order), then the deterministic cross-CTA fold into *dst (common.cuh)
__device__ __forceinline__ void block_det_sum(float v, float* dst, float* scratch, void* ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < kLossWarps; ++w) t += scratch[w];
  det_cta_sum(t, ws, blockIdx.x, gridDim.x, dst);
}

template <int EPT>
__global__ void __launch_bounds__(kLossWarps * 32) feature_dist_loss_kernel(
    const float* __restrict__ student, const float* __restrict__ teacher, const float* __restrict__ pred,
    const long long* __restrict__ label, float* __restrict__ loss, float* __restrict__ d_student,
    float* __restrict__ d_pred, int B, int K, int n_classes, float inv_T, float alpha, float beta, float grad_scale,
    void* __restrict__ ws) {
  __shared__ float scratch[kLossWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kLossWarps + warp;
  float term2_row = 0.f;   // warp-uniform
  float term1_lane = 0.f;  // non-zero in the label's lane only
  if (b < B) {
    // ---- term2: q = softmax(t/T); c = -(q - logsumexp(q)); p = softmax(s/T); loss = sum p c; ds = p (c - <p,c>) / T ----
    float q[EPT], p[EPT];
    float mt = -INFINITY, ms = -INFINITY;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      q[i] = (k < K) ? teacher[size_t(b) * K + k] * inv_T : -INFINITY;
      p[i] = (k < K) ? student[size_t(b) * K + k] * inv_T : -INFINITY;
      mt = fmaxf(mt, q[i]);
      ms = fmaxf(ms, p[i]);
    }
    mt = warp_max(mt);
    ms = warp_max(ms);
    float zt = 0.f, zs = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      q[i] = ex2f_((q[i] - mt) * kL2e);
      p[i] = ex2f_((p[i] - ms) * kL2e);
      zt += q[i];
      zs += p[i];
    }
    const float izt = 1.f / warp_sum(zt), izs = 1.f / warp_sum(zs);
    float se = 0.f;  // sum_k exp(q_k) over the VALID columns (q <= 1: no max subtraction needed)
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      q[i] *= izt;
      p[i] *= izs;
      if (lane + 32 * i < K) se += __expf(q[i]);
    }
    const float lse = __logf(warp_sum(se));
    // row loss = sum_k p_k (lse - q_k) = lse - <p, q>   (sum p = 1);  d/ds_j = p_j (<p, q> - q_j) / T: written without
    // lse so that the difference is taken between numbers in [0, 1], not between numbers near log(K)
    float pq = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) pq = fmaf(p[i], q[i], pq);
    pq = warp_sum(pq);
    term2_row = beta * (lse - pq) / B;
    const float gs = grad_scale * beta * inv_T / B;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      if (k < K) d_student[size_t(b) * K + k] = gs * p[i] * (pq - q[i]);
    }
    // ---- term1: alpha * CE(pred, label), classes strided over the lanes ----
    if (pred) {
      const long long y = label[b];
      float m = -INFINITY;
      for (int c = lane; c < n_classes; c += 32) m = fmaxf(m, pred[size_t(b) * n_classes + c]);
      m = warp_max(m);
      float z = 0.f;
      for (int c = lane; c < n_classes; c += 32) z += __expf(pred[size_t(b) * n_classes + c] - m);
      z = warp_sum(z);
      const float logz = m + __logf(z);
      const float g1 = grad_scale * alpha / B;
      for (int c = lane; c < n_classes; c += 32) {
        const float v = pred[size_t(b) * n_classes + c];
        d_pred[size_t(b) * n_classes + c] = g1 * (__expf(v - logz) - (c == y ? 1.f : 0.f));
        if (c == y) term1_lane = alpha * (logz - v) / B;  // exactly one lane
      }
    }
  }
  float lane_part = term1_lane + (lane == 0 ? term2_row : 0.f);
  lane_part = warp_sum(lane_part);
  block_det_sum(lane_part, loss, scratch, ws);
}

template <int EPT>
__global__ void __launch_bounds__(kLossWarps * 32) cosine_loss_kernel(const float* __restrict__ student,
                                                                    const float* __restrict__ teacher,
                                                                    float* __restrict__ loss, float* __restrict__ d_student,
                                                                    int B, int K, float eps, float grad_scale,
                                                                    void* __restrict__ ws) {
  __shared__ float scratch[kLossWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kLossWarps + warp;
  float contrib = 0.f;
  if (b < B) {
    float s[EPT], t[EPT];
    float ss = 0.f, tt = 0.f, st = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      s[i] = (k < K) ? student[size_t(b) * K + k] : 0.f;
      t[i] = (k < K) ? teacher[size_t(b) * K + k] : 0.f;
      ss = fmaf(s[i], s[i], ss);
      tt = fmaf(t[i], t[i], tt);
      st = fmaf(s[i], t[i], st);
    }
    ss = warp_sum(ss); tt = warp_sum(tt); st = warp_sum(st);
    const float ns = fmaxf(sqrtf(ss), eps), nt = fmaxf(sqrtf(tt), eps);
    const float cosv = st / (ns * nt);
    contrib = (lane == 0) ? -cosv / B : 0.f;
    if (b == 0 && lane == 0) contrib += 1.f;  // loss = 1 - mean cos
    // d(-cos/B)/ds = -(1/B) (t / (ns nt) - cos s / ns^2)   (norm clamp inactive unless ||s|| < eps)
    const float a = -grad_scale / (B * ns * nt), c2 = grad_scale * cosv / (B * ns * ns);
    const bool clamped = sqrtf(ss) < eps;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      if (k < K) d_student[size_t(b) * K + k] = a * t[i] + (clamped ? 0.f : c2 * s[i]);
    }
  }
  contrib = warp_sum(contrib);
  block_det_sum(contrib, loss, scratch, ws);
}

// loss = c_kl * sum_b KL(q_b || p_b) + c_sl1 * sum_{b,k} smooth_l1(s - t) + c_ce * sum_b CE(s_b, label_b)
// q = softmax(t / T), p = softmax(s / T); smooth_l1 with beta = 1 (the F.smooth_l1_loss default).  The host folds the
// reductions (1/B, 1/(B K), T^2, the loss weights) into the three coefficients.
template <int EPT>
__global__ void __launch_bounds__(kLossWarps * 32) kd_loss_kernel(const float* __restrict__ student,
                                                                const float* __restrict__ teacher,
                                                                const long long* __restrict__ label, float* __restrict__ loss,
                                                                float* __restrict__ d_student, int B, int K, float inv_T,
                                                                float c_kl, float c_sl1, float c_ce, float grad_scale,
                                                                void* __restrict__ ws) {
  __shared__ float scratch[kLossWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kLossWarps + warp;
  float contrib = 0.f;
  if (b < B) {
    float s[EPT], t[EPT];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      s[i] = (k < K) ? student[size_t(b) * K + k] : -INFINITY;
      t[i] = (k < K) ? teacher[size_t(b) * K + k] : -INFINITY;
      ms = fmaxf(ms, s[i]);
      mt = fmaxf(mt, t[i]);
    }
    ms = warp_max(ms);
    mt = warp_max(mt);
    // three softmax denominators: student at T, teacher at T, student at 1 (cross-entropy term only)
    float zs = 0.f, zt = 0.f, z1 = 0.f;
    const float a2 = inv_T * kL2e;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      zs += ex2f_((s[i] - ms) * a2);
      zt += ex2f_((t[i] - mt) * a2);
      if (c_ce != 0.f) z1 += ex2f_((s[i] - ms) * kL2e);
    }
    zs = warp_sum(zs);
    zt = warp_sum(zt);
    const float lzs = __logf(zs), lzt = __logf(zt);
    float lz1 = 0.f;
    long long y = -1;
    if (c_ce != 0.f) {
      lz1 = __logf(warp_sum(z1));
      y = label[b];
    }
    float kl = 0.f, sl1 = 0.f, ce = 0.f;
    const float g_kl = grad_scale * c_kl * inv_T, g_sl1 = grad_scale * c_sl1, g_ce = grad_scale * c_ce;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int k = lane + 32 * i;
      if (k < K) {
        const float lp = (s[i] - ms) * inv_T - lzs, lq = (t[i] - mt) * inv_T - lzt;  // log-probabilities
        const float pk = __expf(lp), qk = __expf(lq);
        kl = fmaf(qk, lq - lp, kl);  // q = 0 contributes 0 (nn.KLDivLoss's xlogy convention)
        const float d = s[i] - t[i], ad = fabsf(d);
        sl1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
        float g = g_kl * (pk - qk) + g_sl1 * fminf(fmaxf(d, -1.f), 1.f);
        if (c_ce != 0.f) {
          const float l1 = (s[i] - ms) - lz1;
          g += g_ce * (__expf(l1) - (k == y ? 1.f : 0.f));
          if (k == y) ce = -l1;
        }
        d_student[size_t(b) * K + k] = g;
      }
    }
    contrib = c_kl * kl + c_sl1 * sl1 + c_ce * ce;
  }
  contrib = warp_sum(contrib);
  block_det_sum(contrib, loss, scratch, ws);
}

}  // namespace csn

using namespace csn;

extern "C" int csn_feature_dist_loss_fwd_bwd(const float* student, const float* teacher, const float* pred,
                                             const long long* label, float* loss, float* d_student, float* d_pred, int B,
                                             int K, int n_classes, float temperature, float alpha, float beta,
                                             float grad_scale, void* workspace, void* stream) {
  CSN_REQUIRE(student && teacher && loss && d_student && workspace, "csn_feature_dist_loss_fwd_bwd: null pointer");
  CSN_REQUIRE(B >= 1 && K >= 1 && K <= 1024, "csn_feature_dist_loss_fwd_bwd: need B >= 1 and 1 <= K <= 1024 (got B=%d K=%d)", B, K);
  CSN_REQUIRE(temperature != 0.f, "csn_feature_dist_loss_fwd_bwd: temperature must be non-zero");
  CSN_REQUIRE(!pred || (label && d_pred && n_classes >= 1), "csn_feature_dist_loss_fwd_bwd: pred needs label, d_pred and n_classes");
  cudaStream_t s = as_stream(stream);
  CSN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned), s));  // arrival ticket of the loss fold
  const int blocks = ceil_div(B, kLossWarps);
#define CSN_FD(E) feature_dist_loss_kernel<E><<<blocks, kLossWarps * 32, 0, s>>>(student, teacher, pred, label, loss, d_student, \
                                                                                 d_pred, B, K, n_classes, 1.f / temperature, alpha, beta, grad_scale, workspace)
  if (K <= 128) CSN_FD(4);
  else if (K <= 256) CSN_FD(8);
  else if (K <= 512) CSN_FD(16);
  else CSN_FD(32);
#undef CSN_FD
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_cosine_loss_fwd_bwd(const float* student, const float* teacher, float* loss, float* d_student, int B, int K,
                                       float eps, float grad_scale, void* workspace, void* stream) {
  CSN_REQUIRE(student && teacher && loss && d_student && workspace, "csn_cosine_loss_fwd_bwd: null pointer");
  CSN_REQUIRE(B >= 1 && K >= 1 && K <= 1024, "csn_cosine_loss_fwd_bwd: need B >= 1 and 1 <= K <= 1024 (got B=%d K=%d)", B, K);
  cudaStream_t s = as_stream(stream);
  CSN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned), s));  // arrival ticket of the loss fold
  const int blocks = ceil_div(B, kLossWarps);
#define CSN_CL(E) cosine_loss_kernel<E><<<blocks, kLossWarps * 32, 0, s>>>(student, teacher, loss, d_student, B, K, eps, grad_scale, workspace)
  if (K <= 128) CSN_CL(4);
  else if (K <= 256) CSN_CL(8);
  else if (K <= 512) CSN_CL(16);
  else CSN_CL(32);
#undef CSN_CL
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_kd_loss_fwd_bwd(const float* student, const float* teacher, const long long* label, float* loss,
                                   float* d_student, int B, int K, float temperature, float c_kl, float c_sl1, float c_ce,
                                   float grad_scale, void* workspace, void* stream) {
  CSN_REQUIRE(student && teacher && loss && d_student && workspace, "csn_kd_loss_fwd_bwd: null pointer");
  CSN_REQUIRE(B >= 1 && K >= 1 && K <= 1024, "csn_kd_loss_fwd_bwd: need B >= 1 and 1 <= K <= 1024 (got B=%d K=%d)", B, K);
  CSN_REQUIRE(temperature != 0.f, "csn_kd_loss_fwd_bwd: temperature must be non-zero");
  CSN_REQUIRE(c_ce == 0.f || label, "csn_kd_loss_fwd_bwd: the cross-entropy term needs labels");
  cudaStream_t s = as_stream(stream);
  CSN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned), s));  // arrival ticket of the loss fold
  const int blocks = ceil_div(B, kLossWarps);
#define CSN_KD(E) kd_loss_kernel<E><<<blocks, kLossWarps * 32, 0, s>>>(student, teacher, label, loss, d_student, B, K, \
                                                                       1.f / temperature, c_kl, c_sl1, c_ce, grad_scale, workspace)
  if (K <= 128) CSN_KD(4);
  else if (K <= 256) CSN_KD(8);
  else if (K <= 512) CSN_KD(16);
  else CSN_KD(32);
#undef CSN_KD
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
