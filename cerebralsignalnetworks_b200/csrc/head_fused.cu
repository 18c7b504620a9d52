// Projection head + DINO loss of the single-view train step (LstmDistillFromDinoV2Train.py:323-375: `output = Linear(h_T)`,
// DINOLoss.forward :62-105) forward AND backward in ONE kernel, for the narrow heads of cfg1/2 (384-d targets on a
// 128-wide encoder).  Unfused, this part of the step is five dependent launches of ~microsecond kernels (cast, Linear,
// loss, dX GEMM, plus the dW GEMM) -- ~45 us of launch / drain latency on a 520 us step.  Here a CTA owns R batch rows:
//   W [K, I] is staged once in shared memory with 16-byte cp.async, the 16-byte chunks of row k XOR-swizzled by k % 8:
//   conflict-free both for "thread = output column k" reading float4s in the forward product and for "thread = input
//   column j" reading scalars in the backward product,
//   pre = W h + b, emb = act(pre); teacher softmax ((t - c) / tau_t), student log-softmax (emb / tau_s), loss,
//   d_emb = coef / tau_s (p - q), d_pre = act'(pre) d_emb      (block reductions: one round of maxima, one of sums),
//   d_h = d_pre W (the gradient the BPTT kernel waits for).  The loss is folded across CTAs in a fixed order (det_cta_sum);
//   the batch-centre column sums are a separate fixed-order pass over the teacher (off the step's critical path).
// d_pre is also written out: dW = d_pre^T h and db = sum_b d_pre are NOT on the critical path, the caller runs that GEMM
// beside the backward recurrence.  Same log2-domain arithmetic as dino_loss_rowwarp_kernel.
#include "common.cuh"

namespace csn {

namespace {

constexpr float kL2e_h = 1.4426950408889634f;
constexpr float kLn2_h = 0.6931471805599453f;
constexpr int kHeadMaxWarps = 32;

__device__ __forceinline__ float ex2h(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2h(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

struct HeadParams {
  const void* h_last;   // [B, I] InT
  const float* W;       // [K, I]
  const float* bias;    // [K] or nullptr
  const float* teacher; // [B, K]
  const float* center;  // [K]
  float* loss;          // scalar, written by the last CTA
  void* ws;             // loss fold scratch (ticket + one partial per CTA)
  float* d_hlast;       // [B, I]
  float* d_pre;         // [B, K]
  int B, I, K, act;
  float ss, ts;         // log2(e) / tau_s, log2(e) / tau_t
  float coef, grad_coef;
};

// Block-wide reduction of NV values per thread (max or sum): warp shuffles, one shared-memory slot per warp and value,
// ONE block barrier; every thread then folds the per-warp slots itself.  `buf` is [kHeadMaxWarps][NV], double-buffered
// by the caller (consecutive rounds use different buffers, so no trailing barrier is needed).
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* buf, int warp, int lane, int nwarps) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = IS_MAX ? warp_max(v[i]) : warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[warp * NV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = buf[i];
    for (int w = 1; w < nwarps; ++w) a = IS_MAX ? fmaxf(a, buf[w * NV + i]) : a + buf[w * NV + i];
    v[i] = a;
  }
}

template <int R, typename InT>
__global__ void __launch_bounds__(1024, 1) head_dino_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float smem[];
  const int I = p.I, K = p.K;
  float* Ws = smem;                         // [K][I], chunk c of row k stored at chunk c ^ (k & 7)
  float* hs = Ws + size_t(K) * I;           // [R][I]
  float* dps = hs + R * I;                  // [R][K]
  float* red0 = dps + R * K;                // [kHeadMaxWarps][3 R]
  float* red1 = red0 + kHeadMaxWarps * 3 * R;
  float* part = red1 + kHeadMaxWarps * 3 * R;  // [G][R][I]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int b0 = blockIdx.x * R;
  const int rows = min(R, p.B - b0);

  // ---- stage W (one row per warp and pass, one 16-byte chunk per lane) and the CTA's h rows ----
  {
    const int chunks = I >> 2;
    const uint32_t ws0 = (uint32_t)__cvta_generic_to_shared(Ws);
    for (int kr = warp; kr < K; kr += nwarps) {
      const float* src = p.W + size_t(kr) * I;
      for (int c = lane; c < chunks; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ws0 + uint32_t(kr * I + 4 * (c ^ (kr & 7))) * 4u),
                     "l"(src + 4 * c));
    }
    asm volatile("cp.async.commit_group;\n" ::);
    const InT* h = reinterpret_cast<const InT*>(p.h_last);
    for (int e = tid; e < R * I; e += nthreads) {
      const int r = e / I, j = e - r * I;
      hs[e] = r < rows ? to_f32(h[size_t(b0 + r) * I + j]) : 0.f;
    }
  }
  const int k = tid;
  const bool kv = k < K;
  // teacher / centre of this thread's column, in flight while W arrives
  float tt[R];
  float cen = kv ? __ldg(p.center + k) : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) tt[r] = (kv && r < rows) ? __ldcs(p.teacher + size_t(b0 + r) * K + k) : 0.f;
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  // ---- forward: pre[r] = b[k] + sum_j W[k][j] h[r][j] ----
  float pre[R];
  {
    const float bk = (kv && p.bias) ? p.bias[k] : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) pre[r] = bk;
    if (kv) {
      const float* wrow = Ws + k * I;
      const int sw = k & 7;
#pragma unroll 4
      for (int c = 0; c < (I >> 2); ++c) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + 4 * (c ^ sw));
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 h4 = *reinterpret_cast<const float4*>(hs + r * I + 4 * c);  // broadcast
          pre[r] = fmaf(w4.x, h4.x, pre[r]);
          pre[r] = fmaf(w4.y, h4.y, pre[r]);
          pre[r] = fmaf(w4.z, h4.z, pre[r]);
          pre[r] = fmaf(w4.w, h4.w, pre[r]);
        }
      }
    }
  }

  // ---- loss: log2-domain logits, one round of maxima, one round of sums ----
  float sv[R], tv[R];
  float mx[2 * R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float emb = (p.act == CSN_ACT_RELU) ? fmaxf(pre[r], 0.f) : pre[r];
    sv[r] = kv ? emb * p.ss : -INFINITY;
    tv[r] = kv ? (tt[r] - cen) * p.ts : -INFINITY;
    mx[r] = sv[r];
    mx[R + r] = tv[r];
  }
  block_reduce<2 * R, true>(mx, red0, warp, lane, nwarps);
  float sm[3 * R];
  float es[R], et[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    es[r] = kv ? ex2h(sv[r] - mx[r]) : 0.f;
    et[r] = kv ? ex2h(tv[r] - mx[R + r]) : 0.f;
    sm[r] = es[r];
    sm[R + r] = et[r];
    sm[2 * R + r] = kv ? et[r] * sv[r] : 0.f;
  }
  block_reduce<3 * R, false>(sm, red1, warp, lane, nwarps);
  float loss_cta = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool rv = r < rows;
    const float izs = __fdividef(1.f, sm[r]), izt = __fdividef(1.f, sm[R + r]);
    if (rv) loss_cta += kLn2_h * (mx[r] + lg2h(sm[r]) - sm[2 * R + r] * izt);
    float d = p.grad_coef * (es[r] * izs - et[r] * izt);
    if (p.act == CSN_ACT_RELU && !(pre[r] > 0.f)) d = 0.f;
    if (!rv) d = 0.f;
    if (kv) {
      dps[r * K + k] = d;
      if (rv) p.d_pre[size_t(b0 + r) * K + k] = d;
    }
  }
  det_cta_sum(p.coef * loss_cta, p.ws, blockIdx.x, gridDim.x, p.loss);  // (synchronises the CTA: dps is complete)

  // ---- backward: d_h[r][j] = sum_k d_pre[r][k] W[k][j]; G thread groups split the k range ----
  const int G = nthreads / I;
  const int g = tid / I, j = tid - g * I;
  if (g < G) {
    const int kper = ceil_div(K, G);
    const int kb = g * kper, ke = min(K, kb + kper);
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const int jc = j >> 2, jo = j & 3;
#pragma unroll 8
    for (int kk = kb; kk < ke; ++kk) {
      const float w = Ws[kk * I + 4 * (jc ^ (kk & 7)) + jo];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(dps[r * K + kk], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) part[(g * R + r) * I + j] = acc[r];
  }
  __syncthreads();
  for (int e = tid; e < rows * I; e += nthreads) {
    const int r = e / I, jj = e - r * I;
    float a = 0.f;
    for (int gg = 0; gg < G; ++gg) a += part[(gg * R + r) * I + jj];
    p.d_hlast[size_t(b0 + r) * I + jj] = a;
  }
}

template <int R>
size_t head_smem_bytes(int I, int K, int nthreads) {
  const int G = nthreads / I;
  return sizeof(float) * (size_t(K) * I + size_t(R) * I + size_t(R) * K + 2 * size_t(kHeadMaxWarps) * 3 * R +
                          size_t(G) * R * I);
}

template <int R, typename InT>
int launch_head(const HeadParams& p, int nthreads, cudaStream_t s) {
  const size_t smem = head_smem_bytes<R>(p.I, p.K, nthreads);
  auto kern = head_dino_kernel<R, InT>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  kern<<<ceil_div(p.B, R), nthreads, smem, s>>>(p);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

}  // namespace
}  // namespace csn

using namespace csn;

extern "C" int csn_head_dino_supported(int B, int I, int K) {
  if (B < 1 || I < 32 || I % 32 != 0 || K < 1 || K > 1024) return 0;
  const int nthreads = ceil_div(K, 32) * 32 < I ? I : ceil_div(K, 32) * 32;
  if (nthreads > 1024) return 0;  // one thread per output column / input column
  return head_smem_bytes<4>(I, K, nthreads) <= size_t(220) * 1024 ? 1 : 0;
}

extern "C" int csn_head_dino_fwd_bwd(const void* h_last, int h_dtype, const float* W, const float* bias, int act,
                                     const float* teacher, const float* center, float student_temp, float teacher_temp,
                                     float* loss, float* d_hlast, float* d_pre, float* batch_center, int B, int I, int K,
                                     float grad_scale, void* workspace, void* stream) {
  CSN_REQUIRE(h_last && W && teacher && center && loss && d_hlast && d_pre && workspace,
              "csn_head_dino_fwd_bwd: null pointer");
  CSN_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, "csn_head_dino_fwd_bwd: W must be 16-byte aligned");
  CSN_REQUIRE(h_dtype == CSN_F32 || h_dtype == CSN_BF16, "csn_head_dino_fwd_bwd: h_dtype must be CSN_F32 or CSN_BF16");
  CSN_REQUIRE(act == CSN_ACT_NONE || act == CSN_ACT_RELU, "csn_head_dino_fwd_bwd: act must be NONE or RELU");
  CSN_REQUIRE(student_temp > 0.f && teacher_temp != 0.f, "csn_head_dino_fwd_bwd: bad temperature");
  if (!csn_head_dino_supported(B, I, K)) {
    set_error("csn_head_dino_fwd_bwd: shape B=%d I=%d K=%d not served (needs I %% 32 == 0, K <= 1024 and W in shared memory); "
              "use csn_gemm_f32 + csn_dino_loss_fwd_bwd", B, I, K);
    return CSN_EUNSUPPORTED;
  }
  HeadParams p{};
  p.h_last = h_last; p.W = W; p.bias = bias; p.teacher = teacher; p.center = center; p.loss = loss; p.d_hlast = d_hlast;
  p.d_pre = d_pre; p.ws = workspace; p.B = B; p.I = I; p.K = K; p.act = act;
  p.ss = kL2e_h / student_temp; p.ts = kL2e_h / teacher_temp;
  p.coef = 1.f / B;
  p.grad_coef = p.coef / student_temp * grad_scale;
  cudaStream_t s = as_stream(stream);
  CSN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned), s));  // arrival ticket of the loss fold
  const int kt = ceil_div(K, 32) * 32;
  const int nthreads = kt < I ? I : kt;
  // two rows per CTA while that is at most ~2 waves of CTAs (each CTA re-stages W), else four
  const bool r2 = ceil_div(B, 2) <= 2 * sm_count();
  if (h_dtype == CSN_BF16) CSN_TRY((r2 ? launch_head<2, __nv_bfloat16>(p, nthreads, s) : launch_head<4, __nv_bfloat16>(p, nthreads, s)));
  else CSN_TRY((r2 ? launch_head<2, float>(p, nthreads, s) : launch_head<4, float>(p, nthreads, s)));
  // batch_center [K] += sum_b teacher, fixed order; a caller that overlaps it with other work passes NULL and runs
  // csn_colsum_f32 itself (the train step does, on its side stream)
  if (batch_center) CSN_TRY(colsum_det(teacher, batch_center, B, K, K, 1, s));
  return CSN_OK;
}
