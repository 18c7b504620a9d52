// The optimiser tail of the LstmDistillation step as ONE pass over the parameters (SURVEY.md K8 + K9 + K10):
//   utils.clip_gradients (utils/utils.py:132-141): per-PARAMETER L2 clip, coef = clip / (||g_p|| + 1e-6) applied when < 1
//   torch.optim.AdamW over get_params_groups (LstmDistillation.py:469-471, utils/utils.py:636-647), lr / wd rewritten
//     every iteration from the cosine schedules (:540-544), parameters whose gradient was cancelled skipped entirely
//     (cancel_gradients_last_layer, utils/utils.py:144-149: p.grad = None -> no moments, no decay, no step count)
//   EMA teacher (LstmDistillation.py:616-619): teacher = m teacher + (1 - m) student, for EVERY student parameter
// The reference runs these as three sweeps with a host sync per parameter (`.item()` in clip_gradients).  Here:
//   seg_sumsq   : squared norm partial of every 2048-element chunk of the flat gradient buffer (no atomics)
//   seg_prepare : one warp per parameter -- folds its chunks' partials in chunk order, turns the norm into the clip
//                 coefficient, advances the parameter's own step count and derives the Adam step size from it
//   fused_update: one CTA per chunk -- clip scale, AdamW, EMA in the same registers: 4 reads + 4 writes per parameter
//                 and ONE sweep (24 + 12 B against the reference's 28 + 8 + 12 + 8 B in three)
// lr / wd / EMA momentum live in a small device array, so a captured CUDA graph of the step stays valid while the
// schedules move.  Deterministic: same inputs, same bits.
#include "common.cuh"

namespace csn {

constexpr int kFoChunk = 2048;  // elements per CTA of the update (= the clip chunk: 256 threads x 2 float4)

struct FusedOptim {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;                 // teacher parameters (same flat layout) or NULL
  const long long* seg_off;   // [n_seg + 1] parameter boundaries in the flat buffers (multiples of 4)
  const int* seg_group;       // [n_seg] parameter group of each parameter
  const int* seg_active;      // [n_seg] 0: the optimiser skips the parameter this step (its gradient was cancelled)
  int* seg_step;              // [n_seg] per-parameter step count (advanced here)
  const float* hyper;         // [n_groups][2] = (lr, weight_decay) per group, then the EMA momentum
  float* chunk_sumsq;         // [n_chunks]
  float* seg_consts;          // [n_seg][4] = (step_size, 1/sqrt(bias_correction2), clip coefficient * grad_scale, lr * wd)
  float* sumsq_out;           // [n_seg] squared gradient norms (of the scaled gradients) or NULL
  int n_seg, n_groups;
  float beta1, beta2, eps, clip, grad_scale;
  int decoupled;
};

// chunk -> (segment, first element): chunks are laid per segment, ceil(len / kFoChunk) each
__device__ __forceinline__ bool locate_chunk(const long long* __restrict__ seg_off, int n_seg, long long chunk, int* seg,
                                             long long* beg, long long* end) {
  long long first = 0;
  for (int s = 0; s < n_seg; ++s) {
    const long long len = seg_off[s + 1] - seg_off[s];
    const long long nc = (len + kFoChunk - 1) / kFoChunk;
    if (chunk < first + nc) {
      *seg = s;
      *beg = seg_off[s] + (chunk - first) * kFoChunk;
      *end = min(*beg + (long long)kFoChunk, seg_off[s + 1]);
      return true;
    }
    first += nc;
  }
  return false;
}

__global__ void __launch_bounds__(256) fo_sumsq_kernel(const FusedOptim o) {
  __shared__ float red[8];
  int seg;
  long long beg, end;
  if (!locate_chunk(o.seg_off, o.n_seg, blockIdx.x, &seg, &beg, &end)) {
    if (threadIdx.x == 0) o.chunk_sumsq[blockIdx.x] = 0.f;
    return;
  }
  float acc = 0.f;
  for (long long i = beg + threadIdx.x * 4; i < end; i += 256 * 4) {
    const float4 g = *reinterpret_cast<const float4*>(o.g + i);
    acc = fmaf(g.x, g.x, acc); acc = fmaf(g.y, g.y, acc); acc = fmaf(g.z, g.z, acc); acc = fmaf(g.w, g.w, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    o.chunk_sumsq[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) fo_prepare_kernel(const FusedOptim o) {
  const int seg = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (seg >= o.n_seg) return;
  long long first = 0;
  for (int i = 0; i < seg; ++i) first += (o.seg_off[i + 1] - o.seg_off[i] + kFoChunk - 1) / kFoChunk;
  const long long nc = (o.seg_off[seg + 1] - o.seg_off[seg] + kFoChunk - 1) / kFoChunk;
  float a = 0.f;
  if (o.clip > 0.f || o.sumsq_out)
    for (long long c = lane; c < nc; c += 32) a += o.chunk_sumsq[first + c];
  a = warp_sum(a) * o.grad_scale * o.grad_scale;  // norm of the gradient the reference clips: the rank-averaged one
  if (lane != 0) return;
  if (o.sumsq_out) o.sumsq_out[seg] = a;
  float coef = 1.f;
  if (o.clip > 0.f) {
    const float c = o.clip / (sqrtf(a) + 1e-6f);
    if (c < 1.f) coef = c;
  }
  const int grp = o.seg_group[seg];
  const float lr = o.hyper[2 * grp], wd = o.hyper[2 * grp + 1];
  float step_size = 0.f, inv_sqrt_bc2 = 1.f;
  if (o.seg_active[seg]) {
    const int t = ++o.seg_step[seg];
    const float bc1 = -expm1f((float)t * logf(o.beta1)), bc2 = -expm1f((float)t * logf(o.beta2));
    step_size = lr / bc1;
    inv_sqrt_bc2 = rsqrtf(bc2);
  }
  float* k = o.seg_consts + 4 * seg;
  k[0] = step_size; k[1] = inv_sqrt_bc2; k[2] = coef * o.grad_scale; k[3] = lr * wd;
}

__global__ void __launch_bounds__(256) fo_update_kernel(const FusedOptim o) {
  int seg;
  long long beg, end;
  if (!locate_chunk(o.seg_off, o.n_seg, blockIdx.x, &seg, &beg, &end)) return;
  const bool active = o.seg_active[seg] != 0;
  const float4 k = *reinterpret_cast<const float4*>(o.seg_consts + 4 * seg);
  const float step_size = k.x, inv_sqrt_bc2 = k.y, gscale = k.z, lrwd = k.w;
  const float wd = o.hyper[2 * o.seg_group[seg] + 1];
  const float mom = o.ema ? o.hyper[2 * o.n_groups] : 0.f;
  const float b1 = o.beta1, b2 = o.beta2;
  for (long long i = beg + threadIdx.x * 4; i < end; i += 256 * 4) {
    float4 P = *reinterpret_cast<float4*>(o.p + i);
    if (active) {
      const float4 G = *reinterpret_cast<const float4*>(o.g + i);
      float4 M = *reinterpret_cast<float4*>(o.m + i);
      float4 V = *reinterpret_cast<float4*>(o.v + i);
      float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gr = gg[j] * gscale;
        if (o.decoupled) pp[j] *= (1.f - lrwd); else gr = fmaf(wd, pp[j], gr);
        mm[j] = b1 * mm[j] + (1.f - b1) * gr;
        vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
        pp[j] -= step_size * (mm[j] / (sqrtf(vv[j]) * inv_sqrt_bc2 + o.eps));
      }
      *reinterpret_cast<float4*>(o.p + i) = P;
      *reinterpret_cast<float4*>(o.m + i) = M;
      *reinterpret_cast<float4*>(o.v + i) = V;
    }
    if (o.ema) {
      float4 E = *reinterpret_cast<float4*>(o.ema + i);
      E.x = E.x * mom + P.x * (1.f - mom); E.y = E.y * mom + P.y * (1.f - mom);
      E.z = E.z * mom + P.z * (1.f - mom); E.w = E.w * mom + P.w * (1.f - mom);
      *reinterpret_cast<float4*>(o.ema + i) = E;
    }
  }
}

}  // namespace csn

using namespace csn;

extern "C" int csn_fused_optim_workspace_bytes(int n_seg, long long n_chunks, size_t* bytes) {
  CSN_REQUIRE(n_seg >= 1 && n_chunks >= 1 && bytes, "csn_fused_optim_workspace_bytes: bad arguments");
  *bytes = size_t(n_chunks) * 4 + 256 + size_t(n_seg) * 16;
  return CSN_OK;
}

extern "C" int csn_fused_optim_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_params,
                                    const long long* seg_off_dev, const int* seg_group_dev, const int* seg_active_dev,
                                    int* seg_step_dev, int n_seg, long long n_chunks, const float* hyper_dev, int n_groups,
                                    float beta1, float beta2, float eps, int decoupled, float clip, float grad_scale,
                                    float* sumsq_out, void* workspace, void* stream) {
  CSN_REQUIRE(params && grads && exp_avg && exp_avg_sq && seg_off_dev && seg_group_dev && seg_active_dev && seg_step_dev &&
                  hyper_dev && workspace, "csn_fused_optim_step: null pointer");
  CSN_REQUIRE(n_seg >= 1 && n_chunks >= 1 && n_groups >= 1 && clip >= 0.f, "csn_fused_optim_step: bad arguments");
  CSN_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(ema_params)) & 15) == 0,
              "csn_fused_optim_step: buffers must be 16-byte aligned");
  FusedOptim o{};
  o.p = params; o.g = grads; o.m = exp_avg; o.v = exp_avg_sq; o.ema = ema_params;
  o.seg_off = seg_off_dev; o.seg_group = seg_group_dev; o.seg_active = seg_active_dev; o.seg_step = seg_step_dev;
  o.hyper = hyper_dev;
  o.chunk_sumsq = reinterpret_cast<float*>(workspace);
  o.seg_consts = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((size_t(n_chunks) * 4 + 255) & ~size_t(255)));
  o.sumsq_out = sumsq_out;
  o.n_seg = n_seg; o.n_groups = n_groups;
  o.beta1 = beta1; o.beta2 = beta2; o.eps = eps; o.clip = clip; o.grad_scale = grad_scale; o.decoupled = decoupled;
  cudaStream_t s = as_stream(stream);
  if (clip > 0.f || sumsq_out) {
    fo_sumsq_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(o);
    CSN_LAUNCH_CHECK();
  }
  fo_prepare_kernel<<<ceil_div(n_seg, 8), 256, 0, s>>>(o);
  CSN_LAUNCH_CHECK();
  fo_update_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(o);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
