// LSTM layer, fp32 SIMT path (compute_dtype = CSN_F32): the tight-tolerance parity mode for the restated
// models.lstm.Model (torch.nn.LSTM semantics: gate order i,f,g,o, zero initial state).  One launch per timestep;
// the throughput path is the persistent tcgen05 kernel in lstm_tc.cu.
//
// reserve  (fp32): gates_act [T,B,4H] (i,f,g,o after the nonlinearity) | c [T,B,H]
// workspace(fp32): fwd: W_hh^T [H,4H]      bwd: dG [T,B,4H] | dc [B,H]
#include "common.cuh"

namespace csn {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (r0 + j < R && c < Cc) tile[j][threadIdx.x] = in[size_t(r0 + j) * Cc + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (c0 + j < Cc && r < R) out[size_t(c0 + j) * R + r] = tile[threadIdx.x][j];
}

// block (32 units, 8 batch rows).  gates[t] holds x_t W_ih^T + b_ih on entry, activated gates on exit.
__global__ void __launch_bounds__(256) lstm_fwd_step_f32(float* __restrict__ gates_t, const float* __restrict__ b_hh,
                                                        const float* __restrict__ whhT, const float* __restrict__ h_prev,
                                                        const float* __restrict__ c_prev, float* __restrict__ h_t,
                                                        float* __restrict__ c_t, int B, int H) {
  extern __shared__ float hs[];  // [8][H]
  const int u = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y * 8 + threadIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (h_prev) {
    for (int e = tid; e < 8 * H; e += 256) {
      int bb = blockIdx.y * 8 + e / H;
      hs[e] = (bb < B) ? h_prev[size_t(bb) * H + (e % H)] : 0.f;
    }
    __syncthreads();
  }
  if (u >= H || b >= B) return;
  float* g = gates_t + size_t(b) * 4 * H;
  float pi = g[u] + b_hh[u], pf = g[H + u] + b_hh[H + u], pg = g[2 * H + u] + b_hh[2 * H + u],
        po = g[3 * H + u] + b_hh[3 * H + u];
  if (h_prev) {
    const float* hrow = hs + threadIdx.y * H;
    for (int k = 0; k < H; ++k) {
      const float hv = hrow[k];
      const float* w = whhT + size_t(k) * 4 * H;
      pi = fmaf(hv, w[u], pi);
      pf = fmaf(hv, w[H + u], pf);
      pg = fmaf(hv, w[2 * H + u], pg);
      po = fmaf(hv, w[3 * H + u], po);
    }
  }
  const float i = sigmoid_f(pi), f = sigmoid_f(pf), gg = tanhf(pg), o = sigmoid_f(po);
  const float cp = c_prev ? c_prev[size_t(b) * H + u] : 0.f;
  const float c = fmaf(f, cp, i * gg);
  g[u] = i; g[H + u] = f; g[2 * H + u] = gg; g[3 * H + u] = o;
  c_t[size_t(b) * H + u] = c;
  h_t[size_t(b) * H + u] = o * tanhf(c);
}

// Step t of BPTT.  dG_next = dG[t+1] (NULL at t = T-1) feeds the recurrent term dh += dG[t+1] W_hh.
__global__ void __launch_bounds__(256) lstm_bwd_step_f32(const float* __restrict__ gates_t, const float* __restrict__ c_t,
                                                        const float* __restrict__ c_prev, const float* __restrict__ w_hh,
                                                        const float* __restrict__ dG_next, const float* __restrict__ d_hseq_t,
                                                        const float* __restrict__ d_hlast, float* __restrict__ dc,
                                                        float* __restrict__ dG_t, int B, int H) {
  extern __shared__ float gs[];  // [8][4H]
  const int u = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y * 8 + threadIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int H4 = 4 * H;
  if (dG_next) {
    for (int e = tid; e < 8 * H4; e += 256) {
      int bb = blockIdx.y * 8 + e / H4;
      gs[e] = (bb < B) ? dG_next[size_t(bb) * H4 + (e % H4)] : 0.f;
    }
    __syncthreads();
  }
  if (u >= H || b >= B) return;
  float dh = 0.f;
  if (d_hseq_t) dh += d_hseq_t[size_t(b) * H + u];
  if (d_hlast) dh += d_hlast[size_t(b) * H + u];
  if (dG_next) {
    const float* grow = gs + threadIdx.y * H4;
    float acc = 0.f;
    for (int r = 0; r < H4; ++r) acc = fmaf(grow[r], w_hh[size_t(r) * H + u], acc);
    dh += acc;
  }
  const float* g = gates_t + size_t(b) * H4;
  const float i = g[u], f = g[H + u], gg = g[2 * H + u], o = g[3 * H + u];
  const float c = c_t[size_t(b) * H + u];
  const float cp = c_prev ? c_prev[size_t(b) * H + u] : 0.f;
  const float tc = tanhf(c);
  float dct = dc[size_t(b) * H + u] + dh * o * (1.f - tc * tc);
  float* d = dG_t + size_t(b) * H4;
  d[u] = dct * gg * i * (1.f - i);
  d[H + u] = dct * cp * f * (1.f - f);
  d[2 * H + u] = dct * i * (1.f - gg * gg);
  d[3 * H + u] = dh * tc * o * (1.f - o);
  dc[size_t(b) * H + u] = dct * f;
}

int lstm_layer_fwd_f32(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       float* h_seq, float* reserve, float* workspace, int T, int B, int I, int H, cudaStream_t s) {
  float* gates = reserve;
  float* cbuf = reserve + size_t(T) * B * 4 * H;
  float* whhT = workspace;
  CSN_TRY(csn_gemm_f32(0, 1, T * B, 4 * H, I, 1.f, x, I, w_ih, I, 0.f, gates, 4 * H, b_ih, CSN_ACT_NONE, s));
  transpose_f32_kernel<<<dim3(ceil_div(H, 32), ceil_div(4 * H, 32)), dim3(32, 8), 0, s>>>(w_hh, whhT, 4 * H, H);
  CSN_LAUNCH_CHECK();
  dim3 grid(ceil_div(H, 32), ceil_div(B, 8)), block(32, 8);
  for (int t = 0; t < T; ++t) {
    const size_t o4 = size_t(t) * B * 4 * H, o1 = size_t(t) * B * H;
    lstm_fwd_step_f32<<<grid, block, size_t(8) * H * 4, s>>>(gates + o4, b_hh, whhT, t ? h_seq + o1 - size_t(B) * H : nullptr,
                                                            t ? cbuf + o1 - size_t(B) * H : nullptr, h_seq + o1, cbuf + o1, B, H);
  }
  count_launches(T - 1);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

int lstm_layer_bwd_f32(const float* x, const float* w_ih, const float* w_hh, const float* h_seq, const float* reserve,
                       const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                       float* dx, float* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s) {
  const float* gates = reserve;
  const float* cbuf = reserve + size_t(T) * B * 4 * H;
  float* dG = workspace;
  float* dc = workspace + size_t(T) * B * 4 * H;
  CSN_CUDA(cudaMemsetAsync(dc, 0, size_t(B) * H * 4, s));
  dim3 grid(ceil_div(H, 32), ceil_div(B, 8)), block(32, 8);
  const size_t bwd_smem = size_t(8) * 4 * H * 4;
  CSN_REQUIRE(bwd_smem <= 200 * 1024, "fp32 LSTM path: hidden size %d too large for the step kernel", H);
  if (bwd_smem > 48 * 1024)
    CSN_CUDA(cudaFuncSetAttribute(lstm_bwd_step_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
  for (int t = T - 1; t >= 0; --t) {
    const size_t o4 = size_t(t) * B * 4 * H, o1 = size_t(t) * B * H;
    lstm_bwd_step_f32<<<grid, block, size_t(8) * 4 * H * 4, s>>>(
        gates + o4, cbuf + o1, t ? cbuf + o1 - size_t(B) * H : nullptr, w_hh,
        (t + 1 < T) ? dG + o4 + size_t(B) * 4 * H : nullptr, d_hseq ? d_hseq + o1 : nullptr,
        (t == T - 1) ? d_hlast : nullptr, dc, dG + o4, B, H);
  }
  count_launches(T - 1);
  CSN_LAUNCH_CHECK();
  const float beta = accumulate ? 1.f : 0.f;
  // dW_ih[4H,I] = dG^T x ; dW_hh[4H,H] = dG[1:]^T h_seq[:-1] ; db = colsum(dG)
  CSN_TRY(csn_gemm_f32(1, 0, 4 * H, I, T * B, 1.f, dG, 4 * H, x, I, beta, dw_ih, I, nullptr, CSN_ACT_NONE, s));
  if (T > 1) {
    CSN_TRY(csn_gemm_f32(1, 0, 4 * H, H, (T - 1) * B, 1.f, dG + size_t(B) * 4 * H, 4 * H, h_seq, H, beta, dw_hh, H,
                         nullptr, CSN_ACT_NONE, s));
  } else if (!accumulate) {
    CSN_CUDA(cudaMemsetAsync(dw_hh, 0, size_t(4) * H * H * 4, s));
  }
  CSN_TRY(csn_colsum_f32(dG, db_ih, T * B, 4 * H, 4 * H, accumulate, s));
  CSN_TRY(csn_colsum_f32(dG, db_hh, T * B, 4 * H, 4 * H, accumulate, s));
  if (dx) CSN_TRY(csn_gemm_f32(0, 0, T * B, I, 4 * H, 1.f, dG, 4 * H, w_ih, I, 0.f, dx, I, nullptr, CSN_ACT_NONE, s));
  return CSN_OK;
}

}  // namespace csn
