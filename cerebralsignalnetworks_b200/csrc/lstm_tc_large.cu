// LSTM layer, bf16 tensor-core path for LARGE hidden sizes (H > 128, e.g. cfg 4: H = 512).
//
// W_hh (4H x H bf16 = 2 MB at H = 512) does not fit the tensor memory / shared memory of one SM, so the persistent
// single-CTA recurrence of lstm_tc.cu does not apply.  Here every timestep is ONE tcgen05 GEMM over the whole batch
// (gemm_tc.cu: TMA -> smem ring -> tcgen05.mma -> TMEM) whose epilogue IS the LSTM cell: the gate rows of the weights
// are interleaved (row 4u+g = gate g of unit u) so that the four accumulator columns a lane holds are one cell's
// (i,f,g,o); the epilogue adds the hoisted input projection, applies the nonlinearities, updates c and writes h_t
// as bf16 directly in the K-major layout the next step's A operand is loaded from.  The T launches are captured in
// the step's CUDA graph, so the chain costs one submission.  BPTT mirrors it: a pointwise cell-gradient kernel and
// one GEMM (dh_{t-1} = dG_t W_hh) per step, then three batched GEMMs for dW_ih, dW_hh, dX and a ones-GEMM for db.
//
// A cluster-resident variant (W_hh split over a 16-CTA cluster, h exchanged through DSMEM) is the next step for this
// size class; see DESIGN.md.
#include <mutex>

#include "gemm_tc.cuh"

namespace csn {

// lstm_cluster.cu: persistent cluster recurrence for H = 256 / 512
bool lstm_cluster_supported(int H);
bool lstm_cluster_overlap_ok(int B, int H);
int lstm_cluster_fwd(const float* xp, const __nv_bfloat16* whh_perm, __nv_bfloat16* h_seq, __nv_bfloat16* gates, float* c_seq, int T,
                     int B, int H, const unsigned* xp_flags, int xp_chunk, void* xchg, cudaStream_t s);
size_t lstm_cluster_xchg_bytes(int B, int H);
int lstm_cluster_bwd(const __nv_bfloat16* gates, const float* c_seq, const float* d_hseq, const float* d_hlast,
                     const __nv_bfloat16* whh_t, __nv_bfloat16* dG, int T, int B, int H, unsigned* done, int done_chunk,
                     cudaStream_t s);
int lstm_cluster_participants(int B, int H);

static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

// programmatic dependent launch for the per-timestep kernels (CSN_NO_PDL=1 switches it off for A/B measurements)
static const int kPdl = [] { const char* e = getenv("CSN_NO_PDL"); return (e && e[0] == '1') ? 0 : 1; }();

__device__ __forceinline__ float tanh_fast_l(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rows: dst[4u+g, :] = bf16(src[g*H+u, :])
__global__ void permute_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int N) {
  const int r = blockIdx.x;  // destination row 4u+g
  const int u = r >> 2, g = r & 3;
  const float* s = src + size_t(g * H + u) * N;
  for (int c = threadIdx.x; c < N; c += blockDim.x) dst[size_t(r) * N + c] = __float2bfloat16_rn(s[c]);
}
// dst[m, 4u+g] = bf16(src[g*H+u, m]): W_hh^T with gate-interleaved columns (the cluster BPTT's resident operand rows);
// 32 x 32 tiles through shared memory so that both the reads and the writes are coalesced
__global__ void permute_transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, m0 = blockIdx.x * 32;  // destination columns r = 4u+g in [r0, r0+32), rows m in [m0, m0+32)
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, u = r >> 2, g = r & 3;
    tile[i][threadIdx.x] = src[size_t(g * H + u) * H + m0 + threadIdx.x];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    dst[size_t(m0 + i) * (4 * H) + r0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][i]);
}
__global__ void permute_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ dst, int H) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < 4 * H) {
    const int u = r >> 2, g = r & 3;
    dst[r] = b_ih[g * H + u] + b_hh[g * H + u];
  }
}
// dst[g*H+u, :] (+)= sum_z src[z][4u+g, 0:N] (src leading dimension lds, split-K slabs `stride` floats apart, added in
// split order: the fixed-order tail of the split-K weight-gradient products -- no float atomics)
__global__ void unpermute_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int N, int lds, int accumulate,
                                      int S, size_t stride) {
  const int r = blockIdx.x;
  const int u = r >> 2, g = r & 3;
  float* d = dst + size_t(g * H + u) * N;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float v = src[size_t(r) * lds + c];
    for (int z = 1; z < S; ++z) v += src[size_t(z) * stride + size_t(r) * lds + c];
    d[c] = accumulate ? d[c] + v : v;
  }
}
__global__ void fill_bf16_kernel(__nv_bfloat16* __restrict__ dst, size_t n, float v) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  for (; i < n; i += stride) dst[i] = b;
}

// Cell gradient of one timestep (gate-interleaved layouts).  dG_t = d(pre-activations); dc carried in place.
__global__ void lstm_cell_bwd_kernel(const __nv_bfloat16* __restrict__ gates_t, const float* __restrict__ c_t,
                                     const float* __restrict__ c_prev, const float* __restrict__ dh_rec,
                                     const float* __restrict__ d_hseq_t, const float* __restrict__ d_hlast,
                                     float* __restrict__ dc, __nv_bfloat16* __restrict__ dG_t, int B, int H,
                                     int dh_parts) {
  // launched with programmatic stream serialization: let the next grid (the dh GEMM) start its prologue, then wait for
  // the previous one (the dh GEMM of step t+1) before touching its output
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const size_t cell = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t cells = size_t(B) * H;
  if (cell >= cells) return;
  const uint2 gq = *reinterpret_cast<const uint2*>(gates_t + cell * 4);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&gq.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&gq.y);
  const float i = __bfloat162float(lo.x), f = __bfloat162float(lo.y), g = __bfloat162float(hi.x), o = __bfloat162float(hi.y);
  float dh = 0.f;
  if (dh_rec)  // the split-K partial products of dh = dG_{t+1} W_hh, one slab per split
    for (int z = 0; z < dh_parts; ++z) dh += dh_rec[size_t(z) * cells + cell];
  if (d_hseq_t) dh += d_hseq_t[cell];
  if (d_hlast) dh += d_hlast[cell];
  const float tc = tanh_fast_l(c_t[cell]);
  const float cp = c_prev ? c_prev[cell] : 0.f;
  const float dct = fmaf(dh * o, 1.f - tc * tc, dc[cell]);
  const __nv_bfloat162 q0 = __floats2bfloat162_rn(dct * g * i * (1.f - i), dct * cp * f * (1.f - f));
  const __nv_bfloat162 q1 = __floats2bfloat162_rn(dct * i * (1.f - g * g), dh * tc * o * (1.f - o));
  uint2 out;
  out.x = *reinterpret_cast<const uint32_t*>(&q0);
  out.y = *reinterpret_cast<const uint32_t*>(&q1);
  *reinterpret_cast<uint2*>(dG_t + cell * 4) = out;
  dc[cell] = dct * f;
}

// Helper stream (one per device) for the GEMM-shaped work that runs BESIDE a cluster recurrence: the recurrence occupies
// 64-112 SMs for ~0.8 ms per layer, the rest of the machine takes the projection / weight-gradient GEMMs chunk by chunk.
struct OverlapStream {
  cudaStream_t st;
  cudaEvent_t fork, first, join;
};
static OverlapStream* overlap_stream() {
  static const bool off = [] { const char* e = getenv("CSN_LSTM_NO_OVERLAP"); return e && e[0] == '1'; }();
  if (off) return nullptr;
  static OverlapStream table[64];
  static bool made[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!made[dev]) {
    OverlapStream os{};
    if (cudaStreamCreateWithFlags(&os.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&os.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&os.first, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&os.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    table[dev] = os;
    made[dev] = true;
  }
  return &table[dev];
}
constexpr int kChunks = 8;      // time chunks of the overlapped GEMMs
constexpr int kChunkSplit = 4;  // split-K ranges inside one chunk (products with few output tiles)
__global__ void set_flag_kernel(unsigned* flag) { *flag = 1u; }
// one warp parks on a counter another (running) kernel bumps; everything behind it in its stream starts once the counter
// has reached `target` (20 s watchdog: a stuck producer traps instead of hanging the device)
__global__ void wait_counter_kernel(const unsigned* counter, unsigned target) {
  if (threadIdx.x != 0) return;
  unsigned v, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) return;
    __nanosleep(500);
    if ((++spins & 4095u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}

struct LargeWs {
  float* xp;                 // fwd: [TB,4H] fp32 ; bwd: unused
  __nv_bfloat16* dG;         // bwd: [TB,4H] bf16 (aliases xp)
  __nv_bfloat16* wih;        // [4H, I] permuted bf16
  __nv_bfloat16* whh;        // [4H, H] permuted bf16
  __nv_bfloat16* whh_t;      // [H, 4H] transposed, columns permuted (cluster BPTT only)
  float* bias;               // [4H] permuted b_ih + b_hh
  float* dh_rec;             // 8 x [B, H]: split-K slabs of the recurrent gradient (summed by the cell kernel)
  float* dc;                 // [B, H]
  float* dwp;                // [4H, max(I, H)] permuted dW scratch
  float* dwp2;               // second slab region (cluster BPTT: dW_ih and dW_hh chunk products are in flight together)
  __nv_bfloat16* ones;       // [TB, 8]
  float* dbp;                // [4H, 8]
  unsigned* flags;           // [2 * kChunks] chunk flags of the overlapped GEMMs
  uint8_t* xchg;             // cluster recurrence: global scratch of the per-step exchange (lstm_cluster_xchg_bytes)
  size_t total;
};

// splits of a [M, N] = A^T B product contracting over K rows: enough CTAs to fill the machine (what the GEMM launches)
static int dw_splits(int M, int N, int K) {
  return gemm_tc_splits(K, std::max(1, sm_count() / (ceil_div(M, 128) * ceil_div(N, 128))));
}

static LargeWs carve_ws(void* base, int T, int B, int I, int H) {
  const size_t tb = size_t(T) * B;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* q = p ? p + off : nullptr; off += al256(bytes); return q; };
  LargeWs w;
  w.xp = reinterpret_cast<float*>(take(tb * 4 * H * 4));
  w.dG = reinterpret_cast<__nv_bfloat16*>(w.xp);
  w.wih = reinterpret_cast<__nv_bfloat16*>(take(size_t(4) * H * I * 2));
  w.whh = reinterpret_cast<__nv_bfloat16*>(take(size_t(4) * H * H * 2));
  w.whh_t = reinterpret_cast<__nv_bfloat16*>(take(size_t(4) * H * H * 2));
  w.bias = reinterpret_cast<float*>(take(size_t(4) * H * 4));
  w.dh_rec = reinterpret_cast<float*>(take(size_t(8) * B * H * 4));
  w.dc = reinterpret_cast<float*>(take(size_t(B) * H * 4));
  // split-K slabs of the weight-gradient products (one partial [4H, N] per split, summed by unpermute_rows_kernel)
  // (sized for the REQUESTED split count: the GEMM may round it down, never up)
  // (at least kChunks slabs: with the GEMMs overlapped, every time chunk's product is one slab)
  auto req = [](int M, int N) { return size_t(std::max(kChunks, sm_count() / (ceil_div(M, 128) * ceil_div(N, 128)))); };
  const size_t s_ih = req(4 * H, I) * 4 * H * I, s_hh = req(4 * H, H) * 4 * H * H;
  w.dwp = reinterpret_cast<float*>(take((s_ih > s_hh ? s_ih : s_hh) * 4));
  w.dwp2 = reinterpret_cast<float*>(take(lstm_cluster_supported(H) ? size_t(kChunks) * kChunkSplit * 4 * H * I * 4 : 0));
  w.ones = reinterpret_cast<__nv_bfloat16*>(take(tb * 8 * 2));
  w.dbp = reinterpret_cast<float*>(take(std::max(req(4 * H, 8), size_t(kChunks) * kChunkSplit) * 4 * H * 8 * 4));
  w.flags = reinterpret_cast<unsigned*>(take(2 * kChunks * sizeof(unsigned)));
  w.xchg = reinterpret_cast<uint8_t*>(take(lstm_cluster_supported(H) ? lstm_cluster_xchg_bytes(B, H) : 0));
  w.total = off;
  return w;
}

int lstm_large_bytes(int T, int B, int I, int H, size_t* reserve, size_t* workspace) {
  if (I % 8 != 0 || H % 8 != 0) {
    set_error("bf16 tensor-core LSTM path needs input and hidden sizes that are multiples of 8 (I=%d H=%d)", I, H);
    return CSN_EUNSUPPORTED;
  }
  const size_t tb = size_t(T) * B;
  *reserve = al256(tb * 4 * H * 2) + al256(tb * H * 4);
  *workspace = carve_ws(nullptr, T, B, I, H).total;
  return CSN_OK;
}

static int prep_weights(const LargeWs& w, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int I,
                        int H, cudaStream_t s) {
  permute_rows_bf16_kernel<<<4 * H, 128, 0, s>>>(w_ih, w.wih, H, I);
  CSN_LAUNCH_CHECK();
  permute_rows_bf16_kernel<<<4 * H, 128, 0, s>>>(w_hh, w.whh, H, H);
  CSN_LAUNCH_CHECK();
  if (b_ih) {
    permute_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, s>>>(b_ih, b_hh, w.bias, H);
    CSN_LAUNCH_CHECK();
  }
  return CSN_OK;
}

int lstm_layer_fwd_large(const void* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* h_seq,
                         void* reserve, void* workspace, int T, int B, int I, int H, int training, cudaStream_t s) {
  const size_t tb = size_t(T) * B;
  LargeWs w = carve_ws(workspace, T, B, I, H);
  __nv_bfloat16* gates = reinterpret_cast<__nv_bfloat16*>(reserve);
  float* c_seq = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(reserve) + al256(tb * 4 * H * 2));
  __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(h_seq);
  CSN_TRY(prep_weights(w, w_ih, w_hh, b_ih, b_hh, I, H, s));
  if (lstm_cluster_supported(H)) {
    // The whole time loop in ONE persistent cluster launch.  The hoisted projection GEMM runs BESIDE it on the SMs the
    // clusters leave free, one time chunk per GEMM; a one-thread kernel behind each chunk raises its flag and the
    // recurrence's TMA producer waits for a chunk's flag before it touches that chunk's rows.
    OverlapStream* os = (T >= 4 * kChunks && lstm_cluster_overlap_ok(B, H)) ? overlap_stream() : nullptr;
    if (os) {
      const int tc = ceil_div(T, kChunks);
      CSN_CUDA(cudaMemsetAsync(w.flags, 0, kChunks * sizeof(unsigned), s));
      CSN_CUDA(cudaEventRecord(os->fork, s));
      CSN_CUDA(cudaStreamWaitEvent(os->st, os->fork, 0));
      for (int c = 0; c * tc < T; ++c) {
        const size_t r0 = size_t(c) * tc * B, rows = size_t(std::min(T, (c + 1) * tc) - c * tc) * B;
        CSN_TRY(gemm_tc_run(0, 1, (int)rows, 4 * H, I, reinterpret_cast<const __nv_bfloat16*>(x) + r0 * I, I, w.wih, I, w.xp + r0 * 4 * H,
                            4 * H, CSN_F32, w.bias, 0, 1, nullptr, os->st));
        set_flag_kernel<<<1, 1, 0, os->st>>>(w.flags + c);
        CSN_LAUNCH_CHECK();
        if (c == 0) CSN_CUDA(cudaEventRecord(os->first, os->st));
      }
      CSN_CUDA(cudaEventRecord(os->join, os->st));
      CSN_CUDA(cudaStreamWaitEvent(s, os->first, 0));
      CSN_TRY(lstm_cluster_fwd(w.xp, w.whh, hs, gates, c_seq, T, B, H, w.flags, tc, w.xchg, s));
      CSN_CUDA(cudaStreamWaitEvent(s, os->join, 0));
      return CSN_OK;
    }
    CSN_TRY(gemm_tc_run(0, 1, (int)tb, 4 * H, I, x, I, w.wih, I, w.xp, 4 * H, CSN_F32, w.bias, 0, 1, nullptr, s));
    return lstm_cluster_fwd(w.xp, w.whh, hs, gates, c_seq, T, B, H, nullptr, 1, w.xchg, s);
  }
  // hoisted input projection in gate-interleaved column order
  CSN_TRY(gemm_tc_run(0, 1, (int)tb, 4 * H, I, x, I, w.wih, I, w.xp, 4 * H, CSN_F32, w.bias, 0, 1, nullptr, s));
  for (int t = 0; t < T; ++t) {
    GemmEpi cell{};
    cell.pdl = kPdl;
    cell.zero_acc = (t == 0);
    cell.H = H;
    cell.xp = w.xp + size_t(t) * B * 4 * H;
    cell.c_prev = t ? c_seq + size_t(t - 1) * B * H : nullptr;
    cell.h_out = hs + size_t(t) * B * H;
    cell.gates_out = gates + size_t(t) * B * 4 * H;  // always kept: the reserve is also the scratch for c (c_seq below)
    cell.c_out = c_seq + size_t(t) * B * H;
    const __nv_bfloat16* a = t ? hs + size_t(t - 1) * B * H : hs;  // unused at t = 0 (zero_acc)
    CSN_TRY(gemm_tc_run(0, 1, B, 4 * H, H, a, H, w.whh, H, nullptr, 4 * H, CSN_F32, nullptr, 0, 1, &cell, s));
  }
  (void)training;
  return CSN_OK;
}

int lstm_layer_bwd_large(const void* x, const float* w_ih, const float* w_hh, const void* h_seq, const void* reserve,
                         const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                         float* dx, void* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s) {
  const size_t tb = size_t(T) * B;
  LargeWs w = carve_ws(workspace, T, B, I, H);
  const __nv_bfloat16* gates = reinterpret_cast<const __nv_bfloat16*>(reserve);
  const float* c_seq = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(reserve) + al256(tb * 4 * H * 2));
  CSN_TRY(prep_weights(w, w_ih, w_hh, nullptr, nullptr, I, H, s));
  const bool cluster = lstm_cluster_supported(H);
  if (cluster) {  // the whole BPTT recurrence in ONE persistent cluster launch (lstm_cluster.cu)
    permute_transpose_bf16_kernel<<<dim3(H / 32, 4 * H / 32), dim3(32, 8), 0, s>>>(w_hh, w.whh_t, H);
    CSN_LAUNCH_CHECK();
    OverlapStream* os = (T >= 4 * kChunks && lstm_cluster_overlap_ok(B, H)) ? overlap_stream() : nullptr;
    if (os) {
      // The weight-gradient / input-gradient GEMMs run BESIDE the recurrence on the SMs the clusters leave free, one time
      // chunk per GEMM in the order the recurrence produces dG (last chunk first).  Every (CTA, trial group) of the
      // recurrence counts itself into a chunk's counter once its dG rows of that chunk are stored; a one-warp kernel on
      // the helper stream parks on the counter in front of the chunk's GEMMs.  Each chunk's product is one slab; the
      // slabs are added in chunk order afterwards (fixed order: deterministic).
      const int tc = ceil_div(T, kChunks), nc = ceil_div(T, tc);
      unsigned* done = w.flags + kChunks;
      const unsigned target = (unsigned)lstm_cluster_participants(B, H);
      const int sms = sm_count();
      CSN_CUDA(cudaMemsetAsync(done, 0, kChunks * sizeof(unsigned), s));
      fill_bf16_kernel<<<min(ceil_div<size_t>(tb * 8, 256), size_t(sms) * 8), 256, 0, s>>>(w.ones, tb * 8, 1.f);
      CSN_LAUNCH_CHECK();
      CSN_CUDA(cudaEventRecord(os->fork, s));
      CSN_CUDA(cudaStreamWaitEvent(os->st, os->fork, 0));
      CSN_TRY(lstm_cluster_bwd(gates, c_seq, d_hseq, d_hlast, w.whh_t, w.dG, T, B, H, done, tc, s));
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(h_seq);
      // products with few output tiles are split over K inside a chunk so that a chunk's GEMM still fills the free SMs
      const int sp_ih = std::max(1, std::min(kChunkSplit, 64 / (ceil_div(4 * H, 128) * ceil_div(I, 128))));
      const int sp_b = std::max(1, std::min(kChunkSplit, 64 / ceil_div(4 * H, 128)));
      int n_hh = 0, n_ih = 0, n_b = 0;  // slabs written so far (a chunk's split count follows its own row count)
      GemmEpi slab{};
      for (int c = nc - 1; c >= 0; --c) {
        const size_t r0 = size_t(c) * tc * B, r1 = size_t(std::min(T, (c + 1) * tc)) * B, rr0 = std::max(r0, size_t(B));
        wait_counter_kernel<<<1, 32, 0, os->st>>>(done + c, target);
        CSN_LAUNCH_CHECK();
        slab.split_stride = size_t(4) * H * H;
        CSN_TRY(gemm_tc_run(1, 0, 4 * H, H, (int)(r1 - rr0), w.dG + rr0 * 4 * H, 4 * H, hb + (rr0 - B) * H, H,
                            w.dwp + size_t(n_hh) * slab.split_stride, H, CSN_F32, nullptr, 0, 1, &slab, os->st));
        n_hh += 1;
        if (dx)  // the layer below waits for this one: ahead of the remaining weight-gradient products
          CSN_TRY(gemm_tc_run(0, 0, (int)(r1 - r0), I, 4 * H, w.dG + r0 * 4 * H, 4 * H, w.wih, I, dx + r0 * I, I, CSN_F32, nullptr, 0,
                              1, nullptr, os->st));
        slab.split_stride = size_t(4) * H * I;
        CSN_TRY(gemm_tc_run(1, 0, 4 * H, I, (int)(r1 - r0), w.dG + r0 * 4 * H, 4 * H, xb + r0 * I, I,
                            w.dwp2 + size_t(n_ih) * slab.split_stride, I, CSN_F32, nullptr, 0, sp_ih, &slab, os->st));
        n_ih += gemm_tc_splits((int)(r1 - r0), sp_ih);
        slab.split_stride = size_t(4) * H * 8;
        CSN_TRY(gemm_tc_run(1, 0, 4 * H, 8, (int)(r1 - r0), w.dG + r0 * 4 * H, 4 * H, w.ones + r0 * 8, 8,
                            w.dbp + size_t(n_b) * slab.split_stride, 8, CSN_F32, nullptr, 0, sp_b, &slab, os->st));
        n_b += gemm_tc_splits((int)(r1 - r0), sp_b);
      }
      CSN_CUDA(cudaEventRecord(os->join, os->st));
      CSN_CUDA(cudaStreamWaitEvent(s, os->join, 0));
      unpermute_rows_kernel<<<4 * H, 128, 0, s>>>(w.dwp, dw_hh, H, H, H, accumulate, n_hh, size_t(4) * H * H);
      CSN_LAUNCH_CHECK();
      unpermute_rows_kernel<<<4 * H, 128, 0, s>>>(w.dwp2, dw_ih, H, I, I, accumulate, n_ih, size_t(4) * H * I);
      CSN_LAUNCH_CHECK();
      unpermute_rows_kernel<<<4 * H, 32, 0, s>>>(w.dbp, db_ih, H, 1, 8, accumulate, n_b, size_t(4) * H * 8);
      CSN_LAUNCH_CHECK();
      unpermute_rows_kernel<<<4 * H, 32, 0, s>>>(w.dbp, db_hh, H, 1, 8, accumulate, n_b, size_t(4) * H * 8);
      CSN_LAUNCH_CHECK();
      return CSN_OK;
    }
    CSN_TRY(lstm_cluster_bwd(gates, c_seq, d_hseq, d_hlast, w.whh_t, w.dG, T, B, H, nullptr, 1, s));
  } else {
    CSN_CUDA(cudaMemsetAsync(w.dc, 0, size_t(B) * H * 4, s));
  }
  const int cells = B * H;
  // dh_{t-1}[B,H] = dG_t[B,4H] . W_hh_perm[4H,H]: a contraction of 4H with only H/128 output tiles per batch tile, so
  // it is split over K to put up to 8x more CTAs on each step of the chain; every split stores its partial product to
  // its own slab (no atomics, no memset) and the cell kernel of the next step adds the slabs while it reads dh.
  const int sms0 = sm_count();
  int kparts = max(1, min(8, sms0 / max(1, ceil_div(B, 128) * ceil_div(H, 128))));
  kparts = min(kparts, ceil_div(4 * H, 64));
  {  // the GEMM rounds the split: use the count it will really launch
    const int k_iters = ceil_div(4 * H, 64), per = ceil_div(k_iters, kparts);
    kparts = ceil_div(k_iters, per);
  }
  GemmEpi slabs{};
  slabs.split_stride = size_t(B) * H;
  slabs.pdl = kPdl;
  for (int t = cluster ? -1 : T - 1; t >= 0; --t) {
    {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(ceil_div(cells, 256), 1, 1);
      cfg.blockDim = dim3(256, 1, 1);
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = kPdl;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const __nv_bfloat16* a_gates = gates + size_t(t) * B * 4 * H;
      const float* a_ct = c_seq + size_t(t) * B * H;
      const float* a_cp = t ? c_seq + size_t(t - 1) * B * H : nullptr;
      const float* a_dh = (t + 1 < T) ? w.dh_rec : nullptr;
      const float* a_dhs = d_hseq ? d_hseq + size_t(t) * B * H : nullptr;
      const float* a_dhl = (t == T - 1) ? d_hlast : nullptr;
      __nv_bfloat16* a_dg = w.dG + size_t(t) * B * 4 * H;
      CSN_CUDA(cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel, a_gates, a_ct, a_cp, a_dh, a_dhs, a_dhl, w.dc, a_dg, B, H, kparts));
      count_launches(1);
    }
    if (t > 0)
      CSN_TRY(gemm_tc_run(0, 0, B, H, 4 * H, w.dG + size_t(t) * B * 4 * H, 4 * H, w.whh, H, w.dh_rec, H, CSN_F32, nullptr, 0,
                          kparts, &slabs, s));
  }
  const int sms = sm_count();
  // dW_ih (permuted rows) = dG^T . x as split-K slabs, then add the slabs in split order while scattering the rows back
  // to PyTorch gate order
  GemmEpi dws{};
  int ns = dw_splits(4 * H, I, (int)tb);
  dws.split_stride = size_t(4) * H * I;
  CSN_TRY(gemm_tc_run(1, 0, 4 * H, I, (int)tb, w.dG, 4 * H, x, I, w.dwp, I, CSN_F32, nullptr, 0, ns, &dws, s));
  unpermute_rows_kernel<<<4 * H, 128, 0, s>>>(w.dwp, dw_ih, H, I, I, accumulate, ns, dws.split_stride);
  CSN_LAUNCH_CHECK();
  if (T > 1) {
    ns = dw_splits(4 * H, H, (int)(tb - B));
    dws.split_stride = size_t(4) * H * H;
    CSN_TRY(gemm_tc_run(1, 0, 4 * H, H, (int)(tb - B), w.dG + size_t(B) * 4 * H, 4 * H, h_seq, H, w.dwp, H, CSN_F32, nullptr, 0,
                        ns, &dws, s));
    unpermute_rows_kernel<<<4 * H, 128, 0, s>>>(w.dwp, dw_hh, H, H, H, accumulate, ns, dws.split_stride);
    CSN_LAUNCH_CHECK();
  } else if (!accumulate) {
    CSN_CUDA(cudaMemsetAsync(dw_hh, 0, size_t(4) * H * H * 4, s));
  }
  // db = dG^T . 1  (tensor-core column sums of the bf16 dG)
  fill_bf16_kernel<<<min(ceil_div<size_t>(tb * 8, 256), size_t(sms) * 8), 256, 0, s>>>(w.ones, tb * 8, 1.f);
  CSN_LAUNCH_CHECK();
  ns = dw_splits(4 * H, 8, (int)tb);
  dws.split_stride = size_t(4) * H * 8;
  CSN_TRY(gemm_tc_run(1, 0, 4 * H, 8, (int)tb, w.dG, 4 * H, w.ones, 8, w.dbp, 8, CSN_F32, nullptr, 0, ns, &dws, s));
  unpermute_rows_kernel<<<4 * H, 32, 0, s>>>(w.dbp, db_ih, H, 1, 8, accumulate, ns, dws.split_stride);
  CSN_LAUNCH_CHECK();
  unpermute_rows_kernel<<<4 * H, 32, 0, s>>>(w.dbp, db_hh, H, 1, 8, accumulate, ns, dws.split_stride);
  CSN_LAUNCH_CHECK();
  if (dx) CSN_TRY(gemm_tc_run(0, 0, (int)tb, I, 4 * H, w.dG, 4 * H, w.wih, I, dx, I, CSN_F32, nullptr, 0, 1, nullptr, s));
  return CSN_OK;
}

}  // namespace csn
