// LSTM layer, bf16 tensor-core path: persistent tcgen05 recurrence + hand-written BPTT (H <= 128).
// Replaces torch.nn.LSTM inside models.lstm.Model (call sites LstmDistillFromDinoV2Train.py:323,365,374).
//
// Forward (one CTA per batch tile of NV trials, the whole time loop inside the kernel):
//   * W_hh is converted to bf16 once and stays RESIDENT IN TENSOR MEMORY for all T steps (256 of the 512 columns:
//     lane = hidden unit, gate g at columns 256 + 64 g, packed bf16 pairs) and is the TMEM A operand of tcgen05.mma,
//     so a thread owns all four gates of its unit -- no cross-thread exchange in the epilogue.
//   * per step one elected thread of the issuer warp runs 4*KP/16 tcgen05.mma (M=128 units, N=16 batch slots, K=16):
//     gates^T[g] = W_hh[g] . h_{t-1}^T, fp32 accumulators in TMEM.
//   * the input projection x_t W_ih^T is computed by the same CTA on the tensor pipe's idle slots (W_ih resident in
//     shared memory, 16/NV timesteps per MMA block, a dedicated warp loads x and issues those MMAs; see
//     lstm_fwd_tc_kernel) -- or, for I > 128, read from a hoisted projection GEMM through a TMA ring.
//   * eight epilogue warps tcgen05.ld their lanes, add projection + biases, apply sigmoid/tanh (tanh.approx), update the
//     fp32 cell state held in REGISTERS, write h_t as bf16 straight into the shared-memory operand layout the next
//     step's MMA reads (generic-proxy store -> fence.proxy.async -> named-barrier hand-off), and only then stream
//     h_t / gate activations / c_t to HBM for BPTT.
// Backward mirrors it: W_hh^T resident in tensor memory (M = hidden unit k, gate-interleaved contraction index 4u+g),
//   per step dh_{t-1}^T = W_hh^T . dG_t^T on tcgen05 (two issuer warps), gate derivatives + dc carry in registers, the
//   step's inputs arrive through a TMA ring fed by a producer warp, dG_t is streamed to HBM in bf16; dW_ih / dW_hh / dX
//   are then large tcgen05 GEMMs (gemm_tc.cu, the two dW products side by side), db is accumulated in registers.
//
// The serial chain (T steps) cannot be parallelised; what this design optimises is the per-step latency:
// no HBM round trip, no grid-wide sync, one hand-off in each direction per step (forward ~990, backward ~690 cycles).
#include <cuda_fp16.h>

#include <mutex>
#include <type_traits>

#include "tc.cuh"
#include "gemm_tc.cuh"
#include "lstm_side.cuh"

namespace csn {

using namespace tc;

// warps 0-7: epilogue -- warp w owns TMEM lane quadrant w % 4 (32 hidden units) and half (w / 4) of the CTA's batch
// slots, so the per-step gate math / reserve traffic of a unit is split over two warps; warp 8: MMA issuer (backward:
// warps 8-9); forward warp 9: projection loader / issuer; backward warp 10: TMA producer of the per-step inputs.
constexpr int kEpiWarps = 8, kIssuerWarp = 8;
constexpr int kRecThreads = (kEpiWarps + 1) * 32;     // forward: one issuer warp
constexpr int kRecThreadsBwd = (kEpiWarps + 2) * 32;  // backward: two issuer warps (see issue_bwd); forward: issuer + feeder warp
constexpr int kBwdThreads = kRecThreadsBwd + 64;      // + the TMA producer warp of the per-step BPTT inputs + the publisher warp
constexpr int kDgBatch = 4;                           // backward: steps published to the dW consumers per gpu-scope release
constexpr int kDgLag = 6;                             // backward: a step's dG is published to the dW consumers this many steps later
// epilogue -> issuer hand-off: hardware named barrier 1 over all threads (the 256 epilogue threads ARRIVE without
// blocking, the issuer warp(s) SYNC); ~2x lower latency than an mbarrier round trip, measured on the step stamps.
template <int NTHREADS>
__device__ __forceinline__ void handoff_arrive() { asm volatile("bar.arrive 1, %0;" ::"n"(NTHREADS) : "memory"); }
template <int NTHREADS>
__device__ __forceinline__ void handoff_wait() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }
// 16-row operand: 2 core matrices (2 x 128 B) per k-group.  The k-group stride is 272 B, not the dense 256: the epilogue
// threads write their h / dG values one k-group apart per 2..8 lanes, and a 256-byte stride puts every k-group on the same
// shared-memory banks (4-way conflict for the forward's 2-byte stores, 16-way for the backward's 8-byte stores, and the
// MEMBAR of the hand-off waits for all of it); 272 = 256 + 16 rotates consecutive k-groups over the bank groups.
constexpr uint32_t kLboB = 272, kSboB = 128;
constexpr int kNslots = 16;                    // MMA N (batch slots per CTA); NV <= 16 of them are live
constexpr int kBwdAcc = 4;                     // accumulators the backward K-steps rotate over
constexpr int kProfSteps = 64;                 // bring-up instrumentation: clock64 stamps for the first steps of CTA 0

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

constexpr uint32_t kBarBlock = 256;  // mbarriers + shared-memory step counters between the MMA operand and the rings
struct RecSmem {
  uint8_t* opb;       // B operand: h^T (fwd, KP/8 * 256 B) or dG^T (bwd, 4KP/8 * 256 B), canonical K-major
  uint64_t* bar_acc;  // accumulators ready (issuer -> epilogue)
  uint64_t* bar_pf;   // [kPfStages] prefetch ring stages filled (TMA bulk copies -> epilogue)
  uint64_t* bar_xp;   // [2] fused input projection: Xp accumulator set filled (issuer -> epilogue / x-ring reuse)
  uint64_t* bar_w;    // fused input projection: W_ih operand image landed in shared memory
  uint32_t* tmem_slot;
  uint8_t* ring;      // prefetch ring / fused-projection operands (after a 128-byte barrier block)
};
constexpr int kPfStages = 8;  // prefetch ring depth: seven steps (~2.5 us at 0.35 us per step) ahead of the consumer covers the HBM latency under load
constexpr int kBwdPfRow = 2048;  // backward prefetch ring, bytes per trial and stage: gates (<= 1 KB) | c_prev (<= 512 B) | d_hseq (<= 512 B)

// `raw` is the 128-byte aligned dynamic shared array itself (no runtime alignment arithmetic): its address is a
// link-time constant, so the B-operand descriptors of the unrolled MMA issue become immediates in uniform registers.
__device__ __forceinline__ RecSmem carve(uint8_t* raw, size_t b_bytes) {
  uint8_t* base = raw;
  RecSmem s;
  s.opb = base;
  s.bar_acc = reinterpret_cast<uint64_t*>(s.opb + b_bytes) + 1;  // (first word of the block: spare)
  s.bar_pf = s.bar_acc + 1;
  s.bar_xp = s.bar_pf + kPfStages;
  s.bar_w = s.bar_xp + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bar_w + 1);
  s.ring = s.opb + b_bytes + kBarBlock;
  return s;
}

// Tensor-memory map (512 columns allocated; lane = hidden unit):
//   [0, 64)    fp32 accumulators, 16 columns (batch slots) per gate (forward) / one dh^T accumulator (backward)
//   [256, 512) resident weights as packed bf16 pairs: gate g at columns 256 + 64 g + k/2
//   [64, 192)  (forward, fused input projection) two sets of 4 x 16 columns: x_t W_ih^T of the current / next block of
//              16/NV timesteps, column = (timestep in block) * NV + batch slot
constexpr uint32_t kTmemCols = 512, kAcol0 = 256, kAgate = 64, kXpCol0 = 64;
constexpr uint32_t kLboA = 2048, kSboA = 128;  // W_ih operand in shared memory: 128 rows per k-group (K-major canonical)

// W_ih [4H, I] fp32 -> the bf16 shared-memory image of the four 128 x KI K-major A-operand blocks (zero padded), so a
// recurrence CTA stages it with four bulk copies.
__device__ __forceinline__ void pack_wih_element(const float* __restrict__ w_ih, __nv_bfloat16* __restrict__ img, int H, int I,
                                                 int KI, int idx) {  // idx over 4 * 128 * KI, k fastest
  if (idx >= 4 * 128 * KI) return;
  const int k = idx % KI, r = (idx / KI) & 127, g = idx / (KI * 128);
  const float v = (r < H && k < I) ? w_ih[size_t(g * H + r) * I + k] : 0.f;
  img[(size_t(g) * 128 * KI * 2 + canon_k_off(r, k, kLboA, kSboA)) / 2] = __float2bfloat16_rn(v);
}

// The resident weight operand is prepared once per call by a small packing kernel as the exact TMEM image
// [gate][lane][64 columns] of packed bf16 pairs (column c = K elements 2c, 2c+1; zero padded), so that every
// recurrence CTA fills its tensor memory with 64 coalescable 16-byte loads per thread instead of 512 strided
// scalar fp32 loads.  transposed = 0: forward operand, lane = hidden unit u, K = input unit k  (W_hh[g*H+u][k]);
// transposed = 1: backward operand, lane = k, K = u (W_hh^T).
__device__ __forceinline__ void pack_whh_element(const float* __restrict__ w_hh, uint32_t* __restrict__ img, int H, int transposed,
                                                 int idx) {  // idx over 4 * 128 * 64
  if (idx >= 4 * 128 * 64) return;
  if (!transposed) {
    const int c = idx & 63, lane = (idx >> 6) & 127, g = idx >> 13;  // consecutive threads -> consecutive k
    const int k0 = 2 * c, k1 = 2 * c + 1;
    float v0 = 0.f, v1 = 0.f;
    if (lane < H) {
      if (k0 < H) v0 = w_hh[size_t(g * H + lane) * H + k0];
      if (k1 < H) v1 = w_hh[size_t(g * H + lane) * H + k1];
    }
    __nv_bfloat162 bb = __floats2bfloat162_rn(v0, v1);
    img[(size_t(g) * 128 + lane) * 64 + c] = *reinterpret_cast<uint32_t*>(&bb);
    return;
  }
  // backward operand: lane = m (hidden unit of dh), contraction index K' = 4 u + g (GATE-INTERLEAVED, so the four gate
  // gradients of a cell are 8 contiguous bytes of the dG^T operand and leave the epilogue as one 64-bit store).
  // Column c' in [0, 256) of the lane's row holds K' = 2c', 2c'+1 = gates (0,1) or (2,3) of unit u = c' / 2; the image
  // keeps the [4][128][64] addressing of the forward one: column c' lives in block c' / 64.
  const int lane = idx & 127, cp = idx >> 7;  // consecutive threads -> consecutive m (contiguous in W_hh rows)
  const int u = cp >> 1, g0 = (cp & 1) * 2;
  float v0 = 0.f, v1 = 0.f;
  if (lane < H && u < H) {
    v0 = w_hh[size_t(g0 * H + u) * H + lane];
    v1 = w_hh[size_t((g0 + 1) * H + u) * H + lane];
  }
  __nv_bfloat162 bb = __floats2bfloat162_rn(v0, v1);
  img[(size_t(cp >> 6) * 128 + lane) * 64 + (cp & 63)] = *reinterpret_cast<uint32_t*>(&bb);
}

// ONE preparation launch per layer call: the W_hh tensor-memory image, plus (forward) W_ih as the fused projection's
// shared-memory image or as the plain bf16 [4H, I] matrix the Xp servers' TMA reads, plus the zeroing of the launch's
// flag words (Xp tile READY flags / per-timestep dG publication counters) -- tiny launches folded into one (each
// boundary costs ~2.5 us of the step).
__global__ void lstm_prepare_kernel(const float* __restrict__ w_hh, uint32_t* __restrict__ whh_img, int H, int transposed,
                                    const float* __restrict__ w_ih, __nv_bfloat16* __restrict__ wih_img, int I, int KI,
                                    __nv_bfloat16* __restrict__ wih_bf, unsigned* __restrict__ zero_words, int n_zero) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  pack_whh_element(w_hh, whh_img, H, transposed, idx);
  if (wih_img) pack_wih_element(w_ih, wih_img, H, I, KI, idx);
  if (wih_bf && idx < 4 * H * I) wih_bf[idx] = __float2bfloat16_rn(w_ih[idx]);
  if (zero_words && idx < n_zero) zero_words[idx] = 0u;
}

// image -> tensor memory: thread `row` (TMEM lane) copies its 4 x 64 packed columns
__device__ __forceinline__ void stage_weights_tmem(uint32_t lane_addr, const uint32_t* __restrict__ img, int row) {
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const uint4* src = reinterpret_cast<const uint4*>(img + (size_t(g) * 128 + row) * 64);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(src + half * 8 + q);
        r[q * 4 + 0] = v.x; r[q * 4 + 1] = v.y; r[q * 4 + 2] = v.z; r[q * 4 + 3] = v.w;
      }
      tmem_st32(lane_addr + kAcol0 + g * kAgate + half * 32, r);
    }
  }
  tmem_st_wait();
}

// One timestep's MMAs, fully unrolled so every TMEM address is an immediate (CONST_BASE) or base + immediate.
// All K-steps of a gate accumulate into ONE accumulator: the tensor pipe does not penalise back-to-back MMAs into the
// same D (measured: 14.9 cycles per M=128,N=16,K=16 MMA with A in tensor memory, 43.8 with A in shared memory,
// independent of the accumulator rotation), so split accumulation chains only add tcgen05.ld work.
// Forward: gate g, K-step kk  ->  D = acc[g], A = W_hh block g columns kk*8.., B = h^T k-groups 2kk, 2kk+1.
// KSTEPS > 0: K-steps per gate known at compile time (no predicates in the issue sequence); 0: runtime.
template <bool CONST_BASE, int KSTEPS>
__device__ __forceinline__ void issue_fwd(uint32_t base, uint64_t db0, uint32_t idesc, int ksteps_rt) {
  const uint32_t tb = base;  // CONST_BASE: `base` is a kernel-parameter zero, i.e. a UNIFORM value: addresses stay in the uniform datapath
  const int ksteps = KSTEPS ? KSTEPS : ksteps_rt;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    if (kk < ksteps) {
      const uint64_t db = db0 + uint64_t(kk * ((2 * kLboB) >> 4));
#pragma unroll
      for (int g = 0; g < 4; ++g)
        umma_f16_ts(tb + g * kNslots, tb + kAcol0 + g * kAgate + kk * 8, db, idesc, kk == 0 ? 0u : 1u);
    }
  }
}
// Backward: the contraction index of dh^T = W_hh^T . dG^T is gate-interleaved (K' = 4 u + g), 4*KP elements = KP/4
// K-steps, contiguous from zero.  The K-steps rotate over kBwdAcc accumulators (summed by the epilogue): with all 128
// SMs running, back-to-back MMAs into ONE accumulator measured ~23 cycles each against ~11 when consecutive MMAs
// target different accumulators (the forward kernel's four gate accumulators).
// HALF selects the issuer warp: warp 0 issues the first half of the K-steps into accumulators {0,1}, warp 1 the second
// half into {2,3}.  Everything is unrolled to immediates so each warp's ~50 operands (16 B descriptors, 16 A addresses,
// 2 accumulators) stay in its own uniform register file: one warp issuing all 32 K-steps needs ~100 uniform registers
// and ptxas spills them (MOV.SPILL / R2UR per MMA, ~23 cycles each); a rolled loop with running operands is worse
// still (~37 cycles per MMA: UIADD3 -> UTCHMMA dependency latency).  Measured with scripts/prof_lstm_steps.py.
// KSTEPS = KP / 16 (K-steps per gate in the forward kernel's terms): each half issues 2 * KSTEPS steps.
template <bool CONST_BASE, int KSTEPS, int HALF>
__device__ __forceinline__ void issue_bwd(uint32_t base, uint64_t db0, uint32_t idesc, int ksteps_rt) {
  const uint32_t tb = base;
  const int per_half = 2 * (KSTEPS ? KSTEPS : ksteps_rt);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < per_half) {
      const int kk = HALF * per_half + i;  // compile-time when KSTEPS > 0
      const int acc = HALF * 2 + (i & 1);
      const uint64_t db = db0 + uint64_t(kk * ((2 * kLboB) >> 4));
      umma_f16_ts(tb + acc * kNslots, tb + kAcol0 + kk * 8, db, idesc, i < 2 ? 0u : 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// FX = fused input projection: instead of reading a precomputed Xp[T,B,4H] from HBM, the CTA multiplies its own
// trials' inputs by W_ih on the tensor pipe's idle slots.  W_ih (bf16) is resident in shared memory as the A operand,
// the inputs of a BLOCK of TB = 16/NV timesteps x NV trials form one 16-row B operand (cp.async by the issuer warp,
// two blocks ahead), and the 4*KI/16 MMAs of block n+1 are issued a few per step behind the recurrent MMAs of block
// n into the second of two Xp accumulator sets in tensor memory; the epilogue adds its column of the current set
// with one more tcgen05.ld per gate.  No Xp round trip through HBM (230 MB written + read per cfg2 step), no separate
// projection GEMM.
// Three ways the input projection x_t W_ih^T reaches the chain:
//   FX           : this CTA computes it on its own tensor pipe between the recurrent MMAs (any SM budget);
//   !FX, servers : the launch's spare CTAs (blocks [0, sv.n_srv)) compute it concurrently, tile by tile in time order
//                  (lstm_side.cuh); warp 9 of every recurrence CTA waits on the tiles' READY flags and pulls the CTA's
//                  rows through the TMA ring -- the chain carries no projection MMAs at all;
//   !FX, no servers: a GEMM computed Xp before the launch (inputs wider than 128 features).
// PROF compiles the clock64 stamps of scripts/prof_lstm_steps.py in; the shipped instantiations carry none.
template <int NV, int KSTEPS, bool FX, bool PROF>
__global__ void __launch_bounds__(kRecThreadsBwd, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const side::XpServe sv,
                   const float* __restrict__ xp, const uint32_t* __restrict__ w_img, const float* __restrict__ b_hh,
                   __nv_bfloat16* __restrict__ h_seq, __nv_bfloat16* __restrict__ gates_out, float* __restrict__ c_out,
                   int T, int B, int H, int KP, long long* __restrict__ prof, const uint32_t uz,
                   const __nv_bfloat16* __restrict__ x_in, const uint8_t* __restrict__ wih_img,
                   const float* __restrict__ b_ih, int I, int KI) {
  // Warp 9 is the chain's feeder.  FX: the input-projection issuer / x loader, with its own uniform register file
  // (sharing the recurrent issuer's warp pushed that warp's MMA operands out of uniform registers: one R2UR per MMA on
  // the chain).  !FX: the TMA producer of the Xp ring.
  constexpr int kThreads = kRecThreadsBwd;
  constexpr int kXWarp = kIssuerWarp + 1;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  if constexpr (!FX) {
    if ((int)blockIdx.x < sv.n_srv) {
      side::xp_server_role(smem_raw, &tm_x, &tm_w, sv, (int)blockIdx.x);
      return;
    }
  }
  const int cta = FX ? (int)blockIdx.x : (int)blockIdx.x - sv.n_srv;  // recurrence CTA index
  const size_t b_bytes = size_t(KP / 8) * kLboB;
  RecSmem sm = carve(smem_raw, b_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = cta * NV;
  const int ksteps = KP / 16;
  const bool phase_prof = PROF && prof && cta == 0 && tid == 0;
  if (phase_prof) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    prof[1024 + 0] = (long long)gt;
    prof[1024 + 1] = clock64();
  }

  for (int i = tid; i < (int)(b_bytes / 4); i += kThreads) reinterpret_cast<uint32_t*>(sm.opb)[i] = 0u;
  if (tid == 0) {
    mbar_init(sm.bar_acc, 1);
    for (int i = 0; i < kPfStages; ++i) mbar_init(sm.bar_pf + i, 1);
    mbar_init(sm.bar_xp, 1);
    mbar_init(sm.bar_xp + 1, 1);
    mbar_init(sm.bar_w, 1);
    *reinterpret_cast<volatile int*>(sm.tmem_slot + 1) = 0;
    *reinterpret_cast<volatile int*>(sm.tmem_slot + 2) = 0;
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) tmem_alloc(sm.tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;
  // fused-projection operands: two x blocks (16 rows x KI, same canonical layout as the h operand), then W_ih
  constexpr int TB = kNslots / NV;  // timesteps per Xp block
  const uint32_t x_bytes = uint32_t(KI / 8) * kLboB;
  uint8_t* const xbuf = sm.ring;
  uint8_t* const wih_s = sm.ring + 2 * x_bytes;
  volatile int* const issued_step = reinterpret_cast<volatile int*>(sm.tmem_slot + 1);  // last step whose MMAs are queued
  volatile int* const acc_done = reinterpret_cast<volatile int*>(sm.tmem_slot + 2);     // last step whose accumulators the epilogue has seen complete
  if (FX && tid == kXWarp * 32) {
    const uint32_t gate_bytes = 128u * uint32_t(KI) * 2u;
    mbar_arrive_expect_tx(sm.bar_w, 4 * gate_bytes);
    for (int g = 0; g < 4; ++g) bulk_g2s(wih_s + size_t(g) * gate_bytes, wih_img + size_t(g) * gate_bytes, gate_bytes, sm.bar_w);
  }
  if (phase_prof) prof[1024 + 2] = clock64();
  // W_hh resident in TENSOR MEMORY for the whole sequence: lane = hidden unit u (row of each gate block)
  if (warp < 4) stage_weights_tmem(tmem_base + (uint32_t(warp * 32) << 16), w_img, tid);
  constexpr int NVT = NV / 2;  // cells per epilogue thread
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (phase_prof) prof[1024 + 3] = clock64();

  if (warp == kIssuerWarp) {
    // ================= MMA issuer: whole warp converged, one ELECTED lane issues =================
    // (elect.sync lets ptxas feed tcgen05.mma from uniform registers directly; a plain `lane == 0` guard makes it
    //  emit an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall around every MMA, ~60 cycles each, measured.)
    constexpr uint32_t idesc = make_idesc_bf16(128, kNslots, 0, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(sm.opb), kLboB, kSboB, kLayoutNone);
    const bool base0 = (tmem_base == 0);  // a 512-column allocation owns the whole TMEM: constant addresses
    for (int t = 1; t <= T; ++t) {
      handoff_wait<kRecThreads>();  // h_{t-1} is in shared memory (and TMEM has been drained)
      if (t == T) break;  // the last hand-off only balances the barrier
      tcgen05_fence_after();
      if (elect_one()) {
        if constexpr (PROF) { if (prof && cta == 0 && t < kProfSteps) prof[t * 8 + 4] = clock64(); }
        if (base0) issue_fwd<true, KSTEPS>(uz, db0, idesc, ksteps);
        else issue_fwd<false, KSTEPS>(tmem_base, db0, idesc, ksteps);
        umma_commit(sm.bar_acc);
        if constexpr (PROF) { if (prof && cta == 0 && t < kProfSteps) prof[t * 8 + 5] = clock64(); }
        // the feeder warp follows through this counter: FX queues its projection MMAs behind this step's; !FX knows
        // that every epilogue thread has consumed ring stage (t - 1) % kPfStages (the hand-off above) and refills it
        *issued_step = t;
      }
      __syncwarp();
    }
  } else if (!FX && warp == kXWarp) {
    // ================= Xp ring producer (own warp: a cp.async.bulk costs its issuing thread a few hundred cycles) =========
    // The NV rows Xp[t][b0 .. b0 + NV) (4H floats each, adjacent) are ONE bulk copy into a kPfStages-deep shared-memory
    // ring, up to kPfStages - 1 steps ahead of the chain; completion is counted on an mbarrier per stage, so the prefetch
    // never touches the epilogue warps' scoreboards.  With Xp servers in the launch the rows exist only once their
    // 128-row tile has been published: the lanes check the flags of the next eight timesteps with ONE round trip to L2.
    const int rows_valid = min(NV, B - b0);
    const uint32_t esz = sv.n_srv > 0 ? 2u : 4u;  // the servers write fp16 (without the bias), the hoisted GEMM fp32
    const uint32_t bytes = uint32_t(rows_valid) * uint32_t(4 * H) * esz;
    for (int t0 = 0; t0 < T; t0 += 8) {
      if (sv.n_srv > 0) {
        const int tt = min(t0 + (lane & 7), T - 1);
        const long long r0 = (long long)tt * B + b0;
        const unsigned* f = sv.flags + ((lane & 8) ? ((r0 + rows_valid - 1) >> 7) : (r0 >> 7));
        if (lane < 16) side::wait_ge(f, 1u);
        __syncwarp();
        side::fence_proxy_async_all();
      }
      if (lane == 0) {
        for (int t = t0; t < min(T, t0 + 8); ++t) {
          if (t >= kPfStages) {
            const int need = t - kPfStages + 1;
            while (*issued_step < need) __nanosleep(32);
          }
          uint64_t* bar = sm.bar_pf + (t & (kPfStages - 1));
          mbar_arrive_expect_tx(bar, bytes);
          bulk_g2s(reinterpret_cast<float*>(sm.ring) + size_t(t & (kPfStages - 1)) * NV * 4 * H,
                   reinterpret_cast<const uint8_t*>(xp) + (size_t(t) * B + b0) * 4 * H * esz, bytes, bar);
        }
      }
      __syncwarp();
    }
  } else if (FX && warp == kXWarp) {
    // ================= fused input projection: x loader + Xp MMA issuer (own warp, own uniform registers) ==========
    constexpr uint32_t idesc = make_idesc_bf16(128, kNslots, 0, 0);
    const int n_blocks = (T + TB - 1) / TB;
    const int n_x = 4 * (KI / 16);                 // Xp MMAs per block, k-step major / gate minor
    const int per_step = (n_x + TB - 1) / TB;
    const uint64_t dx0 = make_smem_desc(smem_u32(xbuf), kLboB, kSboB, kLayoutNone);
    const uint64_t da0 = make_smem_desc(smem_u32(wih_s), kLboA, kSboA, kLayoutNone);
    // all lanes: 16-byte cp.async chunks of block m's rows (zero fill out of range); lane = (row parity, k-chunk)
    auto load_x_block = [&](int m) {
      uint8_t* dst = xbuf + size_t(m & 1) * x_bytes;
      const int kc = lane & 15, sub = lane >> 4;
      const bool kc_ok = kc * 8 < I;
#pragma unroll
      for (int i = 0; i < kNslots / 2; ++i) {
        const int row = 2 * i + sub;
        const int tt = row / NV, j = row % NV;  // NV is a power of two
        const int t = m * TB + tt, b = b0 + j;
        const bool ok = kc_ok && (t < T) && (b < B);
        const __nv_bfloat16* src = ok ? x_in + (size_t(t) * B + b) * I + kc * 8 : x_in;
        const uint32_t d = smem_u32(dst + uint32_t(kc) * kLboB + uint32_t(row >> 3) * kSboB + uint32_t(row & 7) * 16);
        const int nbytes = ok ? 16 : 0;
        if (kc * 8 < KI) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto x_landed = [&]() {  // all lanes: every cp.async of this warp is complete and visible to the tensor pipe
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
    };
    auto issue_x_chunk = [&](int m, int c) {  // elected lane: chunk c of block m's projection into set m & 1
      const int set = m & 1;
      const int q1 = min(n_x, (c + 1) * per_step);
      for (int q = c * per_step; q < q1; ++q) {
        const int kk = q >> 2, g = q & 3;
        umma_f16(tmem_base + kXpCol0 + uint32_t(set) * 64u + uint32_t(g) * kNslots,
                 da0 + uint64_t((uint32_t(g) * 128u * uint32_t(KI) * 2u + uint32_t(kk) * 2u * kLboA) >> 4),
                 dx0 + uint64_t((uint32_t(set) * x_bytes + uint32_t(kk) * 2u * kLboB) >> 4), idesc, kk != 0);
      }
    };
    load_x_block(0);
    if (n_blocks > 1) load_x_block(1);
    mbar_wait(sm.bar_w, 0);
    x_landed();
    if (elect_one()) {
      for (int c = 0; c < TB; ++c) issue_x_chunk(0, c);
      umma_commit(sm.bar_xp);
      if (n_blocks > 1) issue_x_chunk(1, 0);
    }
    __syncwarp();
    int blk = 0, c = 0;  // block / position in block of step t
    for (int t = 1; t <= T; ++t) {
      if (++c == TB) { c = 0; ++blk; }
      if (t == T) break;
      const bool next_blk = blk + 1 < n_blocks;
      if (c == 0 && next_blk) x_landed();  // block blk+1's inputs (requested one block ago) are in place
      // Use the tensor pipe's idle window: wait until this step's recurrent MMAs have COMPLETED (the accumulator
      // barrier the epilogue waits on), then queue the chunk; it is done before the next step's MMAs are issued.
      // Queued right behind the recurrent ISSUE instead, the pipe interleaved the chunk with the tail of the recurrent
      // batch (+100 cycles on the chain).  Not the hand-off barrier either: a bar.sync drains this warp's in-flight
      // cp.async (one HBM round trip on the chain per block).  A parity wait can miss a phase if the waiter lags two
      // phases behind, so it is skipped when the epilogue's monotonic counter says the phase is already over.
      // Completion of step t also certifies that every epilogue thread finished step t-1, the last reader of the
      // accumulator set this warp starts to overwrite.
      if (*acc_done < t) mbar_wait(sm.bar_acc, (t - 1) & 1);
      tcgen05_fence_after();
      if (next_blk && elect_one()) {
        issue_x_chunk(blk + 1, c);
        if (c == TB - 1) umma_commit(sm.bar_xp + ((blk + 1) & 1));
      }
      __syncwarp();
      if (c == 1 && blk + 2 < n_blocks) {
        // block blk's projection is complete (its inputs are dead): fetch block blk+2 into the same buffer
        mbar_wait(sm.bar_xp + (blk & 1), (blk >> 1) & 1);
        load_x_block(blk + 2);
      }
    }
  } else {
    // ================= epilogue: thread = (hidden unit u, half of the batch slots) =================
    const int u = tid & 127;
    const int jb = (tid >> 7) * NVT;  // first batch slot of this thread
    const bool active = u < H;
    const uint32_t lane_addr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + jb;
    float bias[4], c[NVT];
    const float* xring = reinterpret_cast<const float*>(sm.ring);  // [4 stages][NV rows][4H] filled by the TMA producer
#pragma unroll
    const bool xp_half = !FX && sv.n_srv > 0;  // served Xp: fp16 without the bias
    for (int g = 0; g < 4; ++g) bias[g] = active ? (b_hh[g * H + u] + ((FX || xp_half) ? b_ih[g * H + u] : 0.f)) : 0.f;
#pragma unroll
    for (int j = 0; j < NVT; ++j) c[j] = 0.f;
    int xblk = 0, xtt = 0;  // FX: Xp block / timestep within the block of step t
    bool valid[NVT];
#pragma unroll
    for (int j = 0; j < NVT; ++j) valid[j] = active && (b0 + jb + j < B);
    // running pointers (advanced by one timestep per iteration): no 64-bit index arithmetic inside the loop
    const size_t step_cells = size_t(B) * H;
    __nv_bfloat16* h_ptr[NVT];        // h_seq[t][b][u]
    uint2* g_ptr[NVT];                // reserve gates [t][b][u][i,f,g,o] bf16 (one 8-byte store per cell)
    float* c_ptr[NVT];                // reserve cell state [t][b][u]
#pragma unroll
    for (int j = 0; j < NVT; ++j) {
      const size_t row = size_t(b0 + jb + j);
      h_ptr[j] = h_seq + (valid[j] ? row * H + u : 0);
      g_ptr[j] = reinterpret_cast<uint2*>(gates_out) + (valid[j] ? row * H + u : 0);
      c_ptr[j] = c_out + (valid[j] ? row * H + u : 0);
    }
    for (int t = 0; t < T; ++t) {
     {
      float pre[4][NVT];
      uint32_t xr[4][NVT];
      if constexpr (FX) {
        if (xtt == 0) {
          mbar_wait(sm.bar_xp + (xblk & 1), (xblk >> 1) & 1);  // this block's x_t W_ih^T is in tensor memory
          tcgen05_fence_after();
        }
        const uint32_t xaddr = lane_addr + kXpCol0 + uint32_t(xblk & 1) * 64u + uint32_t(xtt * NV);
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld<NVT>(xaddr + g * kNslots, xr[g]);  // waited for together with the accumulators
        if (++xtt == TB) { xtt = 0; ++xblk; }
      } else {
        mbar_wait(sm.bar_pf + (t & (kPfStages - 1)), (t / kPfStages) & 1);  // Xp rows of step t have landed (long ago)
        const float* stage = xring + size_t(t & (kPfStages - 1)) * NV * 4 * H;
        if (xp_half) {
          const __half* src = reinterpret_cast<const __half*>(stage) + u;
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < NVT; ++j) pre[g][j] = (valid[j] ? __half2float(src[(jb + j) * 4 * H + g * H]) : 0.f) + bias[g];
        } else {
          const float* src = stage + u;
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < NVT; ++j) pre[g][j] = (valid[j] ? src[(jb + j) * 4 * H + g * H] : 0.f) + bias[g];
        }
      }
      const bool do_prof = PROF && prof && cta == 0 && tid == 0 && t < kProfSteps;
      if (phase_prof && (t == 1 || t == 8 || t == 64 || t == 200 || t == 400)) prof[1024 + 6 + (t == 1 ? 0 : t == 8 ? 1 : t == 64 ? 2 : t == 200 ? 3 : 4)] = clock64();
      if (t > 0) {
        if (do_prof) prof[t * 8 + 6] = clock64();
        mbar_wait(sm.bar_acc, (t - 1) & 1);
        tcgen05_fence_after();
        if (FX && tid == 0) *acc_done = t;
        if (do_prof) prof[t * 8 + 0] = clock64();
        uint32_t r[4][NVT];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld<NVT>(lane_addr + g * kNslots, r[g]);
        if (do_prof) prof[t * 8 + 7] = clock64();
        tmem_ld_wait();
        tcgen05_fence_before();  // our TMEM reads are ordered before the MMAs the next hand-off releases
        if (do_prof) prof[t * 8 + 1] = clock64();
        if constexpr (FX) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < NVT; ++j) pre[g][j] = (__uint_as_float(xr[g][j]) + bias[g]) + __uint_as_float(r[g][j]);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < NVT; ++j) pre[g][j] += __uint_as_float(r[g][j]);
        }
      } else if constexpr (FX) {
        tmem_ld_wait();
        tcgen05_fence_before();
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < NVT; ++j) pre[g][j] = __uint_as_float(xr[g][j]) + bias[g];
      }
      float gi[NVT], gf[NVT], gg[NVT], go[NVT];
      __nv_bfloat16 hb[NVT];
#pragma unroll
      for (int j = 0; j < NVT; ++j) {
        gi[j] = sigmoid_fast(pre[0][j]);
        gf[j] = sigmoid_fast(pre[1][j]);
        gg[j] = tanh_fast(pre[2][j]);
        go[j] = sigmoid_fast(pre[3][j]);
        c[j] = fmaf(gf[j], c[j], gi[j] * gg[j]);
        hb[j] = __float2bfloat16_rn(go[j] * tanh_fast(c[j]));
        if (active) *reinterpret_cast<__nv_bfloat16*>(sm.opb + canon_k_off(jb + j, u, kLboB, kSboB)) = hb[j];
      }
      // publish h_t to the async proxy and hand over: one arrival per warp
      if (do_prof) prof[t * 8 + 2] = clock64();
      fence_proxy_async_smem();
      handoff_arrive<kRecThreads>();
      if (do_prof) prof[t * 8 + 3] = clock64();
      // ---- off the critical path: stream h_t and the BPTT reserve to HBM while the next MMAs run ----
#pragma unroll
      for (int j = 0; j < NVT; ++j) {
        if (valid[j]) {
          *h_ptr[j] = hb[j];
          if (gates_out) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(gi[j], gf[j]), hi = __floats2bfloat162_rn(gg[j], go[j]);
            uint2 packed;
            packed.x = *reinterpret_cast<const uint32_t*>(&lo);
            packed.y = *reinterpret_cast<const uint32_t*>(&hi);
            *g_ptr[j] = packed;
            *c_ptr[j] = c[j];
          }
        }
        h_ptr[j] += step_cells;
        g_ptr[j] += step_cells;
        c_ptr[j] += step_cells;
      }
     }
    }
    if (phase_prof) {
      unsigned long long gt;
      prof[1024 + 4] = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      prof[1024 + 5] = (long long)gt;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// dh_{t-1}^T[k, b] = sum_{kk = g*KP + u} W_hh[g*H + u][k] * dG_t[b, kk]: M = hidden unit k (TMEM lane), K = 4*KP.
// dG is written as [t][b][4u + g] (GATE-INTERLEAVED, 4H bf16 per trial): the same packed 8 bytes per cell that go into
// the MMA operand, one coalesced 8-byte store per thread.  When the launch carries dW consumer CTAs (blocks [n_rec,
// n_rec + cons.n_cons), lstm_side.cuh) every epilogue warp also bumps its own "steps stored" word in shared memory
// (release at CTA scope; per-warp monotonic counters cannot alias however far the reader falls behind), and the
// publisher warp -- its own warp: a gpu-scope release waits for the SM's outstanding writes, about a step's time, and on
// the producer warp that starved the prefetch ring -- publishes kDgBatch steps at a time, kDgLag steps late: one
// gpu-scope fence, then relaxed adds on done[t].
template <int NV, int KSTEPS, bool PROF>
__global__ void __launch_bounds__(kBwdThreads, 1)
lstm_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_dg, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ CUtensorMap tm_h, const side::DwConsume cons,
                   const uint32_t* __restrict__ w_img, const __nv_bfloat16* __restrict__ gates, const float* __restrict__ c_seq,
                   const float* __restrict__ d_hseq, const float* __restrict__ d_hlast, __nv_bfloat16* __restrict__ dG,
                   float* __restrict__ db_part, int T, int B, int H, int KP,
                   long long* __restrict__ prof, const uint32_t uz) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  if ((int)blockIdx.x >= cons.n_rec) {
    side::dw_consumer_role(smem_raw, &tm_dg, &tm_x, &tm_h, cons, (int)blockIdx.x - cons.n_rec);
    return;
  }
  const size_t b_bytes = size_t(4 * 128 / 8) * kLboB;  // gate stride fixed at 128 contraction elements
  RecSmem sm = carve(smem_raw, b_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * NV;
  const int ksteps_gate = KP / 16;
  uint32_t* const dg_stored = sm.tmem_slot + 4;  // [kEpiWarps] steps whose dG rows each epilogue warp has stored

  for (int i = tid; i < (int)(b_bytes / 4); i += kBwdThreads) reinterpret_cast<uint32_t*>(sm.opb)[i] = 0u;
  volatile int* const progress = reinterpret_cast<volatile int*>(sm.tmem_slot + 1);  // hand-offs seen by issuer warp 0
  if (tid == 0) {
    *progress = 0;
    for (int i = 0; i < kEpiWarps; ++i) dg_stored[i] = 0u;
    mbar_init(sm.bar_acc, 2);  // one tcgen05.commit per issuer warp
    for (int i = 0; i < kPfStages; ++i) mbar_init(sm.bar_pf + i, 1);
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) tmem_alloc(sm.tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;
  // W_hh^T resident in tensor memory: lane = m (hidden unit k of dh), column 256 + 64 g + u/2 holds W_hh[g*H+u][m]
  if (warp < 4) stage_weights_tmem(tmem_base + (uint32_t(warp * 32) << 16), w_img, tid);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp == kIssuerWarp + 3) {
    // ================= publisher: step n (t = T - 1 - n) once all epilogue warps have stored its dG rows, kDgLag steps late ====
    if (cons.n_cons > 0) {
      for (int n0 = 0; n0 < T; n0 += kDgBatch) {
        const int n1 = min(T, n0 + kDgBatch);          // steps [n0, n1) are published now ...
        const uint32_t need = (uint32_t)min(T, n1 + kDgLag);  // ... once every epilogue warp has stored this many steps
        if (lane < kEpiWarps) {
          uint32_t seen;
          unsigned long long t0 = 0;
          for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(dg_stored + lane)) : "memory");
            if (seen >= need) break;
            __nanosleep(200);
            if ((spins & 1023u) == 1023u) {  // watchdog: a stuck chain traps instead of hanging the device
              unsigned long long now;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
              if (t0 == 0) t0 = now;
              else if (now - t0 > 20000000000ull) __trap();
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          for (int n = n0; n < n1; ++n)
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(cons.done + (T - 1 - n)), "r"(1u) : "memory");
        }
      }
    }
  } else if (warp == kIssuerWarp + 2) {
    // TMA producer of the per-step BPTT inputs (own warp: on an issuer warp the copies delayed that warp's return to
    // the hand-off barrier, i.e. the chain): for every trial of the tile the row of gate
    // activations (H x 8 B), the row of c_{t-1} (H x 4 B; c_t is carried in a register from the step before) and, when a
    // layer above feeds gradients at every timestep, the row of d_hseq are bulk-copied into a 4-stage shared-memory
    // ring three steps ahead, completion counted on one mbarrier per stage.  (They used to be per-thread 4-byte
    // cp.async: six LDGSTS per cell per step in the epilogue warps, sitting in front of every fence of the hand-off.)
    // The warp follows issuer warp 0 through a shared-memory counter of the hand-offs it has seen: after hand-off n every
    // epilogue thread is done with the ring stage of step T-1-n, which is then refilled with step T-1-n-4.
    const int rows_valid = min(NV, B - b0);
    const uint32_t g_bytes = uint32_t(H) * 8u, c_bytes = uint32_t(H) * 4u;
    uint8_t* const pf_ring = sm.opb + b_bytes + kBarBlock;
    // Lane 0 issues the copies.  The rows of the tile's trials are adjacent in the reserve ([t][b][...]), so one step is
    // TWO bulk copies (gates, c_prev) plus one for d_hseq -- not two or three per trial: a cp.async.bulk costs the
    // issuing thread a few hundred cycles, and with 4-6 of them per step the producer needed more than a step time per
    // step (the ring ran dry after ~15 steps; measured with the ring-wait stamps of scripts/prof_lstm_steps.py).
    auto prefetch_step = [&](int t) {  // stage layout: [gates of all trials | c_prev of all trials | dhs of all trials]
      if (t < 0 || lane != 0) return;
      uint64_t* bar = sm.bar_pf + (t & (kPfStages - 1));
      const uint32_t rv = uint32_t(rows_valid);
      mbar_arrive_expect_tx(bar, rv * (g_bytes + (t > 0 ? c_bytes : 0u) + (d_hseq ? c_bytes : 0u)));
      const size_t row = size_t(t) * B + b0;
      uint8_t* d = pf_ring + size_t(t & (kPfStages - 1)) * NV * kBwdPfRow;
      bulk_g2s(d, gates + row * 4 * H, rv * g_bytes, bar);
      if (t > 0) bulk_g2s(d + NV * 1024, c_seq + (row - B) * H, rv * c_bytes, bar);
      if (d_hseq) bulk_g2s(d + NV * 1536, d_hseq + row * H, rv * c_bytes, bar);
    };
    for (int k = 0; k < kPfStages; ++k) prefetch_step(T - 1 - k);
    for (int n = 0; n + kPfStages < T; ++n) {
      while (*progress <= n) __nanosleep(100);  // (a tight poll steals issue slots and LSU cycles from the epilogue warps)
      prefetch_step(T - 1 - n - kPfStages);
    }
  } else if (warp >= kIssuerWarp) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kNslots, 0, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(sm.opb), kLboB, kSboB, kLayoutNone);
    const bool base0 = (tmem_base == 0);
    // one time loop PER issuer warp: the hoisted operand sets of the two halves must not be live together
    auto issue_loop = [&](auto half_tag) {
      constexpr int HALF = decltype(half_tag)::value;
      int n = 0;
      for (int t = T - 1; t >= 0; --t, ++n) {
        handoff_wait<kRecThreadsBwd>();  // dG_t^T staged
        if (t == 0) break;  // the last hand-off only balances the barrier
        tcgen05_fence_after();
        if (elect_one()) {
          const bool pr = PROF && prof && blockIdx.x == 0 && n + 1 < kProfSteps && HALF == 0;
          if (pr) prof[512 + (n + 1) * 8 + 4] = clock64();
          if (base0) issue_bwd<true, KSTEPS, HALF>(uz, db0, idesc, ksteps_gate);
          else issue_bwd<false, KSTEPS, HALF>(tmem_base, db0, idesc, ksteps_gate);
          umma_commit(sm.bar_acc);
          if (pr) prof[512 + (n + 1) * 8 + 5] = clock64();
          if (HALF == 0) *progress = n + 1;  // the producer warp may refill the ring stage of step t
        }
        __syncwarp();
      }
    };
    if (warp == kIssuerWarp) issue_loop(std::integral_constant<int, 0>{});
    else issue_loop(std::integral_constant<int, 1>{});
  } else {
    constexpr int NVT = NV / 2;  // cells per epilogue thread
    const int u = tid & 127;
    const int jb = (tid >> 7) * NVT;
    const bool active = u < H;
    const uint32_t lane_addr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + jb;
    float dc[NVT], dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NVT; ++j) dc[j] = 0.f;
    bool valid[NVT];
#pragma unroll
    for (int j = 0; j < NVT; ++j) valid[j] = active && (b0 + jb + j < B);

    // Per-step inputs arrive through the TMA ring filled by issuer warp 1 (see prefetch_step): per trial
    // [gates: H x (i,f,g,o) bf16 | c_{t-1}: H fp32 | d_hseq: H fp32]; c_t is carried over from the previous iteration.
    const uint8_t* const pf_ring = sm.opb + b_bytes + kBarBlock;
    uint2* dg_ptr[NVT];  // dG[t][b][4u .. 4u + 3]: one 8-byte element per cell
    float c_cur[NVT], dhl[NVT];  // c_t of the step about to be processed; d_hlast (enters at t = T-1 only)
#pragma unroll
    for (int j = 0; j < NVT; ++j) {
      const size_t row = size_t(b0 + jb + j);
      dg_ptr[j] = reinterpret_cast<uint2*>(dG) + (valid[j] ? (size_t(T - 1) * B + row) * H + u : 0);
      c_cur[j] = valid[j] ? c_seq[(size_t(T - 1) * B + row) * H + u] : 0.f;
      dhl[j] = (valid[j] && d_hlast) ? d_hlast[row * H + u] : 0.f;
    }
    auto bf_lo = [](uint32_t w) { return __uint_as_float(w << 16); };
    auto bf_hi = [](uint32_t w) { return __uint_as_float(w & 0xffff0000u); };
    int n = 0;
    for (int t = T - 1; t >= 0; --t) {
     {
      if constexpr (PROF) { if (prof && blockIdx.x == 0 && tid == 0 && n < kProfSteps) prof[1200 + n * 2] = clock64(); }
      mbar_wait(sm.bar_pf + (t & (kPfStages - 1)), ((T - 1 - t) / kPfStages) & 1);  // the rows of step t have landed (long ago)
      if constexpr (PROF) { if (prof && blockIdx.x == 0 && tid == 0 && n < kProfSteps) prof[1200 + n * 2 + 1] = clock64(); }
      struct { float i[NVT], f[NVT], g[NVT], o[NVT], c[NVT], cp[NVT]; } cur;
      float dh[NVT], pref[NVT], fac[4][NVT];  // fac: everything of dG that does not depend on dh
      // everything that does not need dh is done before the wait
      {
        const uint8_t* src = pf_ring + size_t(t & (kPfStages - 1)) * NV * kBwdPfRow;  // rows packed at their real pitch
#pragma unroll
        for (int j = 0; j < NVT; ++j) {
          const size_t cell = size_t(jb + j) * H + u;
          const uint2 gq = valid[j] ? *reinterpret_cast<const uint2*>(src + cell * 8) : make_uint2(0u, 0u);
          cur.i[j] = bf_lo(gq.x); cur.f[j] = bf_hi(gq.x); cur.g[j] = bf_lo(gq.y); cur.o[j] = bf_hi(gq.y);
          cur.c[j] = c_cur[j];
          cur.cp[j] = (valid[j] && t > 0) ? *reinterpret_cast<const float*>(src + NV * 1024 + cell * 4) : 0.f;
          dh[j] = (valid[j] && d_hseq) ? *reinterpret_cast<const float*>(src + NV * 1536 + cell * 4) : 0.f;
          if (t == T - 1) dh[j] += dhl[j];
          c_cur[j] = cur.cp[j];
          const float tcn = tanh_fast(cur.c[j]);
          pref[j] = cur.o[j] * (1.f - tcn * tcn);
          fac[0][j] = cur.g[j] * cur.i[j] * (1.f - cur.i[j]);
          fac[1][j] = cur.cp[j] * cur.f[j] * (1.f - cur.f[j]);
          fac[2][j] = cur.i[j] * (1.f - cur.g[j] * cur.g[j]);
          fac[3][j] = tcn * cur.o[j] * (1.f - cur.o[j]);
        }
      }
      const bool do_prof = PROF && prof && blockIdx.x == 0 && tid == 0 && n < kProfSteps;
      if (t < T - 1) {
        if (do_prof) prof[512 + n * 8 + 6] = clock64();
        mbar_wait(sm.bar_acc, (n - 1) & 1);
        tcgen05_fence_after();
        if (do_prof) prof[512 + n * 8 + 0] = clock64();
        // accumulators {0,1} belong to issuer warp 0, {2,3} to issuer warp 1 (each half issues >= 2 K-steps)
        uint32_t r[kBwdAcc][NVT];
#pragma unroll
        for (int a = 0; a < kBwdAcc; ++a) tmem_ld<NVT>(lane_addr + a * kNslots, r[a]);
        tmem_ld_wait();
        tcgen05_fence_before();
        if (do_prof) prof[512 + n * 8 + 1] = clock64();
#pragma unroll
        for (int j = 0; j < NVT; ++j)
          dh[j] += (__uint_as_float(r[0][j]) + __uint_as_float(r[1][j])) + (__uint_as_float(r[2][j]) + __uint_as_float(r[3][j]));
      }
      float dg[4][NVT];
#pragma unroll
      for (int j = 0; j < NVT; ++j) {
        const float dct = fmaf(dh[j], pref[j], dc[j]);
        dg[0][j] = dct * fac[0][j];
        dg[1][j] = dct * fac[1][j];
        dg[2][j] = dct * fac[2][j];
        dg[3][j] = dh[j] * fac[3][j];
        dc[j] = dct * cur.f[j];
        if (active) {  // contraction index 4u + g: the cell's four gate gradients are one aligned 8-byte store
          const __nv_bfloat162 lo = __floats2bfloat162_rn(dg[0][j], dg[1][j]), hi = __floats2bfloat162_rn(dg[2][j], dg[3][j]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(sm.opb + canon_k_off(jb + j, 4 * u, kLboB, kSboB)) = pk;
        }
      }
      if (do_prof) prof[512 + n * 8 + 2] = clock64();
      fence_proxy_async_smem();
      handoff_arrive<kRecThreadsBwd>();
      if (do_prof) prof[512 + n * 8 + 3] = clock64();
      // ---- off the critical path: dG_t to HBM (the packed 8 bytes of the operand, coalesced), bias partials ----
#pragma unroll
      for (int j = 0; j < NVT; ++j) {
        if (valid[j]) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(dg[0][j], dg[1][j]), hi = __floats2bfloat162_rn(dg[2][j], dg[3][j]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          *dg_ptr[j] = pk;
#pragma unroll
          for (int g = 0; g < 4; ++g) dbacc[g] += dg[g][j];
        }
        if (valid[j]) dg_ptr[j] -= size_t(B) * H;
      }
      if (cons.n_cons > 0) {  // this warp's rows of step n are stored: tell the producer warp (it publishes the step)
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(dg_stored + warp)), "r"(n + 1) : "memory");
      }
      ++n;
      if (do_prof) prof[512 + (n - 1) * 8 + 7] = clock64();
     }
    }
    // bias-gradient partial of this (CTA, batch half): plain stores, folded in partial order by bptt_finalize_kernel
    if (active) {
      float* dst = db_part + (size_t(blockIdx.x) * 2 + (tid >> 7)) * 4 * H + u;
#pragma unroll
      for (int g = 0; g < 4; ++g) dst[g * H] = dbacc[g];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// The fixed-order tail of BPTT (no float atomics anywhere in the backward pass): dW_ih / dW_hh = sum over the partial
// slabs of the weight-gradient products in slab order (the dW consumer groups of the launch, or the split-K ranges of the
// GEMMs that follow it), db_ih = db_hh = sum over the recurrence CTAs' partials in CTA order.  The slabs' rows are in the
// gate-interleaved order of dG (row 4u + g); the gradients leave in PyTorch's gate-major order (row g*H + u).
struct DwSlabs {
  const float* base;  // slab z at base + z * stride
  int n;              // slabs (0: the product is empty -- a single timestep has no h_{t-1})
  size_t stride;      // floats between slabs
  int ld, col0;       // row pitch and first column of this weight's block inside a slab row
};
struct DwSources {  // a weight gradient = consumer slabs (the rows folded during the sweep) + GEMM slabs (the rest)
  DwSlabs a, b;
};
__global__ void __launch_bounds__(256) bptt_finalize_kernel(const DwSources ih, const DwSources hh, const float* __restrict__ db_part,
                                                           int n_part, float* __restrict__ dw_ih, float* __restrict__ dw_hh,
                                                           float* __restrict__ db_ih, float* __restrict__ db_hh, int H, int I,
                                                           int accumulate, int n_dw_blocks) {
  const int n_ih = 4 * H * I, n_hh = 4 * H * H, n_b = 4 * H;
  if ((int)blockIdx.x < n_dw_blocks) {
    // ---- weight gradients: one element per thread, its slabs are independent loads ----
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ih + n_hh) return;
    const bool is_ih = i < n_ih;
    const DwSources& w = is_ih ? ih : hh;
    const int j = is_ih ? i : i - n_ih, width = is_ih ? I : H;
    const int r = j / width, k = j - r * width;       // r = g*H + u
    const int rp = 4 * (r % H) + r / H;               // slab row 4u + g
    float a = 0.f;
    {
      const float* src = w.a.base + size_t(rp) * w.a.ld + w.a.col0 + k;
      for (int z = 0; z < w.a.n; ++z) a += src[size_t(z) * w.a.stride];
    }
    {
      const float* src = w.b.base + size_t(rp) * w.b.ld + w.b.col0 + k;
      for (int z = 0; z < w.b.n; ++z) a += src[size_t(z) * w.b.stride];
    }
    float* dst = is_ih ? dw_ih + j : dw_hh + j;
    *dst = accumulate ? *dst + a : a;
    return;
  }
  // ---- bias gradients: 32 columns per CTA, eight threads per column each adding every eighth partial (four rotating
  //      accumulators keep the loads independent), folded in thread order: a serial walk over the 256 partials of the
  //      cfg2 launch was the longest chain of the whole kernel (~10 us of dependent L2 round trips) ----
  __shared__ float red[8][33];
  const int col = (blockIdx.x - n_dw_blocks) * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (col < n_b) {
    int q = sub;
    for (; q + 24 < n_part; q += 32) {
      a0 += db_part[size_t(q) * n_b + col];
      a1 += db_part[size_t(q + 8) * n_b + col];
      a2 += db_part[size_t(q + 16) * n_b + col];
      a3 += db_part[size_t(q + 24) * n_b + col];
    }
    for (; q < n_part; q += 8) a0 += db_part[size_t(q) * n_b + col];
  }
  red[sub][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (sub == 0 && col < n_b) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][threadIdx.x];
    db_ih[col] = accumulate ? db_ih[col] + a : a;
    db_hh[col] = accumulate ? db_hh[col] + a : a;
  }
}

// dst[4u + g, :] = bf16(src[g*H + u, :]): W_ih in the gate-interleaved row order of dG (operand of dX = dG W_ih)
__global__ void interleave_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int N) {
  const int r = blockIdx.x, u = r >> 2, g = r & 3;
  const float* s = src + size_t(g * H + u) * N;
  for (int c = threadIdx.x; c < N; c += blockDim.x) dst[size_t(r) * N + c] = __float2bfloat16_rn(s[c]);
}

// ------------------------------------------------------------------------------------------------ host side
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
constexpr size_t kWimgBytes = size_t(4) * 128 * 64 * 4;  // TMEM image of the resident weight operand
constexpr size_t kWihImgBytes = size_t(4) * 128 * 128 * 2;  // shared-memory image of W_ih (fused input projection, I <= 128)
constexpr int kMaxConsumers = 32;  // dW consumer CTAs per launch (16 K groups x 2 M halves): 8 MB of partial slabs

// the recurrence computes x_t W_ih^T itself when the input fits the resident shared-memory operand
static bool fused_projection(int I) {
  static const bool off = [] { const char* e = getenv("CSN_LSTM_NO_FUSED_X"); return e && e[0] == '1'; }();
  return !off && I <= 128 && I % 8 == 0;
}
// spare CTAs of a recurrence launch take the layer's GEMM-shaped work (lstm_side.cuh); CSN_LSTM_NO_SIDE=1 switches both
// side roles off, CSN_LSTM_NO_SERVERS=1 / CSN_LSTM_NO_CONSUMERS=1 one of them (A/B measurements, fallback tests)
static bool side_enabled(const char* which) {
  static const bool off = [] { const char* e = getenv("CSN_LSTM_NO_SIDE"); return e && e[0] == '1'; }();
  const char* e = getenv(which);
  return !off && !(e && e[0] == '1');
}

// Upper bound on the CTAs (= SMs: a recurrence CTA owns a whole SM's tensor memory) one recurrence launch may use; 0 = all.
// A caller that runs independent recurrences side by side on several streams (the crops of different length and the
// teacher pass of LstmDistillation.py:581-589) gives each a share so that they are co-resident instead of queueing.
static int g_cta_budget = 0;

// A launch that runs under a CTA budget (several recurrences side by side) may still carry its side roles when the budget
// has room for them: CSN_LSTM_BUDGET_SIDE=0 switches that off (the side roles then need the whole machine, as before).
static bool budget_side() {
  static const bool on = [] { const char* e = getenv("CSN_LSTM_BUDGET_SIDE"); return !(e && e[0] == '0'); }();
  return on;
}
static int sm_budget() { return g_cta_budget > 0 ? std::min(g_cta_budget, sm_count()) : sm_count(); }
static int servers_needed(int B) { return std::max(2, ceil_div(6 * B, 128)); }

static int pick_nv(int B) {
  // smallest batch tile that still fills the machine: per-step latency falls with NV (fewer MUFU ops per SM)
  static const int forced = [] { const char* e = getenv("CSN_LSTM_NV"); return e ? atoi(e) : 0; }();
  if (forced == 2 || forced == 4 || forced == 8) return forced;
  const int sms = sm_budget();
  if (g_cta_budget > 0 && budget_side() && side_enabled("CSN_LSTM_NO_SERVERS")) {
    // under a budget: the smallest tile that leaves room for the Xp servers / dW consumers of the launch -- a served
    // chain with four trials per CTA steps faster than an unserved one with two (0.43 against 0.53 us per step at T = 300)
    for (int nv = 2; nv <= 8; nv *= 2)
      if (ceil_div(B, nv) + servers_needed(B) <= sms) return nv;
  }
  if (ceil_div(B, 2) <= sms) return 2;
  if (ceil_div(B, 4) <= sms) return 4;
  return 8;
}

// Xp server CTAs of a forward launch over B trials (0: none).  The whole launch must be co-resident (one CTA per SM),
// the servers must keep ahead of the chain (~4k cycles per 128-row tile against ~650 per timestep of B / 128 tiles),
// and the server kernel serves I <= 128, H % 16 == 0.
static int pick_servers(int T, int B, int I, int H, int nv) {
  // An SM writes global memory at 32 B/clk at most (scripts/store_bw.py: STG and TMA alike), so the 20 spare SMs of the
  // cfg2 launch could not deliver an fp32 Xp (230 MB: >= 183 us, measured 355 us against 245 us for the fused in-CTA
  // projection).  In fp16 (115 MB, rounding 2^-11 relative: below the bf16 rounding of h and W_hh in the same sum) they
  // keep ahead of the chain: 202 us.  CSN_LSTM_NO_SERVERS=1 switches the role off.
  if (!side_enabled("CSN_LSTM_NO_SERVERS") || (g_cta_budget > 0 && !budget_side()) || !fused_projection(I) || H % 32 != 0 || H > 128) return 0;
  const int n_rec = ceil_div(B, nv), spare = sm_budget() - n_rec;
  const int n_tiles = ceil_div(T * B, 128);
  const int need = servers_needed(B);
  if (spare < need || T < 16) return 0;
  return std::min(spare, n_tiles);
}
// dW consumer CTAs of a backward launch (0: none; the two dW GEMMs then follow the recurrence)
static int pick_consumers(int T, int B, int I, int H, int nv) {
  if (!side_enabled("CSN_LSTM_NO_CONSUMERS") || (g_cta_budget > 0 && !budget_side()) || I > 128 || I % 8 != 0 || H % 8 != 0 || H > 128) return 0;
  const int n_rec = ceil_div(B, nv), spare = sm_budget() - n_rec;
  const int mh = 4 * H > 256 ? 2 : 1;
  const int i_pad = ceil_div(I, 64) * 64, h_pad = ceil_div(H, 64) * 64;
  if (i_pad + h_pad > 256 || spare < 2 * mh || T < 16) return 0;
  return std::min(spare, kMaxConsumers) / mh * mh;
}

int lstm_large_bytes(int T, int B, int I, int H, size_t* reserve, size_t* workspace);
int lstm_layer_fwd_large(const void* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* h_seq,
                         void* reserve, void* workspace, int T, int B, int I, int H, int training, cudaStream_t s);
int lstm_layer_bwd_large(const void* x, const float* w_ih, const float* w_hh, const void* h_seq, const void* reserve,
                         const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                         float* dx, void* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s);

// scratch of the fixed-order BPTT tail: partial dW slabs (consumer groups, or the split-K ranges of the two dW GEMMs sized
// for the requested split counts -- the GEMM may round them down), one bias-gradient partial per (recurrence CTA, batch
// half) (NV >= 2: at most B partials), and the per-timestep publication counters
static int dw_split_req(int M, int N) { return std::max(1, sm_count() / (ceil_div(M, 128) * ceil_div(N, 128))); }
static size_t bwd_slab_bytes(int I, int H) {
  const size_t gemm = align256(size_t(dw_split_req(4 * H, I)) * 4 * H * I * 4) + align256(size_t(dw_split_req(4 * H, H)) * 4 * H * H * 4);
  return gemm + size_t(kMaxConsumers) * side::dw_slab_floats() * 4;  // (both at once when the consumers take a share)
}
static size_t bwd_tail_bytes(int T, int B, int I, int H) {
  return bwd_slab_bytes(I, H) + align256(size_t(B + 1) * 4 * H * 4) + align256(size_t(T) * 4);
}

int lstm_tc_bytes(int T, int B, int I, int H, size_t* reserve, size_t* workspace) {
  if (H > 128) return lstm_large_bytes(T, B, I, H, reserve, workspace);  // per-step GEMM + fused cell epilogue
  if (I % 8 != 0 || H % 8 != 0) {
    set_error("bf16 tensor-core LSTM path needs input and hidden sizes that are multiples of 8 (I=%d H=%d)", I, H);
    return CSN_EUNSUPPORTED;
  }
  const size_t tb = size_t(T) * B;
  *reserve = align256(tb * 4 * H * 2) + align256(tb * H * 4);
  // forward: Xp | W_ih bf16 | W_hh image | W_ih image | tile flags
  const size_t wf = align256(tb * 4 * H * 4) + align256(size_t(4) * H * I * 2) + kWimgBytes + kWihImgBytes + align256((tb / 128 + 2) * 4);
  // backward: dG | W_ih bf16 | W_hh^T image | tail
  const size_t wb = align256(tb * 4 * H * 2) + align256(size_t(4) * H * I * 2) + kWimgBytes + bwd_tail_bytes(T, B, I, H);
  *workspace = wf > wb ? wf : wb;
  return CSN_OK;
}

void lstm_cluster_set_prof(long long* p);  // lstm_cluster.cu
static long long* g_prof_buf = nullptr;  // set by csn_dbg_lstm_profile_buffer (bring-up only)

// internal helper stream (one per device) for fork / join inside a C-ABI call; nullptr when disabled or on failure
struct SideStream {
  cudaStream_t st;
  cudaEvent_t fork, join;
};
static SideStream* side_stream() {
  static const bool off = [] { const char* e = getenv("CSN_NO_SIDE_STREAM"); return e && e[0] == '1'; }();
  if (off) return nullptr;
  static SideStream table[64];
  static bool made[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!made[dev]) {
    SideStream ss{};
    if (cudaStreamCreateWithFlags(&ss.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    table[dev] = ss;
    made[dev] = true;
  }
  return &table[dev];
}

struct FusedX {  // fused input projection operands (x == nullptr: read Xp through the ring instead)
  const __nv_bfloat16* x;
  const uint8_t* wih_img;
  const float* b_ih;
  int I, KI;
};

// raise-only dynamic shared-memory attribute (the call disappears from steady state and from graph capture)
template <typename K>
static int ensure_smem(K kern, size_t smem, size_t* set) {
  if (smem > *set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *set = smem;
  }
  return CSN_OK;
}

template <typename K, typename... Args>
static int launch_grid(K kern, int grid, int threads, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  CSN_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launches(1);
  return CSN_OK;
}

template <int NV, int KSTEPS, bool PROF>
static int launch_fwd(const float* xp, const float* b_ih_served, const uint32_t* w_hh, const float* b_hh, __nv_bfloat16* h_seq, __nv_bfloat16* gates,
                      float* c_out, int T, int B, int H, int KP, const FusedX& fx, const side::XpServe& sv, const CUtensorMap& tm_x,
                      const CUtensorMap& tm_w, cudaStream_t s) {
  const int n_rec = ceil_div(B, NV);
  if (fx.x) {
    // operand + barriers + two x blocks + W_ih image
    const size_t smem = size_t(KP / 8) * kLboB + kBarBlock + 2 * size_t(fx.KI / 8) * kLboB + size_t(4) * 128 * fx.KI * 2;
    static size_t smem_set = 0;
    auto kern = lstm_fwd_tc_kernel<NV, KSTEPS, true, PROF>;
    CSN_TRY(ensure_smem(kern, smem, &smem_set));
    return launch_grid(kern, n_rec, kRecThreadsBwd, smem, s, tm_x, tm_w, sv, (const float*)nullptr, w_hh, b_hh, h_seq, gates, c_out,
                       T, B, H, KP, g_prof_buf, 0u, fx.x, fx.wih_img, fx.b_ih, fx.I, fx.KI);
  }
  size_t smem = size_t(KP / 8) * kLboB + kBarBlock + size_t(kPfStages) * NV * 4 * H * 4 + 128;  // operand + barriers + Xp ring
  if (sv.n_srv > 0) smem = std::max(smem, side::xp_server_smem(H));
  static size_t smem_set = 0;
  auto kern = lstm_fwd_tc_kernel<NV, KSTEPS, false, PROF>;
  CSN_TRY(ensure_smem(kern, smem, &smem_set));
  return launch_grid(kern, sv.n_srv + n_rec, kRecThreadsBwd, smem, s, tm_x, tm_w, sv, xp, w_hh, b_hh, h_seq, gates, c_out, T, B, H,
                     KP, g_prof_buf, 0u, (const __nv_bfloat16*)nullptr, (const uint8_t*)nullptr, b_ih_served, 0, 16);
}

template <int NV, int KSTEPS, bool PROF>
static int launch_bwd(const uint32_t* w_hh, const __nv_bfloat16* gates, const float* c_seq, const float* d_hseq,
                      const float* d_hlast, __nv_bfloat16* dG, float* db_part, int T, int B, int H, int KP,
                      const side::DwConsume& dc, const CUtensorMap& tm_dg, const CUtensorMap& tm_x, const CUtensorMap& tm_h,
                      cudaStream_t s) {
  // operand + barriers + TMA ring of the per-step inputs
  size_t smem = size_t(4 * 128 / 8) * kLboB + kBarBlock + size_t(kPfStages) * NV * kBwdPfRow + 128;
  if (dc.n_cons > 0) smem = std::max(smem, side::dw_consumer_smem());
  static size_t smem_set = 0;
  auto kern = lstm_bwd_tc_kernel<NV, KSTEPS, PROF>;
  CSN_TRY(ensure_smem(kern, smem, &smem_set));
  return launch_grid(kern, dc.n_rec + dc.n_cons, kBwdThreads, smem, s, tm_dg, tm_x, tm_h, dc, w_hh, gates, c_seq, d_hseq, d_hlast, dG,
                     db_part, T, B, H, KP, g_prof_buf, 0u);
}

int lstm_layer_fwd_tc(const void* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                      void* h_seq, void* reserve, void* workspace, int T, int B, int I, int H, int training,
                      cudaStream_t s) {
  size_t rb, wb;
  CSN_TRY(lstm_tc_bytes(T, B, I, H, &rb, &wb));
  if (H > 128) return lstm_layer_fwd_large(x, w_ih, w_hh, b_ih, b_hh, h_seq, reserve, workspace, T, B, I, H, training, s);
  const size_t tb = size_t(T) * B;
  const int KP = ceil_div(H, 16) * 16;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* xp = reinterpret_cast<float*>(ws);
  __nv_bfloat16* wih_bf = reinterpret_cast<__nv_bfloat16*>(ws + align256(tb * 4 * H * 4));
  uint32_t* w_img = reinterpret_cast<uint32_t*>(ws + align256(tb * 4 * H * 4) + align256(size_t(4) * H * I * 2));
  uint8_t* wih_img = reinterpret_cast<uint8_t*>(w_img) + kWimgBytes;
  unsigned* flags = reinterpret_cast<unsigned*>(wih_img + kWihImgBytes);
  __nv_bfloat16* gates = reinterpret_cast<__nv_bfloat16*>(reserve);
  float* c_out = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(reserve) + align256(tb * 4 * H * 2));
  const int nv = pick_nv(B);
  const int n_srv = pick_servers(T, B, I, H, nv);
  const int KI = ceil_div(I, 16) * 16;
  FusedX fx{nullptr, nullptr, nullptr, 0, 16};
  side::XpServe sv{};
  CUtensorMap tm_x{}, tm_w{};
  const int n_tiles = (int)ceil_div<size_t>(tb, 128);
  if (n_srv > 0) {
    // spare CTAs of the launch compute Xp concurrently: bf16 W_ih for their TMA, READY flags zeroed by the prepare kernel
    sv = side::XpServe{n_srv, n_tiles, (int)tb, I, H, reinterpret_cast<__half*>(xp), flags};
    CSN_TRY(make_tmap_2d(&tm_x, x, (uint64_t)I, (uint64_t)tb, (uint64_t)I, 64, 128));
    CSN_TRY(make_tmap_2d(&tm_w, wih_bf, (uint64_t)I, (uint64_t)(4 * H), (uint64_t)I, 64, (uint32_t)side::xp_w_box_rows(H)));
  } else if (fused_projection(I)) {
    fx = FusedX{reinterpret_cast<const __nv_bfloat16*>(x), wih_img, b_ih, I, KI};
  } else {
    CSN_TRY(csn_cast(w_ih, CSN_F32, wih_bf, CSN_BF16, size_t(4) * H * I, s));
    // hoisted input projection: Xp[T*B, 4H] = x[T*B, I] . W_ih[4H, I]^T + b_ih  (fp32 out)
    CSN_TRY(gemm_tc_run(0, 1, (int)tb, 4 * H, I, x, I, wih_bf, I, xp, 4 * H, CSN_F32, b_ih, 0, 1, nullptr, s));
  }
  __nv_bfloat16* g = training ? gates : nullptr;
  float* c = training ? c_out : nullptr;
  __nv_bfloat16* hs = (__nv_bfloat16*)h_seq;
  {
    const int n_idx = std::max(std::max(4 * 128 * 64, fx.x ? 4 * 128 * fx.KI : 0), n_srv > 0 ? std::max(4 * H * I, n_tiles) : 0);
    lstm_prepare_kernel<<<ceil_div(n_idx, 256), 256, 0, s>>>(
        w_hh, w_img, H, 0, w_ih, fx.x ? reinterpret_cast<__nv_bfloat16*>(wih_img) : nullptr, I, fx.KI,
        n_srv > 0 ? wih_bf : nullptr, n_srv > 0 ? flags : nullptr, n_srv > 0 ? n_tiles : 0);
    CSN_LAUNCH_CHECK();
  }
  // hidden 128 / 96 / 64 get compile-time K-step counts (predicate-free MMA issue); other sizes take the runtime path.
  // The stamped (PROF) instantiation exists for the benchmark shape only.
#define CSN_FWD(KS)                                                                                                     \
  do {                                                                                                                  \
    if (nv == 2) return launch_fwd<2, KS, false>(xp, b_ih, w_img, b_hh, hs, g, c, T, B, H, KP, fx, sv, tm_x, tm_w, s);          \
    if (nv == 4) return launch_fwd<4, KS, false>(xp, b_ih, w_img, b_hh, hs, g, c, T, B, H, KP, fx, sv, tm_x, tm_w, s);          \
    return launch_fwd<8, KS, false>(xp, b_ih, w_img, b_hh, hs, g, c, T, B, H, KP, fx, sv, tm_x, tm_w, s);                       \
  } while (0)
  if (KP == 128 && nv == 2 && g_prof_buf) return launch_fwd<2, 8, true>(xp, b_ih, w_img, b_hh, hs, g, c, T, B, H, KP, fx, sv, tm_x, tm_w, s);
  if (KP == 128) CSN_FWD(8);
  if (KP == 96) CSN_FWD(6);
  if (KP == 64) CSN_FWD(4);
  CSN_FWD(0);
#undef CSN_FWD
}

int lstm_layer_bwd_tc(const void* x, const float* w_ih, const float* w_hh, const void* h_seq, const void* reserve,
                      const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                      float* dx, void* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s) {
  size_t rb, wb;
  CSN_TRY(lstm_tc_bytes(T, B, I, H, &rb, &wb));
  if (H > 128)
    return lstm_layer_bwd_large(x, w_ih, w_hh, h_seq, reserve, d_hseq, d_hlast, dw_ih, dw_hh, db_ih, db_hh, dx, workspace, T,
                                B, I, H, accumulate, s);
  const size_t tb = size_t(T) * B;
  const int KP = ceil_div(H, 16) * 16;
  const __nv_bfloat16* gates = reinterpret_cast<const __nv_bfloat16*>(reserve);
  const float* c_seq = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(reserve) + align256(tb * 4 * H * 2));
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* dG = reinterpret_cast<__nv_bfloat16*>(ws);  // [T, B, 4H], gate-interleaved columns 4u + g
  __nv_bfloat16* wih_bf = reinterpret_cast<__nv_bfloat16*>(ws + align256(tb * 4 * H * 2));
  uint32_t* w_img = reinterpret_cast<uint32_t*>(ws + align256(tb * 4 * H * 2) + align256(size_t(4) * H * I * 2));
  float* slabs = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(w_img) + kWimgBytes);
  float* db_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(slabs) + bwd_slab_bytes(I, H));
  unsigned* done = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(db_part) + align256(size_t(B + 1) * 4 * H * 4));
  const int nv = pick_nv(B);
  const int n_rec = ceil_div(B, nv), n_part = 2 * n_rec;
  const int n_cons = pick_consumers(T, B, I, H, nv);
  const int i_pad = ceil_div(I, 64) * 64;
  // The consumers can leave the LAST timesteps BPTT reaches (rows below chunk c_lo) to the GEMMs below, so that they
  // finish together with the chain: capacity = share * n_cons * T / m_halves chunks (a 64-row unit per M half costs a
  // consumer ~2000 cycles, bound by its SM's ~32 B/clk of L2 traffic, against ~700 cycles per timestep of the chain: a
  // share of ~0.3 would balance them).  Measured at cfg2 (scripts/gpu_lstm_ab.sh): 231 / 225 / 211 / 206 us for shares
  // 0.20 / 0.27 / 0.33 / >= 0.40 against 248 us without consumers -- the tail GEMMs' fixed cost exceeds the ~35 us the
  // consumers trail the chain by, so the default share is 1 (they take everything); CSN_LSTM_CONSUMER_SHARE overrides it.
  const int m_halves = 4 * H > 256 ? 2 : 1;
  const int n_chunks = (int)ceil_div<size_t>(tb, 64);
  int c_lo = 0;
  if (n_cons > 0) {
    static const double share = [] { const char* e = getenv("CSN_LSTM_CONSUMER_SHARE"); return e ? atof(e) : 1.0; }();
    const double capacity = share * double(n_cons) * T / m_halves;  // chunks
    c_lo = std::max(0, n_chunks - (int)capacity);
    if (size_t(c_lo) * 64 < size_t(2) * B) c_lo = 0;  // not worth a separate product
  }
  const int rows_post = n_cons > 0 ? c_lo * 64 : (int)tb;  // rows [0, rows_post) of dG are contracted after the launch
  side::DwConsume dc{n_cons, n_rec, m_halves, T, B, I, H, i_pad, c_lo, done, slabs};
  CUtensorMap tm_dg{}, tm_x{}, tm_h{};
  if (n_cons > 0) {
    CSN_TRY(make_tmap_2d(&tm_dg, dG, (uint64_t)(4 * H), (uint64_t)tb, (uint64_t)(4 * H), 64, 64));
    CSN_TRY(make_tmap_2d(&tm_x, x, (uint64_t)I, (uint64_t)tb, (uint64_t)I, 64, 64));
    CSN_TRY(make_tmap_2d(&tm_h, h_seq, (uint64_t)H, (uint64_t)tb, (uint64_t)H, 64, 64));
  }
  // W_hh^T image (+ the consumers' per-timestep publication counters zeroed)
  lstm_prepare_kernel<<<std::max(4 * 128 * 64, n_cons > 0 ? T : 0) / 256 + 1, 256, 0, s>>>(
      w_hh, w_img, H, 1, nullptr, nullptr, 0, 16, nullptr, n_cons > 0 ? done : nullptr, n_cons > 0 ? T : 0);
  CSN_LAUNCH_CHECK();
#define CSN_BWD(KS)                                                                                                              \
  do {                                                                                                                           \
    if (nv == 2) CSN_TRY((launch_bwd<2, KS, false>(w_img, gates, c_seq, d_hseq, d_hlast, dG, db_part, T, B, H, KP, dc, tm_dg, tm_x, tm_h, s)));      \
    else if (nv == 4) CSN_TRY((launch_bwd<4, KS, false>(w_img, gates, c_seq, d_hseq, d_hlast, dG, db_part, T, B, H, KP, dc, tm_dg, tm_x, tm_h, s))); \
    else CSN_TRY((launch_bwd<8, KS, false>(w_img, gates, c_seq, d_hseq, d_hlast, dG, db_part, T, B, H, KP, dc, tm_dg, tm_x, tm_h, s)));              \
  } while (0)
  if (KP == 128 && nv == 2 && g_prof_buf) CSN_TRY((launch_bwd<2, 8, true>(w_img, gates, c_seq, d_hseq, d_hlast, dG, db_part, T, B, H, KP, dc, tm_dg, tm_x, tm_h, s)));
  else if (KP == 128) CSN_BWD(8);
  else if (KP == 96) CSN_BWD(6);
  else if (KP == 64) CSN_BWD(4);
  else CSN_BWD(0);
#undef CSN_BWD
  DwSources ih{}, hh{};
  if (n_cons > 0) {
    // the consumers' partial dW over [x | h_{t-1}] is already in the slabs: only the fixed-order fold is left
    const int n_groups = n_cons / m_halves;
    ih.a = DwSlabs{slabs, n_groups, side::dw_slab_floats(), 256, 0};
    hh.a = DwSlabs{slabs, T > 1 ? n_groups : 0, side::dw_slab_floats(), 256, i_pad};
  }
  ih.b = DwSlabs{slabs, 0, 0, I, 0};
  hh.b = DwSlabs{slabs, 0, 0, H, 0};
  if (rows_post > 0) {
    // dW_ih[4H, I] += dG^T . x ; dW_hh[4H, H] += dG[1:]^T . h_seq[:-1] over the rows the consumers did not take (all of
    // them without consumers): contraction over time*batch, split-K slabs.  Both products stream the same dG: they run
    // SIDE BY SIDE on half the SMs each (fork / join on an internal stream, also valid under stream capture), so the
    // second reader finds dG in L2.
    float* slabs_ih = slabs + size_t(kMaxConsumers) * side::dw_slab_floats();
    float* slabs_hh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(slabs_ih) + align256(size_t(dw_split_req(4 * H, I)) * 4 * H * I * 4));
    const bool has_hh = rows_post > B;
    SideStream* side = has_hh ? side_stream() : nullptr;
    const int share = side ? 2 : 1;
    if (side) {
      CSN_CUDA(cudaEventRecord(side->fork, s));
      CSN_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    }
    GemmEpi slab_ih{}, slab_hh{};
    slab_ih.split_stride = size_t(4) * H * I;
    slab_hh.split_stride = size_t(4) * H * H;
    const int s_ih = gemm_tc_splits(rows_post, std::max(1, dw_split_req(4 * H, I) / share));
    const int s_hh = has_hh ? gemm_tc_splits(rows_post - B, std::max(1, dw_split_req(4 * H, H) / share)) : 0;
    CSN_TRY(gemm_tc_run(1, 0, 4 * H, I, rows_post, dG, 4 * H, x, I, slabs_ih, I, CSN_F32, nullptr, 0, s_ih, &slab_ih, s));
    if (has_hh) {
      cudaStream_t s2 = side ? side->st : s;
      int r2 = gemm_tc_run(1, 0, 4 * H, H, rows_post - B, dG + size_t(B) * 4 * H, 4 * H, h_seq, H, slabs_hh, H, CSN_F32, nullptr, 0,
                           s_hh, &slab_hh, s2);
      if (side) {  // always rejoin (a capture must not end with a dangling fork)
        CSN_CUDA(cudaEventRecord(side->join, side->st));
        CSN_CUDA(cudaStreamWaitEvent(s, side->join, 0));
      }
      CSN_TRY(r2);
    }
    ih.b = DwSlabs{slabs_ih, s_ih, slab_ih.split_stride, I, 0};
    hh.b = DwSlabs{slabs_hh, s_hh, slab_hh.split_stride, H, 0};
  }
  const int n_dw_blocks = ceil_div(4 * H * (I + H), 256);
  bptt_finalize_kernel<<<n_dw_blocks + ceil_div(4 * H, 32), 256, 0, s>>>(ih, hh, db_part, n_part, dw_ih, dw_hh, db_ih, db_hh, H, I,
                                                                         accumulate, n_dw_blocks);
  CSN_LAUNCH_CHECK();
  if (dx) {  // dX = dG . W_ih with W_ih's rows in dG's gate-interleaved order
    interleave_rows_bf16_kernel<<<4 * H, 128, 0, s>>>(w_ih, wih_bf, H, I);
    CSN_LAUNCH_CHECK();
    CSN_TRY(gemm_tc_run(0, 0, (int)tb, I, 4 * H, dG, 4 * H, wih_bf, I, dx, I, CSN_F32, nullptr, 0, 1, nullptr, s));
  }
  return CSN_OK;
}

extern "C" int csn_lstm_set_cta_budget(int max_ctas) {
  CSN_REQUIRE(max_ctas >= 0, "csn_lstm_set_cta_budget: negative budget");
  g_cta_budget = max_ctas;
  return CSN_OK;
}

}  // namespace csn

using namespace csn;

extern "C" int csn_dbg_lstm_profile_buffer(long long* buf) {
  csn::lstm_cluster_set_prof(buf);
  csn::g_prof_buf = buf;  // device buffer of at least 64*8 int64 (or NULL to switch the stamps off)
  return CSN_OK;
}
