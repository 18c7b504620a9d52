// LSTM layer, bf16 tensor-core path for hidden sizes 256 / 512 (cfg 4: stacked L2 H512, LstmDistillFromDinoV2TrainSpampinato.py:368):
// a PERSISTENT recurrence over a thread-block CLUSTER.
//
// W_hh (4H x H bf16 = 2 MB at H = 512) does not fit one SM, so the hidden units are split over the CS = H / 32 CTAs of a
// cluster: CTA c owns units [32c, 32c + 32), i.e. the 128 gate rows 4u + g of those units -- one M = 128 tile of
// tcgen05.mma whose A operand (128 rows x H, packed bf16 pairs = H / 2 of the 512 tensor-memory columns) stays RESIDENT IN
// TENSOR MEMORY for the whole sequence.  A cluster carries kNT = 16 trials (MMA N = 16) through all T steps; the batch is
// split over clusters (8 clusters of 16 CTAs = 128 SMs at B = 128), which never talk to each other.
//
// Per step every CTA computes its 128 gate rows for the 16 trials (2 CS MMAs of K = 16 over the full h_{t-1}), the
// epilogue applies the cell to its 32 units x 16 trials, and the new h slice (1 KB of bf16, already in the K-major
// operand layout) is PUSHED into the h operand buffer of all CS CTAs with one cp.async.bulk shared::cta ->
// shared::cluster each, completion counted in bytes on the receiver's mbarriers (scripts/cluster_xchg_bench.cu: this
// all-gather is the step's floor, ~1400 cycles at 16 KB per CTA and step, DSMEM moves ~17-21 B/clk per SM).  The
// receiver's MMA issuers wait per GROUP of source CTAs (four groups), so the K-steps of early pieces run under the
// arrival of the late ones.  Nothing on the chain touches HBM: the hoisted input projection arrives through a TMA ring,
// h_t / gates / c_t leave for BPTT after the hand-off.
//
// Backward (lstm_bwd_cluster_kernel) keeps the same ownership: CTA c holds W_hh[gate rows of its units, :]^T as four
// M = 128 tiles (output unit m) x K = 128 (its own gate index 4u + g) in tensor memory, multiplies them with ITS OWN
// dG_t^T (no gather needed) into partial dh^T[512, 16], and the partials are reduce-scattered as bf16 (same 16 KB per
// CTA and step as the forward all-gather): CTA j receives 16 partial [32 x 16] blocks and adds them in source order.
#include <cuda_fp16.h>

#include "tc.cuh"
#include "gemm_tc.cuh"

namespace csn {

using namespace tc;

namespace clus {

constexpr int kNT = 16;                          // trials per cluster = MMA N
constexpr int kUnits = 32;                       // hidden units per CTA: 128 gate rows = one M tile
constexpr uint32_t kPiece = kUnits * kNT * 2;    // bytes of one CTA's h slice (bf16), contiguous in the operand layout
constexpr int kEpiWarps = 8;                     // warp w: TMEM lane quadrant w % 4, trial half w / 4
constexpr int kThreads = 12 * 32;                // per trial group: 8 epilogue + 2 issuer + sender + producer warps
constexpr int kHandoffThreads = (kEpiWarps + 3) * 32;  // epilogue (arrive) + two issuer warps + sender warp (sync)
constexpr int kXStages = 4;                      // TMA ring of the hoisted input projection (one stage = one timestep)
constexpr uint32_t kXStageBytes = kNT * 128 * 4; // 16 trials x this CTA's 128 gate columns, fp32
constexpr uint32_t kLbo = 256, kSbo = 128;       // h operand, canonical K-major: k-group stride 256 B (16 trials x 16 B)
constexpr uint32_t kWcol0 = 256;                 // tensor memory: [64 gi, 64 gi + 64) four accumulators x 16 of trial group gi, [256, 256 + H/2) weights
constexpr uint32_t kTmemCols = 512;
constexpr int kGroups = 4;                       // arrival groups of the all-gather (CS / 4 source CTAs each)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void handoff_arrive_id(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void handoff_wait_id(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// spin (with back-off and a 20 s watchdog) until a flag word written by another kernel becomes non-zero
__device__ __forceinline__ void wait_flag_set(const unsigned* p) {
  unsigned v, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v) return;
    __nanosleep(200);
    if ((++spins & 4095u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}

template <int CS>
struct FwdSmem {
  static constexpr uint32_t h_bytes = CS * kPiece;            // one h operand buffer: [k-group H/8][trial 16][8 k] bf16
  static constexpr uint32_t off_h = 0;                        // two buffers (step parity)
  static constexpr uint32_t off_stage = 2 * h_bytes;          // two staging pieces (step parity)
  static constexpr uint32_t off_x = off_stage + 2 * kPiece;   // Xp ring
  static constexpr uint32_t off_bar = off_x + kXStages * kXStageBytes;
  static constexpr uint32_t total = off_bar + 256;
};

// In-quad 4 x 4 transpose: lanes 4q + g (g = gate) each hold x[j] = gate g of trial j; afterwards lane g holds
// (i, f, g, o) of trial g.  Two exchange rounds (xor 2, xor 1), four shuffles.
__device__ __forceinline__ void quad_transpose(float (&x)[4], int g) {
  const bool hi = (g & 2) != 0, lo = (g & 1) != 0;
  {
    const float s0 = hi ? x[0] : x[2], s1 = hi ? x[1] : x[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (hi) { x[0] = r0; x[1] = r1; } else { x[2] = r0; x[3] = r1; }
  }
  // now x[2a + b] = M[2a + (g & 1)][2 (g >> 1) + b]; second round inside the lane pairs
  {
    const float k0 = lo ? x[1] : x[0], s0 = lo ? x[0] : x[1];
    const float k1 = lo ? x[3] : x[2], s1 = lo ? x[2] : x[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    x[0] = lo ? r0 : k0;
    x[1] = lo ? k0 : r0;
    x[2] = lo ? r1 : k1;
    x[3] = lo ? k1 : r1;
  }
}

// One step's MMAs of issuer warp HALF: K-steps [HALF * CS, (HALF + 1) * CS) (two per source CTA), alternating over the
// warp's two accumulators; PART selects the first / second arrival group of the half.
template <int CS, int HALF, int PART>
__device__ __forceinline__ void issue_part(uint32_t tb, uint32_t acc, uint64_t db, uint32_t idesc) {
#pragma unroll
  for (int i = PART * (CS / 2); i < (PART + 1) * (CS / 2); ++i) {
    const int kk = HALF * CS + i;
    umma_f16_ts(acc + uint32_t(HALF * 2 + (i & 1)) * kNT, tb + kWcol0 + kk * 8, db + uint64_t((uint32_t(kk) * 2u * kLbo) >> 4), idesc,
                i < 2 ? 0u : 1u);
  }
}

// ------------------------------------------------------------------------------------------------ forward
// xp: hoisted input projection incl. both biases, fp32 [T*B, 4H], gate-interleaved columns (4u + g) -- through tm_xp,
// box = this CTA's 128 columns x 16 trials.  w_rows: W_hh as bf16 rows in the same interleaved order ([4H][H], read as
// packed pairs).  Outputs: h_seq bf16 [T,B,H]; BPTT reserve gates bf16 [T,B,4H] (activations, interleaved) and c fp32 [T,B,H].
//
// G = trial GROUPS per cluster.  A step's all-gather keeps the SM idle for most of its ~1250 cycles, and only 7 clusters of
// 16 CTAs are co-resident on a B200 (GPC sizes), so 8 groups of 16 trials (B = 128) would run as two waves.  With G = 2 a
// cluster carries two independent groups of 16 trials, each with its own warps, operand buffers, accumulators and
// barriers, sharing only the resident weights: one group's gather is in flight while the other multiplies and runs its
// cells.  Warp map (12 G warps): epilogue warps [8 gi, 8 gi + 8), issuers 8 G + 2 gi + {0, 1}, sender 10 G + gi,
// producer 11 G + gi.
template <int CS, int G, bool PROF>
__global__ void __launch_bounds__(kThreads * G, 1)
lstm_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tm_xp, const uint32_t* __restrict__ w_rows,
                        __nv_bfloat16* __restrict__ h_seq, __nv_bfloat16* __restrict__ gates_out, float* __restrict__ c_out,
                        int T, int B, const unsigned* __restrict__ xp_flags, int xp_chunk, uint8_t* __restrict__ xchg,
                        long long* __restrict__ prof) {
  constexpr int H = CS * kUnits;
  using L = FwdSmem<CS>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // role and trial group of this warp
  const int gi = warp < 8 * G ? warp / 8 : warp < 10 * G ? (warp - 8 * G) / 2 : warp < 11 * G ? warp - 10 * G : warp - 11 * G;
  uint8_t* const gbase = smem_raw + size_t(gi) * L::total;
  uint8_t* const hbuf = gbase + L::off_h;
  uint8_t* const stage = gbase + L::off_stage;
  uint8_t* const xring = gbase + L::off_x;
  uint64_t* const grp_bar = reinterpret_cast<uint64_t*>(gbase + L::off_bar);  // [2][kGroups]
  uint64_t* const acc_bar = grp_bar + 2 * kGroups;
  uint64_t* const xp_bar = acc_bar + 1;                                           // [kXStages]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + L::off_bar) + 2 * (2 * kGroups + 1 + kXStages);  // group 0's block
  volatile int* const progress = reinterpret_cast<volatile int*>(xp_bar + kXStages) + 1;  // steps whose hand-off the sender has seen

  const uint32_t rank = cluster_ctarank();
  const int b0 = ((int)(blockIdx.x / CS) * G + gi) * kNT;
  const int hbar = 1 + gi;  // named barrier of this group's hand-off
  const bool prof_on = PROF && prof && blockIdx.x == 0 && gi == 0;

  if (tid == 0) {
    for (int k = 0; k < G; ++k) {
      uint64_t* gb = reinterpret_cast<uint64_t*>(smem_raw + size_t(k) * L::total + L::off_bar);
      for (int i = 0; i < 2 * kGroups; ++i) mbar_init(gb + i, 1);
      mbar_init(gb + 2 * kGroups, 2);  // one tcgen05.commit per issuer warp
      for (int i = 0; i < kXStages; ++i) mbar_init(gb + 2 * kGroups + 1 + i, 1);
      *(reinterpret_cast<volatile int*>(gb + 2 * kGroups + 1 + kXStages) + 1) = 0;
    }
    fence_mbar_init();
  }
  if (warp == 8 * G) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_base = tmem_base + uint32_t(gi) * 64u;  // this group's four accumulators
  if (warp < 4) {  // resident weights: TMEM lane = gate row 4u + g of this CTA's units, column kWcol0 + k / 2
    const uint4* src = reinterpret_cast<const uint4*>(w_rows + (size_t(rank) * 128 + tid) * (H / 2));
    const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16) + kWcol0;
#pragma unroll 1
    for (int ch = 0; ch < H / 64; ++ch) {
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(src + ch * 8 + q);
        r[q * 4 + 0] = v.x; r[q * 4 + 1] = v.y; r[q * 4 + 2] = v.z; r[q * 4 + 3] = v.w;
      }
      tmem_st32(lane_addr + ch * 32, r);
    }
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();  // every CTA's barriers exist before a peer's first copy can complete on them

  if (warp >= 8 * G && warp < 10 * G) {
    // ================= MMA issuers: step t consumes h_{t-1} from buffer t & 1 =================
    constexpr uint32_t idesc = make_idesc_bf16(128, kNT, 0, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(hbuf), kLbo, kSbo, kLayoutNone);
    auto loop = [&](auto half_tag) {
      constexpr int HALF = decltype(half_tag)::value;
      for (int t = 1; t < T; ++t) {
        handoff_wait_id(hbar, kHandoffThreads);  // this group's epilogue has drained the accumulators of step t - 1
        const int b = t & 1;
        const uint32_t par = uint32_t((t - 1) >> 1) & 1u;
        const uint64_t db = db0 + uint64_t((uint32_t(b) * L::h_bytes) >> 4);
        const bool pr = prof_on && lane == 0 && t < 128;
        if (pr) prof[HALF ? 1024 + t * 4 + 3 : t * 8 + 3] = clock64();  // hand-off seen
        mbar_wait(grp_bar + b * kGroups + 2 * HALF, par);
        tcgen05_fence_after();
        __syncwarp();
        if (pr) prof[HALF ? 1024 + t * 4 + 0 : t * 8 + 5] = clock64();
        if (elect_one()) issue_part<CS, HALF, 0>(tmem_base, acc_base, db, idesc);
        __syncwarp();
        mbar_wait(grp_bar + b * kGroups + 2 * HALF + 1, par);
        __syncwarp();
        if (pr) prof[HALF ? 1024 + t * 4 + 1 : t * 8 + 6] = clock64();
        if (elect_one()) {
          issue_part<CS, HALF, 1>(tmem_base, acc_base, db, idesc);
          umma_commit(acc_bar);
        }
        __syncwarp();
        if (pr) prof[HALF ? 1024 + t * 4 + 2 : t * 8 + 7] = clock64();
      }
    };
    if (((warp - 8 * G) & 1) == 0) loop(std::integral_constant<int, 0>{});
    else loop(std::integral_constant<int, 1>{});
  } else if (warp >= 10 * G && warp < 11 * G) {
    // ================= all-gather: this CTA's h_t slice -> buffer (t + 1) & 1 of every CTA of the cluster =================
    for (int t = 0; t + 1 < T; ++t) {
      handoff_wait_id(hbar, kHandoffThreads);  // the slice is staged (and fenced towards the async proxy by its writers)
      const int b = (t + 1) & 1;
      if (lane < kGroups) mbar_arrive_expect_tx(grp_bar + b * kGroups + lane, (CS / kGroups) * kPiece);
      if (xchg) {
        // through L2: the slice goes to a global scratch (this warp: 64 chunks of 16 bytes), then ONE multicast bulk copy
        // delivers it to the operand buffer of all CS CTAs and signals each CTA's group barrier.  One copy per CTA and step
        // instead of CS: with two trial groups the copy engine's per-copy cost was the limit (scripts/cluster_xchg_bench.cu:
        // 3300 -> 1950 cycles per round with two groups).
        uint8_t* g = xchg + ((((size_t)(blockIdx.x / CS) * G + gi) * 2 + b) * CS + rank) * kPiece;
#pragma unroll
        for (int c = lane; c < (int)(kPiece / 16); c += 32)
          reinterpret_cast<uint4*>(g)[c] = *reinterpret_cast<const uint4*>(stage + b * kPiece + c * 16);
        asm volatile("fence.proxy.async.global;" ::: "memory");  // generic stores -> the copy engine's read
        __syncwarp();
        if (lane == 0)
          bulk_g2s_multicast(hbuf + b * L::h_bytes + rank * kPiece, g, kPiece, grp_bar + b * kGroups + rank / (CS / kGroups),
                             (uint16_t)((1u << CS) - 1));
      } else if (lane < CS) {
        const uint32_t d = (rank + uint32_t(lane)) % CS;  // staggered: no destination is everybody's last
        bulk_s2c(mapa(smem_u32(hbuf + b * L::h_bytes + rank * kPiece), d), smem_u32(stage + b * kPiece), kPiece,
                 mapa(smem_u32(grp_bar + b * kGroups + rank / (CS / kGroups)), d));
      }
      if (lane == 0) *progress = t + 1;
      __syncwarp();
      if (prof_on && lane == 0 && t + 1 < 128) prof[(t + 1) * 8 + 4] = clock64();  // copies of h_t issued
    }
  } else if (warp >= 11 * G) {
    // ================= TMA ring of the hoisted input projection =================
    if (lane == 0) {
      for (int t = 0; t < T; ++t) {
        if (t >= kXStages) {
          const int need = t - kXStages + 1;  // step t - kXStages has been consumed by every epilogue thread
          while (*progress < need) __nanosleep(64);
        }
        if (xp_flags && t % xp_chunk == 0) {
          // the projection GEMM runs beside this launch, one time chunk per GEMM: its rows exist once the chunk's flag is
          // up (a one-thread kernel behind the GEMM in stream order); acquire, then order the TMA reads behind it
          wait_flag_set(xp_flags + t / xp_chunk);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        uint64_t* bar = xp_bar + (t & (kXStages - 1));
        mbar_arrive_expect_tx(bar, kXStageBytes);
        tma_load_2d(xring + size_t(t & (kXStages - 1)) * kXStageBytes, &tm_xp, bar, (int)rank * 128, t * B + b0);
      }
    }
  } else {
    // ================= epilogue: thread = (gate row 4u + g, eight of the sixteen trials) =================
    const int wl = warp & 7;
    const int q = wl & 3, half = wl >> 2;
    const int row = q * 32 + lane, ul = row >> 2, g = row & 3;
    const int jb = half * 8;
    const uint32_t lane_addr = acc_base + (uint32_t(q * 32) << 16) + jb;
    const float sc = (g == 2) ? 1.f : 0.5f, off = (g == 2) ? 0.f : 0.5f;  // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5
    float c[2] = {0.f, 0.f};
    // after the transposes this thread owns the cells (unit ul, trials jb + g and jb + 4 + g)
    int n[2];
    bool valid[2];
    size_t cell[2];
    const size_t step_cells = size_t(B) * H;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      n[i] = jb + 4 * i + g;
      valid[i] = b0 + n[i] < B;
      cell[i] = valid[i] ? size_t(b0 + n[i]) * H + rank * kUnits + ul : 0;
    }
    uint2* const gates2 = reinterpret_cast<uint2*>(gates_out);
    const bool pr0 = prof_on && (tid & 255) == 0;
    for (int t = 0; t < T; ++t) {
      mbar_wait(xp_bar + (t & (kXStages - 1)), uint32_t(t / kXStages) & 1u);
      const float* xs = reinterpret_cast<const float*>(xring + size_t(t & (kXStages - 1)) * kXStageBytes) + row;
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = xs[(jb + j) * 128];
      if (t > 0) {
        mbar_wait(acc_bar, uint32_t(t - 1) & 1u);
        tcgen05_fence_after();
        if (pr0 && t < 128) prof[t * 8 + 0] = clock64();
        uint32_t r[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) tmem_ld<8>(lane_addr + k * kNT, r[k]);
        tmem_ld_wait();
        tcgen05_fence_before();
        if (pr0 && t < 128) prof[t * 8 + 1] = clock64();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          a[j] += (__uint_as_float(r[0][j]) + __uint_as_float(r[1][j])) + (__uint_as_float(r[2][j]) + __uint_as_float(r[3][j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(tanh_fast(a[j] * sc), sc, off);
      float z[2][4];
      __nv_bfloat16 hb[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) z[i][j] = a[4 * i + j];
        quad_transpose(z[i], g);
        c[i] = fmaf(z[i][1], c[i], z[i][0] * z[i][2]);
        hb[i] = __float2bfloat16_rn(z[i][3] * tanh_fast(c[i]));
      }
      if (t + 1 < T) {
        uint8_t* st = stage + ((t + 1) & 1) * kPiece + (ul >> 3) * kLbo + (ul & 7) * 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<__nv_bfloat16*>(st + n[i] * 16) = hb[i];
        fence_proxy_async_smem();
        handoff_arrive_id(hbar, kHandoffThreads);
        if (pr0 && t < 128) prof[t * 8 + 2] = clock64();
      }
      // ---- off the chain: h_t and the BPTT reserve ----
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (valid[i]) {
          h_seq[cell[i]] = hb[i];
          const __nv_bfloat162 lo = __floats2bfloat162_rn(z[i][0], z[i][1]), hi = __floats2bfloat162_rn(z[i][2], z[i][3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          gates2[cell[i]] = pk;
          c_out[cell[i]] = c[i];
          cell[i] += step_cells;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA retires while a peer's copy may still read its staging piece
  if (warp == 8 * G) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ backward
// Same ownership as the forward kernel: CTA c owns the hidden units [32c, 32c + 32) of its cluster's 16 trials.
//   dh_{t-1}^T[m, n] = sum over gate rows r = 4u + g of W_hh[r][m] dG_t[n][r]
// CTA c contributes the rows of ITS units: a partial [H x 16] product of the resident operand A_c[m][kl] =
// W_hh[gate row 128c + kl][m] (H / 128 tiles of M = 128 output units x K = 128, packed in tensor memory) with its own
// dG_t^T -- no gather before the MMAs.  The partial rows of CTA j's units (a [32 x 16] block per source) are rounded to
// bf16, staged, and pushed to CTA j (one bulk copy per destination, 1 KB); CTA j adds its CS blocks in source order
// (fixed order: deterministic), adds the gradient arriving from above, runs the cell backward for its 32 x 16 cells and
// writes dG_t^T for its next MMAs.  The tiles' products are committed one by one, so the copies of the first tile's
// blocks are in flight while the later tiles still multiply.
// G trial groups per cluster as in the forward kernel.  Warp map (15 G warps): epilogue [8 gi, 8 gi + 8), issuers
// 8 G + 2 gi + {0, 1}, senders 10 G + 4 gi + tile, producer 14 G + gi.
constexpr int kBwdWarps = 15;
constexpr int kBwdHandoff = (kEpiWarps + 2) * 32;                   // epilogue (arrive) + two issuer warps (sync)
constexpr int kBwdSendBar = (kEpiWarps + 1) * 32;                   // epilogue (arrive) + one sender warp (sync)
constexpr uint32_t kLboG = 272;                                     // dG^T operand k-group stride (bank rotation, see lstm_tc.cu)
constexpr uint32_t kGBytes = 16 * kLboG;                            // K = 128 gate rows -> 16 k-groups
constexpr int kBStages = 4;
constexpr uint32_t kBStageBytes = kNT * (256 + 128 + 128);          // gates (8 B per cell) | c_{t-1} | d_hseq, 32 units x 16 trials

template <int CS>
struct BwdSmem {
  static constexpr uint32_t r_bytes = CS * kPiece;                  // received partial blocks of one step: [source][trial][unit] bf16
  static constexpr uint32_t off_r = 0;                              // two buffers (step parity)
  static constexpr uint32_t off_stage = 2 * r_bytes;                // two staging sets [destination][trial][unit] bf16
  static constexpr uint32_t off_g = off_stage + 2 * r_bytes;        // dG^T operand
  static constexpr uint32_t off_ring = off_g + ((kGBytes + 127) & ~127u);
  static constexpr uint32_t off_bar = off_ring + kBStages * kBStageBytes;
  static constexpr uint32_t total = off_bar + 128;
};

// wt_rows: W_hh^T as bf16 [H][4H], column 4u + g (gate-interleaved): row m, columns [128c, 128c + 128) is CTA c's A row.
template <int CS, int G, bool PROF>
__global__ void __launch_bounds__(kBwdWarps * 32 * G, 1)
lstm_bwd_cluster_kernel(const __grid_constant__ CUtensorMap tm_gates, const __grid_constant__ CUtensorMap tm_c,
                        const __grid_constant__ CUtensorMap tm_dh, const uint32_t* __restrict__ wt_rows,
                        const float* __restrict__ c_seq, const float* __restrict__ d_hlast, int has_dhseq,
                        __nv_bfloat16* __restrict__ dG, int T, int B, unsigned* __restrict__ done, int done_chunk,
                        long long* __restrict__ prof) {
  constexpr int H = CS * kUnits;
  constexpr int MT = H / 128;  // M tiles of the partial product; tile mt feeds destinations [4 mt, 4 mt + 4)
  constexpr int kThreadsAll = kBwdWarps * 32 * G;
  using L = BwdSmem<CS>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gi = warp < 8 * G ? warp / 8 : warp < 10 * G ? (warp - 8 * G) / 2 : warp < 14 * G ? (warp - 10 * G) / 4 : warp - 14 * G;
  uint8_t* const gbase = smem_raw + size_t(gi) * L::total;
  uint8_t* const rbuf = gbase + L::off_r;
  uint8_t* const stage = gbase + L::off_stage;
  uint8_t* const gop = gbase + L::off_g;
  uint8_t* const ring = gbase + L::off_ring;
  uint64_t* const recv_bar = reinterpret_cast<uint64_t*>(gbase + L::off_bar);     // [2]
  uint64_t* const acc_bar = recv_bar + 2;                                          // [4] one per M tile
  uint64_t* const pf_bar = acc_bar + 4;                                            // [kBStages]
  volatile int* const progress = reinterpret_cast<volatile int*>(pf_bar + kBStages) + 1;  // hand-offs seen by issuer warp 0
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + L::off_bar) + 2 * (2 + 4 + kBStages);  // group 0's block

  const uint32_t rank = cluster_ctarank();
  const int b0 = ((int)(blockIdx.x / CS) * G + gi) * kNT;
  const int hbar = 1 + gi;            // named barrier of this group's hand-off
  const int sbar0 = 1 + G + 4 * gi;   // first of this group's four send barriers
  const bool prof_on = PROF && prof && blockIdx.x == 0 && gi == 0;

  if (tid == 0) {
    for (int k = 0; k < G; ++k) {
      uint64_t* gb = reinterpret_cast<uint64_t*>(smem_raw + size_t(k) * L::total + L::off_bar);
      for (int i = 0; i < 2 + 4 + kBStages; ++i) mbar_init(gb + i, 1);
      *(reinterpret_cast<volatile int*>(gb + 2 + 4 + kBStages) + 1) = 0;
    }
    fence_mbar_init();
  }
  for (int k = 0; k < G; ++k)
    for (int i = tid; i < (int)(kGBytes / 4); i += kThreadsAll)
      reinterpret_cast<uint32_t*>(smem_raw + size_t(k) * L::total + L::off_g)[i] = 0u;
  if (warp == 8 * G) tmem_alloc(tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_base = tmem_base + uint32_t(gi) * 64u;
  if (warp < 4) {  // resident operand: tile mt, TMEM lane = output unit 128 mt + lane, column kWcol0 + 64 mt + kl / 2
    const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16) + kWcol0;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const uint4* src = reinterpret_cast<const uint4*>(wt_rows + (size_t(mt) * 128 + tid) * (2 * H) + rank * 64);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v = __ldg(src + ch * 8 + q);
          r[q * 4 + 0] = v.x; r[q * 4 + 1] = v.y; r[q * 4 + 2] = v.z; r[q * 4 + 3] = v.w;
        }
        tmem_st32(lane_addr + mt * 64 + ch * 32, r);
      }
    }
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();

  if (warp >= 8 * G && warp < 10 * G) {
    // ================= MMA issuers: warp h multiplies tiles {2h, 2h + 1} (H = 256: one tile each) =================
    constexpr uint32_t idesc = make_idesc_bf16(128, kNT, 0, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(gop), kLboG, kSbo, kLayoutNone);
    const int hw = (warp - 8 * G) & 1;
    constexpr int TPW = MT / 2;  // tiles per issuer warp
    for (int n = 0; n + 1 < T; ++n) {
      handoff_wait_id(hbar, kBwdHandoff);  // dG_t^T staged; every accumulator of the previous step has been read
      tcgen05_fence_after();
      if (elect_one()) {
        if (prof_on && hw == 0 && n < 128) prof[2048 + n * 8 + 4] = clock64();
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          const int mt = hw * TPW + j;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ts(acc_base + uint32_t(mt) * kNT, tmem_base + kWcol0 + uint32_t(mt) * 64 + kk * 8,
                        db0 + uint64_t((uint32_t(kk) * 2u * kLboG) >> 4), idesc, kk ? 1u : 0u);
          umma_commit(acc_bar + mt);
        }
        if (prof_on && hw == 0 && n < 128) prof[2048 + n * 8 + 5] = clock64();
        if (hw == 0) *progress = n + 1;
      }
      __syncwarp();
    }
  } else if (warp >= 10 * G && warp < 14 * G) {
    // ================= reduce-scatter: tile j's blocks -> buffer (n + 1) & 1 of CTAs 4 j .. 4 j + 3 =================
    const int j = (warp - 10 * G) & 3;
    if (j < MT) {
      for (int n = 0; n + 1 < T; ++n) {
        handoff_wait_id(sbar0 + j, kBwdSendBar);  // tile j's blocks are staged and fenced
        const int nb = (n + 1) & 1;
        if (j == 0 && lane == 0) mbar_arrive_expect_tx(recv_bar + nb, L::r_bytes);
        if (lane < 4) {
          const uint32_t d = uint32_t(4 * j + ((lane + (int)rank) & 3));
          bulk_s2c(mapa(smem_u32(rbuf + nb * L::r_bytes + rank * kPiece), d), smem_u32(stage + (n & 1) * L::r_bytes + d * kPiece), kPiece,
                   mapa(smem_u32(recv_bar + nb), d));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 14 * G) {
    // ================= TMA ring of the per-step inputs of this CTA's 32 units x 16 trials =================
    if (lane == 0) {
      const uint32_t bytes = kNT * (256 + 128) + (has_dhseq ? kNT * 128 : 0);
      for (int n = 0; n < T; ++n) {
        const int t = T - 1 - n;
        if (n >= kBStages) {
          const int need = n - kBStages + 1;  // hand-off n - kBStages seen: that stage's readers are done
          while (*progress < need) __nanosleep(64);
        }
        uint64_t* bar = pf_bar + (n & (kBStages - 1));
        uint8_t* dst = ring + size_t(n & (kBStages - 1)) * kBStageBytes;
        mbar_arrive_expect_tx(bar, bytes);
        tma_load_2d(dst, &tm_gates, bar, (int)rank * 128, t * B + b0);
        tma_load_2d(dst + kNT * 256, &tm_c, bar, (int)rank * kUnits, (t - 1) * B + b0);  // t = 0: rows < 0 are zero-filled
        if (has_dhseq) tma_load_2d(dst + kNT * 384, &tm_dh, bar, (int)rank * kUnits, t * B + b0);
      }
    }
  } else {
    // ================= epilogue =================
    // phase B (cells): thread = (unit ul, trials tn and tn + 8);  phase A (partials): thread = (TMEM lane, trial half)
    const int et = tid & 255, wl = warp & 7;
    const int ul = et & 31, tn = et >> 5;
    const int q = wl & 3, half = wl >> 2;
    const uint32_t lane_addr = acc_base + (uint32_t(q * 32) << 16) + half * 8;
    float dc[2] = {0.f, 0.f}, c_cur[2], dhl[2];
    bool valid[2];
    size_t cell[2];
    const size_t step_cells = size_t(B) * H;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int n_i = tn + 8 * i;
      valid[i] = b0 + n_i < B;
      cell[i] = valid[i] ? (size_t(T - 1) * B + b0 + n_i) * H + rank * kUnits + ul : 0;
      c_cur[i] = valid[i] ? c_seq[cell[i]] : 0.f;
      dhl[i] = (valid[i] && d_hlast) ? d_hlast[size_t(b0 + n_i) * H + rank * kUnits + ul] : 0.f;
    }
    uint2* const dG2 = reinterpret_cast<uint2*>(dG);
    auto bf_lo = [](uint32_t w) { return __uint_as_float(w << 16); };
    auto bf_hi = [](uint32_t w) { return __uint_as_float(w & 0xffff0000u); };
    for (int n = 0; n < T; ++n) {
      const int t = T - 1 - n;
      const bool pr = prof_on && et == 0 && n < 128;
      // ---- phase B ----
      mbar_wait(pf_bar + (n & (kBStages - 1)), uint32_t(n / kBStages) & 1u);
      const uint8_t* st = ring + size_t(n & (kBStages - 1)) * kBStageBytes;
      float gi_[2], gf[2], gg[2], go[2], cp[2], dh[2], tcn[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n_i = tn + 8 * i;
        const uint2 gq = *reinterpret_cast<const uint2*>(st + n_i * 256 + ul * 8);
        gi_[i] = bf_lo(gq.x); gf[i] = bf_hi(gq.x); gg[i] = bf_lo(gq.y); go[i] = bf_hi(gq.y);
        cp[i] = *reinterpret_cast<const float*>(st + kNT * 256 + n_i * 128 + ul * 4);
        dh[i] = has_dhseq ? *reinterpret_cast<const float*>(st + kNT * 384 + n_i * 128 + ul * 4) : 0.f;
        if (n == 0) dh[i] += dhl[i];
        tcn[i] = tanh_fast(c_cur[i]);
        c_cur[i] = cp[i];
      }
      if (n > 0) {
        if (pr) prof[2048 + n * 8 + 6] = clock64();
        mbar_wait(recv_bar + (n & 1), uint32_t((n - 1) >> 1) & 1u);
        if (pr) prof[2048 + n * 8 + 0] = clock64();
        const uint8_t* rb = rbuf + (n & 1) * L::r_bytes + ul * 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int n_i = tn + 8 * i;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int s = 0; s < CS; s += 2) {
            a0 += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(rb + s * kPiece + n_i * 64));
            a1 += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(rb + (s + 1) * kPiece + n_i * 64));
          }
          dh[i] += a0 + a1;
        }
      }
      uint2 pk[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float dct = fmaf(dh[i] * go[i], 1.f - tcn[i] * tcn[i], dc[i]);
        const float d0 = dct * gg[i] * gi_[i] * (1.f - gi_[i]);
        const float d1 = dct * cp[i] * gf[i] * (1.f - gf[i]);
        const float d2 = dct * gi_[i] * (1.f - gg[i] * gg[i]);
        const float d3 = dh[i] * tcn[i] * go[i] * (1.f - go[i]);
        dc[i] = dct * gf[i];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1), hi = __floats2bfloat162_rn(d2, d3);
        pk[i].x = *reinterpret_cast<const uint32_t*>(&lo);
        pk[i].y = *reinterpret_cast<const uint32_t*>(&hi);
        // operand element (trial n_i, gate rows 4 ul .. 4 ul + 3): one 8-byte store
        *reinterpret_cast<uint2*>(gop + (ul >> 1) * kLboG + (tn + 8 * i) * 16 + (ul & 1) * 8) = pk[i];
      }
      if (pr) prof[2048 + n * 8 + 1] = clock64();
      if (t > 0) {
        fence_proxy_async_smem();
        handoff_arrive_id(hbar, kBwdHandoff);
      }
      if (pr) prof[2048 + n * 8 + 2] = clock64();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (valid[i]) {
          dG2[cell[i]] = pk[i];
          cell[i] -= step_cells;
        }
      }
      if (done && t % done_chunk == 0) {
        // every dG row of the time chunk that ends here is stored: publish it to the weight-gradient GEMMs that run beside
        // this launch (each writer fences its own stores, one thread of the group counts the group in)
        __threadfence();
        handoff_wait_id(1 + 5 * G + gi, kEpiWarps * 32);
        if (et == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(done + t / done_chunk), "r"(1u) : "memory");
      }
      if (t == 0) break;
      // ---- phase A: the partial product tiles, rounded to bf16, one 1 KB block per destination ----
      uint8_t* const stg = stage + (n & 1) * L::r_bytes;
#pragma unroll
      for (int jj = 0; jj < MT; ++jj) {
        const int mt = (MT == 4) ? ((jj & 1) * 2 + (jj >> 1)) : jj;  // completion order of the two issuer warps: 0, 2, 1, 3
        mbar_wait(acc_bar + mt, uint32_t(n) & 1u);
        tcgen05_fence_after();
        uint32_t r[8];
        tmem_ld<8>(lane_addr + mt * kNT, r);
        tmem_ld_wait();
        tcgen05_fence_before();
        uint8_t* dst = stg + (4 * mt + q) * kPiece + (half * 8) * 64 + lane * 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<__nv_bfloat16*>(dst + k * 64) = __float2bfloat16_rn(__uint_as_float(r[k]));
        fence_proxy_async_smem();
        handoff_arrive_id(sbar0 + mt, kBwdSendBar);
      }
      if (pr) prof[2048 + n * 8 + 3] = clock64();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8 * G) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace clus

static bool cluster_enabled() {
  static const bool off = [] { const char* e = getenv("CSN_LSTM_NO_CLUSTER"); return e && e[0] == '1'; }();
  return !off;
}
bool lstm_cluster_supported(int H) { return cluster_enabled() && (H == 256 || H == 512); }

// A recurrence CTA allocates all 512 tensor-memory columns: a second CTA on the same SM would sit in tcgen05.alloc until the
// first retires.  Asking for more than half of the SM's shared memory keeps the launch at one CTA per SM, so that the
// scheduler spreads the clusters (7 clusters of 16 CTAs are co-resident on this part, scripts/cluster_xchg_bench.cu).
constexpr size_t kOneCtaSmem = 120 * 1024;
static long long* g_clus_prof = nullptr;
void lstm_cluster_set_prof(long long* p) { g_clus_prof = p; }

// One cluster launch: grid = clusters x CS CTAs of `threads` threads; smem padded to one CTA per SM.
template <typename K, typename... Args>
static int launch_cluster(K kern, bool* attr_set, int cs, int clusters, int threads, size_t smem, cudaStream_t s, Args... args) {
  smem = std::max<size_t>(smem, kOneCtaSmem);
  if (!*attr_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cs > 8) CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    *attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * cs), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CSN_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launches(1);
  return CSN_OK;
}

// Trial groups per cluster: one while all groups' clusters are co-resident (lowest step latency), two beyond that -- a
// second wave would double the time, a second group per cluster hides in the first one's exchange.  A B200 holds 7
// clusters of 16 CTAs / 15 of 8 at one CTA per SM (scripts/cluster_xchg_bench.cu: cudaOccupancyMaxActiveClusters).
static int pick_groups(int B, int cs) {
  static const int forced = [] { const char* e = getenv("CSN_CLUSTER_GROUPS"); return e ? atoi(e) : 0; }();
  if (forced == 1 || forced == 2) return forced;
  const int resident = cs > 8 ? 7 : 15;
  return ceil_div(B, clus::kNT) > resident ? 2 : 1;
}

// May GEMM-shaped work run BESIDE the recurrence of B trials?  Only when the recurrence is one wave of clusters that
// leaves at least half of the SMs free: with less, the GEMM chunks crawl (measured at B = 512: 16 clusters in three waves
// of 112 SMs, the chunks on the 36 SMs left, forward 3.2 -> 21 ms) and the later waves queue behind them.
bool lstm_cluster_overlap_ok(int B, int H) {
  if (!lstm_cluster_supported(H)) return false;
  const int cs = H / clus::kUnits;
  const int clusters = ceil_div(B, clus::kNT * pick_groups(B, cs));
  return clusters <= (cs > 8 ? 7 : 15) && 2 * clusters * cs <= sm_count();
}

template <int CS, int G>
static int launch_fwd_cluster(const CUtensorMap& tm_xp, const uint32_t* w_rows, __nv_bfloat16* h_seq, __nv_bfloat16* gates, float* c_seq,
                              int T, int B, const unsigned* xp_flags, int xp_chunk, uint8_t* xchg, cudaStream_t s) {
  static bool attr_set[2] = {false, false};
  const int clusters = ceil_div(B, clus::kNT * G);
  const size_t smem = size_t(G) * clus::FwdSmem<CS>::total;
  if (g_clus_prof)
    return launch_cluster(clus::lstm_fwd_cluster_kernel<CS, G, true>, &attr_set[1], CS, clusters, clus::kThreads * G, smem, s, tm_xp, w_rows,
                          h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, g_clus_prof);
  return launch_cluster(clus::lstm_fwd_cluster_kernel<CS, G, false>, &attr_set[0], CS, clusters, clus::kThreads * G, smem, s, tm_xp, w_rows,
                        h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, g_clus_prof);
}

// Forward recurrence of one layer.  xp: [T*B, 4H] fp32 (input projection + biases, gate-interleaved columns);
// whh_perm: [4H, H] bf16, rows gate-interleaved.  xp_flags (may be NULL): one word per chunk of xp_chunk timesteps, non-zero
// once that chunk of xp has been written by the projection GEMM running beside this launch.
// xchg (may be NULL: DSMEM pushes instead): lstm_cluster_xchg_bytes(B, H) of global scratch for the per-step exchange.
size_t lstm_cluster_xchg_bytes(int B, int H) {
  // per 16-trial group: two buffers x CS slices (forward) / 2 x CS x CS partial blocks (backward)
  const size_t cs = size_t(H / clus::kUnits);
  return size_t(ceil_div(B, clus::kNT) + 1) * 2 * cs * cs * clus::kPiece;
}
int lstm_cluster_fwd(const float* xp, const __nv_bfloat16* whh_perm, __nv_bfloat16* h_seq, __nv_bfloat16* gates, float* c_seq, int T,
                     int B, int H, const unsigned* xp_flags, int xp_chunk, void* xchg_v, cudaStream_t s) {
  static const bool no_mc = [] { const char* e = getenv("CSN_CLUSTER_NO_MULTICAST"); return e && e[0] == '1'; }();
  uint8_t* xchg = no_mc ? nullptr : reinterpret_cast<uint8_t*>(xchg_v);
  CUtensorMap tm{};
  CSN_TRY(make_tmap_2d_plain(&tm, xp, 4, (uint64_t)(4 * H), (uint64_t)T * B, (uint64_t)(4 * H), 128, clus::kNT));
  const uint32_t* w_rows = reinterpret_cast<const uint32_t*>(whh_perm);
  if (H == 512)
    return pick_groups(B, 16) == 2 ? launch_fwd_cluster<16, 2>(tm, w_rows, h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, s)
                                   : launch_fwd_cluster<16, 1>(tm, w_rows, h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, s);
  if (H == 256)
    return pick_groups(B, 8) == 2 ? launch_fwd_cluster<8, 2>(tm, w_rows, h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, s)
                                  : launch_fwd_cluster<8, 1>(tm, w_rows, h_seq, gates, c_seq, T, B, xp_flags, xp_chunk, xchg, s);
  set_error("lstm_cluster_fwd: unsupported hidden size %d", H);
  return CSN_EUNSUPPORTED;
}

template <int CS, int G>
static int launch_bwd_cluster(const CUtensorMap& tm_g, const CUtensorMap& tm_c, const CUtensorMap& tm_dh, const uint32_t* wt_rows,
                              const float* c_seq, const float* d_hlast, int has_dhseq, __nv_bfloat16* dG, int T, int B, unsigned* done,
                              int done_chunk, cudaStream_t s) {
  static bool attr_set[2] = {false, false};
  const int clusters = ceil_div(B, clus::kNT * G);
  const size_t smem = size_t(G) * clus::BwdSmem<CS>::total;
  const int threads = clus::kBwdWarps * 32 * G;
  if (g_clus_prof)
    return launch_cluster(clus::lstm_bwd_cluster_kernel<CS, G, true>, &attr_set[1], CS, clusters, threads, smem, s, tm_g, tm_c, tm_dh,
                          wt_rows, c_seq, d_hlast, has_dhseq, dG, T, B, done, done_chunk, g_clus_prof);
  return launch_cluster(clus::lstm_bwd_cluster_kernel<CS, G, false>, &attr_set[0], CS, clusters, threads, smem, s, tm_g, tm_c, tm_dh,
                        wt_rows, c_seq, d_hlast, has_dhseq, dG, T, B, done, done_chunk, g_clus_prof);
}

// Backward recurrence of one layer: dG [T*B, 4H] bf16 (gate-interleaved) from the reserve (gates, c_seq) and the gradients
// arriving from above (d_hseq [T,B,H] and / or d_hlast [B,H], fp32).  whh_t: W_hh^T as bf16 [H][4H], columns interleaved.
// done (may be NULL): zeroed counters, one per chunk of done_chunk timesteps; every (CTA, trial group) adds 1 to a chunk's
// counter once all its dG rows of that chunk are stored -- lstm_cluster_participants(B, H) of them in total.
int lstm_cluster_participants(int B, int H) {
  const int cs = H / clus::kUnits, g = pick_groups(B, cs);
  return ceil_div(B, clus::kNT * g) * cs * g;
}
int lstm_cluster_bwd(const __nv_bfloat16* gates, const float* c_seq, const float* d_hseq, const float* d_hlast,
                     const __nv_bfloat16* whh_t, __nv_bfloat16* dG, int T, int B, int H, unsigned* done, int done_chunk,
                     cudaStream_t s) {
  CUtensorMap tm_g{}, tm_c{}, tm_dh{};
  const uint64_t tb = (uint64_t)T * B;
  CSN_TRY(make_tmap_2d_plain(&tm_g, gates, 2, (uint64_t)(4 * H), tb, (uint64_t)(4 * H), 128, clus::kNT));
  CSN_TRY(make_tmap_2d_plain(&tm_c, c_seq, 4, (uint64_t)H, tb, (uint64_t)H, clus::kUnits, clus::kNT));
  CSN_TRY(make_tmap_2d_plain(&tm_dh, d_hseq ? d_hseq : c_seq, 4, (uint64_t)H, tb, (uint64_t)H, clus::kUnits, clus::kNT));
  const uint32_t* wt = reinterpret_cast<const uint32_t*>(whh_t);
  const int hd = d_hseq ? 1 : 0;
  if (H == 512)
    return pick_groups(B, 16) == 2 ? launch_bwd_cluster<16, 2>(tm_g, tm_c, tm_dh, wt, c_seq, d_hlast, hd, dG, T, B, done, done_chunk, s)
                                   : launch_bwd_cluster<16, 1>(tm_g, tm_c, tm_dh, wt, c_seq, d_hlast, hd, dG, T, B, done, done_chunk, s);
  if (H == 256)
    return pick_groups(B, 8) == 2 ? launch_bwd_cluster<8, 2>(tm_g, tm_c, tm_dh, wt, c_seq, d_hlast, hd, dG, T, B, done, done_chunk, s)
                                  : launch_bwd_cluster<8, 1>(tm_g, tm_c, tm_dh, wt, c_seq, d_hlast, hd, dG, T, B, done, done_chunk, s);
  set_error("lstm_cluster_bwd: unsupported hidden size %d", H);
  return CSN_EUNSUPPORTED;
}

}  // namespace csn
