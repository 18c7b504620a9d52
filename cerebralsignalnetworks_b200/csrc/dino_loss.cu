// DINO cross-entropy: teacher softmax (centred, sharpened), student log-softmax, loss, dLoss/dstudent and the
// centre statistics, in ONE pass over HBM (every input element is read once, every gradient written once).
//   single-view : LstmDistillFromDinoV2Train.py:62-105
//   multi-crop  : LstmDistillation.py:118-159 (reference behaviour incl. chunk(1) quirk) / dino/main_dino.py:428-481
//
// Decomposition: a thread-block CLUSTER owns one batch row b across all views; CTA c of the cluster owns the
// K/CS-wide column chunk c.  The teacher probabilities q[g] and the current student row live in registers
// (<= 32 values per thread per row), row statistics (max, sum-exp, <Q,s>) are reduced with warp shuffles inside
// the CTA and through distributed shared memory across the cluster.  So the K=65536 head (cfg 3) reads its
// [6,B,K] + [2,B,K] inputs exactly once and never spills a row to HBM.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace csn {

constexpr int kLossThreads = 256;
constexpr int kMaxVt = 2;
constexpr int kMaxVs = 16;
constexpr int kMaxRounds = kMaxVt + kMaxVs;

struct LossParams {
  const float* student;
  const float* teacher;
  const float* center;
  float* loss;
  float* d_student;
  float* batch_center;
  void* ws;         // deterministic loss fold: ticket + one partial per participating CTA (common.cuh)
  int center_rows;  // 1 or B
  int Vs, Vt, B, K, Kc;
  int mode;
  float inv_tau_s, inv_tau_t, coef, grad_coef;
  unsigned mask[kMaxVs];  // bit g set: student view v is matched against teacher view g
};

struct Stat {
  float m, z, d;
};

__device__ __forceinline__ Stat combine(Stat a, Stat b) {
  Stat r;
  r.m = fmaxf(a.m, b.m);
  float ea = (a.m == -INFINITY) ? 0.f : __expf(a.m - r.m);
  float eb = (b.m == -INFINITY) ? 0.f : __expf(b.m - r.m);
  r.z = a.z * ea + b.z * eb;
  r.d = a.d + b.d;
  return r;
}

__device__ __forceinline__ Stat warp_combine(Stat s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stat t;
    t.m = __shfl_xor_sync(0xffffffffu, s.m, o);
    t.z = __shfl_xor_sync(0xffffffffu, s.z, o);
    t.d = __shfl_xor_sync(0xffffffffu, s.d, o);
    s = combine(s, t);
  }
  return s;
}

// CTA-level then cluster-level reduction.  `slot` is this CTA's per-round mailbox in shared memory; peers read
// it through DSMEM after one cluster barrier.  Rounds never reuse a slot, so one barrier per round suffices.
__device__ __forceinline__ Stat cluster_reduce(Stat s, Stat* warp_buf, Stat* slot, cg::cluster_group& cluster, int cs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s = warp_combine(s);
  if (lane == 0) warp_buf[warp] = s;
  __syncthreads();
  if (warp == 0) {
    Stat t = (lane < kLossThreads / 32) ? warp_buf[lane] : Stat{-INFINITY, 0.f, 0.f};
    t = warp_combine(t);
    if (lane == 0) *slot = t;
  }
  if (cs > 1) {
    cluster.sync();
    Stat acc{-INFINITY, 0.f, 0.f};
    for (int r = 0; r < cs; ++r) {
      const Stat* peer = cluster.map_shared_rank(slot, r);
      Stat t = *peer;
      acc = combine(acc, t);
    }
    return acc;
  }
  __syncthreads();
  return *slot;
}

template <int VEC, int NITER>
__global__ void __launch_bounds__(kLossThreads) dino_loss_kernel(const LossParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int k0 = crank * p.Kc;
  constexpr int EPT = VEC * NITER;

  __shared__ Stat warp_buf[kLossThreads / 32];
  __shared__ Stat slots[kMaxRounds];
  int round = 0;

  float q[kMaxVt][EPT];
  float cen[EPT];
  float bc[EPT];

  auto elem = [&](int i) { return ((i / VEC) * kLossThreads + tid) * VEC + (i % VEC); };  // chunk-local index

  auto load_row = [&](const float* base, float (&dst)[EPT]) {
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      int e = (it * kLossThreads + tid) * VEC;
      if (e < p.Kc) {
        if constexpr (VEC == 4) {
          float4 v = __ldcs(reinterpret_cast<const float4*>(base + k0 + e));
          dst[it * 4 + 0] = v.x; dst[it * 4 + 1] = v.y; dst[it * 4 + 2] = v.z; dst[it * 4 + 3] = v.w;
        } else {
          dst[it] = __ldcs(base + k0 + e);
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) dst[it * VEC + j] = 0.f;
      }
    }
  };
  auto store_row = [&](float* base, const float (&src)[EPT]) {
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      int e = (it * kLossThreads + tid) * VEC;
      if (e < p.Kc) {
        if constexpr (VEC == 4) {
          __stcs(reinterpret_cast<float4*>(base + k0 + e),
                 make_float4(src[it * 4 + 0], src[it * 4 + 1], src[it * 4 + 2], src[it * 4 + 3]));
        } else {
          __stcs(base + k0 + e, src[it]);
        }
      }
    }
  };

  {
    const float* cbase = p.center + (p.center_rows > 1 ? size_t(b) * p.K : 0);
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      int e = (it * kLossThreads + tid) * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) cen[it * VEC + j] = (e < p.Kc) ? cbase[k0 + e + j] : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) bc[i] = 0.f;

  // ---- teacher rows: q[g] = softmax((t - c) / tau_t) ----
#pragma unroll
  for (int g = 0; g < kMaxVt; ++g) {
    if (g < p.Vt) {
      load_row(p.teacher + (size_t(g) * p.B + b) * p.K, q[g]);
      Stat s{-INFINITY, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        bc[i] += q[g][i];
        q[g][i] = (q[g][i] - cen[i]) * p.inv_tau_t;
        if (elem(i) < p.Kc) s.m = fmaxf(s.m, q[g][i]);
      }
#pragma unroll
      for (int i = 0; i < EPT; ++i)
        if (elem(i) < p.Kc) s.z += __expf(q[g][i] - s.m);
      s = cluster_reduce(s, warp_buf, &slots[round++], cluster, cs);
      const float inv_z = 1.f / s.z;
#pragma unroll
      for (int i = 0; i < EPT; ++i) q[g][i] = __expf(q[g][i] - s.m) * inv_z;
    } else {
#pragma unroll
      for (int i = 0; i < EPT; ++i) q[g][i] = 0.f;
    }
  }

  // ---- centre statistics (shared-centre modes: a fixed-order column sum over the teacher follows the kernel) ----
  if (p.mode == CSN_DINO_MULTICROP_REF) store_row(p.batch_center + size_t(b) * p.K, bc);

  // ---- student rows ----
  float loss_acc = 0.f;
  for (int v = 0; v < p.Vs; ++v) {
    const unsigned mask = p.mask[v];
    float* gout = p.d_student + (size_t(v) * p.B + b) * p.K;
    float u[EPT];
    if (mask == 0) {  // view not matched against any teacher view: zero gradient (cluster-uniform branch)
#pragma unroll
      for (int i = 0; i < EPT; ++i) u[i] = 0.f;
      store_row(gout, u);
      continue;
    }
    load_row(p.student + (size_t(v) * p.B + b) * p.K, u);
    const float w0 = (mask & 1u) ? 1.f : 0.f, w1 = (mask & 2u) ? 1.f : 0.f;
    const float W = w0 + w1;
    Stat s{-INFINITY, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      u[i] *= p.inv_tau_s;
      if (elem(i) < p.Kc) {
        s.m = fmaxf(s.m, u[i]);
        s.d += (w0 * q[0][i] + w1 * q[1][i]) * u[i];
      }
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (elem(i) < p.Kc) s.z += __expf(u[i] - s.m);
    s = cluster_reduce(s, warp_buf, &slots[round++], cluster, cs);
    loss_acc += p.coef * (W * (s.m + __logf(s.z)) - s.d);
    const float inv_z = 1.f / s.z;
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      u[i] = p.grad_coef * (W * __expf(u[i] - s.m) * inv_z - (w0 * q[0][i] + w1 * q[1][i]));
    store_row(gout, u);
  }
  if (cs > 1) cluster.sync();  // peers may still be reading our mailboxes
  if (crank == 0) det_cta_sum(loss_acc, p.ws, (unsigned)b, (unsigned)p.B, p.loss);
}


// ---------------------------------------------------------------------------------------------------------------
// Staged variant (the HBM-bound shapes: K % 4 == 0, 16-byte aligned rows).  Same decomposition as dino_loss_kernel,
// rebuilt around three rules measured with ncu on the cfg3 shape (profiles/ncu_loss_r01d.md):
//   * row chunks are pulled by the TMA unit: one thread issues 1-D bulk copies (cp.async.bulk global -> shared,
//     mbarrier complete_tx) for the next S rows into a shared-memory ring while the CTA works on the current row;
//   * instruction diet: logits are scaled into the log2 domain once, every element costs ONE ex2 (taken against the
//     thread-local maximum and rescaled by a per-thread scalar once the row statistics are known), the teacher
//     probabilities are pre-summed when every student view is matched against all teacher views (ALLSAME);
//   * the cluster exchange is a PUSH: warp 0 sends the CTA's (max, sum, dot) to every peer's mailbox with st.async
//     (data + mbarrier complete_tx in one DSMEM packet), everybody waits on a local mbarrier.  No cluster.sync (and
//     its L1 flush) and no remote loads on the critical path.
// Measured and dropped: a two-pass variant that takes the exchange off the row chain (pass 1 of row r -> push, then pass 2
// of row r - 1 re-reads its ring stage against the global statistics).  Parity-green but 60 -> 70 us at the cfg3 shape:
// the stage of row r - 1 stays occupied one row longer, so only one row per CTA is in flight instead of two, and the
// memory-level parallelism lost costs more than the exchange wait it hides.  Keeping the row's 2^(x - max) in registers
// instead (gradient of row r deferred behind pass 1 of row r + 1, ring untouched) also measured 70 us: 32 more live
// registers per thread spill at the 128-register cap of two CTAs per SM.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// 16 bytes into a peer CTA's shared memory + complete_tx(16) on the peer's mbarrier, one DSMEM packet
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cluster_arrive_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kNegBig = -1.0e30f;   // padding logit: never the maximum, ex2(kNegBig - m) == 0, 0 * kNegBig == -0
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct RowStat {
  float m, z, d;  // log2-domain maximum, sum of 2^(x - m), <Q, x>
};

template <int NITER, bool ALLSAME>
__global__ void __launch_bounds__(kLossThreads, 2) dino_loss_staged_kernel(const LossParams p, const int S) {
  uint32_t cs, crank;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k0 = (int)crank * p.Kc;
  constexpr int EPT = 4 * NITER;
  constexpr int NQ = ALLSAME ? 1 : kMaxVt;
  const uint32_t row_bytes = (uint32_t)p.Kc * 4u;

  extern __shared__ __align__(128) uint8_t smem_dyn[];
  float* stages = reinterpret_cast<float*>(smem_dyn);                   // [S][Kc]
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + size_t(S) * p.Kc);  // [S] row barriers
  __shared__ __align__(16) float4 mailbox[kMaxRounds][8];  // [round][sender rank] = (m, z, d, -)
  __shared__ __align__(16) float4 warp_buf[kLossThreads / 32];
  __shared__ __align__(8) uint64_t xbar[kMaxRounds];       // one exchange barrier per round, never reused
  __shared__ int vlist[kMaxVs];

  // Prologue, spread over threads so that nothing serial sits in front of the first bulk copy: every thread counts the
  // active views itself (a few constant-bank reads), warp 2 builds the view list with a ballot, the first threads of
  // warp 0 initialise one exchange barrier each, and the PRODUCER thread initialises its own ring barriers and starts
  // the first S row copies before the CTA even synchronises.
  int n_active = 0;
  for (int v = 0; v < p.Vs; ++v) n_active += p.mask[v] ? 1 : 0;
  const int nrows = p.Vt + n_active;
  auto nth_active = [&](int a) -> int {
    int v = 0;
    for (; v < p.Vs; ++v)
      if (p.mask[v] && a-- == 0) break;
    return v;
  };
  auto row_src = [&](int r) -> const float* {  // producer only
    if (r < p.Vt) return p.teacher + (size_t(r) * p.B + b) * p.K + k0;
    return p.student + (size_t(nth_active(r - p.Vt)) * p.B + b) * p.K + k0;
  };
  constexpr int kProducer = 32;  // warp 1 lane 0 feeds the ring; warp 0 owns the exchange
  auto issue = [&](int r) {
    uint64_t* bar = &full[r % S];
    mbar_expect_tx_(bar, row_bytes);
    bulk_g2s(stages + size_t(r % S) * p.Kc, row_src(r), row_bytes, bar);
  };
  if (tid == kProducer) {
    for (int i = 0; i < S; ++i) mbar_init_(&full[i], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the inits are visible to the TMA unit
    for (int r = 0; r < S && r < nrows; ++r) issue(r);
  } else if (tid < nrows) {
    mbar_init_(&xbar[tid], 1);
    mbar_expect_tx_(&xbar[tid], cs * 16u);  // the phase completes when all cs mailboxes of the round have landed
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    const bool act = lane < p.Vs && p.mask[lane < p.Vs ? lane : 0] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (act) vlist[__popc(bal & ((1u << lane) - 1u))] = lane;
  }
  __syncthreads();
  // peers may push into our mailboxes once they have seen this arrive (waited for before push 0); the barrier inits
  // were released by fence.mbarrier_init, so the arrive itself carries no memory fence
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  auto in_range = [&](int it) { return (it * kLossThreads + tid) * 4 < p.Kc; };
  auto store_row = [&](float* base, const float (&src)[EPT]) {
#pragma unroll
    for (int it = 0; it < NITER; ++it)
      if (in_range(it))
        __stcs(reinterpret_cast<float4*>(base + k0 + (it * kLossThreads + tid) * 4),
               make_float4(src[it * 4 + 0], src[it * 4 + 1], src[it * 4 + 2], src[it * 4 + 3]));
  };

  // CTA reduction of the per-thread (max, sum 2^(x - max), dot), push to every peer, combine the cs mailboxes.
  // The __syncthreads also certifies that every thread has pulled its values of row r out of the ring, so the
  // producer refills that stage right behind it.
  bool first_push = true;
  auto reduce = [&](float mt, float zt, float dt, int r) -> RowStat {
    const float wm = warp_max(mt);
    zt *= ex2_(mt - wm);
    zt = warp_sum(zt);
    dt = warp_sum(dt);
    if (lane == 0) warp_buf[warp] = make_float4(wm, zt, dt, 0.f);
    if (first_push) { cluster_wait_(); first_push = false; }
    __syncthreads();
    if (tid == kProducer && r + S < nrows) issue(r + S);
    if (warp == 0) {
      float4 t = (lane < kLossThreads / 32) ? warp_buf[lane] : make_float4(kNegBig, 0.f, 0.f, 0.f);
      float cm = t.x;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
      float cz = t.y * ex2_(t.x - cm), cd = t.z;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
        cd += __shfl_xor_sync(0xffffffffu, cd, o);
      }
      if (lane < (int)cs)
        st_async_f4(map_to_cta(smem_u32(&mailbox[r][crank]), lane), cm, cz, cd, 0.f, map_to_cta(smem_u32(&xbar[r]), lane));
    }
    mbar_wait_(&xbar[r], 0);
    RowStat o{kNegBig, 0.f, 0.f};
    float4 mb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mb[j] = (j < (int)cs) ? mailbox[r][j] : make_float4(kNegBig, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.m = fmaxf(o.m, mb[j].x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o.z = fmaf(mb[j].y, ex2_(mb[j].x - o.m), o.z);
      o.d += mb[j].z;
    }
    return o;
  };

  // student views that are not matched against any teacher view: zero gradient
  {
    float z[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) z[i] = 0.f;
    for (int v = 0; v < p.Vs; ++v)
      if (p.mask[v] == 0) store_row(p.d_student + (size_t(v) * p.B + b) * p.K, z);
  }

  float q[NQ][EPT];  // teacher probabilities (ALLSAME: their sum over the teacher views)
  float bc[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) bc[i] = 0.f;
  const float ts = p.inv_tau_t * kLog2e, ss = p.inv_tau_s * kLog2e;
  const float* cbase = p.center + (p.center_rows > 1 ? size_t(b) * p.K : 0) + k0;

  // ---- teacher rows: q[g] = softmax((t - c) / tau_t) ----
#pragma unroll
  for (int g = 0; g < kMaxVt; ++g) {
    if (g < p.Vt) {
      float e[EPT];
      // centre chunk straight from L2/HBM while the row's bulk copy is still in flight
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        const float4 c4 = in_range(it) ? __ldg(reinterpret_cast<const float4*>(cbase + (it * kLossThreads + tid) * 4))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        e[it * 4 + 0] = c4.x; e[it * 4 + 1] = c4.y; e[it * 4 + 2] = c4.z; e[it * 4 + 3] = c4.w;
      }
      mbar_wait_(&full[g % S], (uint32_t)((g / S) & 1));
      const float* src = stages + size_t(g % S) * p.Kc;
      float mx[4] = {kNegBig, kNegBig, kNegBig, kNegBig};
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        if (in_range(it)) {
          const float4 t4 = *reinterpret_cast<const float4*>(src + (it * kLossThreads + tid) * 4);
          const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = it * 4 + j;
            bc[i] += tt[j];
            e[i] = (tt[j] - e[i]) * ts;
            mx[j] = fmaxf(mx[j], e[i]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) e[it * 4 + j] = kNegBig;
        }
      }
      const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float zs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        e[i] = ex2_(e[i] - mt);
        zs[i & 3] += e[i];
      }
      const RowStat st = reduce(mt, (zs[0] + zs[1]) + (zs[2] + zs[3]), 0.f, g);
      const float A = __fdividef(ex2_(mt - st.m), st.z);
      if (ALLSAME && g > 0) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) q[0][i] = fmaf(e[i], A, q[0][i]);
      } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) q[ALLSAME ? 0 : g][i] = e[i] * A;
      }
    } else if (!ALLSAME) {
#pragma unroll
      for (int i = 0; i < EPT; ++i) q[ALLSAME ? 0 : g][i] = 0.f;
    }
  }

  // ---- centre statistics (shared-centre modes: a fixed-order column sum over the teacher follows the kernel) ----
  if (p.mode == CSN_DINO_MULTICROP_REF) store_row(p.batch_center + size_t(b) * p.K, bc);

  // ---- student rows ----
  float loss_acc = 0.f;
  for (int a = 0; a < n_active; ++a) {
    const int r = p.Vt + a;
    const int v = vlist[a];
    const unsigned mask = p.mask[v];
    const float w0 = (mask & 1u) ? 1.f : 0.f, w1 = (mask & 2u) ? 1.f : 0.f;
    const float W = ALLSAME ? float(p.Vt) : (w0 + w1);
    auto Qv = [&](int i) -> float {
      if constexpr (ALLSAME) return q[0][i];
      else return fmaf(w1, q[NQ - 1][i], w0 * q[0][i]);
    };
    float e[EPT];
    mbar_wait_(&full[r % S], (uint32_t)((r / S) & 1));
    const float* src = stages + size_t(r % S) * p.Kc;
    float mx[4] = {kNegBig, kNegBig, kNegBig, kNegBig};
    float ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      if (in_range(it)) {
        const float4 u4 = *reinterpret_cast<const float4*>(src + (it * kLossThreads + tid) * 4);
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = it * 4 + j;
          e[i] = uu[j] * ss;
          mx[j] = fmaxf(mx[j], e[i]);
          ds[j] = fmaf(Qv(i), e[i], ds[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) e[it * 4 + j] = kNegBig;
      }
    }
    const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    float zs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      e[i] = ex2_(e[i] - mt);
      zs[i & 3] += e[i];
    }
    const RowStat st = reduce(mt, (zs[0] + zs[1]) + (zs[2] + zs[3]), (ds[0] + ds[1]) + (ds[2] + ds[3]), r);
    loss_acc += p.coef * kLn2 * (W * (st.m + lg2_(st.z)) - st.d);
    const float A = W * __fdividef(ex2_(mt - st.m), st.z);
#pragma unroll
    for (int i = 0; i < EPT; ++i) e[i] = p.grad_coef * fmaf(e[i], A, -Qv(i));
    store_row(p.d_student + (size_t(v) * p.B + b) * p.K, e);
  }
  if (first_push) cluster_wait_();  // (no rows at all: still pair the arrive)
  if (crank == 0) det_cta_sum(loss_acc, p.ws, (unsigned)b, (unsigned)p.B, p.loss);
  // No closing cluster barrier: a CTA gets here only after the LAST round's cs packets have landed in its mailbox, i.e.
  // after every peer has issued its last push to it, and it never reads remote memory -- nobody can still touch the
  // shared memory of a CTA that exits.  (A release-arrive here made every CTA wait for its gradient stores to drain.)
}

template <int NITER, bool ALLSAME>
static int launch_loss_staged(const LossParams& p, int cs, int S, cudaStream_t s) {
  const size_t smem = size_t(S) * p.Kc * 4 + size_t(S) * 8 + 128;
  auto kern = dino_loss_staged_kernel<NITER, ALLSAME>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cs, p.B, 1);
  cfg.blockDim = dim3(kLossThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA(cudaLaunchKernelEx(&cfg, kern, p, S));
  return CSN_OK;
}

template <bool ALLSAME>
static int dispatch_loss_staged(int per_thread, const LossParams& p, int cs, int S, cudaStream_t s) {
  if (per_thread <= 1) return launch_loss_staged<1, ALLSAME>(p, cs, S, s);
  if (per_thread <= 2) return launch_loss_staged<2, ALLSAME>(p, cs, S, s);
  if (per_thread <= 4) return launch_loss_staged<4, ALLSAME>(p, cs, S, s);
  return launch_loss_staged<8, ALLSAME>(p, cs, S, s);
}


// ---------------------------------------------------------------------------------------------------------------
// Narrow heads (K <= 1024: the 384 / 768-d DINOv2 feature targets of LstmDistillFromDinoV2Train.py): ONE WARP per
// batch row, values strided over the lanes (coalesced 128-byte loads), every row statistic is a warp-shuffle
// reduction -- no shared-memory staging, no block or cluster barrier on the row path.  Same log2-domain arithmetic
// as the staged kernel.  The centre statistics of the CTA's rows are folded in shared memory first (one atomic per
// column per CTA instead of one per row).
constexpr int kRowWarps = 8;

template <int EPT>
__global__ void __launch_bounds__(kRowWarps * 32) dino_loss_rowwarp_kernel(const LossParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kRowWarps + warp;
  const bool row_ok = b < p.B;
  const int K = p.K;
  __shared__ float loss_part[kRowWarps];
  const float ts = p.inv_tau_t * kLog2e, ss = p.inv_tau_s * kLog2e;
  float loss_acc = 0.f;
  float bc[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) bc[i] = 0.f;

  if (row_ok) {
    float q[kMaxVt][EPT];
    const float* cbase = p.center + (p.center_rows > 1 ? size_t(b) * K : 0);
#pragma unroll
    for (int g = 0; g < kMaxVt; ++g) {
      if (g < p.Vt) {
        const float* trow = p.teacher + (size_t(g) * p.B + b) * K;
        float mt = kNegBig;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int k = lane + 32 * i;
          if (k < K) {
            const float t = __ldcs(trow + k);
            bc[i] += t;
            q[g][i] = (t - __ldg(cbase + k)) * ts;
            mt = fmaxf(mt, q[g][i]);
          } else {
            q[g][i] = kNegBig;
          }
        }
        const float m = warp_max(mt);
        float z = 0.f;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          q[g][i] = ex2_(q[g][i] - m);
          z += q[g][i];
        }
        const float inv_z = __fdividef(1.f, warp_sum(z));
#pragma unroll
        for (int i = 0; i < EPT; ++i) q[g][i] *= inv_z;
      } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) q[g][i] = 0.f;
      }
    }
    if (p.mode == CSN_DINO_MULTICROP_REF) {  // per-row centre statistics (reference quirk Q3): plain stores
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int k = lane + 32 * i;
        if (k < K) p.batch_center[size_t(b) * K + k] = bc[i];
      }
    }
    for (int v = 0; v < p.Vs; ++v) {
      const unsigned mask = p.mask[v];
      float* gout = p.d_student + (size_t(v) * p.B + b) * K;
      if (mask == 0) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int k = lane + 32 * i;
          if (k < K) __stcs(gout + k, 0.f);
        }
        continue;
      }
      const float w0 = (mask & 1u) ? 1.f : 0.f, w1 = (mask & 2u) ? 1.f : 0.f;
      const float W = w0 + w1;
      const float* srow = p.student + (size_t(v) * p.B + b) * K;
      float e[EPT];
      float mt = kNegBig, d = 0.f;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int k = lane + 32 * i;
        e[i] = (k < K) ? __ldcs(srow + k) * ss : kNegBig;
        mt = fmaxf(mt, e[i]);
        d = fmaf(fmaf(w1, q[1][i], w0 * q[0][i]), e[i], d);
      }
      const float m = warp_max(mt);
      float z = 0.f;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        e[i] = ex2_(e[i] - m);
        z += e[i];
      }
      z = warp_sum(z);
      d = warp_sum(d);
      loss_acc += p.coef * kLn2 * (W * (m + lg2_(z)) - d);
      const float A = W * __fdividef(1.f, z);
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int k = lane + 32 * i;
        if (k < K) __stcs(gout + k, p.grad_coef * fmaf(e[i], A, -fmaf(w1, q[1][i], w0 * q[0][i])));
      }
    }
  }
  // ---- loss: CTA fold in warp order, then the deterministic cross-CTA fold (the shared-centre column sums are a
  //      fixed-order pass over the teacher after the kernel) ----
  if (lane == 0) loss_part[warp] = loss_acc;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < kRowWarps; ++w) t += loss_part[w];
  det_cta_sum(t, p.ws, blockIdx.x, gridDim.x, p.loss);
}

template <int EPT>
static int launch_loss_rowwarp(const LossParams& p, cudaStream_t s) {
  dino_loss_rowwarp_kernel<EPT><<<ceil_div(p.B, kRowWarps), kRowWarps * 32, 0, s>>>(p);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

template <int VEC, int NITER>
static int launch_loss(const LossParams& p, int cs, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cs, p.B, 1);
  cfg.blockDim = dim3(kLossThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA(cudaLaunchKernelEx(&cfg, dino_loss_kernel<VEC, NITER>, p));
  return CSN_OK;
}

__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ bc, size_t n, float mom,
                                  float scale) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) center[i] = center[i] * mom + bc[i] * scale * (1.f - mom);
}

}  // namespace csn

using namespace csn;

extern "C" int csn_dino_loss_fwd_bwd(const float* student, const float* teacher, const float* center, int center_rows,
                                     float student_temp, float teacher_temp, float* loss, float* d_student,
                                     float* batch_center, int Vs, int Vt, int B, int K, int mode, float grad_scale,
                                     void* workspace, void* stream) {
  CSN_REQUIRE(student && teacher && center && loss && d_student && workspace, "csn_dino_loss_fwd_bwd: null pointer");
  CSN_REQUIRE(batch_center || mode != CSN_DINO_MULTICROP_REF, "csn_dino_loss_fwd_bwd: MULTICROP_REF needs batch_center");
  CSN_REQUIRE(Vs >= 1 && Vs <= kMaxVs && Vt >= 1 && Vt <= kMaxVt, "csn_dino_loss_fwd_bwd: need 1<=Vs<=%d, 1<=Vt<=%d", kMaxVs, kMaxVt);
  CSN_REQUIRE(B >= 1 && K >= 1 && B <= 65535, "csn_dino_loss_fwd_bwd: bad B/K");
  CSN_REQUIRE(center_rows == 1 || center_rows == B, "csn_dino_loss_fwd_bwd: center_rows must be 1 or B");
  CSN_REQUIRE(student_temp > 0.f && teacher_temp != 0.f, "csn_dino_loss_fwd_bwd: bad temperature");
  LossParams p{};
  p.student = student; p.teacher = teacher; p.center = center; p.loss = loss; p.d_student = d_student;
  p.batch_center = batch_center; p.center_rows = center_rows; p.ws = workspace;
  p.Vs = Vs; p.Vt = Vt; p.B = B; p.K = K; p.mode = mode;
  p.inv_tau_s = 1.f / student_temp; p.inv_tau_t = 1.f / teacher_temp;
  int n_terms = 0;
  if (mode == CSN_DINO_SINGLE) {
    CSN_REQUIRE(Vs == 1 && Vt == 1, "csn_dino_loss_fwd_bwd: SINGLE mode needs Vs == Vt == 1");
    p.mask[0] = 1u;
    p.coef = 1.f / B;
  } else if (mode == CSN_DINO_MULTICROP_REF) {
    // teacher.chunk(1) -> one "view" holding all Vt teacher views; student view 0 is skipped (v == iq),
    // every other view is averaged over all Vt*B rows (LstmDistillation.py:128,136-144)
    CSN_REQUIRE(Vs >= 2, "csn_dino_loss_fwd_bwd: MULTICROP_REF needs Vs >= 2");
    p.mask[0] = 0u;
    for (int v = 1; v < Vs; ++v) p.mask[v] = (1u << Vt) - 1u;
    n_terms = Vs - 1;
    p.coef = 1.f / (float(n_terms) * float(Vt) * float(B));
  } else if (mode == CSN_DINO_MULTICROP_CANONICAL) {
    for (int v = 0; v < Vs; ++v) {
      p.mask[v] = ((1u << Vt) - 1u) & ~(1u << v);
      n_terms += __builtin_popcount(p.mask[v]);
    }
    CSN_REQUIRE(n_terms > 0, "csn_dino_loss_fwd_bwd: no loss terms");
    p.coef = 1.f / (float(n_terms) * float(B));
  } else {
    CSN_REQUIRE(false, "csn_dino_loss_fwd_bwd: bad mode %d", mode);
  }
  p.grad_coef = p.coef * p.inv_tau_s * grad_scale;

  cudaStream_t s = as_stream(stream);
  CSN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned), s));  // arrival ticket of the loss fold
  // shared-centre modes: batch_center[K] += sum over (view, row) of the teacher, in a fixed order (runs after the loss)
  auto center_sums = [&]() -> int {
    if (mode == CSN_DINO_MULTICROP_REF || !batch_center) return CSN_OK;
    return colsum_det(teacher, batch_center, Vt * B, K, K, 1, s);
  };

  static const bool no_rowwarp = [] { const char* e = getenv("CSN_LOSS_NO_ROWWARP"); return e && e[0] == '1'; }();
  if (K <= 1024 && !no_rowwarp) {  // narrow heads: one warp per batch row
    p.Kc = K;
    if (K <= 128) CSN_TRY(launch_loss_rowwarp<4>(p, s));
    else if (K <= 256) CSN_TRY(launch_loss_rowwarp<8>(p, s));
    else if (K <= 512) CSN_TRY(launch_loss_rowwarp<16>(p, s));
    else CSN_TRY(launch_loss_rowwarp<32>(p, s));
    return center_sums();
  }

  const bool aligned = ((reinterpret_cast<uintptr_t>(student) | reinterpret_cast<uintptr_t>(teacher) |
                         reinterpret_cast<uintptr_t>(d_student) | reinterpret_cast<uintptr_t>(batch_center)) & 15) == 0;
  int cs = 1;
  // widen the cluster until a chunk fits the register budget (and to spread small batches over more SMs)
  auto fits = [&](int c, int vec) { return K % (c * vec) == 0 && K / c <= kLossThreads * vec * 8; };
  int vec = (aligned && K % 4 == 0) ? 4 : 1;
  while (cs <= 8 && !fits(cs, vec)) cs *= 2;
  if (cs > 8 && vec == 4) { vec = 1; cs = 1; while (cs <= 8 && !fits(cs, vec)) cs *= 2; }
  CSN_REQUIRE(cs <= 8, "csn_dino_loss_fwd_bwd: K=%d not supported (needs K %% cluster == 0 and K <= 65536)", K);
  while (cs < 8 && B * cs < sm_count() && fits(cs * 2, vec) && K / (cs * 2) >= kLossThreads * vec) cs *= 2;
  p.Kc = K / cs;
  const int per_thread = ceil_div(p.Kc, kLossThreads * vec);
  int r;
  static const bool no_stage = [] { const char* e = getenv("CSN_LOSS_NO_STAGE"); return e && e[0] == '1'; }();
  if (vec == 4 && !no_stage && (reinterpret_cast<uintptr_t>(center) & 15) == 0) {
    int n_act = 0;
    bool all_same = true;  // every matched student view is matched against ALL teacher views (REF and SINGLE modes)
    for (int v = 0; v < Vs; ++v) {
      n_act += p.mask[v] ? 1 : 0;
      if (p.mask[v] && p.mask[v] != (1u << Vt) - 1u) all_same = false;
    }
    const int nrows = Vt + n_act;
    // ring depth: row chunks that fit in ~100 KB (two CTAs per SM overlap each other's exchange latency), at most 4
    // and at most nrows.  CSN_LOSS_STAGE_KB overrides the budget.
    static const unsigned budget_kb = [] { const char* e = getenv("CSN_LOSS_STAGE_KB"); return e ? (unsigned)atoi(e) : 100u; }();
    int S = (int)((budget_kb * 1024u) / (size_t(p.Kc) * 4u));
    S = S > 4 ? 4 : S;
    S = S > nrows ? nrows : S;
    if (S >= 1) {
      r = all_same ? dispatch_loss_staged<true>(per_thread, p, cs, S, s) : dispatch_loss_staged<false>(per_thread, p, cs, S, s);
      if (r == CSN_OK) count_launches(1);
      return r == CSN_OK ? center_sums() : r;
    }
  }
  if (vec == 4) {
    if (per_thread <= 1) r = launch_loss<4, 1>(p, cs, s);
    else if (per_thread <= 2) r = launch_loss<4, 2>(p, cs, s);
    else if (per_thread <= 4) r = launch_loss<4, 4>(p, cs, s);
    else r = launch_loss<4, 8>(p, cs, s);
  } else {
    if (per_thread <= 1) r = launch_loss<1, 1>(p, cs, s);
    else if (per_thread <= 2) r = launch_loss<1, 2>(p, cs, s);
    else if (per_thread <= 4) r = launch_loss<1, 4>(p, cs, s);
    else r = launch_loss<1, 8>(p, cs, s);
  }
  if (r == CSN_OK) count_launches(1);
  return r == CSN_OK ? center_sums() : r;
}

extern "C" int csn_loss_workspace_bytes(int rows, size_t* bytes) {
  CSN_REQUIRE(rows >= 1 && bytes, "csn_loss_workspace_bytes: bad arguments");
  *bytes = det_scratch_bytes(size_t(rows));
  return CSN_OK;
}

extern "C" int csn_center_ema(float* center, const float* batch_center, size_t n, float momentum, float scale,
                              void* stream) {
  CSN_REQUIRE(center && batch_center, "csn_center_ema: null pointer");
  if (n == 0) return CSN_OK;
  int blocks = (int)std::min<size_t>(ceil_div<size_t>(n, 256), size_t(sm_count()) * 8);
  center_ema_kernel<<<blocks, 256, 0, as_stream(stream)>>>(center, batch_center, n, momentum, scale);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
