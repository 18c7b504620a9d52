// Data-parallel gradient exchange fused into the optimiser: every rank reads the flat [gradients | centre sums] buffer
// of EVERY rank straight over NVLink (peer-mapped symmetric memory), sums in rank order -- so all replicas compute
// bit-identical sums -- and applies Adam + the DINO centre EMA in the same pass.  Replaces ncclAllReduce + adam +
// centre-EMA launches of the step (DDP gradient averaging, LstmDistillation.py:445; update_center all-reduce,
// LstmDistillation.py:154-156).  Cross-rank ordering uses epoch flags in peer memory (system-scope release/acquire):
//   READY[src] = t : rank src finished writing its step-t gradients        (set at the start of the fused kernel)
//   DONE[src]  = t : rank src finished reading everybody's step-t gradients (set by its last CTA)
// A rank may overwrite its gradient buffer for step t+1 only after every DONE[src] >= t (csn_dp_wait_done_zero,
// which also zeroes the centre-sum tail in place of a memset).  Epochs come from the device-side Adam step counter,
// so the kernels are CUDA-graph replay safe.
#include "common.cuh"

namespace csn {

constexpr int kMaxWorld = 8;
constexpr int kFlagReady = 0, kFlagDone = kMaxWorld;  // uint32 slots in each rank's flag block (>= 2 * kMaxWorld words)

struct PeerSet {
  const float* grad[kMaxWorld];
  unsigned* flags[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {  // never served from a stale local cache line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// Watchdog of the flag waits.  A peer that never arrives (crashed rank) must not hang the GPU for ever, but ranks may
// legitimately drift apart by minutes between steps (a rank-0-only checkpoint or evaluation, a slow torch.save, a
// debugger): the limit is CSN_DP_TIMEOUT_S seconds (default 600, the order of NCCL's own watchdog; 0 = wait for
// ever).  Before the trap the waiter records (1, rank, flag index, awaited epoch) in a pinned host slot that
// csn_dp_last_timeout() reads back after the context has died.
struct Watchdog {
  unsigned long long limit_ns;  // 0: no limit
  unsigned* host_slot;          // pinned, mapped; may be NULL
  int rank;
};
__device__ __forceinline__ void wait_flag_ge(const unsigned* flag, unsigned target, const Watchdog& wd, int flag_index) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while ((int)(ld_acquire_sys(flag) - target) < 0) {
    if ((++spins & 1023u) == 0 && wd.limit_ns) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > wd.limit_ns) {
        if (wd.host_slot) {
          wd.host_slot[1] = (unsigned)wd.rank;
          wd.host_slot[2] = (unsigned)flag_index;
          wd.host_slot[3] = target;
          __threadfence_system();
          wd.host_slot[0] = 1u;
          __threadfence_system();
        }
        __trap();
      }
    }
  }
}

static Watchdog make_watchdog(int rank) {
  static const unsigned long long limit = [] {
    const char* e = getenv("CSN_DP_TIMEOUT_S");
    const double sec = e ? atof(e) : 600.0;
    return sec > 0 ? (unsigned long long)(sec * 1e9) : 0ull;
  }();
  static unsigned* slot = [] {
    unsigned* h = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), 4 * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return (unsigned*)nullptr;
    }
    h[0] = h[1] = h[2] = h[3] = 0u;
    return h;
  }();
  return Watchdog{limit, slot, rank};
}

__global__ void dp_wait_done_zero_kernel(const unsigned* __restrict__ flags_local, int world, const int* __restrict__ step_dev,
                                         float* __restrict__ zero_ptr, size_t n, const Watchdog wd) {
  if (world > 1) {
    if (threadIdx.x < world) wait_flag_ge(flags_local + kFlagDone + threadIdx.x, (unsigned)*step_dev, wd, kFlagDone + threadIdx.x);
    __syncthreads();
  }
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) zero_ptr[i] = 0.f;
}

__global__ void __launch_bounds__(256) dp_adam_peer_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                          size_t n_param, const PeerSet ps, int world, int rank,
                                                          float* __restrict__ center, size_t K, float center_momentum,
                                                          float center_scale, int* __restrict__ step_dev, float lr, float b1,
                                                          float b2, float eps, float wd, int decoupled, float grad_scale,
                                                          unsigned* __restrict__ ticket, const Watchdog dog) {
  __shared__ float s_consts[2];
  __shared__ int s_last;
  const int t = *step_dev + 1;  // this step's 1-based count (the counter is advanced by the last CTA)
  if (world > 1) {  // (one rank: stream order already is the ordering)
    if (blockIdx.x == 0 && threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(ps.flags[threadIdx.x] + kFlagReady + rank, (unsigned)t);
    }
    if (threadIdx.x < world) wait_flag_ge(ps.flags[rank] + kFlagReady + threadIdx.x, (unsigned)t, dog, kFlagReady + threadIdx.x);
  }
  // bias corrections 1 - beta^t = -expm1(t log beta): accurate in fp32 for small and large t alike, and cheap enough
  // to recompute in every CTA (a double-precision pow here, or in the last CTA, costs ~3 us of every step)
  if (threadIdx.x == 32) {
    const float bc1 = -expm1f((float)t * logf(b1)), bc2 = -expm1f((float)t * logf(b2));
    s_consts[0] = lr / bc1;
    s_consts[1] = rsqrtf(bc2);
  }
  __syncthreads();
  const float step_size = s_consts[0], inv_sqrt_bc2 = s_consts[1];

  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t n4 = n_param >> 2;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 G[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) G[r] = ld_peer_f4(ps.grad[r] + 4 * i);
    float4 S = G[0];
#pragma unroll
    for (int r = 1; r < kMaxWorld; ++r)
      if (r < world) { S.x += G[r].x; S.y += G[r].y; S.z += G[r].z; S.w += G[r].w; }
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float* pp = &P.x; float* gg = &S.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = gg[j] * grad_scale;
      if (decoupled) pp[j] *= (1.f - lr * wd); else gr = fmaf(wd, pp[j], gr);
      mm[j] = b1 * mm[j] + (1.f - b1) * gr;
      vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  // centre EMA over the tail of the exchanged buffer: center = center * mom + (sum over ranks) * scale * (1 - mom)
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < K; k += stride) {
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) sum += ld_peer_f(ps.grad[r] + n_param + k);
    center[k] = center[k] * center_momentum + sum * center_scale * (1.f - center_momentum);
  }
  // last CTA: advance the step counter and tell every rank that this rank is done reading
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) {
      *ticket = 0u;
      *step_dev = t;
    }
    if (world > 1 && threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(ps.flags[threadIdx.x] + kFlagDone + rank, (unsigned)t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-shot all-reduce over peer memory for LARGE buffers (cfg3: 22.3 M parameters + the per-row centre statistics, 106 MB):
// the one-shot exchange above makes every rank read every rank's whole buffer (4 P world bytes per rank: fine for the
// 0.8 MB of cfg1/2, 8x the minimum at cfg3).  Here rank r REDUCES slice r (reads it from every peer, adds in rank order,
// writes the sum into its own buffer), then every rank GATHERS the other ranks' reduced slices: 2 (n-1)/n 4 P bytes per
// rank over NVLink, the ring-equivalent minimum, in two kernels ordered by epoch flags in peer memory:
//   READY[src]   = e : rank src has written its part of the buffer for exchange e        (start of the reduce kernel)
//   REDUCED[src] = e : rank src has reduced its slice                                      (last CTA of the reduce kernel)
//   DONE[src]    = e : rank src has gathered every slice, nobody's buffer is read any more (last CTA of the gather kernel)
// Every slice is summed by exactly one rank and copied: all replicas hold bit-identical results.  `lane` selects one of
// the flag sets of the block, so independent exchanges (the DINO head's gradients, which are final long before BPTT
// ends, and the rest) can be in flight at the same time on different streams.
constexpr int kFlagReduced = 2 * kMaxWorld;
constexpr int kLaneStride = 4 * kMaxWorld;  // uint32 words per flag set (>= 3 * kMaxWorld)

struct PeerBufs {
  float* buf[kMaxWorld];
  unsigned* flags[kMaxWorld];
};

__device__ __forceinline__ void slice_range(size_t n4, int world, int r, size_t* b, size_t* e) {
  const size_t per = (n4 + world - 1) / world;  // float4 units
  *b = per * r < n4 ? per * r : n4;
  *e = per * (r + 1) < n4 ? per * (r + 1) : n4;
}

__global__ void __launch_bounds__(256) dp_reduce_slice_kernel(const PeerBufs ps, int world, int rank, size_t off, size_t n4,
                                                             const int* __restrict__ epoch_dev, int lane_id,
                                                             unsigned* __restrict__ ticket, const Watchdog dog) {
  __shared__ int s_last;
  const unsigned e = (unsigned)(*epoch_dev + 1);
  const int fl = lane_id * kLaneStride;
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(ps.flags[threadIdx.x] + fl + kFlagReady + rank, e);
  }
  if (threadIdx.x < world) wait_flag_ge(ps.flags[rank] + fl + kFlagReady + threadIdx.x, e, dog, fl + kFlagReady + threadIdx.x);
  __syncthreads();
  size_t b, en;
  slice_range(n4, world, rank, &b, &en);
  for (size_t i = b + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < en; i += size_t(gridDim.x) * blockDim.x) {
    float4 G[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) G[r] = ld_peer_f4(ps.buf[r] + off + 4 * i);
    float4 S = G[0];
#pragma unroll
    for (int r = 1; r < kMaxWorld; ++r)
      if (r < world) { S.x += G[r].x; S.y += G[r].y; S.z += G[r].z; S.w += G[r].w; }
    *reinterpret_cast<float4*>(ps.buf[rank] + off + 4 * i) = S;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *ticket = 0u;
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(ps.flags[threadIdx.x] + fl + kFlagReduced + rank, e);
    }
  }
}

__global__ void __launch_bounds__(256) dp_gather_slices_kernel(const PeerBufs ps, int world, int rank, size_t off, size_t n4,
                                                              int* __restrict__ epoch_dev, int lane_id,
                                                              unsigned* __restrict__ ticket, const Watchdog dog) {
  __shared__ int s_last;
  const unsigned e = (unsigned)(*epoch_dev + 1);
  const int fl = lane_id * kLaneStride;
  if (threadIdx.x < world) wait_flag_ge(ps.flags[rank] + fl + kFlagReduced + threadIdx.x, e, dog, fl + kFlagReduced + threadIdx.x);
  __syncthreads();
  for (int k = 1; k < world; ++k) {
    const int r = (rank + k) % world;  // start with different peers on different ranks
    size_t b, en;
    slice_range(n4, world, r, &b, &en);
    for (size_t i = b + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < en; i += size_t(gridDim.x) * blockDim.x)
      *reinterpret_cast<float4*>(ps.buf[rank] + off + 4 * i) = ld_peer_f4(ps.buf[r] + off + 4 * i);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) {
      *ticket = 0u;
      *epoch_dev = (int)e;
    }
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(ps.flags[threadIdx.x] + fl + kFlagDone + rank, e);
    }
  }
}

// wait until every rank has finished exchange `*epoch_dev` of flag set `lane_id` (its DONE flags): this rank's part of the
// buffer may be overwritten after it
__global__ void dp_wait_done_kernel(const unsigned* __restrict__ flags_local, int world, const int* __restrict__ epoch_dev,
                                    int lane_id, const Watchdog dog) {
  if ((int)threadIdx.x < world)
    wait_flag_ge(flags_local + lane_id * kLaneStride + kFlagDone + threadIdx.x, (unsigned)*epoch_dev, dog,
                 lane_id * kLaneStride + kFlagDone + threadIdx.x);
}

}  // namespace csn

using namespace csn;

extern "C" int csn_dp_wait_done_zero(const void* flags_local, int world, const int* step_counter, float* zero_ptr,
                                     size_t n, void* stream) {
  CSN_REQUIRE(flags_local && step_counter, "csn_dp_wait_done_zero: null pointer");
  CSN_REQUIRE(world >= 1 && world <= kMaxWorld, "csn_dp_wait_done_zero: world must be in [1, %d]", kMaxWorld);
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>(ceil_div<size_t>(n, 256), size_t(sm_count())));
  dp_wait_done_zero_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned*>(flags_local), world,
                                                                  step_counter, zero_ptr, zero_ptr ? n : 0, make_watchdog(-1));
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_dp_adam_step_peer(float* params, float* exp_avg, float* exp_avg_sq, size_t n_param,
                                     const void* const* grad_ptrs, void* const* flag_ptrs, int world, int rank,
                                     float* center, size_t K, float center_momentum, float center_scale,
                                     int* step_counter, unsigned* ticket, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, int decoupled, float grad_scale, void* stream) {
  CSN_REQUIRE(params && exp_avg && exp_avg_sq && grad_ptrs && flag_ptrs && step_counter && ticket,
              "csn_dp_adam_step_peer: null pointer");
  CSN_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
              "csn_dp_adam_step_peer: need 1 <= world <= %d and 0 <= rank < world", kMaxWorld);
  CSN_REQUIRE(n_param % 4 == 0, "csn_dp_adam_step_peer: parameter count must be a multiple of 4 (flat buffers are padded)");
  CSN_REQUIRE(K == 0 || center, "csn_dp_adam_step_peer: centre pointer missing");
  PeerSet ps{};
  for (int r = 0; r < world; ++r) {
    CSN_REQUIRE(grad_ptrs[r] && flag_ptrs[r], "csn_dp_adam_step_peer: null peer pointer for rank %d", r);
    CSN_REQUIRE((reinterpret_cast<uintptr_t>(grad_ptrs[r]) & 15) == 0, "csn_dp_adam_step_peer: peer buffer %d not 16-byte aligned", r);
    ps.grad[r] = reinterpret_cast<const float*>(grad_ptrs[r]);
    ps.flags[r] = reinterpret_cast<unsigned*>(flag_ptrs[r]);
  }
  CSN_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(exp_avg) |
                reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "csn_dp_adam_step_peer: buffers must be 16-byte aligned");
  // every CTA spins on the READY flags: the whole grid must be resident (<= 8 CTAs of 256 threads per SM)
  const size_t want = std::max<size_t>(1, ceil_div<size_t>(std::max(n_param / 4, K), 256));
  const unsigned blocks = (unsigned)std::min<size_t>(want, size_t(sm_count()) * 4);
  dp_adam_peer_kernel<<<blocks, 256, 0, as_stream(stream)>>>(params, exp_avg, exp_avg_sq, n_param, ps, world, rank, center, K,
                                                             center_momentum, center_scale, step_counter, lr, beta1, beta2, eps,
                                                             weight_decay, decoupled, grad_scale, ticket, make_watchdog(rank));
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_dp_last_timeout(int* out4) {
  CSN_REQUIRE(out4, "csn_dp_last_timeout: null pointer");
  const Watchdog wd = make_watchdog(-1);
  for (int i = 0; i < 4; ++i) out4[i] = wd.host_slot ? (int)reinterpret_cast<volatile unsigned*>(wd.host_slot)[i] : 0;
  return CSN_OK;
}

extern "C" int csn_dp_allreduce_twoshot(void* const* buf_ptrs, void* const* flag_ptrs, int world, int rank, size_t offset,
                                        size_t n, int* epoch_counter, int flag_set, unsigned* ticket, void* stream) {
  CSN_REQUIRE(buf_ptrs && flag_ptrs && epoch_counter && ticket, "csn_dp_allreduce_twoshot: null pointer");
  CSN_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
              "csn_dp_allreduce_twoshot: need 1 <= world <= %d and 0 <= rank < world", kMaxWorld);
  CSN_REQUIRE(offset % 4 == 0 && n % 4 == 0, "csn_dp_allreduce_twoshot: offset and length must be multiples of 4 floats");
  CSN_REQUIRE(flag_set >= 0 && flag_set < 2, "csn_dp_allreduce_twoshot: flag_set must be 0 or 1 (flag block of >= 64 words)");
  PeerBufs ps{};
  for (int r = 0; r < world; ++r) {
    CSN_REQUIRE(buf_ptrs[r] && flag_ptrs[r], "csn_dp_allreduce_twoshot: null peer pointer for rank %d", r);
    CSN_REQUIRE((reinterpret_cast<uintptr_t>(buf_ptrs[r]) & 15) == 0, "csn_dp_allreduce_twoshot: peer buffer %d not 16-byte aligned", r);
    ps.buf[r] = reinterpret_cast<float*>(buf_ptrs[r]);
    ps.flags[r] = reinterpret_cast<unsigned*>(flag_ptrs[r]);
  }
  if (n == 0) return CSN_OK;
  cudaStream_t s = as_stream(stream);
  const size_t n4 = n / 4;
  const Watchdog dog = make_watchdog(rank);
  // enough loads in flight to fill the NVLink pipe (latency ~2 us, 770 GB/s: ~1.5 MB) without taking every SM
  const size_t per = ceil_div<size_t>(n4, world);
  // CTAs of the two kernels.  The exchange usually runs BESIDE a latency-bound backward recurrence: more CTAs finish it
  // sooner but take L2 / NVLink bandwidth and issue slots from the chain.  CSN_DP_XCHG_BLOCKS overrides the default.
  static const int forced_blocks = [] { const char* e = getenv("CSN_DP_XCHG_BLOCKS"); return e ? atoi(e) : 0; }();
  const size_t cap = forced_blocks > 0 ? size_t(forced_blocks) : size_t(sm_count()) * 2;
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>(ceil_div<size_t>(per, 256), cap));
  dp_reduce_slice_kernel<<<blocks, 256, 0, s>>>(ps, world, rank, offset, n4, epoch_counter, flag_set, ticket, dog);
  CSN_LAUNCH_CHECK();
  dp_gather_slices_kernel<<<blocks, 256, 0, s>>>(ps, world, rank, offset, n4, epoch_counter, flag_set, ticket + 1, dog);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_dp_wait_done(const void* flags_local, int world, const int* epoch_counter, int flag_set, void* stream) {
  CSN_REQUIRE(flags_local && epoch_counter, "csn_dp_wait_done: null pointer");
  CSN_REQUIRE(world >= 1 && world <= kMaxWorld && flag_set >= 0 && flag_set < 2, "csn_dp_wait_done: bad arguments");
  if (world == 1) return CSN_OK;
  dp_wait_done_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned*>(flags_local), world, epoch_counter, flag_set,
                                                       make_watchdog(-1));
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
