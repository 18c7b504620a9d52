// Side roles of the persistent LSTM launches: the CTAs a recurrence launch does NOT need (a recurrence CTA owns a whole
// SM's tensor memory and the batch gives ceil(B / NV) of them: 128 of 148 SMs at the cfg2 batch of 256) do the layer's
// two big GEMM-shaped jobs WHILE the serial chain runs, instead of before / after it:
//   forward : "Xp servers" compute the input projection  Xp[t, b, :] = x[t, b, :] W_ih^T + b_ih  tile by tile in time
//             order on tcgen05 (W_ih resident in shared memory, TMA-fed x tiles, double-buffered TMEM accumulators) and
//             publish one READY flag per 128-row tile; the recurrence CTAs' producer warp waits on the flag of the
//             timestep it prefetches (seven steps ahead of the chain) and pulls its rows through the TMA ring.
//             The chain itself then carries no projection MMAs (they cost it ~300 of 990 cycles per step when the same
//             CTA issued them between the recurrent MMAs).
//   backward: "dW consumers" accumulate  dW[4H, I + H] = sum_{t, b} dG[t, b, :]^T [x[t, b, :] | h[t-1, b, :]]  in tensor
//             memory over the whole reverse sweep, 64 (t, b) rows at a time, as soon as every recurrence CTA has
//             published that timestep's dG (a counter per timestep, released after the CTA's TMA store of its dG rows
//             has completed).  dG is consumed out of L2 right behind its producers; the two weight-gradient GEMMs that
//             used to follow the recurrence (2 x 42 us, each re-reading the 115 MB dG from HBM) are gone.
// Replaces the hoisted x W_ih^T of nn.LSTM's forward and the weight-gradient sums of its autograd backward
// (LstmDistillFromDinoV2Train.py:365,374).  Cross-CTA ordering is release / acquire at gpu scope on flags in global
// memory; all CTAs of the launch are co-resident (grid <= SM count, one CTA per SM; launched cooperatively), and every
// spin carries a watchdog so that a scheduling surprise traps instead of hanging the device.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "tc.cuh"

namespace csn {
namespace side {

using namespace tc;

__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy acquire -> async-proxy (TMA) reads of the data it guards
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// spin until *p >= target.  The writers are CTAs of the same (co-resident) launch, so the wait is short; the watchdog
// (20 s) turns a scheduling surprise into a trap instead of a hung device.
__device__ __forceinline__ void wait_ge(const unsigned* p, unsigned target) {
  if (ld_acquire_gpu(p) >= target) return;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(64);
    if ((++spins & 4095u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}

// TMA store: shared -> global bulk copy, tracked in bulk async-groups of the issuing thread
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() {  // all but the N most recent groups have COMPLETED (writes performed)
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // ... have finished READING their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

constexpr int kEpiTileStride = 36;  // floats per staged row: 16-byte aligned, conflict-free float4 reads

// One 32-column chunk of a [128-lane x N] fp32 accumulator -> global rows (row-contiguous 128-byte stores): the warp
// drains its 32 lanes through a private padded shared-memory tile, like the GEMM epilogue of gemm_tc.cu.
// dst: pointer to (row 0 of this warp's 32 rows, column 0 of the chunk); ld: row pitch in floats; rows_ok: rows of this
// warp that exist; bias: this chunk's 32 bias values or NULL.
__device__ __forceinline__ void drain_chunk(uint32_t taddr, float* tile, int lane, float* dst, size_t ld, int rows_ok,
                                            const float* bias) {
  uint32_t r[32];
  tmem_ld<32>(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4)
    *reinterpret_cast<uint4*>(tile + lane * kEpiTileStride + q4 * 4) = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
  __syncwarp();
  const int col4 = (lane & 7) * 4;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float4*>(bias + col4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    float4 v = *reinterpret_cast<const float4*>(tile + rr * kEpiTileStride + col4);
    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    if (rr < rows_ok) *reinterpret_cast<float4*>(dst + size_t(rr) * ld + col4) = v;
  }
  __syncwarp();
}

// 64 columns of a [128-lane x N] fp32 accumulator -> fp16 global rows, 128 contiguous bytes per row and store: the same
// staging, half the bytes (an SM writes global memory at 32 B/clk at most, scripts/store_bw.py: the output format decides
// what the Xp servers can deliver).  dst: (row 0 of this warp's 32 rows, column 0 of the chunk); ld in halves.
__device__ __forceinline__ void drain_chunk_half(uint32_t taddr, uint8_t* tile, int lane, __half* dst, size_t ld, int rows_ok) {
  uint32_t r0[32], r1[32];
  tmem_ld<32>(taddr, r0);
  tmem_ld<32>(taddr + 32, r1);
  tmem_ld_wait();
  uint32_t h[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const __half2 a = __floats2half2_rn(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
    const __half2 b = __floats2half2_rn(__uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1]));
    h[i] = *reinterpret_cast<const uint32_t*>(&a);
    h[16 + i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  constexpr int kRow = kEpiTileStride * 4;  // 144 B per staged row: 128 B of data + 16 B bank rotation
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4)
    *reinterpret_cast<uint4*>(tile + lane * kRow + q4 * 16) = make_uint4(h[q4 * 4], h[q4 * 4 + 1], h[q4 * 4 + 2], h[q4 * 4 + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    const uint4 v = *reinterpret_cast<const uint4*>(tile + rr * kRow + (lane & 7) * 16);
    if (rr < rows_ok) *reinterpret_cast<uint4*>(dst + size_t(rr) * ld + (lane & 7) * 8) = v;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ forward: Xp servers
struct XpServe {
  int n_srv;        // server CTAs = the first n_srv blocks of the launch (0: Xp is precomputed, nobody waits)
  int n_tiles;      // ceil(rows / 128)
  int rows;         // T * B
  int I, H;
  __half* xp;       // [rows, 4H] fp16, gate-major columns g*H + u, WITHOUT the bias (the recurrence adds b_ih + b_hh)
  unsigned* flags;  // [n_tiles], zeroed before the launch; 1 = the tile's rows of Xp are in memory
};

// rows per TMA box of the W_ih image: the largest of 256 / 128 / 64 that divides 4H (H % 16 == 0), or 4H itself when it
// fits one box -- a box never reaches past the last gate row, so the shared-memory image is exactly 4H rows
__host__ __device__ inline int xp_w_box_rows(int H) {
  const int n = 4 * H;
  if (n <= 256) return n;
  return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64);
}
constexpr int kSrvStages = 3;                    // x ring: 128 rows x 64 columns (one K chunk of one tile) per stage
constexpr uint32_t kSrvStageBytes = 128 * 128;   // 128 rows x 128 B
__host__ __device__ inline size_t xp_server_smem(int H) {
  return 1024 /* alignment slack */ + size_t(2) * 4 * H * 128 /* W_ih, two K chunks */ + size_t(kSrvStages) * kSrvStageBytes +
         size_t(8) * 32 * kEpiTileStride * 4 + 256;
}

// blockDim >= 320: warps 0-7 epilogue (TMEM lane quadrant w % 4, 32-column chunks of parity w / 4), warp 8 MMA issuer
// (+ tensor-memory owner), warp 9 TMA producer.  Requires I <= 128, I % 8 == 0, H <= 128, H % 32 == 0.
__device__ __forceinline__ void xp_server_role(uint8_t* smem_raw, const CUtensorMap* tm_x, const CUtensorMap* tm_w,
                                               const XpServe& sv, int srv) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = sv.H, I = sv.I;
  const int nkc = (I + 63) / 64;                       // K chunks of 64 input features
  const int nks = (I + 15) / 16;                       // K steps of 16
  const uint32_t w_chunk_bytes = uint32_t(4 * H) * 128u;
  uint8_t* ws = base;                                  // [2][4H rows][128 B]  (K-major, 128B swizzle)
  uint8_t* xs = ws + 2 * w_chunk_bytes;                // [kSrvStages][128 rows][128 B]
  float* epi = reinterpret_cast<float*>(xs + kSrvStages * kSrvStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(epi) + 8 * 32 * kEpiTileStride * 4);
  uint64_t* x_full = bars;                 // [kSrvStages]
  uint64_t* x_empty = bars + kSrvStages;   // [kSrvStages]
  uint64_t* acc_full = x_empty + kSrvStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kSrvStages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 9 && lane == 0) { prefetch_tmap(tm_x); prefetch_tmap(tm_w); }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int N2 = 2 * H;  // accumulator width: two gates

  if (warp == 9) {
    if (lane == 0) {
      // W_ih once: per K chunk, 4H rows in boxes of (up to) 256 rows
      mbar_arrive_expect_tx(w_full, uint32_t(nkc) * w_chunk_bytes);
      const int box_rows = xp_w_box_rows(H);
      for (int kc = 0; kc < nkc; ++kc)
        for (int r0 = 0; r0 < 4 * H; r0 += box_rows)
          tma_load_2d(ws + size_t(kc) * w_chunk_bytes + size_t(r0) * 128, tm_w, w_full, kc * 64, r0);
      uint32_t xi = 0;
      for (int tile = srv; tile < sv.n_tiles; tile += sv.n_srv) {
        for (int kc = 0; kc < nkc; ++kc, ++xi) {
          const uint32_t s = xi % kSrvStages;
          if (xi >= kSrvStages) mbar_wait(&x_empty[s], ((xi / kSrvStages) - 1) & 1);
          mbar_arrive_expect_tx(&x_full[s], kSrvStageBytes);
          tma_load_2d(xs + size_t(s) * kSrvStageBytes, tm_x, &x_full[s], kc * 64, tile * 128);
        }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N2, 0, 0);
      mbar_wait(w_full, 0);
      uint32_t xi = 0, ti = 0;
      for (int tile = srv; tile < sv.n_tiles; tile += sv.n_srv, ++ti) {
        for (int half = 0; half < 2; ++half) {
          if (ti >= 1) mbar_wait(&acc_empty[half], (ti - 1) & 1);  // the epilogue has drained this buffer's previous tile
          tcgen05_fence_after();
          for (int ks = 0; ks < nks; ++ks) {
            const int kc = ks >> 2, kk = ks & 3;
            const uint32_t s = (xi + kc) % kSrvStages;
            if (half == 0 && kk == 0) {
              mbar_wait(&x_full[s], ((xi + kc) / kSrvStages) & 1);
              tcgen05_fence_after();
            }
            // K-major SW128: rows of 128 B, 8-row groups 1024 B apart, 16 K-elements = 32 B along the row
            const uint64_t da = make_smem_desc(smem_u32(xs + size_t(s) * kSrvStageBytes) + kk * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(smem_u32(ws + size_t(kc) * w_chunk_bytes + size_t(half) * N2 * 128) + kk * 32, 16,
                                               1024, kLayoutSw128);
            umma_f16(tmem_base + half * 256, da, db, idesc, ks != 0);
          }
          umma_commit(&acc_full[half]);
        }
        for (int kc = 0; kc < nkc; ++kc) umma_commit(&x_empty[(xi + kc) % kSrvStages]);
        xi += nkc;
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, par = warp >> 2;
    float* tile_s = epi + warp * (32 * kEpiTileStride);
    const int nch = N2 / 64;
    uint32_t ti = 0;
    for (int tile = srv; tile < sv.n_tiles; tile += sv.n_srv, ++ti) {
      const int row0 = tile * 128 + q * 32;
      const int rows_ok = min(32, sv.rows - row0);
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&acc_full[half], ti & 1);
        tcgen05_fence_after();
        for (int c = par; c < nch; c += 2) {
          const int n0 = half * N2 + c * 64;
          drain_chunk_half(tmem_base + (uint32_t(q * 32) << 16) + half * 256 + c * 64, reinterpret_cast<uint8_t*>(tile_s), lane,
                           sv.xp + size_t(row0) * 4 * H + n0, size_t(4) * H, rows_ok);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[half]);
      }
      // every epilogue warp's rows of this tile are written: publish it (barrier, then ONE release store)
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (tid == 0) {
        __threadfence();
        st_release_gpu(sv.flags + tile, 1u);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dW consumers
struct DwConsume {
  int n_cons;        // consumer CTAs = the LAST n_cons blocks of the launch (0: none)
  int n_rec;         // recurrence CTAs (each adds 1 to done[t] per timestep)
  int m_halves;      // 1 (4H <= 256) or 2: a consumer owns 256 of the 4H gate rows
  int T, B, I, H;
  int i_pad;         // ceil(I / 64) * 64: column of the first h_{t-1} feature in the [x | h] operand
  int c_lo;          // first 64-row chunk the consumers take; rows below c_lo * 64 (the LAST timesteps BPTT reaches) are
                     // left to a GEMM after the launch, so that the consumers finish together with the chain
  unsigned* done;    // [T], zeroed before the launch
  float* slabs;      // [n_groups][512 rows (interleaved gate index 4u + g)][256 columns] fp32 partial dW
};
constexpr int kConsStages = 3;
constexpr uint32_t kConsABytes = 256 * 64 * 2, kConsBBytes = 256 * 64 * 2;  // per 64-row chunk: dG^T half, [x | h]
constexpr uint32_t kConsStageBytes = kConsABytes + kConsBBytes;
__host__ __device__ inline size_t dw_consumer_smem() {  // (the final drain stages through the idle operand ring)
  return 1024 + size_t(kConsStages) * kConsStageBytes + 256;
}
__host__ __device__ inline size_t dw_slab_floats() { return size_t(512) * 256; }

// blockDim >= 320: warp 8 MMA issuer (+ tensor-memory owner); warps 0-7 are the TMA producers -- ONE 64 x 64 box of the
// stage each (a TMA instruction blocks its issuing warp for a few hundred cycles: eight boxes from one thread cost more
// than the stage's MMAs) -- and drain the accumulators at the end.
// cons = index of this consumer in [0, n_cons): M half = cons % m_halves, K group = cons / m_halves.
// The contraction runs over the rows (t, b) of dG in 64-row chunks, DESCENDING (the order BPTT produces them in);
// group g takes every n_groups-th chunk.
__device__ __forceinline__ void dw_consumer_role(uint8_t* smem_raw, const CUtensorMap* tm_dg, const CUtensorMap* tm_x,
                                                 const CUtensorMap* tm_h, const DwConsume& dc, int cons) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = base;
  float* epi = reinterpret_cast<float*>(stages);  // the accumulators are drained after the last MMA has read the ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + size_t(kConsStages) * kConsStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kConsStages;
  uint64_t* acc_full = empty + kConsStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mh = cons % dc.m_halves, grp = cons / dc.m_halves, n_groups = dc.n_cons / dc.m_halves;
  const int rows = dc.T * dc.B;
  const int n_chunks = (rows + 63) / 64;
  // chunks of this group: q = grp, grp + n_groups, ... in descending-row order; chunk index c = n_chunks - 1 - q >= c_lo
  const int n_mine = n_chunks - dc.c_lo;
  const int my_chunks = n_mine > grp ? (n_mine - 1 - grp) / n_groups + 1 : 0;

  if (tid == 0) {
    for (int i = 0; i < kConsStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(tm_dg); prefetch_tmap(tm_x); prefetch_tmap(tm_h); }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    if (lane == 0) {
      // box `warp` of every stage.  A = dG^T (MN-major): boxes 0-3 = four 64 (gate rows) x 64 (t,b rows) tiles; B = [x | h_{t-1}]
      // (MN-major): boxes 4-7 = 64-column tiles, x first, then h shifted back by one timestep (rows of t = 0 read out of
      // bounds: zero fill = h_{-1} = 0).  Box 0's thread also posts the stage's byte count.
      const int j = warp & 3;
      // Every recurrence CTA publishes its steps in descending t, so done[t] == n_rec implies the same for every later
      // timestep: ONE flag per chunk (its earliest timestep), and none while the chunk lies above the lowest timestep
      // already seen complete.  A consumer that is behind looks eight timesteps further down in the same round trip.
      int known_t = dc.T;  // timesteps >= known_t are complete
      for (int i = 0; i < my_chunks; ++i) {
        const int c = n_chunks - 1 - (grp + i * n_groups);
        const int r0 = c * 64;
        const int t_lo = r0 / dc.B;
        if (t_lo < known_t) {
          const int t_far = max(0, t_lo - 8);
          const bool far_ok = ld_acquire_gpu(dc.done + t_far) >= (unsigned)dc.n_rec;
          if (!far_ok) wait_ge(dc.done + t_lo, (unsigned)dc.n_rec);
          known_t = far_ok ? t_far : t_lo;
          fence_proxy_async_all();
        }
        const uint32_t s = uint32_t(i) % kConsStages;
        if (i >= kConsStages) mbar_wait(&empty[s], ((uint32_t(i) / kConsStages) - 1) & 1);
        uint8_t* sa = stages + size_t(s) * kConsStageBytes;
        uint8_t* sb = sa + kConsABytes;
        if (warp == 0) mbar_arrive_expect_tx(&full[s], kConsStageBytes);
        if (warp < 4) {
          tma_load_2d(sa + j * 8192, tm_dg, &full[s], mh * 256 + j * 64, r0);
        } else {
          const int col = j * 64;
          if (col < dc.i_pad) tma_load_2d(sb + j * 8192, tm_x, &full[s], col, r0);
          else tma_load_2d(sb + j * 8192, tm_h, &full[s], col - dc.i_pad, r0 - dc.B);
        }
      }
    }
    __syncwarp();
  }
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 256, 1, 1);
      for (int i = 0; i < my_chunks; ++i) {
        const uint32_t s = uint32_t(i) % kConsStages;
        mbar_wait(&full[s], (uint32_t(i) / kConsStages) & 1);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(stages + size_t(s) * kConsStageBytes), sb = sa + kConsABytes;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // MN-major SW128: 8 k-rows x 128 B atoms 1024 B apart along K (SBO), 64-element MN chunks 8192 B apart (LBO)
            const uint64_t da = make_smem_desc(sa + mt * 16384 + kk * 2048, 8192, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(sb + kk * 2048, 8192, 1024, kLayoutSw128);
            umma_f16(tmem_base + mt * 256, da, db, idesc, (i | kk) != 0);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  }
  if (warp < 8) {
    // partial dW of this (group, M half): lanes = gate rows, 256 columns -> slab rows mh*256 + mt*128 + lane
    const int q = warp & 3, mt = warp >> 2;
    float* tile_s = epi + warp * (32 * kEpiTileStride);
    float* dst = dc.slabs + size_t(grp) * dw_slab_floats() + size_t(mh * 256 + mt * 128 + q * 32) * 256;
    if (my_chunks > 0) {
      mbar_wait(acc_full, 0);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < 8; ++c)
        drain_chunk(tmem_base + (uint32_t(q * 32) << 16) + mt * 256 + c * 32, tile_s, lane, dst + c * 32, 256, 32, nullptr);
    } else {
      for (int e = lane; e < 32 * 256 / 4; e += 32) reinterpret_cast<float4*>(dst)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace side
}  // namespace csn
