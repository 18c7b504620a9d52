// Band-pass along time as a cascade of DF2T biquads (replaces scipy sosfilt / filtfilt as called at
// utils/Utilities.py:421-427 with the designs of utils/EEGFilters.py:26-39).
//
// Layout: x is [B, C, T] fp32, i.e. N = B*C independent series of T contiguous samples.  One thread owns one
// series (the recurrence is sequential in t); a block owns R consecutive series.  Global <-> shared traffic is
// tile-wise and coalesced:
//   * causal kernel: streams time in chunks of 32 samples, [R x 32] tiles double-buffered with cp.async (16 B),
//     row stride 36 floats so the per-thread float4 row reads are bank-conflict free; results leave through a
//     second tile written time-major when the output is the encoder's [T, B, C] layout (fused transpose + cast).
//   * zero-phase kernel (sosfiltfilt semantics): whole series resident in shared memory, forward then backward
//     sweep in place, odd extension of 3*ntaps samples generated on the fly, steady-state initial conditions.
// HBM-bound by design: 8 bytes/sample of traffic against 5*n_sections FMA/sample.
#include "common.cuh"

namespace csn {

constexpr int kMaxSec = 8;
constexpr int kTC = 32;  // samples per streamed chunk
constexpr int kRS = 36;  // shared row stride (floats): 144 B = 16 B * 9 -> conflict-free float4 rows
constexpr int kStages = 4;  // cp.async ring depth: three 32-sample chunks in flight per block hide the HBM latency

struct SosCoef {
  float b0[kMaxSec], b1[kMaxSec], b2[kMaxSec], a1[kMaxSec], a2[kMaxSec], zi1[kMaxSec], zi2[kMaxSec];
  int padlen;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NSEC>
__device__ __forceinline__ float biquad_cascade(float v, float (&s1)[NSEC], float (&s2)[NSEC], const SosCoef& c) {
#pragma unroll
  for (int k = 0; k < NSEC; ++k) {
    float y = fmaf(c.b0[k], v, s1[k]);
    s1[k] = fmaf(-c.a1[k], y, fmaf(c.b1[k], v, s2[k]));
    s2[k] = fmaf(-c.a2[k], y, c.b2[k] * v);
    v = y;
  }
  return v;
}

// One 32-sample chunk through the cascade with the sections SKEWED in time: in iteration j section k works on sample
// j - k, fed by what section k-1 produced one iteration earlier, so the NSEC section updates of an iteration are
// independent of each other (instruction-level parallelism NSEC instead of one dependent chain per sample; a warp
// issues in order and only ~1.7 warps share a scheduler at cfg2).  Same operations on the same operands as
// biquad_cascade sample by sample: results are bit-identical.  `emit(idx, y)` receives the outputs in order.
// UNIT_B2: sections 1.. have b2 == 1.0f exactly (scipy's sos form keeps the gain in section 0), so their b2 * v is v --
// one multiply per section and sample less (17 instead of 20 arithmetic instructions per sample at 4 sections in a
// kernel that is instruction bound), same bits.
template <int NSEC, bool UNIT_B2, typename Emit>
__device__ __forceinline__ void cascade_chunk_skewed(const float (&x)[kTC], float (&s1)[NSEC], float (&s2)[NSEC],
                                                     const SosCoef& c, Emit emit) {
  float p[NSEC];
#pragma unroll
  for (int j = 0; j < kTC + NSEC - 1; ++j) {
#pragma unroll
    for (int k = NSEC - 1; k >= 0; --k) {
      const int idx = j - k;
      if (idx >= 0 && idx < kTC) {
        const float v = (k == 0) ? x[idx] : p[k - 1];
        const float y = fmaf(c.b0[k], v, s1[k]);
        s1[k] = fmaf(-c.a1[k], y, fmaf(c.b1[k], v, s2[k]));
        s2[k] = fmaf(-c.a2[k], y, (UNIT_B2 && k > 0) ? v : c.b2[k] * v);
        p[k] = y;
        if (k == NSEC - 1) emit(idx, y);
      }
    }
  }
}

__device__ __forceinline__ size_t out_index(int layout, long long r, int t, int Bn, int C, int T) {
  // r = b*C + c
  if (layout == CSN_LAYOUT_BCT) return size_t(r) * T + t;
  if (layout == CSN_LAYOUT_TBC) return size_t(t) * (size_t(Bn) * C) + r;
  long long b = r / C, cc = r - b * C;  // BTC
  return (size_t(b) * T + t) * C + cc;
}

template <typename OutT>
__device__ __forceinline__ void store_out(OutT* y, size_t i, float v) {
  if constexpr (sizeof(OutT) == 4) y[i] = v; else y[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------ causal
template <int NSEC, typename OutT>
__global__ void __launch_bounds__(128) sosfilt_stream_kernel(const float* __restrict__ x, OutT* __restrict__ y,
                                                             const SosCoef coef, long long n_series, int T, int Bn,
                                                             int C, int layout, int vec_ok) {
  extern __shared__ __align__(16) float smem[];
  const int R = blockDim.x;
  float* in_tile = smem;                        // [kStages][R][kRS]
  float* out_tile = smem + kStages * R * kRS;   // [kTC][R]   (transposed layouts only)
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * R;
  const int n_chunks = (T + kTC - 1) / kTC;
  const bool transposed = (layout != CSN_LAYOUT_BCT);

  float s1[NSEC], s2[NSEC];
#pragma unroll
  for (int k = 0; k < NSEC; ++k) s1[k] = s2[k] = 0.f;

  auto issue_load = [&](int chunk, int buf) {
    const int t0 = chunk * kTC;
    float* dst = in_tile + buf * R * kRS;
    if (vec_ok) {
#pragma unroll
      for (int j = 0; j < kTC / 4; ++j) {
        int v = tid + j * R;
        int row = v >> 3, c4 = v & 7;
        long long r = r0 + row;
        int t = t0 + c4 * 4;
        bool ok = (r < n_series) && (t < T);
        const float* src = ok ? (x + size_t(r) * T + t) : x;
        cp_async16(dst + row * kRS + c4 * 4, src, ok);
      }
    } else {
      for (int j = 0; j < kTC; ++j) {
        int e = tid + j * R;
        int row = e >> 5, cc = e & 31;
        long long r = r0 + row;
        int t = t0 + cc;
        dst[row * kRS + cc] = (r < n_series && t < T) ? x[size_t(r) * T + t] : 0.f;
      }
    }
    cp_async_commit();
  };

  // prologue: kStages-1 chunks in flight (one cp.async group per chunk, empty groups past the end keep the count uniform)
  for (int c = 0; c < kStages - 1; ++c) {
    if (c < n_chunks) issue_load(c, c); else cp_async_commit();
  }
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int buf = chunk % kStages;
    cp_async_wait<kStages - 2>();
    __syncthreads();  // tile `buf` visible to all; everyone is done with the tile of chunk-1 and with out_tile
    if (chunk + kStages - 1 < n_chunks) issue_load(chunk + kStages - 1, (chunk + kStages - 1) % kStages);
    else cp_async_commit();

    float* row = in_tile + buf * R * kRS + tid * kRS;
#pragma unroll
    for (int q = 0; q < kTC / 4; ++q) {
      float4 v = *reinterpret_cast<float4*>(row + q * 4);
      v.x = biquad_cascade<NSEC>(v.x, s1, s2, coef);
      v.y = biquad_cascade<NSEC>(v.y, s1, s2, coef);
      v.z = biquad_cascade<NSEC>(v.z, s1, s2, coef);
      v.w = biquad_cascade<NSEC>(v.w, s1, s2, coef);
      if (transposed) {
        out_tile[(q * 4 + 0) * R + tid] = v.x;
        out_tile[(q * 4 + 1) * R + tid] = v.y;
        out_tile[(q * 4 + 2) * R + tid] = v.z;
        out_tile[(q * 4 + 3) * R + tid] = v.w;
      } else {
        *reinterpret_cast<float4*>(row + q * 4) = v;
      }
    }
    __syncthreads();

    const int t0 = chunk * kTC;
    if (transposed) {
      // out_tile[t][s]: consecutive series -> consecutive addresses in [T, B*C]
      const long long N = (long long)Bn * C;
      if (layout == CSN_LAYOUT_TBC && (N & 7) == 0 && r0 + R <= n_series && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        // vectorised: each thread moves 8 consecutive series of one time sample (16 B of bf16 / 2 x 16 B of fp32)
        const int groups = R >> 3;  // 8-series groups per sample
        for (int e = tid; e < kTC * groups; e += R) {
          const int j = e / groups, g8 = e - j * groups;
          const int t = t0 + j;
          if (t >= T) break;
          const float4 a = *reinterpret_cast<const float4*>(out_tile + j * R + g8 * 8);
          const float4 b = *reinterpret_cast<const float4*>(out_tile + j * R + g8 * 8 + 4);
          OutT* dst = y + size_t(t) * N + r0 + g8 * 8;
          if constexpr (sizeof(OutT) == 4) {
            *reinterpret_cast<float4*>(dst) = a;
            *reinterpret_cast<float4*>(dst + 4) = b;
          } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst) = o;
          }
        }
      } else {
        for (int j = 0; j < kTC; ++j) {
          int t = t0 + j;
          long long r = r0 + tid;
          if (t < T && r < n_series) store_out<OutT>(y, out_index(layout, r, t, Bn, C, T), out_tile[j * R + tid]);
        }
      }
    } else if (vec_ok && sizeof(OutT) == 4) {
      const float* src = in_tile + buf * R * kRS;
#pragma unroll
      for (int j = 0; j < kTC / 4; ++j) {
        int v = tid + j * R;
        int rr = v >> 3, c4 = v & 7;
        long long r = r0 + rr;
        int t = t0 + c4 * 4;
        if (r < n_series && t < T)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + size_t(r) * T + t) =
              *reinterpret_cast<const float4*>(src + rr * kRS + c4 * 4);
      }
    } else {
      const float* src = in_tile + buf * R * kRS;
      for (int j = 0; j < kTC; ++j) {
        int e = tid + j * R;
        int rr = e >> 5, cc = e & 31;
        long long r = r0 + rr;
        int t = t0 + cc;
        if (r < n_series && t < T) store_out<OutT>(y, size_t(r) * T + t, src[rr * kRS + cc]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------- causal, lean fast path
// ncu on the general kernel above at cfg2 (profiles/ncu_full_r01d.md): 15.1 M warp instructions, 56 % issue-slot
// utilisation, no dominant stall -- the kernel is INSTRUCTION bound, and only 20 of its 33 instructions per sample
// are the biquad FMAs; the rest is generic index / predicate arithmetic around the cp.async issue (250 instructions
// per 32-sample chunk) and the tile store (200).  This variant serves the common case (16-byte aligned rows, full
// blocks of 32 series, time-major bf16/fp32 or channel-major fp32 output) with running pointers, a predicate-free
// body for full chunks and one warp per block (warp-level barriers only): ~24 instructions per sample.
// Optional fused GATHER (GPU-resident dataset): with `g.idx` set, batch row b of the input is trial
// g.idx[b] of a [n_trials, C, row_stride] array, samples [t_off, t_off + T), optionally (x - mean) * inv_std -- the crop
// / normalisation of EEGDataset.__getitem__ (utils/PerilsEEGDataset.py:541-573) folded into the filter's loads, so the
// gathered batch is never written to HBM.
struct GatherSrc {
  const long long* idx;  // nullptr: x is the dense [B, C, T] batch
  int n_trials, C, row_stride, t_off;
  float mean, inv_std;   // applied when norm != 0
  int norm;
};

template <int NSEC, typename OutT, bool TIME_MAJOR, bool UNIT_B2, bool GATHER>
__global__ void __launch_bounds__(32) sosfilt_warp_kernel(const float* __restrict__ x, OutT* __restrict__ y,
                                                         const SosCoef coef, int T, long long N, const GatherSrc g) {
  constexpr int R = 32;
  extern __shared__ __align__(16) float smem[];
  float* in_tile = smem;                        // [kStages][R][kRS]
  float* out_tile = smem + kStages * R * kRS;   // [kTC][R]   (time-major output only)
  const int lane = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * R;
  const int n_chunks = (T + kTC - 1) / kTC;

  float s1[NSEC], s2[NSEC];
#pragma unroll
  for (int k = 0; k < NSEC; ++k) s1[k] = s2[k] = 0.f;

  // cp.async of a chunk: lane -> (row lane/8 + 4 j, 16-byte column lane%8), j = 0..7
  const int row8 = lane >> 3, c4 = lane & 7;
  const float* src = x + size_t(r0 + row8) * T + c4 * 4;   // advanced by kTC floats per issued chunk
  size_t src_jstride = size_t(4) * T;
  if constexpr (GATHER) {  // the block's 32 series are channels [c0, c0 + 32) of ONE batch row (C % 32 == 0)
    const long long bi = r0 / g.C;
    long long trial = g.idx[bi];
    if (trial < 0) trial += g.n_trials;
    const int c0 = (int)(r0 - bi * g.C);
    src = x + (size_t(trial) * g.C + c0 + row8) * g.row_stride + g.t_off + c4 * 4;
    src_jstride = size_t(4) * g.row_stride;
  }
  const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(in_tile + row8 * kRS + c4 * 4);
  int t_issue = c4 * 4;                                     // first sample this lane fetches in the next chunk
  auto issue_load = [&](int buf) {
    const uint32_t d = dst0 + uint32_t(buf) * (R * kRS * 4);
    const int bytes = (t_issue < T) ? 16 : 0;               // T % 4 == 0: a 16-byte piece is in range or not at all
    const float* s = (t_issue < T) ? src : x;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d + uint32_t(j) * (4 * kRS * 4)),
                   "l"(s + (bytes ? j * src_jstride : 0)), "r"(bytes));
    cp_async_commit();
    src += kTC;
    t_issue += kTC;
  };

  // store side
  OutT* yrow = nullptr;        // channel-major: this lane's piece of row (row8 + 4 j)
  OutT* ytm = nullptr;         // time-major: 8 consecutive series (lane % 4) of sample (lane / 4 + 8 i)
  if constexpr (TIME_MAJOR) ytm = y + size_t(lane >> 2) * N + r0 + (lane & 3) * 8;
  else yrow = y + size_t(r0 + row8) * T + c4 * 4;

  for (int c = 0; c < kStages - 1; ++c) {
    if (c < n_chunks) issue_load(c); else cp_async_commit();
  }
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int buf = chunk % kStages;
    cp_async_wait<kStages - 2>();
    __syncwarp();  // tile `buf` visible to every lane; everyone is done with the tile of chunk-1 and with out_tile
    if (chunk + kStages - 1 < n_chunks) issue_load((chunk + kStages - 1) % kStages);
    else cp_async_commit();

    float* row = in_tile + buf * R * kRS + lane * kRS;
    float xin[kTC];
#pragma unroll
    for (int q = 0; q < kTC / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(row + q * 4);
      xin[q * 4 + 0] = v.x; xin[q * 4 + 1] = v.y; xin[q * 4 + 2] = v.z; xin[q * 4 + 3] = v.w;
    }
    if (GATHER && g.norm) {  // same two roundings as csn_gather_trials: the fused and the two-kernel path agree bit for bit
#pragma unroll
      for (int i = 0; i < kTC; ++i) xin[i] = (xin[i] - g.mean) * g.inv_std;
    }
    if constexpr (TIME_MAJOR) {
      cascade_chunk_skewed<NSEC, UNIT_B2>(xin, s1, s2, coef, [&](int idx, float yv) { out_tile[idx * R + lane] = yv; });
    } else {
      cascade_chunk_skewed<NSEC, UNIT_B2>(xin, s1, s2, coef, [&](int idx, float yv) { row[idx] = yv; });
    }
    __syncwarp();

    const int t0 = chunk * kTC;
    if constexpr (TIME_MAJOR) {
      const float* ot = out_tile + (lane >> 2) * R + (lane & 3) * 8;
      const int tl = t0 + (lane >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (tl + 8 * i < T) {
          const float4 a = *reinterpret_cast<const float4*>(ot + i * 8 * R);
          const float4 b = *reinterpret_cast<const float4*>(ot + i * 8 * R + 4);
          OutT* dst = ytm + size_t(8 * i) * N;
          if constexpr (sizeof(OutT) == 4) {
            *reinterpret_cast<float4*>(dst) = a;
            *reinterpret_cast<float4*>(dst + 4) = b;
          } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst) = o;
          }
        }
      }
      ytm += size_t(kTC) * N;
    } else {
      const float* st = in_tile + buf * R * kRS + row8 * kRS + c4 * 4;
      if (t0 + c4 * 4 < T) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(yrow) + j * src_jstride) =
              *reinterpret_cast<const float4*>(st + j * 4 * kRS);
      }
      yrow += kTC;
    }
  }
}

// -------------------------------------------------------------------------------------------- zero phase
template <int NSEC, typename OutT>
__global__ void __launch_bounds__(32) sosfiltfilt_kernel(const float* __restrict__ x, OutT* __restrict__ y,
                                                         const SosCoef coef, long long n_series, int T, int Bn, int C,
                                                         int layout, int rows, int row_stride) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows;
  const int p = coef.padlen;

  // coalesced load: one row at a time, lanes along time
  for (int rr = 0; rr < rows; ++rr) {
    long long r = r0 + rr;
    if (r >= n_series) break;
    const float* src = x + size_t(r) * T;
    float* dst = smem + rr * row_stride;
    for (int t = lane; t < T; t += 32) dst[t] = src[t];
  }
  __syncwarp();

  if (lane < rows && r0 + lane < n_series) {
    float* s = smem + lane * row_stride;  // odd stride -> lanes hit distinct banks
    float s1[NSEC], s2[NSEC];
    const float x0 = s[0], xl = s[T - 1];
    // right-edge odd extension inputs, saved before the forward sweep overwrites the tail
    for (int j = 1; j <= p; ++j) s[T + j - 1] = 2.f * xl - s[T - 1 - j];
    // forward: steady-state initial condition scaled by the first extended sample
    const float xe0 = 2.f * x0 - s[p];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) { s1[k] = coef.zi1[k] * xe0; s2[k] = coef.zi2[k] * xe0; }
    for (int j = p; j >= 1; --j) (void)biquad_cascade<NSEC>(2.f * x0 - s[j], s1, s2, coef);
    for (int t = 0; t < T + p; ++t) s[t] = biquad_cascade<NSEC>(s[t], s1, s2, coef);
    // backward
    const float ye0 = s[T + p - 1];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) { s1[k] = coef.zi1[k] * ye0; s2[k] = coef.zi2[k] * ye0; }
    for (int t = T + p - 1; t >= T; --t) (void)biquad_cascade<NSEC>(s[t], s1, s2, coef);
    for (int t = T - 1; t >= 0; --t) s[t] = biquad_cascade<NSEC>(s[t], s1, s2, coef);
  }
  __syncwarp();

  if (layout == CSN_LAYOUT_BCT) {
    for (int rr = 0; rr < rows; ++rr) {
      long long r = r0 + rr;
      if (r >= n_series) break;
      const float* src = smem + rr * row_stride;
      for (int t = lane; t < T; t += 32) store_out<OutT>(y, size_t(r) * T + t, src[t]);
    }
  } else {
    // lanes along series: consecutive addresses in [T, B*C]; odd row stride keeps the column read conflict-free
    for (int t = 0; t < T; ++t) {
      long long r = r0 + lane;
      if (lane < rows && r < n_series)
        store_out<OutT>(y, out_index(layout, r, t, Bn, C, T), smem[lane * row_stride + t]);
    }
  }
}

template <int NSEC, typename OutT>
static int launch_sosfilt(const float* x, void* y, const SosCoef& coef, int B, int C, int T, int zero_phase,
                          int layout, cudaStream_t s, const GatherSrc& gsrc = GatherSrc{}) {
  const long long n_series = (long long)B * C;
  if (gsrc.idx) {  // fused gather: fast path only (the caller falls back to csn_gather_trials + csn_sosfilt_f32)
    const bool ok = !zero_phase && T % 4 == 0 && C % 32 == 0 && gsrc.row_stride % 4 == 0 && gsrc.t_off % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                    (layout == CSN_LAYOUT_TBC || (layout == CSN_LAYOUT_BCT && sizeof(OutT) == 4));
    if (!ok) {
      set_error("csn_sosfilt_gather_f32: needs causal filtering, C %% 32 == 0, T / T_raw / time_low multiples of 4, 16-byte "
                "aligned arrays and a TBC (or fp32 BCT) output");
      return CSN_EUNSUPPORTED;
    }
  }
  static const bool no_fast = [] { const char* e = getenv("CSN_FILTER_NO_FAST"); return e && e[0] == '1'; }();
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (!zero_phase && (!no_fast || gsrc.idx) && aligned16 && T % 4 == 0 && n_series % 32 == 0 &&
      (layout == CSN_LAYOUT_TBC || (layout == CSN_LAYOUT_BCT && sizeof(OutT) == 4))) {
    const unsigned grid = (unsigned)(n_series / 32);
    bool unit_b2 = NSEC > 1;
    for (int k = 1; k < NSEC; ++k) unit_b2 = unit_b2 && coef.b2[k] == 1.0f;
    const bool tm = layout == CSN_LAYOUT_TBC;
    const size_t smem = size_t(kStages) * 32 * kRS * 4 + (tm ? size_t(kTC) * 32 * 4 : 0);
#define CSN_FILT(TM, UB, GA) sosfilt_warp_kernel<NSEC, OutT, TM, UB, GA><<<grid, 32, smem, s>>>(x, (OutT*)y, coef, T, n_series, gsrc)
    if (gsrc.idx) {
      if (tm) { if (unit_b2) CSN_FILT(true, true, true); else CSN_FILT(true, false, true); }
      else    { if (unit_b2) CSN_FILT(false, true, true); else CSN_FILT(false, false, true); }
    } else {
      if (tm) { if (unit_b2) CSN_FILT(true, true, false); else CSN_FILT(true, false, false); }
      else    { if (unit_b2) CSN_FILT(false, true, false); else CSN_FILT(false, false, false); }
    }
#undef CSN_FILT
    CSN_LAUNCH_CHECK();
    return CSN_OK;
  }
  if (!zero_phase) {
    // rows per block: keep >= 4 blocks per SM in flight when the problem allows, warps of 32 series
    int R = 64;
    if (n_series >= (long long)sm_count() * 4 * 128) R = 128;
    if (n_series < (long long)sm_count() * 256) R = 32;  // cfg2 (32768 series): 1024 single-warp blocks balance the 148 SMs best
    static const int forced_r = [] { const char* e = getenv("CSN_FILTER_R"); return e ? atoi(e) : 0; }();
    if (forced_r == 32 || forced_r == 64 || forced_r == 128) R = forced_r;
    size_t smem = size_t(kStages) * R * kRS * 4 + size_t(kTC) * R * 4;
    int vec_ok = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                 ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    unsigned grid = (unsigned)ceil_div<long long>(n_series, R);
    if (smem > 48 * 1024) {
      static bool attr_set = false;
      if (!attr_set) {
        CSN_CUDA(cudaFuncSetAttribute(sosfilt_stream_kernel<NSEC, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
      }
    }
    sosfilt_stream_kernel<NSEC, OutT><<<grid, R, smem, s>>>(x, (OutT*)y, coef, n_series, T, B, C, layout, vec_ok);
  } else {
    int row_stride = (T + coef.padlen) | 1;
    int rows = 32;
    while (rows > 1 && size_t(rows) * row_stride * 4 > 200 * 1024) rows >>= 1;
    size_t smem = size_t(rows) * row_stride * 4;
    if (smem > 227 * 1024) {
      set_error("csn_sosfilt_f32: zero-phase series too long for shared memory (T=%d)", T);
      return CSN_EUNSUPPORTED;
    }
    auto kern = sosfiltfilt_kernel<NSEC, OutT>;
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)ceil_div<long long>(n_series, rows);
    kern<<<grid, 32, smem, s>>>(x, (OutT*)y, coef, n_series, T, B, C, layout, rows, row_stride);
  }
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

template <typename OutT>
static int dispatch_nsec(int nsec, const float* x, void* y, const SosCoef& coef, int B, int C, int T, int zp,
                         int layout, cudaStream_t s, const GatherSrc& g = GatherSrc{}) {
  switch (nsec) {
    case 1: return launch_sosfilt<1, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 2: return launch_sosfilt<2, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 3: return launch_sosfilt<3, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 4: return launch_sosfilt<4, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 5: return launch_sosfilt<5, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 6: return launch_sosfilt<6, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 7: return launch_sosfilt<7, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
    case 8: return launch_sosfilt<8, OutT>(x, y, coef, B, C, T, zp, layout, s, g);
  }
  set_error("csn_sosfilt_f32: n_sections must be in [1, 8], got %d", nsec);
  return CSN_EINVAL;
}

}  // namespace csn

using namespace csn;

static int make_coef(const double* sos, int n_sections, SosCoef& c, const char* who) {
  CSN_REQUIRE(n_sections >= 1 && n_sections <= kMaxSec, "%s: n_sections must be in [1, %d]", who, kMaxSec);
  int nb2 = 0, na2 = 0;
  double scale = 1.0;
  for (int k = 0; k < n_sections; ++k) {
    const double* r = sos + 6 * k;
    CSN_REQUIRE(r[3] != 0.0, "%s: a0 == 0 in section %d", who, k);
    double b0 = r[0] / r[3], b1 = r[1] / r[3], b2 = r[2] / r[3], a1 = r[4] / r[3], a2 = r[5] / r[3];
    c.b0[k] = (float)b0; c.b1[k] = (float)b1; c.b2[k] = (float)b2; c.a1[k] = (float)a1; c.a2[k] = (float)a2;
    if (r[2] == 0.0) ++nb2;
    if (r[5] == 0.0) ++na2;
    // steady-state (unit-step) state of this DF2T section, scaled by the DC gain of the sections before it
    double asum = 1.0 + a1 + a2;
    double yss = (asum != 0.0) ? (b0 + b1 + b2) / asum : 0.0;
    double z2 = b2 - a2 * yss;
    double z1 = b1 - a1 * yss + z2;
    c.zi1[k] = (float)(scale * z1);
    c.zi2[k] = (float)(scale * z2);
    scale *= yss;
  }
  int ntaps = 2 * n_sections + 1 - (nb2 < na2 ? nb2 : na2);
  c.padlen = 3 * ntaps;
  return CSN_OK;
}

extern "C" int csn_sosfilt_f32(const float* x, void* y, const double* sos, int n_sections, int B, int C, int T,
                               int zero_phase, int out_layout, int out_dtype, void* stream) {
  CSN_REQUIRE(B >= 0 && C >= 0 && T >= 0, "csn_sosfilt_f32: negative dimension");
  if (B == 0 || C == 0 || T == 0) return CSN_OK;  // empty batch: nothing to do (pointers may be null)
  CSN_REQUIRE(x && y && sos, "csn_sosfilt_f32: null pointer");
  CSN_REQUIRE(out_layout >= CSN_LAYOUT_BCT && out_layout <= CSN_LAYOUT_TBC, "csn_sosfilt_f32: bad out_layout");
  CSN_REQUIRE(out_dtype == CSN_F32 || out_dtype == CSN_BF16, "csn_sosfilt_f32: bad out_dtype");
  SosCoef c{};
  if (int rc = make_coef(sos, n_sections, c, "csn_sosfilt_f32")) return rc;
  if (zero_phase) CSN_REQUIRE(T > c.padlen, "csn_sosfilt_f32: zero-phase needs T > padlen (%d), got T=%d", c.padlen, T);
  cudaStream_t s = as_stream(stream);
  if (out_dtype == CSN_F32) return dispatch_nsec<float>(n_sections, x, y, c, B, C, T, zero_phase, out_layout, s);
  return dispatch_nsec<__nv_bfloat16>(n_sections, x, y, c, B, C, T, zero_phase, out_layout, s);
}

extern "C" int csn_sosfilt_gather_f32(const float* src, const long long* idx, int n_trials, int C, int T_raw, int time_low,
                                      int time_high, float mean, float std, void* y, const double* sos, int n_sections,
                                      int B, int out_layout, int out_dtype, void* stream) {
  CSN_REQUIRE(n_trials >= 1 && C >= 1 && T_raw >= 1 && B >= 0, "csn_sosfilt_gather_f32: bad sizes");
  CSN_REQUIRE(time_low >= 0 && time_high > time_low && time_high <= T_raw,
              "csn_sosfilt_gather_f32: need 0 <= time_low < time_high <= T_raw (got %d, %d, T_raw=%d)", time_low, time_high, T_raw);
  CSN_REQUIRE(std != 0.f, "csn_sosfilt_gather_f32: std must be non-zero");
  if (B == 0) return CSN_OK;
  CSN_REQUIRE(src && idx && y && sos, "csn_sosfilt_gather_f32: null pointer");
  CSN_REQUIRE(out_layout >= CSN_LAYOUT_BCT && out_layout <= CSN_LAYOUT_TBC, "csn_sosfilt_gather_f32: bad out_layout");
  CSN_REQUIRE(out_dtype == CSN_F32 || out_dtype == CSN_BF16, "csn_sosfilt_gather_f32: bad out_dtype");
  SosCoef c{};
  if (int rc = make_coef(sos, n_sections, c, "csn_sosfilt_gather_f32")) return rc;
  GatherSrc g{};
  g.idx = idx; g.n_trials = n_trials; g.C = C; g.row_stride = T_raw; g.t_off = time_low;
  g.mean = mean; g.inv_std = 1.f / std; g.norm = (mean != 0.f || std != 1.f) ? 1 : 0;
  const int T = time_high - time_low;
  cudaStream_t s = as_stream(stream);
  if (out_dtype == CSN_F32) return dispatch_nsec<float>(n_sections, src, y, c, B, C, T, 0, out_layout, s, g);
  return dispatch_nsec<__nv_bfloat16>(n_sections, src, y, c, B, C, T, 0, out_layout, s, g);
}
