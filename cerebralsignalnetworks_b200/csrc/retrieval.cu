// Exact top-k retrieval of EEG-encoder embeddings against an image-feature gallery: replaces the faiss IndexFlatL2
// add/search of utils/Utilities.py:45-58 (LstmDistillFromDinoV2Eval.py:333-380).  One kernel computes the score tile
// of 32 queries x 128 gallery rows from shared memory (fp32 FMAs: the ranking must match an exact fp32/fp64 search,
// so no bf16 tensor-core scores here) and keeps each query's k best in a warp-distributed sorted list -- the
// [nq, nb] score matrix never exists in HBM.  The gallery may be split over gridDim.y; a second small kernel merges
// the per-split lists.  metric 0: squared L2 (sum of squared differences, as IndexFlatL2 reports), ascending;
// metric 1: inner product (cosine on normalised rows), descending.  Ties go to the lower gallery index.
#include <float.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace csn {

constexpr int kQT = 32, kGT = 128, kDK = 32, kRowStride = kDK + 4;  // tile sizes; 36-float rows: conflict-free float4
constexpr int kTopkThreads = 256;
constexpr int kMaxK = 32;

// Warp-distributed sorted list (best first), one entry per lane; entries past k are never read.
// `better(a, ia, b, ib)`: candidate a (index ia) ranks strictly before b (index ib).
__device__ __forceinline__ bool ranks_before(float a, long long ia, float b, long long ib) {
  return a < b || (a == b && ia < ib);
}
// insert (val, idx) into the list held in (lv, li) across lanes [0, k); all lanes call with the same candidate
__device__ __forceinline__ void list_insert(float& lv, long long& li, float val, long long idx, int k, int lane) {
  const bool mine_after = ranks_before(val, idx, lv, li);       // candidate goes before this lane's entry
  const unsigned m = __ballot_sync(0xffffffffu, mine_after && lane < k);
  if (m == 0) return;
  const int pos = __ffs(m) - 1;                                  // first lane whose entry ranks after the candidate
  const float up_v = __shfl_up_sync(0xffffffffu, lv, 1);
  const long long up_i = __shfl_up_sync(0xffffffffu, li, 1);
  if (lane == pos) { lv = val; li = idx; }
  else if (lane > pos) { lv = up_v; li = up_i; }
}

template <int METRIC>
__global__ void __launch_bounds__(kTopkThreads) topk_scan_kernel(const float* __restrict__ gallery, const float* __restrict__ query,
                                                                int nb, int nq, int d, int k, int rows_per_split,
                                                                float* __restrict__ part_val, long long* __restrict__ part_idx) {
  __shared__ __align__(16) float Qs[kQT][kRowStride];
  __shared__ __align__(16) float Gs[kGT][kRowStride];
  __shared__ float Ss[kQT][kGT + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * kQT;
  const int split = blockIdx.y;
  const int g_begin = split * rows_per_split, g_end = min(nb, g_begin + rows_per_split);

  // thread -> 4 queries (warp * 4 + i) x 4 gallery rows (lane + 32 j) of the tile
  float lv[4];
  long long li[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { lv[i] = FLT_MAX; li[i] = LLONG_MAX; }

  for (int g0 = g_begin; g0 < g_end; g0 += kGT) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += kDK) {
      // stage the [32 x 32] query chunk and the [128 x 32] gallery chunk (zero padded)
      for (int e = tid; e < kQT * kDK; e += kTopkThreads) {
        const int r = e / kDK, c = e % kDK;
        Qs[r][c] = (q0 + r < nq && k0 + c < d) ? query[size_t(q0 + r) * d + k0 + c] : 0.f;
      }
      for (int e = tid; e < kGT * kDK; e += kTopkThreads) {
        const int r = e / kDK, c = e % kDK;
        Gs[r][c] = (g0 + r < g_end && k0 + c < d) ? gallery[size_t(g0 + r) * d + k0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kDK; kk += 4) {
        float4 qv[4], gv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[warp * 4 + i][kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) gv[j] = *reinterpret_cast<const float4*>(&Gs[lane + 32 * j][kk]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (METRIC == 0) {
              float t;
              t = qv[i].x - gv[j].x; acc[i][j] = fmaf(t, t, acc[i][j]);
              t = qv[i].y - gv[j].y; acc[i][j] = fmaf(t, t, acc[i][j]);
              t = qv[i].z - gv[j].z; acc[i][j] = fmaf(t, t, acc[i][j]);
              t = qv[i].w - gv[j].w; acc[i][j] = fmaf(t, t, acc[i][j]);
            } else {
              acc[i][j] = fmaf(qv[i].x, gv[j].x, acc[i][j]);
              acc[i][j] = fmaf(qv[i].y, gv[j].y, acc[i][j]);
              acc[i][j] = fmaf(qv[i].z, gv[j].z, acc[i][j]);
              acc[i][j] = fmaf(qv[i].w, gv[j].w, acc[i][j]);
            }
          }
      }
      __syncthreads();
    }
    // scores -> shared tile in "smaller is better" form; rows past the split end can never win
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = g0 + lane + 32 * j;
        Ss[warp * 4 + i][lane + 32 * j] = (g < g_end) ? (METRIC == 0 ? acc[i][j] : -acc[i][j]) : FLT_MAX;
      }
    __syncwarp();  // a warp only reads back the rows it wrote
    // each warp folds its 4 query rows: candidates in index order, only those beating the current k-th best
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      for (int c0 = 0; c0 < kGT; c0 += 32) {
        const float v = Ss[warp * 4 + i][c0 + lane];
        const long long gi = (long long)g0 + c0 + lane;
        const float kth_v = __shfl_sync(0xffffffffu, lv[i], k - 1);
        const long long kth_i = __shfl_sync(0xffffffffu, li[i], k - 1);
        unsigned cand = __ballot_sync(0xffffffffu, v != FLT_MAX && ranks_before(v, gi, kth_v, kth_i));
        while (cand) {
          const int src = __ffs(cand) - 1;
          cand &= cand - 1;
          const float cv = __shfl_sync(0xffffffffu, v, src);
          const long long ci = __shfl_sync(0xffffffffu, gi, src);
          list_insert(lv[i], li[i], cv, ci, k, lane);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + warp * 4 + i;
    if (q < nq && lane < k) {
      const size_t o = (size_t(q) * gridDim.y + split) * k + lane;
      part_val[o] = lv[i];
      part_idx[o] = li[i];
    }
  }
}

// one warp per query: merge `splits` sorted lists of k, write distances (sign restored) and indices (-1 when fewer than k rows)
__global__ void topk_merge_kernel(const float* __restrict__ part_val, const long long* __restrict__ part_idx, int nq, int splits,
                                  int k, int metric, float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  float lv = FLT_MAX;
  long long li = LLONG_MAX;
  for (int s = 0; s < splits; ++s)
    for (int e = 0; e < k; ++e) {
      const size_t o = (size_t(q) * splits + s) * k + e;
      const float v = part_val[o];
      const long long i = part_idx[o];
      if (i != LLONG_MAX) list_insert(lv, li, v, i, k, lane);
    }
  if (lane < k) {
    const bool ok = li != LLONG_MAX;
    out_val[size_t(q) * k + lane] = ok ? (metric == 0 ? lv : -lv) : (metric == 0 ? FLT_MAX : -FLT_MAX);
    out_idx[size_t(q) * k + lane] = ok ? li : -1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core path for galleries of thousands of rows: a tcgen05 distance GEMM PRE-SELECTS, exact fp32
// arithmetic DECIDES.
//   1. rows are split into bf16 pairs x = hi + lo (lo = bf16(x - hi)); the scores' dot products are ONE GEMM over the
//      concatenated operands [Q_hi | Q_hi | Q_lo] . [G_hi | G_lo | G_hi]^T (K = 3 d): every bf16 x bf16 product is exact in
//      the fp32 accumulator and only the lo.lo terms (<= 2^-16 relative) are dropped -- the accuracy class of an fp32
//      BLAS search, which is what faiss' IndexFlatL2 itself runs for batched queries;
//   2. one warp per query scans its row of approximate scores (||g||^2 - 2 q.g, or -q.g) and keeps the kCand = 32 best
//      (a warp-distributed sorted list; ties to the lower index; the gallery is processed in chunks so the [nq, chunk]
//      score block stays small);
//   3. the 32 candidates' distances are recomputed in fp32 in the direct form sum (q - g)^2 -- exact zeros for duplicates,
//      no cancellation -- and the k best of those are the answer.  The ranking therefore follows exact fp32 distances;
//      the GEMM only has to be right about WHICH 32 rows are worth the exact look (approximate error ~1e-6 of the norms
//      against a margin of 32 - k ranks).
constexpr int kCand = 32;
constexpr int kTcChunk = 32768;  // gallery rows per GEMM (score block nq x chunk fp32)

// dst [n_pad, 3 d] bf16 = (hi | hi | lo) (queries) or (hi | lo | hi) (gallery) of src [n, d]; rows >= n are zero;
// norm[r] = sum x^2 in fp32 (lane-strided partials, fixed shuffle tree) when requested.  One warp per row.
__global__ void split_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, float* __restrict__ norm,
                                  int n, int n_pad, int d, int gallery_order) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  __nv_bfloat16* o = dst + size_t(row) * 3 * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float x = row < n ? src[size_t(row) * d + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    o[c] = hi;
    o[d + c] = gallery_order ? lo : hi;
    o[2 * d + c] = gallery_order ? hi : lo;
    acc = fmaf(x, x, acc);
  }
  if (norm) {
    acc = warp_sum(acc);
    if (lane == 0) norm[row] = acc;
  }
}

// one warp per query: fold the approximate scores of gallery rows [g0, g0 + nbc) into the query's running candidate list
template <int METRIC>
__global__ void select_candidates_kernel(const float* __restrict__ dots, int ldd, const float* __restrict__ gnorm, int nq, int nbc,
                                         long long g0, int first, float* __restrict__ cand_val, long long* __restrict__ cand_idx) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  float lv = first ? FLT_MAX : cand_val[size_t(q) * kCand + lane];
  long long li = first ? LLONG_MAX : cand_idx[size_t(q) * kCand + lane];
  const float* row = dots + size_t(q) * ldd;
  // eight 32-column groups per round: all sixteen loads of a round are in flight together (the scan is a chain of L2
  // round trips otherwise: 313 of them per query at 10 k gallery rows)
  constexpr int U = 8;
  for (int c0 = 0; c0 < nbc; c0 += 32 * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 32 * u + lane;
      float dot = 0.f, gn = 0.f;
      if (c < nbc) {
        dot = __ldg(row + c);
        if (METRIC == 0) gn = __ldg(gnorm + c);
      }
      v[u] = c < nbc ? (METRIC == 0 ? fmaf(-2.f, dot, gn) : -dot) : FLT_MAX;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 32 * u + lane;
      const long long gi = g0 + c;
      const float kth_v = __shfl_sync(0xffffffffu, lv, kCand - 1);
      const long long kth_i = __shfl_sync(0xffffffffu, li, kCand - 1);
      unsigned cand = __ballot_sync(0xffffffffu, c < nbc && ranks_before(v[u], gi, kth_v, kth_i));
      while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        list_insert(lv, li, __shfl_sync(0xffffffffu, v[u], src), __shfl_sync(0xffffffffu, gi, src), kCand, lane);
      }
    }
  }
  cand_val[size_t(q) * kCand + lane] = lv;
  cand_idx[size_t(q) * kCand + lane] = li;
}

// one warp per query: exact fp32 scores of its candidates (direct form), the k best of them in rank order
template <int METRIC>
__global__ void rerank_kernel(const float* __restrict__ gallery, const float* __restrict__ query, const long long* __restrict__ cand_idx,
                              int nq, int d, int k, float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  // lane e scores candidate e (the 32 rows are independent streams of 16-byte loads; sequential fp32 accumulation over
  // the features, the order the fused scan kernel uses), then the warp inserts the 32 exact scores into the result list
  const float4* qv = reinterpret_cast<const float4*>(query + size_t(q) * d);
  const long long mine = cand_idx[size_t(q) * kCand + lane];
  float sc = FLT_MAX;
  if (mine != LLONG_MAX) {
    const float4* gv = reinterpret_cast<const float4*>(gallery + size_t(mine) * d);
    float acc = 0.f;
    for (int c = 0; c < d / 4; ++c) {
      const float4 a = __ldg(qv + c), b = __ldg(gv + c);
      if (METRIC == 0) {
        float t;
        t = a.x - b.x; acc = fmaf(t, t, acc);
        t = a.y - b.y; acc = fmaf(t, t, acc);
        t = a.z - b.z; acc = fmaf(t, t, acc);
        t = a.w - b.w; acc = fmaf(t, t, acc);
      } else {
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
    }
    sc = METRIC == 0 ? acc : -acc;
  }
  float lv = FLT_MAX;
  long long li = LLONG_MAX;
  for (int e = 0; e < kCand; ++e) {
    const long long gi = __shfl_sync(0xffffffffu, mine, e);
    const float v = __shfl_sync(0xffffffffu, sc, e);
    if (gi == LLONG_MAX) break;  // the candidate list is sorted: empty slots are last (warp-uniform)
    list_insert(lv, li, v, gi, k, lane);
  }
  if (lane < k) {
    const bool ok = li != LLONG_MAX;
    out_val[size_t(q) * k + lane] = ok ? (METRIC == 0 ? lv : -lv) : (METRIC == 0 ? FLT_MAX : -FLT_MAX);
    out_idx[size_t(q) * k + lane] = ok ? li : -1;
  }
}

static inline size_t al256r(size_t x) { return (x + 255) & ~size_t(255); }
struct TcWs {
  __nv_bfloat16* q3;   // [nq_pad, 3d]
  __nv_bfloat16* g3;   // [chunk_pad, 3d]
  float* gnorm;        // [chunk_pad]
  float* dots;         // [nq_pad, chunk_pad]
  float* cand_val;     // [nq, 32]
  long long* cand_idx; // [nq, 32]
  size_t total;
};
static TcWs carve_tc(void* base, int nq, int nb, int d) {
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* r = p ? p + off : nullptr; off += al256r(bytes); return r; };
  const size_t nq_pad = size_t(ceil_div(nq, 128)) * 128, chunk = size_t(ceil_div(std::min(nb, kTcChunk), 128)) * 128;
  TcWs w;
  w.q3 = reinterpret_cast<__nv_bfloat16*>(take(nq_pad * 3 * d * 2));
  w.g3 = reinterpret_cast<__nv_bfloat16*>(take(chunk * 3 * d * 2));
  w.gnorm = reinterpret_cast<float*>(take(chunk * 4));
  w.dots = reinterpret_cast<float*>(take(nq_pad * chunk * 4));
  w.cand_val = reinterpret_cast<float*>(take(size_t(nq) * kCand * 4));
  w.cand_idx = reinterpret_cast<long long*>(take(size_t(nq) * kCand * 8));
  w.total = off + 256;
  return w;
}
// shapes the tensor-core path serves (TMA rows of 16 bytes, enough rows to be worth three launches)
static bool tc_search_ok(int nq, int nb, int d, int k) { return nb >= 2048 && nq >= 8 && d % 8 == 0 && d <= 8192 && k <= kCand / 2; }

static int topk_splits(int nq, int nb) {
  const int q_tiles = ceil_div(nq, kQT);
  int splits = std::max(1, (2 * sm_count()) / std::max(1, q_tiles));
  splits = std::min(splits, std::max(1, ceil_div(nb, 4 * kGT)));  // at least four gallery tiles per split
  return std::min(splits, 64);
}

}  // namespace csn

using namespace csn;

extern "C" int csn_topk_workspace_bytes(int nq, int nb, int k, size_t* bytes) {
  CSN_REQUIRE(bytes, "csn_topk_workspace_bytes: null pointer");
  CSN_REQUIRE(nq >= 0 && nb >= 0 && k >= 1 && k <= kMaxK, "csn_topk_workspace_bytes: need k in [1, %d]", kMaxK);
  *bytes = size_t(std::max(nq, 1)) * topk_splits(nq, nb) * k * (sizeof(float) + sizeof(long long)) + 256;
  return CSN_OK;
}

extern "C" int csn_topk_tc_workspace_bytes(int nq, int nb, int d, int k, size_t* bytes) {
  CSN_REQUIRE(bytes, "csn_topk_tc_workspace_bytes: null pointer");
  CSN_REQUIRE(nq >= 0 && nb >= 0 && d >= 1 && k >= 1 && k <= kMaxK, "csn_topk_tc_workspace_bytes: need d >= 1 and k in [1, %d]", kMaxK);
  size_t simt = 0;
  CSN_TRY(csn_topk_workspace_bytes(nq, nb, k, &simt));
  *bytes = tc_search_ok(nq, nb, d, k) ? std::max(simt, carve_tc(nullptr, nq, nb, d).total) : simt;
  return CSN_OK;
}

extern "C" int csn_topk_search(const float* gallery, const float* query, int nb, int nq, int d, int k, int metric,
                               float* out_dist, long long* out_idx, void* workspace, void* stream);

extern "C" int csn_topk_search_tc(const float* gallery, const float* query, int nb, int nq, int d, int k, int metric,
                                  float* out_dist, long long* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  CSN_REQUIRE(nq >= 0 && nb >= 0 && d >= 1, "csn_topk_search_tc: bad sizes");
  CSN_REQUIRE(k >= 1 && k <= kMaxK, "csn_topk_search_tc: k must be in [1, %d], got %d", kMaxK, k);
  CSN_REQUIRE(metric == 0 || metric == 1, "csn_topk_search_tc: metric must be 0 (squared L2) or 1 (inner product)");
  if (nq == 0) return CSN_OK;
  // other shapes (small galleries, odd feature sizes, k > 16) are served by the fused fp32 scan kernel: same results
  if (!tc_search_ok(nq, nb, d, k)) return csn_topk_search(gallery, query, nb, nq, d, k, metric, out_dist, out_idx, workspace, stream);
  CSN_REQUIRE(query && gallery && out_dist && out_idx && workspace, "csn_topk_search_tc: null pointer");
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  TcWs w = carve_tc(base, nq, nb, d);
  CSN_REQUIRE(workspace_bytes >= w.total, "csn_topk_search_tc: workspace of %zu bytes, need %zu (csn_topk_tc_workspace_bytes)",
              workspace_bytes, w.total);
  cudaStream_t s = as_stream(stream);
  const int nq_pad = ceil_div(nq, 128) * 128;
  const int chunk_cap = ceil_div(std::min(nb, kTcChunk), 128) * 128;
  split_rows_kernel<<<ceil_div(nq_pad, 8), 256, 0, s>>>(query, w.q3, nullptr, nq, nq_pad, d, 0);
  CSN_LAUNCH_CHECK();
  for (int g0 = 0; g0 < nb; g0 += kTcChunk) {
    const int nbc = std::min(kTcChunk, nb - g0), nbc_pad = ceil_div(nbc, 128) * 128;
    split_rows_kernel<<<ceil_div(nbc_pad, 8), 256, 0, s>>>(gallery + size_t(g0) * d, w.g3, w.gnorm, nbc, nbc_pad, d, 1);
    CSN_LAUNCH_CHECK();
    CSN_TRY(gemm_tc_run(0, 1, nq_pad, nbc_pad, 3 * d, w.q3, 3 * d, w.g3, 3 * d, w.dots, chunk_cap, CSN_F32, nullptr, 0, 1, nullptr, s));
    if (metric == 0)
      select_candidates_kernel<0><<<ceil_div(nq, 8), 256, 0, s>>>(w.dots, chunk_cap, w.gnorm, nq, nbc, g0, g0 == 0, w.cand_val, w.cand_idx);
    else
      select_candidates_kernel<1><<<ceil_div(nq, 8), 256, 0, s>>>(w.dots, chunk_cap, w.gnorm, nq, nbc, g0, g0 == 0, w.cand_val, w.cand_idx);
    CSN_LAUNCH_CHECK();
  }
  if (metric == 0) rerank_kernel<0><<<ceil_div(nq, 8), 256, 0, s>>>(gallery, query, w.cand_idx, nq, d, k, out_dist, out_idx);
  else rerank_kernel<1><<<ceil_div(nq, 8), 256, 0, s>>>(gallery, query, w.cand_idx, nq, d, k, out_dist, out_idx);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_topk_search(const float* gallery, const float* query, int nb, int nq, int d, int k, int metric,
                               float* out_dist, long long* out_idx, void* workspace, void* stream) {
  CSN_REQUIRE(nq >= 0 && nb >= 0 && d >= 1, "csn_topk_search: bad sizes");
  CSN_REQUIRE(k >= 1 && k <= kMaxK, "csn_topk_search: k must be in [1, %d], got %d", kMaxK, k);
  CSN_REQUIRE(metric == 0 || metric == 1, "csn_topk_search: metric must be 0 (squared L2) or 1 (inner product)");
  if (nq == 0) return CSN_OK;
  CSN_REQUIRE(query && out_dist && out_idx && workspace && (gallery || nb == 0), "csn_topk_search: null pointer");
  cudaStream_t s = as_stream(stream);
  const int splits = topk_splits(nq, nb);
  const int rows_per_split = ceil_div(ceil_div(std::max(nb, 1), splits), kGT) * kGT;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  long long* part_idx = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  float* part_val = reinterpret_cast<float*>(part_idx + size_t(nq) * splits * k);
  dim3 grid(ceil_div(nq, kQT), splits);
  if (metric == 0) topk_scan_kernel<0><<<grid, kTopkThreads, 0, s>>>(gallery, query, nb, nq, d, k, rows_per_split, part_val, part_idx);
  else topk_scan_kernel<1><<<grid, kTopkThreads, 0, s>>>(gallery, query, nb, nq, d, k, rows_per_split, part_val, part_idx);
  CSN_LAUNCH_CHECK();
  topk_merge_kernel<<<ceil_div(nq, 8), 256, 0, s>>>(part_val, part_idx, nq, splits, k, metric, out_dist, out_idx);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
