// Exact top-k retrieval of EEG-encoder embeddings against an image-feature gallery: replaces the faiss IndexFlatL2
// add/search of utils/Utilities.py:45-58 (LstmDistillFromDinoV2Eval.py:333-380).  One kernel computes the score tile
// of 32 queries x 128 gallery rows from shared memory (fp32 FMAs: the ranking must match an exact fp32/fp64 search,
// so no bf16 tensor-core scores here) and keeps each query's k best in a warp-distributed sorted list -- the
// [nq, nb] score matrix never exists in HBM.  The gallery may be split over gridDim.y; a second small kernel merges
// the per-split lists.  metric 0: squared L2 (sum of squared differences, as IndexFlatL2 reports), ascending;
// metric 1: inner product (cosine on normalised rows), descending.  Ties go to the lower gallery index.
#include <float.h>

#include "common.cuh"

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
