// GPU-resident EEG dataset: the whole "dataset" list of the .pth file written by
// ConvertToPth.py:170-201 lives in HBM as one [N, C, T_raw] fp32 tensor; a batch is ONE gather kernel that does what
// EEGDataset.__getitem__ does per item on the host (utils/PerilsEEGDataset.py:541-573): pick trial idx[b], crop
// [time_low, time_high), optionally (x - mean) / std with the dataset-level scalars (:572-573), and emit either the
// stored layout [B, C, T] (what the fused band-pass kernel consumes) or the DataLoader layout [B, T, C] (`eeg.t()`).
#include "common.cuh"

namespace csn {

// one CTA per (trial, channel tile of 32); threads sweep time: coalesced reads of src rows, and for the [B,T,C] output
// a 32 x 33 shared tile turns them into coalesced writes along C
__global__ void __launch_bounds__(256) gather_trials_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                                           float* __restrict__ out, int N, int C, int T_raw, int t_low, int T,
                                                           float mean, float inv_std, int layout) {
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * 32;
  long long n = idx[b];
  if (n < 0) n += N;  // python-style negative indices
  const float* base = src + (size_t(n) * C) * T_raw + t_low;
  if (layout == CSN_LAYOUT_BCT) {
    for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
      const int c = c0 + cc;
      if (c >= C) break;
      const float* row = base + size_t(c) * T_raw;
      float* orow = out + (size_t(b) * C + c) * T;
      for (int t = threadIdx.x & 31; t < T; t += 32) orow[t] = (row[t] - mean) * inv_std;
    }
    return;
  }
  __shared__ float tile[32][33];
  for (int t0 = 0; t0 < T; t0 += 32) {
    for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
      const int c = c0 + cc, t = t0 + (threadIdx.x & 31);
      tile[cc][threadIdx.x & 31] = (c < C && t < T) ? (base[size_t(c) * T_raw + t] - mean) * inv_std : 0.f;
    }
    __syncthreads();
    for (int tt = threadIdx.x >> 5; tt < 32; tt += 8) {
      const int t = t0 + tt, c = c0 + (threadIdx.x & 31);
      if (t < T && c < C) out[(size_t(b) * T + t) * C + c] = tile[threadIdx.x & 31][tt];
    }
    __syncthreads();
  }
}

// Channel selection of EEGDataset.__getitem__ (utils/PerilsEEGDataset.py:554-565, `filter_channels`): keep the listed
// channels, crop [t_low, t_low + T), and -- `apply_channel_wise_norm` -- z-score every (trial, channel) row over the
// cropped window with the POPULATION standard deviation (normlizeEEG :454-461 runs on a numpy array there).  One warp
// per output row; done ONCE when the dataset is uploaded, the result is the resident tensor the batches are drawn from.
__global__ void __launch_bounds__(256) select_crop_zscore_kernel(const float* __restrict__ src, const int* __restrict__ channels,
                                                                float* __restrict__ out, long long rows, int C, int T_raw,
                                                                int n_sel, int t_low, int T, int zscore) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long n = row / n_sel;
  const int s = (int)(row - n * n_sel);
  const float* x = src + (size_t(n) * C + channels[s]) * T_raw + t_low;
  float* y = out + size_t(row) * T;
  float mean = 0.f, sd = 1.f;
  if (zscore) {
    float a = 0.f;
    for (int t = lane; t < T; t += 32) a += x[t];
    mean = warp_sum(a) / float(T);
    float v = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float d = x[t] - mean;
      v = fmaf(d, d, v);
    }
    sd = sqrtf(warp_sum(v) / float(T));
  }
  for (int t = lane; t < T; t += 32) y[t] = (x[t] - mean) / sd;
}

}  // namespace csn

using namespace csn;

extern "C" int csn_select_crop_zscore(const float* src, const int* channels, float* out, int N, int C, int T_raw, int n_sel,
                                      int time_low, int time_high, int zscore, void* stream) {
  CSN_REQUIRE(N >= 0 && C >= 1 && T_raw >= 1 && n_sel >= 1, "csn_select_crop_zscore: bad sizes");
  CSN_REQUIRE(time_low >= 0 && time_high > time_low && time_high <= T_raw,
              "csn_select_crop_zscore: need 0 <= time_low < time_high <= T_raw (got %d, %d, T_raw=%d)", time_low, time_high, T_raw);
  if (N == 0) return CSN_OK;
  CSN_REQUIRE(src && channels && out, "csn_select_crop_zscore: null pointer");
  const long long rows = (long long)N * n_sel;
  const long long blocks = ceil_div<long long>(rows, 8);
  CSN_REQUIRE(blocks <= 2147483647LL, "csn_select_crop_zscore: grid too large");
  select_crop_zscore_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, channels, out, rows, C, T_raw, n_sel, time_low,
                                                                            time_high - time_low, zscore ? 1 : 0);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

extern "C" int csn_gather_trials(const float* src, const long long* idx, float* out, int N, int C, int T_raw, int B,
                                 int time_low, int time_high, float mean, float std, int out_layout, void* stream) {
  CSN_REQUIRE(N >= 0 && C >= 1 && T_raw >= 1 && B >= 0, "csn_gather_trials: bad sizes");
  CSN_REQUIRE(time_low >= 0 && time_high > time_low && time_high <= T_raw,
              "csn_gather_trials: need 0 <= time_low < time_high <= T_raw (got %d, %d, T_raw=%d)", time_low, time_high, T_raw);
  CSN_REQUIRE(out_layout == CSN_LAYOUT_BCT || out_layout == CSN_LAYOUT_BTC, "csn_gather_trials: out_layout must be BCT or BTC");
  CSN_REQUIRE(std != 0.f, "csn_gather_trials: std must be non-zero");
  if (B == 0) return CSN_OK;
  CSN_REQUIRE(src && idx && out, "csn_gather_trials: null pointer");
  CSN_REQUIRE(B <= 2147483647 / 1 && ceil_div(C, 32) <= 65535, "csn_gather_trials: grid too large");
  dim3 grid(B, ceil_div(C, 32));
  gather_trials_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, idx, out, N, C, T_raw, time_low, time_high - time_low, mean,
                                                            1.f / std, out_layout);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}
