// C-ABI entry points of the LSTM encoder layer: dispatch on compute dtype (precision mode, not a fallback --
// an unsupported shape in the requested mode is an error, never a silent switch).
#include "common.cuh"

namespace csn {
int lstm_layer_fwd_f32(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       float* h_seq, float* reserve, float* workspace, int T, int B, int I, int H, cudaStream_t s);
int lstm_layer_bwd_f32(const float* x, const float* w_ih, const float* w_hh, const float* h_seq, const float* reserve,
                       const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                       float* dx, float* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s);
int lstm_tc_bytes(int T, int B, int I, int H, size_t* reserve, size_t* workspace);
int lstm_layer_fwd_tc(const void* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                      void* h_seq, void* reserve, void* workspace, int T, int B, int I, int H, int training,
                      cudaStream_t s);
int lstm_layer_bwd_tc(const void* x, const float* w_ih, const float* w_hh, const void* h_seq, const void* reserve,
                      const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                      float* dx, void* workspace, int T, int B, int I, int H, int accumulate, cudaStream_t s);
}  // namespace csn

using namespace csn;

static int check_dims(const char* fn, int T, int B, int I, int H) {
  CSN_REQUIRE(T >= 1 && B >= 1 && I >= 1 && H >= 1, "%s: dimensions must be positive (T=%d B=%d I=%d H=%d)", fn, T, B, I, H);
  CSN_REQUIRE((long long)T * B * 4 * H < (1ll << 40), "%s: problem too large", fn);
  return CSN_OK;
}

extern "C" int csn_lstm_layer_bytes(int T, int B, int I, int H, int compute_dtype, size_t* reserve_bytes,
                                    size_t* workspace_bytes) {
  CSN_TRY(check_dims("csn_lstm_layer_bytes", T, B, I, H));
  size_t r = 0, w = 0;
  if (compute_dtype == CSN_F32) {
    r = size_t(T) * B * 5 * H * 4;
    size_t wf = size_t(H) * 4 * H * 4;
    size_t wb = size_t(T) * B * 4 * H * 4 + size_t(B) * H * 4;
    w = wf > wb ? wf : wb;
  } else if (compute_dtype == CSN_BF16) {
    CSN_TRY(lstm_tc_bytes(T, B, I, H, &r, &w));
  } else {
    CSN_REQUIRE(false, "csn_lstm_layer_bytes: bad compute_dtype %d", compute_dtype);
  }
  if (reserve_bytes) *reserve_bytes = r;
  if (workspace_bytes) *workspace_bytes = w;
  return CSN_OK;
}

extern "C" int csn_lstm_layer_fwd(const void* x, const float* w_ih, const float* w_hh, const float* b_ih,
                                  const float* b_hh, void* h_seq, void* reserve, void* workspace, int T, int B, int I,
                                  int H, int compute_dtype, int training, void* stream) {
  CSN_TRY(check_dims("csn_lstm_layer_fwd", T, B, I, H));
  CSN_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && h_seq && reserve && workspace, "csn_lstm_layer_fwd: null pointer");
  if (compute_dtype == CSN_F32)
    return lstm_layer_fwd_f32((const float*)x, w_ih, w_hh, b_ih, b_hh, (float*)h_seq, (float*)reserve,
                              (float*)workspace, T, B, I, H, as_stream(stream));
  if (compute_dtype == CSN_BF16)
    return lstm_layer_fwd_tc(x, w_ih, w_hh, b_ih, b_hh, h_seq, reserve, workspace, T, B, I, H, training,
                             as_stream(stream));
  CSN_REQUIRE(false, "csn_lstm_layer_fwd: bad compute_dtype %d", compute_dtype);
  return CSN_EINVAL;
}

extern "C" int csn_lstm_layer_bwd(const void* x, const float* w_ih, const float* w_hh, const void* h_seq,
                                  const void* reserve, const float* d_hseq, const float* d_hlast, float* dw_ih,
                                  float* dw_hh, float* db_ih, float* db_hh, float* dx, void* workspace, int T, int B,
                                  int I, int H, int compute_dtype, int accumulate, void* stream) {
  CSN_TRY(check_dims("csn_lstm_layer_bwd", T, B, I, H));
  CSN_REQUIRE(x && w_ih && w_hh && h_seq && reserve && dw_ih && dw_hh && db_ih && db_hh && workspace,
              "csn_lstm_layer_bwd: null pointer");
  CSN_REQUIRE(d_hseq || d_hlast, "csn_lstm_layer_bwd: need d_hseq and/or d_hlast");
  if (compute_dtype == CSN_F32)
    return lstm_layer_bwd_f32((const float*)x, w_ih, w_hh, (const float*)h_seq, (const float*)reserve, d_hseq,
                              d_hlast, dw_ih, dw_hh, db_ih, db_hh, dx, (float*)workspace, T, B, I, H, accumulate,
                              as_stream(stream));
  if (compute_dtype == CSN_BF16)
    return lstm_layer_bwd_tc(x, w_ih, w_hh, h_seq, reserve, d_hseq, d_hlast, dw_ih, dw_hh, db_ih, db_hh, dx,
                             workspace, T, B, I, H, accumulate, as_stream(stream));
  CSN_REQUIRE(false, "csn_lstm_layer_bwd: bad compute_dtype %d", compute_dtype);
  return CSN_EINVAL;
}
