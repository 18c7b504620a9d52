/* csn_b200_debug.h -- bring-up / self-test hooks of libcsn_b200.so.  NOT part of the product ABI (include/csn_b200.h):
 * nothing on the train step calls these; tests/test_gpu_gemm.py and scripts/umma_bench*.py / prof_lstm_steps.py do. */
#ifndef CSN_B200_DEBUG_H_
#define CSN_B200_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bring-up / self-test hooks (tests only) --------------------------------------------------------------
 * One tcgen05.mma tile D[128,N] = A[128,K] * B[N,K]^T with operands staged in the no-swizzle canonical layouts
 * the recurrence kernel uses; a_mn_major / b_mn_major exercise the MN-major descriptors. */
int csn_dbg_umma_tile(const void* A, const void* B, float* D, int N, int K, int a_mn_major, int b_mn_major, void* stream);
/* a_mn_major = 2 stages A in tensor memory instead (the TS form the recurrence uses for the resident W_hh).
 * csn_dbg_lstm_profile_buffer: device buffer of >= 2*64*8 + 16 int64 that receives clock64 stamps of the first 64 forward
 * ([0,512)) and backward ([512,1024)) recurrence steps of CTA 0 (NULL switches the stamps off). */
int csn_dbg_lstm_profile_buffer(long long* buf);
/* tcgen05.mma issue/completion cost microbenchmark: out[2*rep] = issue cycles, out[2*rep+1] = cycles until commit arrives */
int csn_dbg_umma_bench(long long* out, int M, int N, int n_acc, int a_mode, int reps, void* stream);

/* per-SM global store bandwidth: n_cta CTAs each write bytes_per_cta (multiple of 16 KB) reps times; mode 0 STG.128
 * contiguous, 1 STG.128 in the GEMM-epilogue pattern (4 rows x 128 B at a 2 KB pitch), 2 TMA bulk stores; out[cta] = cycles */
int csn_dbg_store_bw(float* dst, long long* out, size_t bytes_per_cta, int n_cta, int mode, int reps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSN_B200_DEBUG_H_ */
