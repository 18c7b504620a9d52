/* csn_b200.h -- C-ABI of libcsn_b200.so: the B200 (sm_100a) kernels behind the EEG distillation train step.
 *
 * The reference (Vi-Sri/CerebralSignalNetworks) has NO native/FFI layer: its seams are Python nn.Module
 * signatures.  This header is therefore the NEW drop-in boundary; each entry point
 * cites the reference lines whose arithmetic it replaces.  The Python mirror of the reference interface
 * (models.lstm.Model, DINOHead, DINOLoss, EEGFilters) binds these symbols through ctypes
 * (cerebralsignalnetworks_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CSN_E* code on failure; csn_last_error() gives a
 *     thread-local message.  No exception crosses the ABI.
 *   - all data pointers are DEVICE pointers owned by the caller (torch allocator), contiguous, 16-byte
 *     aligned, unless the parameter says "host".  The library never allocates device memory in the step
 *     path; scratch is passed in (sizes come from the *_bytes queries).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), and thread-safe across streams.
 *   - NO floating-point atomics: every cross-CTA sum (loss scalars, centre column sums, bias / weight gradients,
 *     split-K products) is folded in a fixed order, so the same inputs give the same bits on every run -- like the
 *     reference's CPU reductions.  Entry points that need scratch for that take a `workspace` pointer.
 *   - tensors are row-major; "time-major" means [T, B, *].
 */
#ifndef CSN_B200_H_
#define CSN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSN_VERSION 200 /* round 2: deterministic reductions (workspace arguments), overlapped recurrence */

enum { CSN_OK = 0, CSN_EINVAL = -1, CSN_ECUDA = -2, CSN_EUNSUPPORTED = -3, CSN_EARCH = -4 };
enum { CSN_F32 = 0, CSN_BF16 = 1 };
enum { CSN_LAYOUT_BCT = 0, CSN_LAYOUT_BTC = 1, CSN_LAYOUT_TBC = 2 };
enum { CSN_ACT_NONE = 0, CSN_ACT_RELU = 1, CSN_ACT_GELU = 2 };
enum { CSN_DINO_SINGLE = 0, CSN_DINO_MULTICROP_REF = 1, CSN_DINO_MULTICROP_CANONICAL = 2 };

int csn_version(void);
const char* csn_last_error(void);
/* number of kernels this library has launched in this process so far (bench.py's gpu_launches evidence) */
unsigned long long csn_launch_count(void);
/* sm count / compute capability of the current device; CSN_EARCH if it is not sm_100. */
int csn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- band-pass ---------------------------------------------------------------------------------------
 * Replaces scipy.signal.filtfilt / sosfilt as called at utils/Utilities.py:421-427 (apply) with the designs of
 * utils/EEGFilters.py:26-39, as a cascade of DF2T biquads along time, one series per thread, tiles staged
 * through shared memory.  x: fp32 [B, C, T] (the stored "channel first" layout, ConvertToPth.py:186).
 * sos: HOST pointer, float64 [n_sections, 6] (scipy SOS rows b0 b1 b2 a0 a1 a2), n_sections <= 8.
 * zero_phase=0 -> causal sosfilt; 1 -> sosfiltfilt (odd extension, padlen = 3*ntaps, steady-state zi).
 * y: out_layout BCT | BTC | TBC, out_dtype f32 | bf16 (the fused transpose+cast that feeds the encoder). */
int csn_sosfilt_f32(const float* x, void* y, const double* sos, int n_sections, int B, int C, int T,
                    int zero_phase, int out_layout, int out_dtype, void* stream);

/* [B, T, C] fp32 (batch_first, what models.lstm.Model.forward receives) -> [T, B, C] f32|bf16 */
int csn_btc_to_tbc(const float* x, void* y, int B, int T, int C, int out_dtype, void* stream);
/* elementwise cast between f32 and bf16 (n elements) */
int csn_cast(const void* x, int x_dtype, void* y, int y_dtype, size_t n, void* stream);

/* ---- dense helpers (fp32 SIMT) -------------------------------------------------------------------------
 * C[M,N] = alpha * op(A) * op(B) + beta * C (+ bias[N]) (+ activation).  op(A) is [M,K]: transA=0 -> A stored
 * [M,K] (lda), transA=1 -> A stored [K,M].  op(B) is [K,N]: transB=0 -> B stored [K,N], transB=1 -> B stored [N,K].
 * Replaces the nn.Linear calls of the restated Model (LSTMDistillRetreival.py:92,108) and DINOHead
 * (LstmDistillation.py:65-99) in fp32 mode. */
int csn_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                 const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int act, void* stream);
/* Same product, plus rowsum[M] (+)= row sums of op(A) from the same pass: with transA=1, A = dY^T this is the bias
 * gradient that rides on the weight-gradient product dW = dY^T X (autograd of nn.Linear). */
int csn_gemm_f32_rowsum(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                        const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int act,
                        float* rowsum, int rowsum_accumulate, void* stream);
/* out[n] (+)= sum_m x[m, n]   (bias gradients, DINO centre column sums); fixed summation order */
int csn_colsum_f32(const float* x, float* out, int M, int N, int ldx, int accumulate, void* stream);
/* y = act(x); dx = dy * act'(x)  (x is the PRE-activation) */
int csn_act_fwd(const float* x, float* y, size_t n, int act, void* stream);
int csn_act_bwd(const float* x, const float* dy, float* dx, size_t n, int act, void* stream);
/* x *= (*scale_dev if scale_dev else 1) * scale_host, in place (autograd glue: upstream scalar gradient) */
int csn_scale_f32(float* x, size_t n, const float* scale_dev, float scale_host, void* stream);
/* row-wise L2 normalise (F.normalize(p=2, eps=1e-12), LstmDistillation.py:97): y = x / max(||x||, eps); inv_norm[M] out */
int csn_l2norm_fwd(const float* x, float* y, float* inv_norm, int M, int N, void* stream);
int csn_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int M, int N, void* stream);

/* nn.BatchNorm1d(N) on x [M, N] (the optional use_bn layers of DINOHead, LstmDistillation.py:72-80).  training != 0: batch
 * statistics over the M rows (biased variance for the normalisation), running_mean / running_var (may be NULL) moved with
 * `momentum` (unbiased variance), save_mean / save_rstd [N] kept for the backward pass; training == 0: the running
 * statistics normalise.  gamma / beta may be NULL (affine=False).  Backward: dx (may be NULL), dgamma, dbeta (may be NULL);
 * fixed-order column reductions. */
int csn_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* y, float* save_mean, float* save_rstd, int M, int N, float eps, float momentum, int training,
                      void* stream);
int csn_batchnorm_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_rstd,
                      float* dx, float* dgamma, float* dbeta, int M, int N, int training, void* stream);
/* weight-norm rows (nn.utils.weight_norm dim=0, LstmDistillation.py:86): w[n,:] = g[n] * v[n,:] / ||v[n,:]||; inv_norm[N] out */
int csn_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm, int N, int K, void* stream);
/* dv[n,:] = g/||v|| * (dw[n,:] - (dw[n,:].v[n,:]) v[n,:]/||v||^2); dg[n] = dw[n,:].v[n,:]/||v|| (dg may be NULL) */
int csn_weight_norm_bwd(const float* v, const float* g, const float* inv_norm, const float* dw, float* dv, float* dg,
                        int N, int K, void* stream);

/* ---- tensor-core GEMM (tcgen05 / TMEM / TMA, bf16 operands, fp32 accumulate) ----------------------------
 * D[M,N] (fp32 or bf16) = op(A) * op(B) (+ bias[N]); same op() convention as csn_gemm_f32 but A and B are bf16.
 * accumulate != 0 adds into D (fp32 only).  split_k > 1 (fp32 only) cuts the contraction into split_k ranges: each range
 * stores its partial product to its own slab of `workspace` (>= split_k * M * N floats, required then, else may be NULL)
 * and a second kernel adds the slabs in range order.
 * Used for the hoisted LSTM input projection, dW / dX of BPTT and the DINO head. */
int csn_gemm_bf16_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                     void* D, int ldd, int d_dtype, const float* bias, int accumulate, int split_k, void* workspace,
                     void* stream);

/* ---- LSTM encoder layer ---------------------------------------------------------------------------------
 * Replaces torch.nn.LSTM inside the (missing) models.lstm.Model -- call sites LstmDistillFromDinoV2Train.py:323,365;
 * analogue LSTMDistillRetreival.py:91,103.  Gate order i,f,g,o, zero initial state, one layer per call.
 * compute_dtype CSN_F32 : SIMT fp32 path (tight-tolerance parity mode, any H).
 * compute_dtype CSN_BF16: H <= 128: persistent tcgen05 recurrence (W_hh resident in tensor memory, gate accumulators
 *                         in TMEM, fused sigmoid/tanh/cell epilogue), hoisted input projection on tcgen05.
 *                         H > 128 : one tcgen05 GEMM per timestep with the LSTM cell fused into its epilogue.
 *                         I and H must be multiples of 8 (16-byte TMA rows).
 * x [T,B,I] in x_dtype (must equal compute_dtype); h_seq [T,B,H] in compute_dtype; weights fp32 masters.
 * reserve: opaque per-step state kept for BPTT (training != 0); sizes from csn_lstm_layer_bytes. */
int csn_lstm_layer_bytes(int T, int B, int I, int H, int compute_dtype, size_t* reserve_bytes, size_t* workspace_bytes);
int csn_lstm_layer_fwd(const void* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       void* h_seq, void* reserve, void* workspace, int T, int B, int I, int H,
                       int compute_dtype, int training, void* stream);
/* Hand-written BPTT.  d_hseq [T,B,H] fp32 (gradient arriving at every timestep, e.g. from the layer above) and/or
 * d_hlast [B,H] fp32 (gradient at t = T-1 only, the encoder head); either may be NULL.  dw_* / db_* are
 * ACCUMULATED into when accumulate != 0, overwritten otherwise.  dx [T,B,I] fp32 may be NULL (first layer). */
int csn_lstm_layer_bwd(const void* x, const float* w_ih, const float* w_hh, const void* h_seq, const void* reserve,
                       const float* d_hseq, const float* d_hlast, float* dw_ih, float* dw_hh, float* db_ih,
                       float* db_hh, float* dx, void* workspace, int T, int B, int I, int H,
                       int compute_dtype, int accumulate, void* stream);

/* ---- DINO cross-entropy, forward + backward + centre statistics in one pass ------------------------------
 * Replaces DINOLoss.forward/update_center: single-view LstmDistillFromDinoV2Train.py:62-105, multi-crop
 * LstmDistillation.py:118-159 (mode MULTICROP_REF reproduces its chunk(1) behaviour: student view 0 skipped,
 * every other view matched against both teacher views, centre statistics kept per batch row), and upstream
 * DINO (dino/main_dino.py:428-481) as MULTICROP_CANONICAL.
 * student [Vs,B,K], teacher [Vt,B,K] fp32; center [K] (center_rows==1) or [B,K] (center_rows==B, the
 * reference's post-first-step shape).  loss: device scalar, overwritten.  d_student [Vs,B,K] = dLoss/dstudent
 * * grad_scale.  batch_center: SINGLE/CANONICAL -> [K] += sum over (view,row) of teacher (caller zeroes; may be NULL);
 * MULTICROP_REF -> [B,K] = sum over views.  The EMA itself is csn_center_ema.
 * workspace: csn_loss_workspace_bytes(B) bytes of scratch for the fixed-order loss fold (contents irrelevant). */
int csn_loss_workspace_bytes(int rows, size_t* bytes);
int csn_dino_loss_fwd_bwd(const float* student, const float* teacher, const float* center, int center_rows,
                          float student_temp, float teacher_temp, float* loss, float* d_student,
                          float* batch_center, int Vs, int Vt, int B, int K, int mode, float grad_scale,
                          void* workspace, void* stream);
/* center = center*momentum + batch_center*scale*(1-momentum)   (scale = 1/(rows*world)) */
int csn_center_ema(float* center, const float* batch_center, size_t n, float momentum, float scale, void* stream);

/* ---- optimiser ---------------------------------------------------------------------------------------------
 * torch.optim.Adam / AdamW semantics (LSTMDistill.py:322, LstmDistillFromDinoV2TrainSpampinato.py:378,
 * LstmDistillation.py:469-471) over a flat fp32 buffer.  grads are multiplied by grad_scale first (1/world for
 * data-parallel averaging).  decoupled=1 -> AdamW.  step is the 1-based step count. */
int csn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int decoupled, int step,
                  float grad_scale, void* stream);

/* CUDA-graph friendly variant: the 1-based step count lives on the DEVICE (`step_counter`, int32, incremented by this
 * call) and the bias corrections are computed there into consts2 (2 floats), so a captured graph replays correctly. */
int csn_adam_step_graph(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int decoupled, int* step_counter,
                        float* consts2, float grad_scale, void* stream);

/* Cap on the CTAs (one per SM) a persistent recurrence launch may occupy; 0 (default) = all SMs.  Lets independent
 * recurrences issued on different streams run side by side -- the crops of different length and the teacher pass of
 * LstmDistillation.py:581-589 -- instead of queueing behind each other.  Process-wide setting. */
int csn_lstm_set_cta_budget(int max_ctas);

/* ---- data-parallel exchange fused into the optimiser (peer-mapped buffers over NVLink) --------------------
 * Replaces DDP's gradient all-reduce (LstmDistillation.py:445) + optimizer.step() + DINOLoss.update_center's all-reduce
 * and EMA (LstmDistillation.py:149-159) by ONE kernel per rank.  grad_ptrs[r] is rank r's flat fp32 buffer
 * [n_param gradients | K centre column sums], mapped into this process (symmetric memory); flag_ptrs[r] is rank r's
 * flag block (>= 16 zero-initialised uint32, same mapping).  Sums run in rank order on every rank (bit-identical
 * replicas); grads are scaled by grad_scale (1/world), the centre by center_scale (1/(B_local*world)).  step_counter
 * is the device-side count of completed steps (advanced by the call); ticket is one zero-initialised device word (CTA arrival
 * counter).
 * csn_dp_wait_done_zero must run on the stream before a rank writes the next step's gradients: it waits until every
 * peer has finished reading the previous ones, then zeroes n floats at zero_ptr (the centre-sum tail).
 * Watchdog: a flag wait longer than CSN_DP_TIMEOUT_S seconds (environment, default 600, 0 = no limit) records
 * (1, rank, flag index, awaited epoch) in a pinned host slot and traps; csn_dp_last_timeout copies the slot to out4
 * (host int[4]; out4[0] == 0: no timeout so far).  Ranks may therefore drift apart by up to that long between steps (a
 * rank-0-only checkpoint or evaluation); put a barrier in front of anything longer. */
int csn_dp_last_timeout(int* out4);
int csn_dp_wait_done_zero(const void* flags_local, int world, const int* step_counter, float* zero_ptr, size_t n,
                          void* stream);
int csn_dp_adam_step_peer(float* params, float* exp_avg, float* exp_avg_sq, size_t n_param,
                          const void* const* grad_ptrs, void* const* flag_ptrs, int world, int rank, float* center,
                          size_t K, float center_momentum, float center_scale, int* step_counter, unsigned* ticket,
                          float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                          float grad_scale, void* stream);

/* Two-shot all-reduce (SUM, in place) of floats [offset, offset + n) of the ranks' peer-mapped buffers, for the large
 * exchanges of the LstmDistillation step (DDP's gradient all-reduce, LstmDistillation.py:445, 22.3 M parameters at cfg3,
 * and the per-row centre statistics of DINOLoss.update_center, :154-156): rank r reduces slice r from every peer in rank
 * order and writes it back, then every rank gathers the reduced slices -- 2 (n-1)/n 4 n bytes per rank over NVLink (the
 * one-shot csn_dp_adam_step_peer reads world * 4 n).  Two kernels, ordered across ranks by epoch flags in peer memory;
 * every replica ends with bit-identical sums.  buf_ptrs[r] / flag_ptrs[r]: rank r's buffer and flag block (>= 64 zeroed
 * uint32) mapped into this process; epoch_counter: device int32 per flag set (advanced by the call); flag_set 0 / 1:
 * two exchanges may be in flight at once on different streams; ticket: 2 zeroed device words per flag set.  Before a
 * rank overwrites its part of the buffer again it must run csn_dp_wait_done (every peer has gathered). */
int csn_dp_allreduce_twoshot(void* const* buf_ptrs, void* const* flag_ptrs, int world, int rank, size_t offset, size_t n,
                             int* epoch_counter, int flag_set, unsigned* ticket, void* stream);
int csn_dp_wait_done(const void* flags_local, int world, const int* epoch_counter, int flag_set, void* stream);

/* ---- the losses LstmDistillFromDinoV2Train.py runs today, forward + backward -------------
 * FeatureDistributionLoss.forward (LstmDistillFromDinoV2Train.py:118-140, live at :371):
 *   loss = alpha * cross_entropy(pred, label) + beta * F.cross_entropy(softmax(teacher / T), softmax(student / T))
 * -- the teacher probabilities are log-softmaxed once more and the student probabilities are the soft target, as
 * written in the reference.  student/teacher [B,K] fp32 (K <= 1024), pred [B,n_classes] fp32 or NULL, label [B] int64.
 * loss: device scalar (overwritten); d_student [B,K], d_pred [B,n_classes] = gradients * grad_scale.
 * workspace (all three losses below): csn_loss_workspace_bytes(B) bytes of scratch for the fixed-order loss fold.
 * CosineSimilarityLoss.forward (LstmDistillFromDinoV2Train.py:36-43): loss = 1 - mean_b cos(student_b, teacher_b),
 * nn.CosineSimilarity semantics (dim 1, each norm clamped at eps). */
int csn_feature_dist_loss_fwd_bwd(const float* student, const float* teacher, const float* pred, const long long* label,
                                  float* loss, float* d_student, float* d_pred, int B, int K, int n_classes,
                                  float temperature, float alpha, float beta, float grad_scale, void* workspace,
                                  void* stream);
int csn_cosine_loss_fwd_bwd(const float* student, const float* teacher, float* loss, float* d_student, int B, int K,
                            float eps, float grad_scale, void* workspace, void* stream);
/* loss_fn_kd of the older training scripts (LSTMDistillRetreival.py:40-70, LstmDistillFromDinoV2TrainSpampinato.py:107-121):
 *   loss = c_kl * sum_b KL(softmax(teacher_b / T) || softmax(student_b / T)) + c_sl1 * sum_{b,k} smooth_l1(student - teacher)
 *        + c_ce * sum_b cross_entropy(student_b, label_b)
 * (smooth_l1 with beta = 1).  The caller folds the reductions and weights into the coefficients, e.g. Retreival:
 * c_kl = w_soft T^2 / B, c_sl1 = w_ce / (B K), c_ce = 0; Spampinato: c_kl = alpha T^2 / (B K), c_ce = (1 - alpha) / B.
 * label [B] int64, may be NULL when c_ce == 0.  K <= 1024.  d_student [B,K] = gradient * grad_scale. */
int csn_kd_loss_fwd_bwd(const float* student, const float* teacher, const long long* label, float* loss, float* d_student,
                        int B, int K, float temperature, float c_kl, float c_sl1, float c_ce, float grad_scale,
                        void* workspace, void* stream);

/* ---- projection head + single-view DINO loss, forward and backward in one kernel ----------------------------------
 * The chain `output = Linear(h_T)` -> DINOLoss.forward -> d_output -> d_h_T of LstmDistillFromDinoV2Train.py:365-375 for the
 * narrow heads (K <= 1024 targets on an I <= 128-wide encoder; csn_head_dino_supported says whether a shape is served):
 *   pre = W h + bias, emb = act(pre) (CSN_ACT_NONE / CSN_ACT_RELU); loss = mean_b CE(softmax((t - c) / tau_t), emb / tau_s)
 *   d_pre [B,K] = dLoss/dpre * grad_scale, d_hlast [B,I] = d_pre W, batch_center [K] += sum_b teacher (caller zeroes it;
 *   NULL skips it: a caller that wants it off the critical path runs csn_colsum_f32 over the teacher on another stream).
 * h_last [B,I] in h_dtype (CSN_F32 / CSN_BF16: the recurrence's own output, no cast pass), W [K,I], bias [K] or NULL,
 * teacher [B,K], center [K].  loss: device scalar (overwritten).  The weight gradient dW = d_pre^T h, db = sum_b d_pre is
 * left to the caller (csn_gemm_f32_rowsum): it is off the critical path of the step.
 * workspace: csn_loss_workspace_bytes(B) bytes of scratch for the fixed-order loss fold. */
int csn_head_dino_supported(int B, int I, int K);
int csn_head_dino_fwd_bwd(const void* h_last, int h_dtype, const float* W, const float* bias, int act, const float* teacher,
                          const float* center, float student_temp, float teacher_temp, float* loss, float* d_hlast,
                          float* d_pre, float* batch_center, int B, int I, int K, float grad_scale, void* workspace,
                          void* stream);

/* ---- GPU-resident dataset batches ------------------------------------------------------
 * src [N, C, T_raw] fp32: the stacked "eeg" tensors of the .pth file ConvertToPth.py:170-201 writes.  For every b:
 * trial idx[b] (int64, negative counts from the end), samples [time_low, time_high), (x - mean) / std (pass 0 / 1 for no
 * normalisation) -- the per-item work of EEGDataset.__getitem__, utils/PerilsEEGDataset.py:541-573.  out is
 * [B, C, T] (CSN_LAYOUT_BCT, what csn_sosfilt_f32 consumes) or [B, T, C] (CSN_LAYOUT_BTC, the DataLoader layout). */
int csn_gather_trials(const float* src, const long long* idx, float* out, int N, int C, int T_raw, int B, int time_low,
                      int time_high, float mean, float std, int out_layout, void* stream);
/* Channel selection of EEGDataset.__getitem__ (`filter_channels`, utils/PerilsEEGDataset.py:554-565), applied once at
 * upload: out [N, n_sel, T] = src[n, channels[s], time_low : time_high]; zscore != 0 additionally normalises every
 * (trial, channel) row over the cropped window with the population standard deviation (`apply_channel_wise_norm`,
 * normlizeEEG :454-461).  channels: device int32 [n_sel], values in [0, C) (validated by the caller). */
int csn_select_crop_zscore(const float* src, const int* channels, float* out, int N, int C, int T_raw, int n_sel,
                           int time_low, int time_high, int zscore, void* stream);
/* The same gather FUSED into the causal band-pass: y = sosfilt(crop / normalise(src[idx[b]])) without writing the gathered
 * batch to HBM (the filter's loads do the gather).  Serves the train step's shape only -- C % 32 == 0, T_raw, time_low and
 * time_high - time_low multiples of 4, 16-byte aligned arrays, out_layout TBC (or fp32 BCT) -- and returns
 * CSN_EUNSUPPORTED otherwise (callers then run csn_gather_trials + csn_sosfilt_f32).  Indices are NOT range-checked on
 * the device (negative values wrap once): validate them on the host. */
int csn_sosfilt_gather_f32(const float* src, const long long* idx, int n_trials, int C, int T_raw, int time_low,
                           int time_high, float mean, float std, void* y, const double* sos, int n_sections, int B,
                           int out_layout, int out_dtype, void* stream);

/* ---- exact top-k retrieval -------------------------------------------------------
 * Replaces faiss.IndexFlatL2(d).add(gallery) / .search(query, k) as called by utils/Utilities.py:45-58 (evaluate) from
 * LstmDistillFromDinoV2Eval.py:333-380.  gallery [nb, d], query [nq, d] fp32 row-major on the device.  metric 0: squared
 * L2 distance, ascending (what IndexFlatL2 returns in D); metric 1: inner product, descending (IndexFlatIP; cosine on
 * L2-normalised rows).  out_dist [nq, k] fp32, out_idx [nq, k] int64 (faiss' idx_t), best first, ties to the lower
 * gallery index; slots beyond nb hold index -1.  1 <= k <= 32.  workspace: csn_topk_workspace_bytes(). */
int csn_topk_workspace_bytes(int nq, int nb, int k, size_t* bytes);
int csn_topk_search(const float* gallery, const float* query, int nb, int nq, int d, int k, int metric,
                    float* out_dist, long long* out_idx, void* workspace, void* stream);
/* Same contract, tensor-core pre-selection for galleries of thousands of rows: a tcgen05 GEMM over bf16
 * hi/lo splits of the rows scores every (query, gallery row) pair, each query's 32 best by that score are re-scored in
 * exact fp32 (direct sum of squared differences / dot product) and the k best of those are returned -- the ranking follows
 * exact fp32 distances, the GEMM only decides which 32 rows get the exact look.  Shapes it does not serve (nb < 2048,
 * d % 8 != 0, k > 16) run csn_topk_search.  workspace_bytes >= csn_topk_tc_workspace_bytes(). */
int csn_topk_tc_workspace_bytes(int nq, int nb, int d, int k, size_t* bytes);
int csn_topk_search_tc(const float* gallery, const float* query, int nb, int nq, int d, int k, int metric,
                       float* out_dist, long long* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* EMA teacher update over flat buffers: dst = momentum * dst + (1 - momentum) * src  (LstmDistillation.py:616-619) */
int csn_ema_update(float* dst, const float* src, size_t n, float momentum, void* stream);

/* The optimiser tail of the LstmDistillation step in ONE sweep over the flat buffers (SURVEY.md K8 + K9 + K10): per-parameter
 * gradient clipping (utils/utils.py:132-141), AdamW / Adam over parameter groups whose lr / weight decay move every
 * iteration (LstmDistillation.py:469-471, :540-544), parameters with a cancelled gradient skipped like torch skips
 * p.grad = None (utils/utils.py:144-149), and the EMA teacher (LstmDistillation.py:616-619).
 * The flat buffers are cut into n_seg parameters [seg_off[i], seg_off[i+1]) (device int64, offsets multiples of 4);
 * n_chunks = sum_i ceil(len_i / 2048).  seg_group[i]: parameter group; seg_active[i]: 0 skips the parameter this step;
 * seg_step[i]: its own step count (device int32, advanced by the call).  hyper (device floats): (lr, weight_decay) per
 * group, then the EMA momentum -- rewritten by the caller between steps, so a captured graph of the step stays valid.
 * grads are scaled by grad_scale (1/world) BEFORE the norm, as DDP averages before clip_gradients runs.  clip = 0: no
 * clipping.  ema_params: the teacher's flat parameters (same layout) or NULL.  sumsq_out: n_seg squared norms or NULL. */
int csn_fused_optim_workspace_bytes(int n_seg, long long n_chunks, size_t* bytes);
int csn_fused_optim_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_params,
                         const long long* seg_off_dev, const int* seg_group_dev, const int* seg_active_dev, int* seg_step_dev,
                         int n_seg, long long n_chunks, const float* hyper_dev, int n_groups, float beta1, float beta2, float eps,
                         int decoupled, float clip, float grad_scale, float* sumsq_out, void* workspace, void* stream);

/* Per-parameter gradient clipping without host syncs (utils/utils.py:132-141).  The flat gradient buffer is cut into
 * n_seg segments [seg_off[i], seg_off[i+1]) (device int64 array of n_seg+1 offsets); n_chunks = sum_i ceil(len_i/2048);
 * sumsq_dev: n_seg + n_chunks floats of scratch; the first n_seg are left holding the squared norms (the norms the
 * reference returns), the rest the per-chunk partial sums they were folded from (chunk order, no atomics). */
int csn_clip_grad_segments(float* grads, const long long* seg_off_dev, int n_seg, long long n_chunks, float* sumsq_dev,
                           float clip, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSN_B200_H_ */
