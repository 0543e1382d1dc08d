/*
 * lstc_vad_b200 — C-ABI of the B200 (sm_100a) hot path of LSTC_VAD.
 *
 * The reference (shengyangsun/LSTC_VAD) has no FFI layer: its hot path is pure PyTorch
 * (`models/*.py` + the loss functions inside `Train/*.py`).  The entry points below are what a
 * binding for that path would call; each one cites the reference expression it replaces.  The Python
 * host code in `lstc_vad_b200/` binds them with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - matrices are row-major; `ld*` are leading dimensions in ELEMENTS;
 *   - bf16 tensors are passed as `void*` (raw __nv_bfloat16), fp32 as `float*`;
 *   - every function returns 0 on success or a LSTC_ERR_* code; `lstc_last_error()` returns a
 *     thread-local human-readable message for the last failure.  Nothing falls back to the CPU.
 *   - dropout masks are a pure function of (seed, offset, element index) — Philox4x32-7, 16 random
 *     bits per element — so backward kernels regenerate the forward mask instead of loading it.
 */
#ifndef LSTC_VAD_B200_H_
#define LSTC_VAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSTC_ABI_VERSION 2

int lstc_abi_version(void);
const char* lstc_last_error(void);
/* Registers (NULL: clears) a device-resident uint64 dropout step counter on the current device.  Every kernel that
 * draws a dropout mask adds *counter to its `offset` argument, so a CUDA graph — which bakes (seed, offset) by value —
 * draws fresh masks on every replay once the graph itself bumps the counter.  Not stream-ordered: call it outside
 * graph capture, before the first launch that should see it. */
int lstc_set_rng_step(const void* counter_dev);

/* ---------------------------------------------------------------------------------------------
 * tcgen05/TMEM + TMA GEMM:  C[M,N] = epilogue(A * B^T), bf16 operands, fp32 accumulate.
 * Replaces nn.Linear + F.relu + nn.Dropout + residual add of
 *   models/MultiHeadAttention.py:97-99 (w_qs/w_ks/w_vs), :123-124 (fc, dropout, += residual),
 *   models/FFN.py:17-19 (w_1, relu, w_2, dropout, += residual), models/Classifier.py:8 and
 *   models/Regressor.py:7 (first Linear+ReLU+Dropout), and their autograd backward (dgrad/wgrad).
 *
 *   a_mn_major = 0 : A is [M,K] (lda >= K);  1 : A is stored [K,M] (lda >= M)
 *   b_mn_major = 0 : B is [N,K] (ldb >= K);  1 : B is stored [K,N] (ldb >= N)
 *   epilogue order: + bias[N] -> relu -> (relu_mask > 0 ? v : 0) -> dropout(p) -> + residual -> store
 *   c_is_f32 = 0 : C bf16, 1 : C fp32.   split_k > 1 (fp32 C only, no epilogue): C is zero-filled and
 *   partial sums are combined with red.add;  accumulate = 1 adds into the existing C.
 *   colsum (fp32 [N], 8-byte aligned, or NULL): colsum[n] += sum_m C[m,n] of the stored bf16 values, accumulated in the
 *   epilogue (the bias gradient of the Linear whose input gradient this product is, models/FFN.py:17 w_1); only where
 *   lstc_gemm_bf16_fuses_colsum() returns 1, and the caller zero-fills it.
 * Requirements: lda, ldb multiples of 8; A, B, C 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int lstc_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                   int64_t M, int64_t N, int64_t K, void* C, int64_t ldc, int c_is_f32, const float* bias,
                   int relu, const void* relu_mask, int64_t ld_mask, const void* residual, int64_t ld_res,
                   float dropout_p, uint64_t seed, uint64_t offset, int split_k, int accumulate, float* colsum,
                   void* stream);
/* 1 when the product above can fuse the column sums for this output (bf16 C, ldc % 8 == 0, M >= 32, N >= 64: the
 * shared-memory + TMA-store epilogue), else 0 (run lstc_colsum on C instead). */
int lstc_gemm_bf16_fuses_colsum(int64_t M, int64_t N, int64_t ldc, int c_is_f32);

/* ---------------------------------------------------------------------------------------------
 * Fused short-sequence attention, one CTA per (window-group, head).
 * Replaces models/MultiHeadAttention.py:100-122:
 *   S = (q / sqrt(dk)) k^T ; S[1:,1:] += rel-pos bias ; P = dropout(softmax(S)) ; O = P v ; merge heads.
 *   qkv  : bf16 [W*L, ld]; row w*L+i holds q (cols h*dk..), k (cols H*dk + h*dk..), v (cols 2*H*dk + h*dk..)
 *   bias : fp32 dense [H, L, L] (row/col 0 zero) or NULL
 *   out  : bf16 [W*L, ld_out], head h in cols h*dk.. (already "transpose(1,2).contiguous().view")
 *   probs: optional fp32 [W, H, L, L] post-dropout probabilities (return_attn, :127-132), or NULL
 * Supported: 1 <= L <= 96, dk in {64,128,256}, d_v == d_k.
 * ------------------------------------------------------------------------------------------- */
int lstc_attn_fwd(const void* qkv, int64_t ld, int64_t W, int L, int H, int dk, const float* bias, float scale,
                  float dropout_p, uint64_t seed, uint64_t offset, void* out, int64_t ld_out, float* probs,
                  void* stream);

/* Backward of the above (autograd of models/MultiHeadAttention.py:100-122).
 *   dout : bf16 [W*L, ld_dout] ; dqkv : bf16 [W*L, ld_dqkv] same column layout as qkv
 *   dbias: fp32 [H, L, L], zero-filled by this call then accumulated over windows, or NULL */
int lstc_attn_bwd(const void* qkv, int64_t ld, const void* dout, int64_t ld_dout, int64_t W, int L, int H, int dk,
                  const float* bias, float scale, float dropout_p, uint64_t seed, uint64_t offset, void* dqkv,
                  int64_t ld_dqkv, float* dbias, void* stream);

/* CLS-query attention of the LAST encoder layer (opt-in fast path, `Encoder.forward_cls`): every reference caller
 * consumes only `feats[:, 0, :]` (Train/temporal_transformer_shanghaitech.py:123), so in the last layer only the
 * CLS query attends: o = softmax(q_cls K^T * scale) V per (window, head); the CLS row of the rel-pos bias is zero
 * (models/MultiHeadAttention.py:111).  Exact dead-work elimination of models/MultiHeadAttention.py:103-122.
 *   q   : bf16 [W, ld_q] (head h at cols h*dk) ; k, v : bf16 [W*L, ld_kv] ; out : bf16 [W, ld_out]
 * The dropout mask is the row-0 slice of the mask lstc_attn_fwd draws for the same (seed, offset).
 * Backward: dout bf16 [W, ld_do] -> dq bf16 [W, ld_dq], dk / dv bf16 [W*L, ld_dkv] (every key / value row). */
int lstc_attn_cls_fwd(const void* q, int64_t ld_q, const void* k, const void* v, int64_t ld_kv, int64_t W, int L,
                      int H, int dk, float scale, float dropout_p, uint64_t seed, uint64_t offset, void* out,
                      int64_t ld_out, void* stream);
int lstc_attn_cls_bwd(const void* q, int64_t ld_q, const void* k, const void* v, int64_t ld_kv, const void* dout,
                      int64_t ld_do, int64_t W, int L, int H, int dk, float scale, float dropout_p, uint64_t seed,
                      uint64_t offset, void* dq, int64_t ld_dq, void* dk_out, void* dv_out, int64_t ld_dkv,
                      void* stream);
/* dst[r, c] += src[r, c] on bf16 row-strided views (cols, lds multiples of 8) */
int lstc_add_rows_bf16(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows, int64_t cols,
                       void* stream);

/* Relative-position bias: dense[h,i,j] = (i>0 && j>0) ? table[index[(i-1)*index_ld + (j-1)], h] : 0
 * (models/MultiHeadAttention.py:107-117; index is the int64 `relative_position_index` buffer, table is
 * `relative_position_bias_table` [T,H]).  The scatter is its transpose: dtable is zero-filled then
 * dtable[index[..], h] += ddense[h,i,j]. */
int lstc_relbias_gather(const float* table, const int64_t* index, int64_t index_ld, int L, int H, int64_t T,
                        float* dense, void* stream);
int lstc_relbias_scatter(const float* ddense, const int64_t* index, int64_t index_ld, int L, int H, int64_t T,
                         float* dtable, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (nn.LayerNorm(D, eps=1e-6): models/Encoder.py:31,
 * models/MultiHeadAttention.py:47,125-126, models/FFN.py:10,20-21).  fp32 statistics.
 *   x : bf16 or fp32 [rows, D] ; y : bf16 or fp32 ; mean/rstd : fp32 [rows] saved for backward.
 * Backward: dx (bf16 or fp32) and, when drop_p > 0 and dx_drop != NULL, a second bf16 output
 * dx_drop = dropout_mask(seed,offset) * dx / (1-p) (the gradient entering the preceding
 * Linear whose output was dropped out).  dgamma/dbeta fp32 [D] are overwritten.  dxsum (nullable) fp32 [D]
 * receives the column sums of dx_drop (of dx when there is no dropout) = the bias gradient of that Linear.
 *   workspace: fp32, at least lstc_layernorm_bwd_workspace(rows, D) bytes.
 * Requirements: D % 8 == 0, D <= 8192.
 * ------------------------------------------------------------------------------------------- */
int lstc_layernorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, void* y, int y_is_f32,
                       float* mean, float* rstd, int64_t rows, int64_t D, float eps, void* stream);
int64_t lstc_layernorm_bwd_workspace(int64_t rows, int64_t D);
int lstc_layernorm_bwd(const void* dy, int dy_is_f32, const void* x, int x_is_f32, const float* gamma,
                       const float* mean, const float* rstd, void* dx, int dx_is_f32, void* dx_drop,
                       float drop_p, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta, float* dxsum,
                       void* workspace, int64_t rows, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CLS-token prepend (+ optional learned absolute position encoding and its dropout):
 * models/Encoder.py:51-59 and models/PatchEmbedding.py:12-19.
 *   x   : fp32 or bf16 [W, L0, D]
 *   cls : fp32 [D] learned token, or NULL => CLS = mean over the L0 tokens
 *   pos : fp32 [>= L0+1, D] or NULL
 *   out : bf16 [W, L0+1, D]
 * Backward: g bf16 [W, L0+1, D] -> dx fp32 [W, L0, D] (nullable), dcls fp32 [D] (nullable, learned
 * token only), dpos fp32 [L0+1, D] (nullable); dcls / dpos are zero-filled then accumulated.
 * ------------------------------------------------------------------------------------------- */
int lstc_cls_prepend_fwd(const void* x, int x_is_f32, const float* cls, const float* pos, float drop_p,
                         uint64_t seed, uint64_t offset, void* out, int64_t W, int64_t L0, int64_t D,
                         void* stream);
int lstc_cls_prepend_bwd(const void* g, int cls_learned, float drop_p, uint64_t seed, uint64_t offset, float* dx,
                         float* dcls, float* dpos, int64_t W, int64_t L0, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heads: layers 2 and 3 of Classifier / Regressor (models/Classifier.py:9-10, models/Regressor.py:8-9):
 *   h2 = dropout(h1 W2^T + b2) ; z = h2 W3^T + b3 ; out = softmax(z) (C=2) | sigmoid(z) (C=1)
 *   h1 : bf16 [n, K1] (output of the first Linear+ReLU+Dropout, computed by lstc_gemm_bf16)
 *   W2 : fp32 [32, K1], b2 [32], W3 fp32 [C, 32], b3 [C] ; h2 : bf16 [n, 32] (post-dropout, saved)
 *   out: fp32 [n, C]
 * Backward: dout fp32 [n,C], out fp32 [n,C] -> dh2 bf16 [n,32] (gradient w.r.t. the pre-dropout layer-2
 * output), dW3 fp32 [C,32] and db3 fp32 [C] (overwritten).
 * Requirements: K1 % 8 == 0, K1 <= 1024, C in {1,2}.
 * ------------------------------------------------------------------------------------------- */
int lstc_head_tail_fwd(const void* h1, int64_t n, int K1, const float* W2, const float* b2, const float* W3,
                       const float* b3, int C, int sigmoid, float drop_p, uint64_t seed, uint64_t offset,
                       void* h2, float* out, void* stream);
int lstc_head_tail_bwd(const float* dout, const float* out, const void* h2, const float* W3, int64_t n, int C,
                       int sigmoid, float drop_p, uint64_t seed, uint64_t offset, void* dh2, float* dW3,
                       float* db3, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses (all fp32, single launch each, forward value + unit-upstream gradient in one pass).
 *
 * lstc_mil_loss: get_MIL_loss, LTN form Train/temporal_transformer_shanghaitech.py:25-36 and STN form
 * Train/spatio_transformer_shanghaitech.py:21-32.
 *   scores : element i at scores[i*stride]; n_scores = 2*B*P*T ; bag b = scores [b*P*T, (b+1)*P*T)
 *   part score = mean over T consecutive scores; bag score = mean of the top-k part scores (k=1: max,
 *   lowest index wins ties); err = sum_{i<B, j<B} relu(1 - abn_j + nor_i) / B^2 ;
 *   spar = mean(scores[spar_start:]) ; loss = err + lambda1 * spar.
 *   out3   : {loss, err, spar}; top_idx : int32 [2B, k] selected part indices (nullable);
 *   dscores: d loss / d scores, written with the same stride (other elements untouched), nullable.
 *
 * lstc_soft_ce_loss: get_CE_loss = F.cross_entropy(probs, soft_labels), i.e. the mean over rows of
 * -sum_c lab[r,c] * log_softmax(probs[r,:])[c] — taken on ALREADY-softmaxed classifier outputs
 * (Train/temporal_transformer_shanghaitech.py:21-23,130).
 *
 * lstc_bce_loss: get_BCE_loss, Train/spatio_transformer_MIL_CE.py:23-26:
 *   mean(-w_n * lab[:,0] * log(1 - o + 1e-8) - w_a * lab[:,1] * log(o + 1e-8)), o = mean over T scores.
 *
 * lstc_threshold_labels: where(s > thr, s, 0), Train/pseudo_labels_generator_temporal.py:103-104.
 * ------------------------------------------------------------------------------------------- */
int lstc_mil_loss(const float* scores, int64_t stride, int B, int P, int T, int topk, float lambda1,
                  int64_t spar_start, float* out3, int32_t* top_idx, float* dscores, void* stream);
int lstc_soft_ce_loss(const float* probs, const float* labels, int64_t n, int C, float* out1, float* dprobs,
                      void* stream);
int lstc_bce_loss(const float* scores, const float* labels, int64_t n_parts, int T, float w_normal,
                  float w_abnormal, float* out1, float* dscores, void* stream);
int lstc_threshold_labels(const float* scores, float thr, float* out, int64_t n, void* stream);
/* Frame-level ROC-AUC (utils/eval_utils.py:21-24: sklearn roc_curve + auc over per-frame scores) computed from
 * per-WINDOW scores: every frame of a window carries the window's score
 * (Test/evaluation_shanghaitech_ubnormal.py:92-94), so window i contributes pos_w[i] positive and neg_w[i] negative
 * frames at threshold scores[i]; equal scores form one ROC step, exactly like sklearn's distinct-threshold curve.
 * out3 (device, fp64) = {auc, total positive frames, total negative frames}.  n <= 16384 windows. */
int lstc_weighted_auc(const float* scores, const float* pos_w, const float* neg_w, int64_t n, double* out3,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small utilities used between the kernels above.
 * ------------------------------------------------------------------------------------------- */
int lstc_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int lstc_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* dst[c, r] = bf16(src[r, c]) for an fp32 [rows, cols] matrix (pitches in elements): the transposed bf16 weight copy
 * that the input-gradient products read as a K-major operand (dX = dY W of nn.Linear's backward; the reference runs it
 * inside autograd for models/FFN.py:17-19, models/MultiHeadAttention.py:97-99,123). */
int lstc_cast_f32_to_bf16_transposed(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                                     void* stream);
/* out[c] = sum_r x[r, c] for bf16 x [rows, cols] (bias gradients); deterministic two-stage reduce.
 * workspace >= lstc_colsum_workspace(rows, cols) bytes. */
int64_t lstc_colsum_workspace(int64_t rows, int64_t cols);
int lstc_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* workspace,
                     void* stream);
/* y = dropout_mask(seed, offset) * x / (1-p), bf16 [rows, cols] contiguous (cols % 8 == 0) */
int lstc_dropout_apply_bf16(const void* x, void* y, int64_t rows, int64_t cols, float p, uint64_t seed,
                            uint64_t offset, void* stream);
/* mask[r, c] = 1 if kept else 0 (uint8) — the exact mask the fused kernels use; test/debug aid */
int lstc_dropout_mask(uint8_t* mask, int64_t rows, int64_t cols, float p, uint64_t seed, uint64_t offset,
                      void* stream);
/* dst[i] = src[i] * (*scalar_dev) */
int lstc_scale_by_device_scalar(const float* src, const float* scalar_dev, float* dst, int64_t n, void* stream);
/* dst[i, :] = src[idx[i], :] for rows of row_bytes bytes (multiple of 16; idx int64 on the device): the clip
 * selection `feat[chosen]` of the training-window sampler (utils/load_dataset.py:56-88) on a corpus resident in HBM. */
int lstc_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int64_t m, void* dst, void* stream);
/* Temporal pooling of a video into n_bins pseudo-clips + optional L2 normalisation of every token
 * (Test/evaluation_UCF.py:52-77: r = linspace(0, n_clips, 33); bin b = mean(feats[r[b]:r[b+1]]), or the single clip
 * feats[r[b]] when the bin is empty; F.normalize(p=2, dim=-1)).
 *   feats fp32 [n_clips, n_patch, D] ; bounds int32 [n_bins+1] (device) ; out fp32 [n_bins, n_patch, D] */
int lstc_segment_mean(const float* feats, const int32_t* bounds, int n_bins, int64_t n_clips, int n_patch, int D,
                      int l2norm, float* out, void* stream);
/* Fused multi-tensor-free Adagrad step on one flat fp32 tensor (torch.optim.Adagrad semantics,
 * Train/temporal_transformer_shanghaitech.py:83-85,142): g += wd * p ; state += g*g ;
 * p -= lr * g / (sqrt(state) + eps).
 * grad_scale_dev (nullable) is a device scalar multiplied into grad_scale — the clip coefficient of
 * torch.nn.utils.clip_grad_norm_ (Train/temporal_transformer_shanghaitech.py:139-141), produced without a host
 * sync by lstc_sumsq_accumulate (accum += sum x^2, accum zeroed by the caller) and lstc_clip_coef
 * (coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))). */
int lstc_sumsq_accumulate(const float* x, int64_t n, float* accum, void* stream);
int lstc_clip_coef(const float* sumsq, float max_norm, float* coef, void* stream);
int lstc_adagrad_step(float* param, const float* grad, float* state_sum, int64_t n, float lr, float weight_decay,
                      float eps, float grad_scale, const float* grad_scale_dev, void* stream);

/* ---- multi-tensor launches: ONE kernel per call walks a table of up to 48 tensors (longer tables are split); the
 * pointer / size tables are HOST arrays (copied into the kernel parameters), every tensor pointer is a device pointer.
 *
 * lstc_multi_pack_bf16 / lstc_multi_unpack_bf16: the fp32 gradients of one data-parallel bucket <-> one flat bf16
 * buffer, the payload of the NCCL all-reduce that replaces nn.DataParallel's reduce_add_coalesced
 * (Train/temporal_transformer_shanghaitech.py:76-78).  dst / src pointers of the flat side are (flat + offset), 16-byte
 * aligned.
 * lstc_multi_adagrad: lstc_adagrad_step for every parameter in one launch; lrs[i] per tensor (the scripts use two
 * parameter groups, :83-85); grads are fp32 tensors or (grad_is_bf16 != 0) slices of a reduced bf16 bucket. */
int lstc_multi_pack_bf16(const void* const* src_f32, const int64_t* numel, int n, void* const* dst_bf16, void* stream);
int lstc_multi_unpack_bf16(const void* const* src_bf16, const int64_t* numel, int n, void* const* dst_f32,
                           void* stream);
int lstc_multi_adagrad(void* const* params, const void* const* grads, int grad_is_bf16, void* const* states,
                       const int64_t* numel, const float* lrs, int n, float weight_decay, float eps, float grad_scale,
                       const float* grad_scale_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSTC_VAD_B200_H_ */
