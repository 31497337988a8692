/*
 * dmt_b200.h — C-ABI of the B200-native DMTCDR hot path (libdmt_b200.so, sm_100a only).
 *
 * The reference (diaoenmao/Decentralized-Multi-Target-Cross-Domain-Recommendation-...) is pure Python on
 * PyTorch; it has no FFI of its own. These entry points are what a binding for its hot path attaches to
 * (INTEGRATION.md shows the ctypes stub). Each one cites the reference code it replaces as
 * `src/<file>:<lines>`.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - every function returns int: 0 = ok, < 0 = DMT_E_* below, > 0 = a cudaError_t value.
 *     Nothing throws or aborts. dmt_last_error() returns a static message for the calling thread.
 *   - all kernels are asynchronous on the given stream and re-entrant across streams; the only state is
 *     inside the opaque handles (dmt_plan_t, dmt_org_t).
 *   - floating point is fp32 (as the reference); indices are int32 after ingestion (the reference's int64
 *     indices are narrowed once at the API edge); pointer-offset arrays (indptr) are int32, nnz < 2^31.
 *   - "rows" are the aligned entity (users in data_mode 'user', items in 'item'); "columns" the other one.
 */
#ifndef DMT_B200_H
#define DMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMT_E_ARG (-1)      /* invalid argument */
#define DMT_E_STATE (-2)    /* handle used in the wrong state */
#define DMT_E_NOMEM (-3)    /* host allocation failed */
#define DMT_E_ARCH (-4)     /* device is not sm_100 */

#define DMT_LOSS_MSE 0 /* explicit: (o-y)^2            src/models/utils.py:11 */
#define DMT_LOSS_BCE 1 /* implicit: BCE with logits     src/models/utils.py:9  */

const char* dmt_last_error(void);
int dmt_version(void);
/* 0 when the current device is a compute-capability 10.x part, DMT_E_ARCH otherwise. */
int dmt_check_device(void);

/* ------------------------------------------------------------------ MTAL coordinator (src/assist.py) */

/* Pseudo-residual r = -dL/dF of the summed loss (src/assist.py:45-58):
 * MSE: 2(y-F); BCE: y - sigmoid(F); clamped to [-clamp, clamp] when clamp > 0 (src/assist.py:51-56). */
int dmt_residual(const float* F, const float* y, float* r, int64_t n, int loss_kind, float clamp, void* stream);

/* One streaming pass of Assist.update's apply step for ALL owners (src/assist.py:131-176 +
 * src/models/assist.py:36-37):  F_new[p] = F_old[p] + rate_col[col[p]] * sum_j S[owner[col[p]]][j] * O'[j][p]
 * with O'[j][p] = O[j][p] if p < match_end[owner] else O[owner][p]   (partial alignment, src/assist.py:95-103).
 * O is org-major [K][nnz]; S is the row-wise softmax of every owner's assistance weights, [K][K] row-major;
 * rate_col[c] = assist_rate of the owner of column c at c's local index.
 * org_row (may be NULL = identity): row of O that holds organization j (rank-blocked layouts of the sharded exchange).
 * S_cold (may be NULL): cold start (src/assist.py:109-117,150-157; src/models/assist.py:28-34) - wherever organization
 * 0's output is NaN (aligned rows organization 0 never saw) the sum runs over j >= 1 with S_cold[owner][j] =
 * softmax(w_owner[1:])[j-1], S_cold[owner][0] = 0. */
int dmt_assist_combine(const float* F_old, const float* O, const int32_t* col, const int32_t* owner,
                       const float* rate_col, const float* S, const int64_t* match_end, float* F_new,
                       int64_t nnz, int K, const int32_t* org_row, const float* S_cold, void* stream);

/* Gather one owner's view for the L-BFGS fit (src/assist.py:93-117): for e in [0,n): p = pos[e];
 * h[e]=F_old[p], t[e]=y[p], V[j][e] = (rank[e] < n_match ? O[j][p] : O[owner][p]). pos is the owner's entries in
 * column-sorted order, rank[e] the entry's rank in the original (row-major) order. V is [K][n]. */
int dmt_assist_gather_view(const float* F_old, const float* y, const float* O, const int32_t* pos,
                           const int32_t* rank, int64_t nnz, int64_t n, int K, int owner, int64_t n_match, float* h,
                           float* t, float* V, const int32_t* org_row, void* stream);

/* Fused loss + gradient of models.Assist for one owner (src/models/assist.py:25-40, closure src/assist.py:121-126).
 * Entries are column-sorted; seg_off[n_rate+1] delimits the entries of each owned column (local index = segment).
 * loss = mean_e l(h_e + rate[c(e)] * sum_j softmax(w)_j V[j][e], t_e);
 * out[0]=loss, d_rate[n_rate], d_w[K] (gradient w.r.t. the UN-normalised weights w). scratch >= dmt_assist_scratch_floats(K). */
int64_t dmt_assist_scratch_floats(int K);
int dmt_assist_loss_grad(const float* h, const float* t, const float* V, const int32_t* seg_off, const float* rate,
                         const float* w, int64_t n, int n_rate, int K, int loss_kind, float* out_loss, float* d_rate,
                         float* d_w, float* scratch, void* stream);

/* The whole L-BFGS fit of one owner's assisted learning rates / assistance weights (src/assist.py:118-129:
 * `steps` x optimizer.step(closure) of torch.optim.LBFGS(lr, max_iter, history, no line search), src/utils.py:255-256)
 * enqueued on `stream` without any host round trip: the closure is dmt_assist_loss_grad, the two-loop recursion, the
 * step and torch's stopping rules (tolerance_grad 1e-7, tolerance_change 1e-9, max_eval = 5/4 max_iter) run in one
 * single-block kernel per inner iteration; evaluations after a stopping rule skip themselves on a device flag.
 * params = [rate (n_rate) | w (K)], initialised by the caller and updated in place; only the slices selected by
 * ar_optim / aw_optim move. work >= dmt_assist_fit_work_floats(n_rate, K, history), scratch as dmt_assist_loss_grad. */
int64_t dmt_assist_fit_work_floats(int n_rate, int K, int history);
int dmt_assist_fit(const float* h, const float* t, const float* V, const int32_t* seg_off, int64_t n, int n_rate, int K,
                   int loss_kind, float* params, int ar_optim, int aw_optim, float lr, int steps, int max_iter,
                   int history, float* work, float* scratch, void* stream);

/* models.Assist as a differentiable module (src/models/assist.py:25-40; the closure at src/assist.py:121-126 calls
 * loss.backward() through it). out is the [n x K] stack of the organizations' outputs addressed as
 * out[e*stride_e + j*stride_j]; target[e] = history[e] + rate[idx[e]] * sum_j softmax(w)_j out[e][j]; entries whose slot 0
 * is NaN (cold start) use sum_{j>=1} softmax(w[1:])_j out[e][j] instead. q[e] (may be NULL) keeps the weighted sum.
 * bwd: given delta[e] = dLoss/dtarget[e]: d_rate[c] = sum_{e: idx[e]=c} delta[e] q[e] (perm / seg_* = dmt_sort_segments
 * of idx; columns without entries get 0), d_w = gradient w.r.t. the UN-normalised weights through both softmaxes.
 * d_rate or d_w may be NULL. scratch >= dmt_assist_rows_scratch_floats(K). Deterministic (no atomics). */
int dmt_assist_rows_fwd(const float* out, int64_t stride_e, int64_t stride_j, const float* history, const int32_t* idx,
                        const float* rate, const float* w, int64_t n, int K, float* target, float* q, void* stream);
int64_t dmt_assist_rows_scratch_floats(int K);
int dmt_assist_rows_bwd(const float* out, int64_t stride_e, int64_t stride_j, const int32_t* idx, const float* rate,
                        const float* w, const float* q, const float* delta, const int32_t* perm, const int32_t* seg_key,
                        const int32_t* seg_off, const int32_t* n_seg, int64_t n, int K, int n_rate, float* d_rate,
                        float* d_w, float* scratch, void* stream);

/* Round-0 predictor models.Base (src/models/base.py:22-60). fit: base[idx] += rating, count[idx] += 1.
 * predict (explicit): base/(count+1e-10), unseen columns -> mean of the seen means (fill computed here);
 * predict (implicit): base / scalar_count, scalar_count = sum over fitted batches of #distinct rows (base.py:35-37). */
int dmt_base_fit(const int32_t* idx, const float* rating, int64_t n, float* base, float* count, void* stream);
int dmt_base_predict(const float* base, const float* count, int32_t n_cols, const int32_t* target_idx, int64_t n,
                     int implicit, float implicit_count, float* out, float* scratch /* >= 2 floats */, void* stream);

/* ------------------------------------------------------------------ sort-by-index + segmented reduction */

/* Stable sort of keys (values in [0, key_bound)) -> perm (source positions in sorted order), unique keys,
 * segment offsets and the segment count (device int32). Replaces torch.sort/unique_consecutive at
 * src/models/ae.py:103-104,137-139 and the implicit sort inside embedding_dense_backward.
 * temp: device scratch of dmt_sort_segments_temp_bytes(n) bytes. seg_key/seg_off need n and n+1 slots. */
int64_t dmt_sort_segments_temp_bytes(int64_t n);
int dmt_sort_segments(const int32_t* keys, int64_t n, int32_t key_bound, int32_t* perm, int32_t* seg_key,
                      int32_t* seg_off, int32_t* n_seg, void* temp, int64_t temp_bytes, void* stream);

/* grad[seg_key[s]][:] = sum_{e in segment s} coef[perm[e]] * src[src_row[perm[e]]][:]   (width floats per row),
 * bias_grad[seg_key[s]] = sum coef (may be NULL). One warp per segment, no atomics: the dense-gradient form of
 * autograd's embedding / index backward (src/models/mf.py:37,44; src/models/ae.py:102,135). Rows of grad that
 * own no segment are NOT touched (zero them first). n_seg is read from device memory. */
int dmt_segment_reduce_rows(const int32_t* perm, const int32_t* seg_key, const int32_t* seg_off, const int32_t* n_seg,
                            int64_t n_seg_max, const float* coef, const int32_t* src_row, const float* src, int width,
                            float* grad, float* bias_grad, void* stream);

/* ------------------------------------------------------------------ optimizer (src/utils.py:253-254, src/organization.py:161-162) */

/* sum of squares of g[0..n) -> out[0] (deterministic two-stage); scratch >= dmt_sqnorm_scratch_floats() floats. */
int64_t dmt_sqnorm_scratch_floats(void);
int dmt_sqnorm(const float* g, int64_t n, float* out, float* scratch, void* stream);
/* clip_grad_norm_(params, max_norm) fused into torch.optim.Adam(lr, betas, eps, weight_decay) (dense, L2):
 * coef = min(1, max_norm/(sqrt(*sqnorm)+1e-6)); g' = coef*g + wd*w; m,v update; w -= lr/(1-b1^t) * m/(sqrt(v)/sqrt(1-b2^t)+eps).
 * sqnorm (device) may be NULL = no clipping. step is the 1-based step count t. scratch >= 4 floats (device). */
int dmt_adam_clip_step(float* w, const float* g, float* m, float* v, int64_t n, const float* sqnorm, float max_norm,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       float* scratch, void* stream);

/* ------------------------------------------------------------------ MF / GMF (src/models/mf.py, GMF branch of nmf.py) */

/* MF / GMF forward + loss + per-rating dloss/dpred in one pass (src/models/mf.py:36-48,79-92; nmf.py:122-146):
 *   q_d  = (Wu[u,d]+bu[u]) * (Wi[i,d]+bi[i] [+ pu[e,d]])  [+ (Wi[i,d]+bi[i]) * pi[e,d]]
 *   pred = sum_d colscale[d] * q_d + bias [+ add[e]]        (colscale NULL = all ones, bias NULL = 0)
 * pu/pi: optional per-rating side-information projections [n x H] (outputs of the user_profile / item_attr Linear).
 * NCF: colscale = the GMF part of the affine weight, add[e] = the MLP-tower part of the logit, q_out [n x H] (may be
 * NULL) receives q for the affine-weight gradient. H in {128,256,384,512}. dpred may be NULL (eval).
 * sums[0] = sum of per-rating losses (mean = /n), sums[1] = sum of dloss/dpred. scratch >= dmt_mf_scratch_floats(). */
int64_t dmt_mf_scratch_floats(void);
int dmt_mf_fwd(const int32_t* user, const int32_t* item, const float* rating, int64_t n, const float* Wu,
               const float* Wi, const float* bu, const float* bi, const float* bias, const float* pu, const float* pi,
               const float* colscale, const float* add, int H, int loss_kind, float* pred, float* dpred, float* q_out,
               float* sums, float* scratch, void* stream);
/* Dense gradient of one embedding table and its bias table from the ratings sorted by that table's index
 * (perm/seg_* from dmt_sort_segments on user[] or item[]):
 *   dW[r] = colscale * sum_{e: key(e)=r} g_e * (W_other[other[e]] + b_other[other[e]] [+ p_side[e]]),
 *   db[r] = row-sum of dW[r],  g_e = dpred[e]*grad_scale. Rows without ratings are not touched (zero dW/db first). */
int dmt_mf_bwd_table(const int32_t* other, const float* W_other, const float* b_other, const float* p_side, int H,
                     const float* dpred, float grad_scale, const int32_t* perm, const int32_t* seg_key,
                     const int32_t* seg_off, const int32_t* n_seg, int64_t n_seg_max, const float* colscale, float* dW,
                     float* db, void* stream);
/* d_p[e][:] = colscale * g_e * (W[idx[e]] + b[idx[e]]): gradient w.r.t. a side-information projection. */
int dmt_mf_bwd_side(const int32_t* idx, int64_t n, const float* W, const float* b, int H, const float* dpred,
                    float grad_scale, const float* colscale, float* d_p, void* stream);

/* ------------------------------------------------------------------ MLP / NCF tower pieces (src/models/mlp.py, nmf.py) */

/* out[e][col_off + d] = W[idx[e]][d] + b[idx[e]]: embedding with its bias broadcast-added, written into one slice of
 * the concatenated tower input (row stride ld floats; ld and col_off multiples of 4). */
int dmt_embed_fwd(const int32_t* idx, int64_t n, const float* W, const float* b, int H, float* out, int ld,
                  int col_off, void* stream);
/* Dense gradient of that embedding: dW[r] = sum_{e in segment r} dOut[perm[e]][col_off : col_off+H], db[r] = row-sum. */
int dmt_embed_bwd(const float* dOut, int ld, int col_off, int H, const int32_t* perm, const int32_t* seg_key,
                  const int32_t* seg_off, const int32_t* n_seg, int64_t n_seg_max, float* dW, float* db, void* stream);
/* out[d] = sum_e g[e]*scale*Q[e][d], d < width (row stride ld): gradient of a one-row weight (the affine layer) and
 * of the GMF column scale. Deterministic two-stage. scratch >= dmt_weighted_colsum_scratch_floats(width). */
int64_t dmt_weighted_colsum_scratch_floats(int width);
int dmt_weighted_colsum(const float* g, float scale, const float* Q, int64_t n, int width, int ld, float* out,
                        float* scratch, void* stream);
/* loss_fn on a finished prediction vector (src/models/utils.py:7-14): sums[0] = sum of losses, sums[1] = sum of
 * dloss/dpred, dpred[e] (may be NULL) = dloss/dpred. scratch >= dmt_mf_scratch_floats(). */
int dmt_loss_fwd(const float* pred, const float* y, int64_t n, int loss_kind, float* dpred, float* sums,
                 float* scratch, void* stream);

/* ------------------------------------------------------------------ AAE (src/models/ae.py:98-157) */

/* Dense layer forward  Y = act(X W^T + b) [* keep*keep_scale]: X [m x k], W [n x k] (nn.Linear layout), act: 0 none,
 * 1 tanh, 2 relu. keep (uint8 [m x n], 0/1) may be NULL; when given Y_pre receives the pre-dropout activation
 * and Y the dropped one (src/models/ae.py:14-19,44-45,132). */
int dmt_dense_fwd(const float* X, const float* W, const float* b, float* Y, float* Y_pre, const uint8_t* keep,
                  float keep_scale, int m, int n, int k, int act, void* stream);
/* dX = (dY W) * dact(A_prev) [* keep*keep_scale]:  dY [m x n], W [n x k], A_prev [m x k] = activation that fed
 * this layer (act_prev 1: 1-a^2; 2: a>0; 0 or A_prev NULL: none). */
int dmt_dense_bwd_x(const float* dY, const float* W, const float* A_prev, const uint8_t* keep, float keep_scale,
                    float* dX, int m, int n, int k, int act_prev, void* stream);
/* dW = dY^T X  [n x k], db = column sums of dY [n] (db may be NULL). */
int dmt_dense_bwd_w(const float* dY, const float* X, float* dW, float* db, int m, int n, int k, void* stream);

/* Encoder first layer as CSR SpMM (src/models/ae.py:101-110): for batch row j (row id rows[j]):
 * A1[j] = tanh(b1 + sum_{e in CSR row} val[e] * W1t[col[e]])   W1t = encoder_linear.weight^T, [n_cols x H]. */
int dmt_ae_encoder_fwd(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                       const float* val, const float* W1t, const float* b1, int H, float* A1, void* stream);
/* Decoder last layer + loss + first backward product in one pass over the target rows (src/models/ae.py:135-142,153-156):
 * for each target entry e (CSR position) of batch row j: o = A3[j].W4[c_e] + b4[c_e]; pred[e] = o (pred may be NULL);
 * train mode (gout != NULL): gout[e] = dloss/do / *n_targets, dZ3[j] = (sum_e gout[e] W4[c_e]) [* (1 - A3[j]^2) when
 * tanh_deriv != 0: A3 is the tanh output of the last decoder block], loss_rows[j] = sum_e loss. H in {128,256,384,512}. */
int dmt_ae_decoder_fwd(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                       const float* target, const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                       const int32_t* n_targets, float* pred, float* gout, float* dZ3, float* loss_rows, int tanh_deriv,
                       void* stream);

/* Tensor-core forms (tcgen05.mma kind::tf32, accumulators in TMEM; sm_100a only). passes = 3: 3xTF32 — every operand
 * is split into hi + lo TF32 parts while it is staged into shared memory and three MMAs accumulate per k-step, which
 * keeps fp32-level accuracy (~2^-21 relative); passes = 1: one TF32 pass (reduced precision, NOT the parity mode).
 * Same contracts as dmt_dense_fwd / dmt_dense_bwd_x / dmt_dense_bwd_w above (nn.Linear, src/models/ae.py:14-19,44-45). */
int dmt_dense_fwd_tc(const float* X, const float* W, const float* b, float* Y, float* Y_pre, const uint8_t* keep,
                     float keep_scale, int m, int n, int k, int act, int passes, void* stream);
int dmt_dense_bwd_x_tc(const float* dY, const float* W, const float* A_prev, const uint8_t* keep, float keep_scale,
                       float* dX, int m, int n, int k, int act_prev, int passes, void* stream);
int dmt_dense_bwd_w_tc(const float* dY, const float* X, float* dW, float* db, int m, int n, int k, int passes,
                       void* stream);

/* Decoder last layer as dense tensor-core GEMMs with a masked loss epilogue (src/models/ae.py:135-142,153-156 and
 * its autograd backward): O = A3 W4^T is formed tile by tile in TMEM and only the batch's target entries leave the
 * SM (pred[e], gout[e] = dloss/do / *n_targets, loss); train mode (gout != NULL) then runs dZ3 = (G W4) [*(1-A3^2)]
 * and dW4 = G^T A3 [n_dec x H], db4 = column sums of G [n_dec] (db4 may be NULL), with the sparse G scattered from
 * the CSR into the shared-memory operand tiles. loss_rows[0] = sum of the loss over all entries, loss_rows[1..] = 0.
 * Column indices must ascend inside every CSR row. H % 128 == 0. scratch >= dmt_ae_decoder_tc_scratch_floats(). */
int64_t dmt_ae_decoder_tc_scratch_floats(int n_rows, int n_dec, int H);
int dmt_ae_decoder_tc(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                      const float* target, const float* A3, const float* W4, const float* b4, int H, int n_dec,
                      int loss_kind, const int32_t* n_targets, int passes, float* pred, float* gout, float* dZ3,
                      float* dW4, float* db4, float* loss_rows, int tanh_deriv, float* scratch, void* stream);

/* ------------------------------------------------------------------ evaluation and privacy */

/* Global test metrics of src/train_recsys_assist.py:175-217 on the device: the split (CSR rows, pred/target aligned
 * with its storage order) is walked in blocks of block_rows rows; out[3*b + {0,1,2}] = sum of the loss
 * (src/models/utils.py:7-14), sum of squared errors (RMSE, src/metrics/metrics.py:8-11) and sum over the block's
 * non-empty rows of DCG@k/IDCG@k (src/metrics/metrics.py:63-84; unobserved columns rank last with gain 0, k =
 * block_k[b] = min(topk, #distinct columns of the block); want_ndcg = 0 skips it). The caller divides by the block's
 * entry / row counts and forms the entry-weighted mean over blocks (src/logger.py:35-55). No atomics: reproducible. */
int dmt_eval_blocks(const int32_t* indptr, const float* pred, const float* target, int n_rows, int block_rows,
                    int loss_kind, int want_ndcg, const int32_t* block_k, float* out, void* stream);

/* make_privacy of src/privacy.py:6-58 (applied to the pseudo-residuals at src/assist.py:59-60): clip range [a, b] =
 * the 2.5 % / 97.5 % quantiles of y (numpy's linear interpolation), mode 0 = dp: clip(y) + Laplace((b-a)/param),
 * mode 1 = ip: mean over int(param) uniform thresholds t of (2t-b if y<t else 2t-a). Noise: counter-based generator
 * keyed on (seed, element, draw) — reproducible, but NOT numpy's stream (the host path replays that). out may alias y.
 * quantiles (may be NULL) receives a, b. temp >= dmt_privacy_temp_bytes(n) bytes. */
int64_t dmt_privacy_temp_bytes(int64_t n);
int dmt_privacy(const float* y, int64_t n, int mode, float param, uint64_t seed, float* out, float* quantiles,
                void* temp, int64_t temp_bytes, void* stream);

/* ------------------------------------------------------------------ device-resident organization engine */

typedef struct dmt_org dmt_org_t;

/* One organization's AAE with its data resident in HBM (Organization.train/predict, src/organization.py:140-217).
 * data CSR: n_rows x n_enc (the organization's own columns, local ids); target CSR: n_rows x n_dec (ALL columns;
 * values = residuals, set every round with dmt_org_set_target). The CSR arrays are borrowed device pointers that
 * must outlive the handle; hidden sizes are the reference's [H1=256, H2=128] (src/utils.py:166-171).
 * stream NULL: the handle creates a private non-blocking stream (organizations then run concurrently).
 * plan_epochs >= 1: how many local epochs one plan / one graph may cover (sizes the plan buffers: ~40 B per target
 * entry and planned epoch — 180 GB of HBM buy whole-round plans at ML1M/Douban/Amazon shape). */
int dmt_org_create(dmt_org_t** out, int n_rows, int n_enc, int n_dec, int H1, int H2, const int32_t* d_indptr,
                   const int32_t* d_indices, const float* d_val, int64_t d_nnz, const int32_t* t_indptr,
                   const int32_t* t_indices, int64_t t_nnz, int batch_rows, int loss_kind, int plan_epochs,
                   void* stream);
int dmt_org_destroy(dmt_org_t* org);
/* How the decoder's last layer runs in dmt_org_train_epoch / dmt_org_predict: mode 0 = row-gather SDDMM + segmented
 * reductions (any CSR), mode 1 = tcgen05 GEMMs (dmt_ae_decoder_tc; needs ascending column indices inside every row of
 * the target CSR and of the CSRs passed to dmt_org_predict). passes: 3 (3xTF32, parity mode) or 1. Default: mode 0. */
int dmt_org_set_decoder_mode(dmt_org_t* org, int mode, int passes);
/* Shape of a training step inside the epoch graph. on = 0 (default): one chain of kernels. on != 0: after the decoder
 * the backward pass is enqueued as the DAG it is ({dW4,db4} | {dW3,db3} | dZ2 -> ({dW2,db2} | dZ1 -> (dW1 | db1)))
 * on auxiliary streams, i.e. parallel branches of the captured graph: a shorter critical path per step, which pays
 * when a GPU holds few organizations (org-sharded runs); with many organizations per GPU their graphs already fill the
 * machine and the extra cross-branch dependencies cost more than they save (measured). Same results. */
int dmt_org_set_fanout(dmt_org_t* org, int on);
/* Grid of the decoder chunk kernel inside a training step: 0 (default) = two blocks per SM, the fastest for one
 * organization alone; a rank that runs many organizations concurrently gets a shorter ROUND with one block per SM
 * (148), because the kernel's register footprint then leaves room for the other organizations' kernels (measured at
 * ML1M shape, 18 organizations: 214.7 -> 205.6 ms per round while the kernel itself goes from 23.6 to 32.9 us). */
int dmt_org_set_decoder_blocks(dmt_org_t* org, int blocks);
/* How the gather kernels of the fused step (decoder SDDMM of src/models/ae.py:135-142 and the dW4 segments of its
 * backward) bring their 1 KB weight / activation rows to the SM. mode 0 (default): 128-bit loads into registers,
 * software-pipelined in groups of four rows. mode 1 (DMT_GATHER=bulk): one cp.async.bulk copy per row into
 * warp-private shared-memory rings completed on mbarriers (csrc/bulk.cuh); parity-tested, measured SLOWER at every
 * grid (decoder 28-45 us against 23 us at ML1M shape: the copy engine's per-operation cost dominates at 1 KB), kept
 * as the measured alternative. Same sums in a different fixed order. */
int dmt_org_set_gather_mode(dmt_org_t* org, int mode);
/* on = 1: the same-stream kernels of the fused step are launched with programmatic stream serialization (PDL): a
 * kernel's blocks are scheduled while its predecessor drains and wait at cudaGridDependencySynchronize before touching
 * memory, which hides launch latency between the six dependent launches of a batch (src/organization.py:149-162 is a
 * strictly sequential loop). Pays on ranks with few organizations; default off (DMT_PDL=1). Results are unchanged. */
int dmt_org_set_pdl(dmt_org_t* org, int on);
/* Batch rows per CTA of the row-local forward / backward kernels of the fused step at 500-row batches (4, 8 or 16;
 * default 8). With 4 rows per CTA two warps share every row's encoder entries and the per-CTA dense work halves: the
 * kernels take 14 / 12 us instead of 19 / 16 us (ML1M shape), which shortens the round of a rank that holds one or two
 * organizations (28.6 vs 30.5 ms) and lengthens it when more organizations share the GPU (twice the CTAs and weight
 * traffic: 18 organizations 205.7 vs 197.7 ms). Same arithmetic, but the encoder sum of a row is split in two ordered
 * halves and the bias-gradient partials group four rows: results differ from the 8-row tile in the last bit, so the
 * Python layers never switch it by themselves (results must not depend on how organizations are sharded). */
int dmt_org_set_row_tile(dmt_org_t* org, int rows);
int dmt_org_gather_mode(const dmt_org_t* org);
/* How one iteration of the batch loop (src/organization.py:149-162) is cut into launches. mode 1 (default whenever
 * H1 = 256, H2 = 128, batch_rows <= 512 and decoder mode 0): the fused step of csrc/fused.cu — six dependent launches
 * (row-local forward, decoder chunks, {row-local backward | dW4 segments}, {dW3/dW2 tiles | bias gradients | dW1
 * segments}, norm + step scalars, Adam), finish passes replaced by last-arriver sums in chunk order. mode 0: the
 * classic step of twenty kernels (one per layer / reduction). Same arithmetic contract; sums are taken in a different
 * but fixed order. dmt_org_step_mode returns the mode in effect (0 while the tensor-core decoder is on). */
int dmt_org_set_step_mode(dmt_org_t* org, int mode);
int dmt_org_step_mode(const dmt_org_t* org);
/* number of fp32 parameters; flat layout: W1t[n_enc*H1] b1[H1] W2[H2*H1] b2[H2] W3[H1*H2] b3[H1] W4[n_dec*H1] b4[n_dec] */
int64_t dmt_org_num_params(const dmt_org_t* org);
/* Copy parameters in/out (device pointers, flat layout above). set also resets the Adam state: the reference builds
 * a fresh model and optimizer every round (src/organization.py:144-148). */
int dmt_org_set_params(dmt_org_t* org, const float* flat);
int dmt_org_get_params(const dmt_org_t* org, float* flat);
/* Target values for the coming rounds (borrowed device pointer aligned with t_indices; keep it stable, its address
 * is part of the captured graph). */
int dmt_org_set_target(dmt_org_t* org, const float* t_val);
/* One local epoch = one CUDA-graph launch. Batch b = rows[row_off[b] .. row_off[b+1]) (device int32; every batch
 * sorted ascending, rows with neither data nor targets removed = the reference's `total_user`, ae.py:101).
 * Batches without DATA entries are skipped on device (src/organization.py:153-155). n_t_entries / n_d_entries:
 * total target / data entries of the listed rows. keep: optional uint8 [n_rows_total x H2] dropout keep-masks in
 * batch-row order (parity mode); NULL = counter-based on-device generator seeded with `seed`.
 * epoch_loss[b] (device, may be NULL) receives each batch's mean loss. Asynchronous on the handle's stream. */
int dmt_org_train_epoch(dmt_org_t* org, const int32_t* rows, const int32_t* row_off, int n_rows_total, int n_batches,
                        int64_t n_t_entries, int64_t n_d_entries, const uint8_t* keep, uint64_t seed, double lr,
                        double beta1, double beta2, double eps, double weight_decay, float max_norm,
                        float* epoch_loss);
/* Organization.predict: eval forward at every target position of the given (data, target-structure) pair — pass the
 * test split's CSR to predict it — writing pred aligned with t_indices. */
int dmt_org_predict(dmt_org_t* org, const int32_t* d_indptr, const int32_t* d_indices, const float* d_val,
                    const int32_t* t_indptr, const int32_t* t_indices, int n_rows, float* pred);
int dmt_org_sync(dmt_org_t* org);
void* dmt_org_stream(dmt_org_t* org);
/* Stream ordering with the caller's streams: wait = the handle's stream waits for work already enqueued on
 * `stream`; signal = `stream` waits for work already enqueued on the handle's stream. */
int dmt_org_wait_stream(dmt_org_t* org, void* stream);
int dmt_org_signal_stream(dmt_org_t* org, void* stream);

/* ------------------------------------------------------------------ organization groups */

typedef struct dmt_group dmt_group_t;

/* A group trains several organizations of one rank in lockstep: every step kernel is launched ONCE with the
 * organization as grid dimension z (same rows, hidden sizes, batch size and target width required). This is the
 * same arithmetic as dmt_org_train_epoch per organization — organizations never share data — but ~20 launches per
 * step for all of them instead of ~20 per organization (the reference loops organizations in Python,
 * src/train_recsys_assist.py:148-149). */
int dmt_group_create(dmt_group_t** out, dmt_org_t* const* orgs, int n, void* stream);
int dmt_group_destroy(dmt_group_t* g);
/* All local epochs of one round for every organization of the group: organization i trains on batches
 * rows[i][row_off[i][b] .. row_off[i][b+1]) for b < n_batches (n_batches = epochs x batches per epoch; the layout
 * rules of dmt_org_train_epoch apply; device pointers in host arrays of length n). Dropout comes from the on-device
 * generator seeded with seeds[i]. batch_loss[i] (device, may be NULL) receives every batch's mean loss.
 * Call dmt_org_set_params / dmt_org_set_target on the members first; afterwards their streams are ordered after
 * the group's work (dmt_org_predict / dmt_org_get_params see the trained parameters). */
int dmt_group_train(dmt_group_t* g, const int32_t* const* rows, const int32_t* const* row_off, int n_rows_total,
                    int n_batches, const int64_t* n_t_entries, const int64_t* n_d_entries, const uint64_t* seeds,
                    double lr, double beta1, double beta2, double eps, double weight_decay, float max_norm,
                    float* const* batch_loss);
int dmt_group_sync(dmt_group_t* g);
void* dmt_group_stream(dmt_group_t* g);
int dmt_group_wait_stream(dmt_group_t* g, void* stream);

/* ------------------------------------------------------------------ measurement support (bench.py) */

/* Number of kernels of THIS library launched so far in the process (kernels captured in a CUDA graph are counted
 * once per graph launch). */
int64_t dmt_launch_count(void);
/* Average duration in ms of each kernel class of one training step on batch b of the handle's current epoch plan
 * (call after dmt_org_train_epoch): every class is launched `reps` times back to back between two CUDA events on
 * the handle's stream; parameters and optimizer state are restored afterwards. ms_out (host) has
 * dmt_org_profile_classes() entries: 0 encoder SpMM, 1 dense forward (2 GEMMs), 2 gradient zeroing,
 * 3 decoder+loss+dZ3, 4 dW4 segmented reduction, 5 dense backward (4 GEMMs + 2 column sums),
 * 6 dW1 segmented reduction + column sum, 7 gradient norm, 8 clip+Adam. */
int dmt_org_profile_classes(void);
int dmt_org_profile_step(dmt_org_t* org, int b, int reps, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* DMT_B200_H */
