/* mimrl_b200 — C ABI of the B200-native MI / CMI hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream
 * (passed as void*, i.e. a cudaStream_t).  No ownership is transferred: all
 * buffers are allocated by the caller (PyTorch on the Python side).  Return
 * value: 0 = ok, non-zero = error, text via mimrl_last_error().
 *
 * The reference (kiva12138/MIMRL) has no FFI of its own — its boundary is a
 * set of Python symbols.  Each group below names the
 * reference code it replaces; mimrl_b200/*.py binds these with ctypes and
 * re-exposes the reference's Python signatures on top.
 *
 * Row-block convention (single- and multi-GPU alike): a rank OWNS n_own rows
 * with global indices [own_offset, own_offset + n_own) and sweeps them against
 * ALL n_all rows of the other operand.  On one GPU n_own == n_all, offset 0.
 */
#ifndef MIMRL_B200_H
#define MIMRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMRL_ABI_VERSION 1

/* --bound_type values, Model.py:121-146 / Parameters.py:41-51 */
enum mimrl_bound {
  MIMRL_BOUND_DV = 0,
  MIMRL_BOUND_MINE = 1,
  MIMRL_BOUND_TUBA = 2,
  MIMRL_BOUND_NWJ = 3,
  MIMRL_BOUND_INFONCE = 4,
  MIMRL_BOUND_JS_FGAN = 5,
  MIMRL_BOUND_JS = 6,
  MIMRL_BOUND_SMILE = 7,
  MIMRL_BOUND_INTERPOLATE = 8
};

/* statistics requested from the score sweep */
#define MIMRL_STAT_CLAMP 1    /* exp-sum runs on clamp(S,-1,1)  (smile, VMI.py:186-191) */
#define MIMRL_STAT_SOFTPLUS 2 /* also accumulate sum softplus(S) (js_fgan/js/smile, VMI.py:169-174) */
#define MIMRL_STAT_MAXONLY 4  /* tcgen05 path: row_max only, from ONE fp16 product (approximate, ~2^-11 |y||x|): the
                                reference point of mimrl_sep_fused_forward; row_sum / row_sp come back as 0 */

/* per-pair weight families of the backward sweep */
#define MIMRL_WEIGHT_EXP 0     /* w_ij = exp(S_ij - shift) */
#define MIMRL_WEIGHT_SIGMOID 1 /* w_ij = sigmoid(S_ij)     */
#define MIMRL_WEIGHT_INTERP 2  /* w_ij = p (a1 + a2 / (1 - sg p)), p = exp(S_ij - lse): the row-parameterised part of the
                                  interpolated bound's gradient (VMI.py:229-250); shift = four vectors [4][n]: lse, a1, a2,
                                  sg of the score's row, scaled by the caller so that |w| <= 1 (tcgen05 path only) */

/* kernel implementation selector (both are CUDA kernels of this library) */
#define MIMRL_IMPL_AUTO 0
#define MIMRL_IMPL_FFMA 1    /* fp32 CUDA-core tiles, any embed width <= 256 */
#define MIMRL_IMPL_TCGEN05 2 /* tcgen05 + TMA + TMEM, fp16x3 split (fp32-class accuracy), embed <= 128 */

int mimrl_version(void);
const char *mimrl_last_error(void);
/* number of kernels this library has launched so far in this process */
uint64_t mimrl_launch_count(void);

/* ------------------------------------------------------------------------
 * Separable critic: fused score + bound.  Replaces VMI.py:57
 * (scores = y_ @ x_.T) together with the bound functions VMI.py:136-198 as
 * they are called from VMIEstimator.forward (Model.py:115-148); the B x B
 * matrix is never written to memory.
 * ---------------------------------------------------------------------- */

size_t mimrl_sep_workspace_bytes(int n_own, int n_all, int embed);
/* which implementation a request resolves to: MIMRL_IMPL_FFMA or MIMRL_IMPL_TCGEN05 */
int mimrl_sep_selected_impl(int n_own, int n_all, int embed, int impl);

/* Forward sweep.  S_ij = own_i . all_j for the owned rows; for every owned row
 * i, over the columns j != own_offset + i:
 *   row_max[i] = max_j t(S_ij),  row_sum[i] = sum_j exp(t(S_ij) - row_max[i]),
 *   row_sp[i]  = sum_j softplus(S_ij)        (only with MIMRL_STAT_SOFTPLUS)
 * and diag[i] = S_{i, own_offset+i}.  t = clamp(.,-1,1) with MIMRL_STAT_CLAMP. */
int mimrl_sep_row_stats(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                        int own_offset, int flags, int impl, float *row_max, float *row_sum, float *row_sp,
                        float *diag, void *workspace, size_t workspace_bytes, void *stream);

/* Backward sweep.  out_i = coef[0] * sum_{j} w_ij * all_j + dcoef[i] * all_{own_offset+i}
 * with w_ij from the weight family; the diagonal pair j == own_offset+i is left
 * out of the sum unless include_diag.  shift is indexed by the owned row
 * (shift[i], i < n_own) or, with shift_by_swept, by the swept row (shift[j],
 * j < n_all).  Called twice per backward: rows = h(y) to get d/dh(y), then with
 * the operands swapped to get d/dg(x). */
/* Fused forward sweep (tcgen05 path): for every owned row i, in ONE pass over the score tiles,
 *   row_sum[i] = sum_{j != i} exp(S_ij - shift[i])          (the off-diagonal statistic of mimrl_sep_row_stats)
 *   wsum[i,:]  = sum_j exp(S_ij - shift[i]) all_emb[j,:]    (j = i included iff include_diag)
 * shift is any per-row reference point near the row maximum (MIMRL_STAT_MAXONLY pre-pass).  For the exp-family bounds
 * (dv, mine, tuba, nwj, infonce) this replaces the exact statistics sweep and the owned-row gradient sweep
 * (the gradient is wsum rescaled by exp(shift - final shift)): 4 1/3 tensor-core units per step instead of 5. */
int mimrl_sep_fused_forward(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                            int own_offset, int include_diag, const float *shift, float *wsum, float *row_sum,
                            void *workspace, size_t workspace_bytes, void *stream);
/* Online-softmax forward sweep (tcgen05 path): mimrl_sep_fused_forward without a supplied reference point.  The sweep
 * keeps a running reference per row (the row maximum seen so far, lazily updated) and returns it:
 *   row_ref[i]  = the reference point the sums are referred to (a value near the row maximum; 0 for an empty row)
 *   row_sum[i]  = sum_{j != i} exp(S_ij - row_ref[i])
 *   wsum[i,:]   = sum_j exp(S_ij - row_ref[i]) all_emb[j,:]      (j = i included iff include_diag)
 *   diag[i]     = S_{i, own_offset+i}                             (optional, may be NULL)
 * This is the whole forward of the exp-family bounds of VMI.py:136-166 in ONE pass over the score tiles (no max
 * pre-pass): 4 tensor-core units per step instead of 4 1/3. */
int mimrl_sep_online_forward(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                             int own_offset, int include_diag, float *row_ref, float *wsum, float *row_sum, float *diag,
                             void *workspace, size_t workspace_bytes, void *stream);
/* Second forward sweep of the interpolated bound (VMI.py:201-250; tcgen05 path): with p_ij = exp(S_ij - row_lse[i]), over
 * the off-diagonal columns: row_q[i] = sum_j p_ij / (1 - row_sigma[i] p_ij), row_t[i] = sum_j log(1 - row_sigma[i] p_ij). */
int mimrl_sep_interp_stats(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed, int own_offset,
                           const float *row_lse, const float *row_sigma, float *row_q, float *row_t, void *workspace,
                           size_t workspace_bytes, void *stream);
int mimrl_sep_weighted_sum(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                           int own_offset, int weight_family, int include_diag, const float *shift,
                           int shift_by_swept, const float *coef, const float *dcoef, int impl, float *out,
                           void *workspace, size_t workspace_bytes, void *stream);

/* Bound value from the per-row statistics of ALL n rows (VMI.py:136-198 and
 * the MINE branch Model.py:121-124).  log_baseline may be NULL (constant
 * baseline).  result[0] = mi, result[1] = mi_loss, result[2..15] = scalars kept
 * for the backward pass. */
int mimrl_bound_finalize(int bound, const float *row_max, const float *row_sum, const float *row_sp,
                         const float *diag, const float *log_baseline, int n, float *result, void *stream);

/* Coefficients of the backward sweep from the saved scalars and the incoming
 * gradients grad[0] = d/d mi, grad[1] = d/d mi_loss:
 *   coef[0], shift[n], dcoef[n] as consumed by mimrl_sep_weighted_sum and
 *   dbaseline[n] = d/d log_baseline (written only if log_baseline != NULL).
 * weight_family / include_diag for the bound are returned by
 * mimrl_bound_weight_family(). */
int mimrl_bound_backward_coef(int bound, const float *result, const float *grad, const float *row_max,
                              const float *row_sum, const float *diag, const float *log_baseline, int n,
                              float *coef, float *shift, float *dcoef, float *dbaseline, void *stream);
int mimrl_bound_weight_family(int bound, int *weight_family, int *include_diag, int *stat_flags);

/* ------------------------------------------------------------------------
 * Materialised-score entry points: the free bound functions of VMI.py:136-250
 * applied to a score matrix that already exists in memory (API parity for
 * CriticModel.forward -> bound(scores); also the concat-critic path).
 * scores is [n_rows, n_cols] row-major, the owned row block of the global
 * n_cols x n_cols matrix.
 * ---------------------------------------------------------------------- */
int mimrl_scores_row_stats(const float *scores, int n_rows, int n_cols, int own_offset, int flags,
                           float *row_max, float *row_sum, float *row_sp, float *diag, void *stream);
/* grad_scores[i][j] = coef[0] * w_ij (off-diagonal, or all with include_diag) + dcoef[i] on the diagonal */
int mimrl_scores_grad(const float *scores, int n_rows, int n_cols, int own_offset, int weight_family,
                      int include_diag, const float *shift, const float *coef, const float *dcoef,
                      float *grad_scores, void *stream);

/* ------------------------------------------------------------------------
 * fp32-class GEMM on the tensor cores (fp16 hi/lo split, three products) for the
 * relu MLP stacks of the critics, baselines and the CMI classifier.  Replaces
 * the nn.Linear calls inside VMI.py:13-22 `mlps` and Model.py:52-57.
 *   mode 0: C[M,N] = A[M,K] . B[N,K]^T (+ bias[N], relu)    Linear forward
 *   mode 1: C[M,N] = A[M,K] . B[K,N]                         input gradient  dz W
 *   mode 2: C[M,N] = A[K,M]^T . B[K,N]                       weight gradient dz^T x
 * a_mask (nullable, same shape as A): A is multiplied by (a_mask > 0) first
 * (ReLU backward with the saved layer output as mask).
 * ---------------------------------------------------------------------- */
size_t mimrl_gemm_workspace_bytes(int mode, int M, int N, int K);
int mimrl_gemm_f32x3(int mode, const float *A, const float *a_mask, const float *B, int M, int N, int K,
                     const float *bias, int relu, float *C, void *workspace, size_t workspace_bytes, void *stream);
/* Same product on operands split once and reused (a layer's input serves its forward and its weight gradient, a
 * weight serves forward and input gradient).  mimrl_split_f32 writes [scale header | fp16 hi | fp16 lo] for a
 * row-major [rows, cols] matrix, optionally masked by (mask > 0); colsum (nullable) ACCUMULATES the column sums of
 * the masked matrix (bias gradient).  Stored operand shapes as listed for the three modes above. */
size_t mimrl_split_bytes(int rows, int cols);
int mimrl_split_f32(const float *src, const float *mask, int rows, int cols, void *out, float *colsum, void *stream);
/* mask from the hi half of another split buffer of the same shape (entries with hi <= 0 are zeroed) */
int mimrl_split_f32_hmask(const float *src, const void *mask_split, int rows, int cols, void *out, float *colsum,
                          void *stream);
size_t mimrl_gemm_split_workspace_bytes(int mode, int M, int N, int K);
int mimrl_gemm_split(int mode, const void *a_split, const void *b_split, int M, int N, int K, const float *bias,
                     int relu, float *C, void *workspace, size_t workspace_bytes, void *stream);
/* C[M,N] = A[M,K] . B[N,K]^T with both operands in BLOCKED-K order: tiles of 64 consecutive k, each [rows][64]
 * contiguous (same header, hi / lo offsets and total size as mimrl_split_f32 of a [rows, K] matrix; K % 64 == 0).
 * Every TMA box is then one contiguous run in memory: the layout for contractions over millions of entries (the
 * weight gradients over pairs / fibres).  Workspace: mimrl_gemm_split_workspace_bytes(0, M, N, K). */
int mimrl_gemm_split_blocked(const void *a_split, const void *b_split, int M, int N, int K, float *C, void *workspace,
                             size_t workspace_bytes, void *stream);
/* Same contraction ADDED into C[M,N] (zero-filled by the caller or holding a running sum): every split-K work item
 * adds its partial sum with red.global.add.f32 -- no partial buffer, no reduction launch.  The order of the additions
 * is not fixed (last-bit differences from run to run); used for the CubeMLP weight gradients. */
int mimrl_gemm_split_blocked_acc(const void *a_split, const void *b_split, int M, int N, int K, float *C, void *stream);

/* The same three products on the CUDA cores (exact fp32 FFMA), for the layers outside the tensor-core envelope: small
 * batches (the reference trains at bs = 128, README.md:16-26) and the 1- / 2-wide heads of the unnormalized baseline
 * (VMI.py:86-89) and the CMI classifier (Model.py:52-57).  Modes, a_mask, bias and relu as mimrl_gemm_f32x3; colsum
 * (nullable, mode 2 only): colsum[m] += sum_k A'[k, m] of the masked matrix, i.e. the bias gradient. */
size_t mimrl_linear_small_workspace_bytes(int mode, int M, int N, int K);
int mimrl_linear_small(int mode, const float *A, const float *a_mask, const float *B, int M, int N, int K,
                       const float *bias, int relu, float *C, float *colsum, void *workspace, size_t workspace_bytes,
                       void *stream);

/* The whole relu MLP of VMI.py:13-22 / Model.py:52-57 (Linear+ReLU x3, Linear; hidden 256) for SMALL batches in three
 * launches: forward of all four layers (h1..h3 [M,256] saved), data gradients of all four layers (dz1..dz3 [M,256],
 * dz4 [M,d_out]: caller-allocated scratch; gx nullable), and one grouped launch for the weight gradients
 * gw_l = dz_l^T input_l and bias gradients gb_l (written, not accumulated; any of them nullable).  d_in <= 384,
 * d_out <= 256, fp32 FFMA. */
int mimrl_mlp4_small_supported(int d_in, int hidden, int d_out);
int mimrl_mlp4_small_fwd(const float *x, int M, int d_in, const float *w1, const float *b1, const float *w2, const float *b2,
                         const float *w3, const float *b3, const float *w4, const float *b4, int d_out, float *h1, float *h2,
                         float *h3, float *y, void *stream);
int mimrl_mlp4_small_bwd(const float *gy, const float *x, int M, int d_in, int d_out, const float *w1, const float *w2,
                         const float *w3, const float *w4, const float *h1, const float *h2, const float *h3, float *dz1,
                         float *dz2, float *dz3, float *dz4, float *gx, float *gw1, float *gb1, float *gw2, float *gb2,
                         float *gw3, float *gb3, float *gw4, float *gb4, void *stream);

/* ------------------------------------------------------------------------
 * k-NN conditional-MI sampler.  Replaces the neighbour search and gathers of
 * prod_knn_sample (Model.py:75-106), i.e. sklearn NearestNeighbors.kneighbors.
 * ---------------------------------------------------------------------- */

size_t mimrl_knn_workspace_bytes(int n_keys, int n_queries, int width, int k);

/* keys [n_keys, width] (the Z pool, fp32), query_ids [n_queries] (int64 row ids
 * into keys, drawn on the host from numpy's global RNG).  Keys listed in
 * query_ids are excluded from the search (Model.py:83-84).  exact_form: 1 =
 * float64 GEMM form ||q||^2 - 2q.z + ||z||^2 (sklearn 'brute' route), 0 =
 * float64 direct differences ('kd_tree' route).  Output neighbours [n_queries,k]
 * int64, nearest first, exact ties -> lowest index:
 *   nbr_orig  = indices into keys,
 *   nbr_comp  = indices into the pool with the query rows removed (what
 *               sklearn returns in the reference).
 * key_index_offset is added to nbr_orig (multi-GPU key shards). radius is
 * accepted and ignored, like the reference. */
int mimrl_knn_search(const float *keys, int n_keys, int width, const int64_t *query_ids, int n_queries, int k,
                     float radius, int exact_form, int64_t *nbr_orig, int64_t *nbr_comp, double *nbr_dist,
                     void *workspace, size_t workspace_bytes, void *stream);
/* Host-side query draw of the sampler (Model.py:81, np.random.choice(range(N), size=m, replace=False), which numpy's
 * legacy RandomState turns into permutation(N)[:m]).  key[624] / *pos: the MT19937 words and position of
 * np.random.get_state(); both are advanced exactly as numpy advances them (N - 1 masked-rejection draws), out[m]
 * receives numpy's first m entries.  Pure host code (no CUDA call): 4.5x faster than numpy's own shuffle at
 * N = 2^20 on the B200 host because the swap targets are prefetched a block ahead. */
int mimrl_legacy_permutation_head(uint32_t *key, int *pos, int64_t n, int64_t m, int64_t *out);

/* Fitted pool: the part of a search that depends on the keys alone (squared norms, fp16 hi / lo planes with one scale per
 * 128-key tile), computed once per pool instead of once per search.  The reference refits on every call
 * (Model.py:82-85, NearestNeighbors(...).fit(Z2) on the pool minus the drawn rows); here the drawn rows are masked per
 * search, so one fit of the WHOLE pool serves every search until the pool's contents change (Model.py:323,329 search
 * T_F_all twice per step; the *_F_all pools are constant over an epoch).  mimrl_knn_fit_bytes returns 0 for pools the
 * search does not prepare (width <= 15: direct-difference / sorted routes; fewer than 2048 keys): search those with
 * mimrl_knn_search.  mimrl_knn_search_fitted: `fitted` filled by mimrl_knn_fit for the same keys; everything else as
 * mimrl_knn_search, results bit-identical to it. */
size_t mimrl_knn_fit_bytes(int n_keys, int width);
int mimrl_knn_fit(const float *keys, int n_keys, int width, void *fitted, size_t fitted_bytes, void *stream);
int mimrl_knn_search_fitted(const float *keys, int n_keys, int width, const void *fitted, size_t fitted_bytes,
                            const int64_t *query_ids, int n_queries, int k, float radius, int exact_form,
                            int64_t *nbr_orig, int64_t *nbr_comp, double *nbr_dist, void *workspace,
                            size_t workspace_bytes, void *stream);
/* Same search with explicit query rows (multi-GPU: queries live on another rank).
 * excluded_sorted: ascending global ids to skip, may be NULL. */
int mimrl_knn_search_rows(const float *keys, int n_keys, int width, int64_t key_index_offset,
                          const float *queries, int n_queries, const int64_t *excluded_sorted, int n_excluded,
                          int k, int exact_form, int64_t *nbr_orig, double *nbr_dist, void *workspace,
                          size_t workspace_bytes, void *stream);

/* out[r, c] = src[idx[r / repeat], c % width] for c < out_width: row gather +
 * neighbour repeat + column tiling (Model.py:97-104) in one pass. */
int mimrl_gather_rows(const float *src, int n_src, int width, const int64_t *idx, int n_idx, int repeat,
                      int out_width, float *out, void *stream);

/* ------------------------------------------------------------------------
 * Conditional-MI classifier head.  Replaces Model.py:69-70 (clamp + final
 * activation), Model.py:198 (BCE) and estimate_cmi Model.py:203-219.
 * logits [2n, 2]; rows [0,n) joint, [n,2n) product.  act: 0 hardtanh, 1 sigmoid.
 * result[0] = cmi, result[1] = loss.
 * ---------------------------------------------------------------------- */
int mimrl_vcmi_head_fwd(const float *logits, int n, int act, float *result, void *stream);
/* grad[0] = d/d cmi, grad[1] = d/d loss -> grad_logits [2n, 2] */
int mimrl_vcmi_head_bwd(const float *logits, int n, int act, const float *grad, float *grad_logits,
                        void *stream);

/* ------------------------------------------------------------------------
 * CubeMLP axis mix.  Replaces one third of MLPsBlock.forward_ln_last /
 * forward_ln_first (MLPProcess.py:64-122): for x [outer, A, inner] (A = the
 * mixed axis, inner = product of the faster axes)
 *   y = LN_{A'}( W2 act(W1 x + b1) + b2 + (Wres x | x) )            ln_first = 0
 *   y = W2 act(W1 LN_A(x) + b1) + b2 + (Wres x | x)                 ln_first = 1
 * without any permute copy.  act: 0 gelu(erf), 1 relu, 2 tanh.  Biases and wres
 * may be NULL.  `saved` receives the LayerNorm (mean, rstd) per fibre
 * (mimrl_cubemlp_saved_floats).  The backward recomputes the forward per tile,
 * writes gx, ACCUMULATES gln_w / gln_b (caller zero-fills) and leaves
 * s_gz [outer,a_out,inner], s_h = act(W1 u + b1) and s_gpre [outer,a_hid,inner]
 * (and s_u = LN(x) [outer,a_in,inner] when ln_first) for the weight gradients:
 *   gW2 = sum s_gz s_h^T, gb2 = sum s_gz, gW1 = sum s_gpre u^T, gb1 = sum s_gpre,
 *   gWres = sum s_gz x^T.
 * ---------------------------------------------------------------------- */
size_t mimrl_cubemlp_saved_floats(int outer, int a_in, int a_hid, int a_out, int inner);
int mimrl_cubemlp_mix_fwd(const float *x, int outer, int a_in, int inner, const float *w1, const float *b1,
                          int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                          const float *ln_w, const float *ln_b, int ln_first, int act, float *y, float *saved,
                          void *stream);
int mimrl_cubemlp_mix_bwd(const float *x, const float *gy, int outer, int a_in, int inner, const float *w1,
                          const float *b1, int a_hid, const float *w2, const float *b2, int a_out,
                          const float *wres, const float *ln_w, const float *ln_b, int ln_first, int act,
                          const float *saved, float *gx, float *s_gz, float *s_h, float *s_gpre, float *s_u,
                          float *gln_w, float *gln_b, void *stream);

/* Tiny mixed axis (all three sizes <= 4, the modality mix K = 3 of MLPProcess.py:106-112): the complete backward in
 * one register-resident kernel.  Writes gx; accumulates (+=) gw1 [a_hid,a_in], gb1, gw2 [a_out,a_hid], gb2,
 * gwres [a_out,a_in] (NULL without res_projection), gln_w, gln_b. */
int mimrl_cubemlp_small_supported(int a_in, int a_hid, int a_out);
int mimrl_cubemlp_small_bwd(const float *x, const float *gy, int outer, int a_in, int inner, const float *w1,
                            const float *b1, int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                            const float *ln_w, const float *ln_b, int ln_first, int act, float *gx, float *gw1, float *gb1,
                            float *gw2, float *gb2, float *gwres, float *gln_w, float *gln_b, void *stream);

/* Tensor-core forward of the same mix (ln_first = 0, axis sizes <= 128, not the tiny-axis case): fibres in TMEM
 * lanes, W1 / W2 / Wres resident in shared memory, LayerNorm thread-local in the epilogue.  Writes the same
 * `saved` statistics, so mimrl_cubemlp_mix_bwd applies unchanged.  prev_ln_w / prev_ln_b [prev_n] (nullable): x is the
 * unmodified output of a LayerNorm with these parameters over prev_n features (the previous mix of the block); its
 * operand scale then comes from the bound max|ln_w| sqrt(prev_n - 1) + max|ln_b| instead of a pass over x.  Without
 * such a bound the compile-time specialised sequence-mix kernel (README shapes) scales every fibre by its own max|x|
 * inside the kernel; other shapes take one pass over x.  prepared != 0: mimrl_cubemlp_prep_many has filled `workspace`. */
int mimrl_cubemlp_tc_supported(int a_in, int a_hid, int a_out, int ln_first, int act);
size_t mimrl_cubemlp_tc_workspace_bytes(int a_in, int a_hid, int a_out);
int mimrl_cubemlp_mix_fwd_tc(const float *x, int outer, int a_in, int inner, const float *w1, const float *b1,
                             int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                             const float *ln_w, const float *ln_b, int act, float *y, float *saved, void *workspace,
                             size_t workspace_bytes, const float *prev_ln_w, const float *prev_ln_b, int prev_n,
                             int prepared, void *stream);
/* The preparation (weight split, operand scales) of n <= 8 mixes of an encoder forward in ONE launch.  Every array has n
 * entries, in execution order; workspace[m] must be ZERO-FILLED (mimrl_cubemlp_tc_workspace_bytes each).  Mix m > 0 must
 * carry prev_ln_w / prev_ln_b / prev_n (its input is the previous mix's LayerNorm output); x and n_cols[0] describe the
 * input of mix 0; inner[m] / act[m] are the `inner` and `act` arguments the forward call of mix m will get (they
 * decide whether that mix takes its input scale inside the kernel).  Afterwards mimrl_cubemlp_mix_fwd_tc is called with
 * prepared = 1 on the same workspaces. */
int mimrl_cubemlp_prep_many(int n, const float *x, const long long *n_cols, const int *a_in, const int *a_hid,
                            const int *a_out, const int *inner, const int *act, const float *const *w1,
                            const float *const *b1, const float *const *w2,
                            const float *const *wres, const float *const *ln_w, const float *const *prev_ln_w,
                            const float *const *prev_ln_b, const int *prev_n, void *const *workspace, void *stream);

/* Tensor-core backward of the same mix.  Writes gx; accumulates (+=, caller zero-fills) g_b1 [a_hid], g_b2 [a_out],
 * gln_w, gln_b [a_out] and the weight gradients gw1 [a_hid, a_in], gw2 [a_out, a_hid], gwres [a_out, a_in] (NULL
 * without res_projection).  op_x, op_h, op_gz, op_gpre: scratch for the fp16 hi/lo weight-gradient operands the data
 * pass leaves behind (a_in / a_hid / a_out / a_hid features over R = mimrl_cubemlp_tc_fibre_rows(outer, inner) fibre
 * rows; mimrl_cubemlp_tc_op_bytes(features, R) bytes each).  The three contractions over the fibres
 * (gW1 = gpre^T x, gW2 = gz^T h, gWres = gz^T x) are launched by this call: split over K, partial sums added in place.
 * ws_from_forward != 0: `workspace` is the buffer mimrl_cubemlp_mix_fwd_tc filled for the same x and weights; the
 * weight split, max|x| and the largest rstd are then taken from it instead of being recomputed.
 * Both directions take ONE preparation launch (operand maxima, weight split, operand scales) and one main kernel;
 * the mixes of the reference configuration (sequence mix 100->50->50 and 50->10->10 over inner = 384, channel mix
 * 128->128->128; gelu, residual projection) run on compile-time specialised kernels (csrc/cubemlp_tc2.cu, _tc3.cu). */
long long mimrl_cubemlp_tc_fibre_rows(int outer, int inner);
size_t mimrl_cubemlp_tc_op_bytes(int features, long long R);
int mimrl_cubemlp_mix_bwd_tc(const float *x, const float *gy, int outer, int a_in, int inner, const float *w1,
                             const float *b1, int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                             const float *ln_w, const float *ln_b, int act, const float *saved, float *gx, float *g_b1,
                             float *g_b2, float *gln_w, float *gln_b, float *gw1, float *gw2, float *gwres, void *op_x,
                             void *op_h, void *op_gz, void *op_gpre, void *workspace, size_t workspace_bytes,
                             int ws_from_forward, void *stream);

/* ---- the critic MLP in one forward kernel (reference VMI.py:13-22 `mlps(dim, 256, out, layers=2, 'relu')`) ----
 * y = W4 relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3) + b4 for x [M, d_in], d_in <= 128, hidden 256, d_out <= 128:
 * activations stay in TMEM between the layers.  Also written (caller-allocated with mimrl_split_bytes, mimrl_split_f32
 * format): op_x [M,d_in], op_h1..3 [M,256] (weight-gradient operands; the hi half of op_h* is the ReLU mask for
 * mimrl_split_f32_hmask) and ws_w1..4 (the split weights, reusable by mimrl_gemm_split in the backward).
 * scratch256: 256 bytes of device scratch. */
int mimrl_mlp4_supported(int d_in, int hidden, int d_out);
int mimrl_mlp4_fwd(const float *x, int M, int d_in, const float *w1, const float *b1, const float *w2, const float *b2,
                   const float *w3, const float *b3, const float *w4, const float *b4, int d_out, float *y, void *op_x,
                   void *op_h1, void *op_h2, void *op_h3, void *ws_w1, void *ws_w2, void *ws_w3, void *ws_w4,
                   void *scratch256, void *stream);

/* Data-gradient pass of mimrl_mlp4_fwd in one kernel: gx [M,d_in] (nullable), the row-major operands dz1..3 [M,256] and
 * dz4 [M,d_out] (mimrl_split_f32 format, caller-allocated; gW_l = dz_l^T input_l through mimrl_gemm_split mode 2) and
 * the bias gradients g_b1..4 (+=, nullable).  ws_w*, op_h* as left by the forward. */
int mimrl_mlp4_bwd(const float *gy, int M, int d_in, int d_out, const float *w2, const float *w3, const float *w4,
                   const void *ws_w1, const void *ws_w2, const void *ws_w3, const void *ws_w4, const void *op_h1,
                   const void *op_h2, const void *op_h3, float *gx, void *dz1, void *dz2, void *dz3, void *dz4, float *g_b1,
                   float *g_b2, float *g_b3, float *g_b4, void *scratch256, void *stream);

/* ---- concat critic, all pairs on the tensor cores (reference VMI.py:58-65 with mlps of VMI.py:13-22) ----
 * The first layer factorises over the concatenation: u = x W1x^T + b1 [n_own, 256], vt = (y W1y^T)^T [256, ldv]
 * (both from the caller).  scores[i, j] = w4 . relu(W3 relu(W2 relu(u_i + v_j) + b2) + b3) + b4 for every pair,
 * without the [n_own * n_all, 256] activations ever reaching HBM.  hidden = 256, layers = 2 only. */
int mimrl_concat_tc_supported(int hidden, int layers);
size_t mimrl_concat_workspace_bytes(int hidden);
int mimrl_concat_scores(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden, const float *w2,
                        const float *b2, const float *w3, const float *b3, const float *w4, const float *b4,
                        float *scores, void *workspace, size_t workspace_bytes, void *stream);
/* Backward for g = dL/dscores.  Accumulates (+=) into g_u [n_own,256], g_vt [256,ldv], g_b2, g_b3, g_w4 [256] and
 * writes the weight-gradient operands h1, h2, g2, g3 (mimrl_split_f32 format of a [256, mimrl_concat_pair_rows()]
 * matrix, feature-major in blocked-K order): gW2 = g2 h1^T and gW3 = g3 h2^T through mimrl_gemm_split_blocked. */
long long mimrl_concat_pair_rows(int n_own, int n_all);
int mimrl_concat_grad(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden, const float *w2,
                      const float *b2, const float *w3, const float *b3, const float *w4, const float *g, float *g_u,
                      float *g_vt, float *g_b2, float *g_b3, float *g_w4, void *op_h1, void *op_h2, void *op_g2,
                      void *op_g3, void *workspace, size_t workspace_bytes, void *stream);

/* ---- concat critic fused with its bound: neither the score matrix nor dL/dscores exists in memory ----
 * mimrl_concat_row_stats: the off-diagonal row statistics of mimrl_sep_row_stats (row_max, row_sum, row_sp; flags
 * MIMRL_STAT_CLAMP | MIMRL_STAT_SOFTPLUS) of scores[i, j] = f([x_i, y_j]) for the owned rows, reduced inside the forward
 * kernel; the diagonal scores come from the same MLP applied to the n_own diagonal pairs (an ordinary row batch).
 * mimrl_concat_grad_fused: mimrl_concat_grad with g_ij = coef[0] * w(s_ij) (exp family: exp(s_ij - shift[i]); sigmoid
 * family: sigmoid(s_ij)) for j != own_offset + i and 0 on the diagonal, formed in the kernel from the recomputed score
 * (VMI.py:58-65 + VMI.py:136-198 in two passes over the pair tiles); g_b4[0] += sum_ij g_ij. */
size_t mimrl_concat_stats_workspace_bytes(int hidden, int n_own, int n_all);
int mimrl_concat_row_stats(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden, int own_offset,
                           int flags, const float *w2, const float *b2, const float *w3, const float *b3, const float *w4,
                           const float *b4, float *row_max, float *row_sum, float *row_sp, void *workspace,
                           size_t workspace_bytes, void *stream);
int mimrl_concat_grad_fused(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden, int own_offset,
                            const float *w2, const float *b2, const float *w3, const float *b3, const float *w4,
                            const float *b4, int family, const float *coef, const float *shift, float *g_u, float *g_vt,
                            float *g_b2, float *g_b3, float *g_w4, float *g_b4, void *op_h1, void *op_h2, void *op_g2,
                            void *op_g3, void *workspace, size_t workspace_bytes, void *stream);

/* ---- feature heads either side of the fusion encoder (reference Model.py:466-475 and 489-507) ----
 * stack: t [bs,len_t,d], a [bs,len_a,d], v [bs,len_v,d] (len_* <= time_len) ->
 *   mean_t/a/v [bs,d] = the unmasked temporal means T_F, A_F, V_F (Model.py:466) and
 *   x [bs,time_len,3,d] = stack of the three sequences zero-padded to time_len (Model.py:468-475), in one pass.
 * Backward: g_src[b,l,:] = g_x[b,l,m,:] + g_mean_m[b,:] / len_m; g_x, g_mean_* and g_t/g_a/g_v may be NULL. */
int mimrl_feature_stack_fwd(const float *t, const float *a, const float *v, int bs, int len_t, int len_a, int len_v,
                            int time_len, int d, float *x, float *mean_t, float *mean_a, float *mean_v, void *stream);
int mimrl_feature_stack_bwd(const float *g_x, const float *g_mean_t, const float *g_mean_a, const float *g_mean_v, int bs,
                            int len_t, int len_a, int len_v, int time_len, int d, float *g_t, float *g_a, float *g_v,
                            void *stream);
/* reduce: out[b,:] = scale * sum_r x[b,r,:] over the rows = L' * K' (time x modality) rows of the encoder output:
 * features_compose_k / features_compose_t in {mean, sum} (Model.py:489-504) with scale = 1 / (L' K'), 1 / L', 1 / K' or 1. */
int mimrl_feature_reduce_fwd(const float *x, int bs, int rows, int d, float scale, float *out, void *stream);
int mimrl_feature_reduce_bwd(const float *g_out, int bs, int rows, int d, float scale, float *g_x, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MIMRL_B200_H */
