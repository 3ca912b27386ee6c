// Shared helpers for the mimrl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mimrl_b200.h"

namespace mimrl {

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

#define MIMRL_REQUIRE(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      mimrl::set_error(__VA_ARGS__); \
      return 2;                      \
    }                                \
  } while (0)

// ---- small device math ----------------------------------------------------
__device__ __forceinline__ float softplusf(float z) {
  // log(1 + exp(z)), stable on both tails
  return fmaxf(z, 0.f) + log1pf(__expf(-fabsf(z)));
}
__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + __expf(-z)); }

// Gaussian cdf and pdf of z for the exact-erf GELU (Utils.py:88: GELU(z) = z Phi(z), GELU'(z) = Phi(z) + z phi(z)).
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 absolute, i.e. fp32 rounding level) instead of libm's erff: a
// dozen instructions and two MUFU ops against ~60 -- the CubeMLP epilogues evaluate 7e7 of these per forward pass and
// are bound by instruction issue.  The tail is formed without cancellation (Phi(z) = q for z < 0, 1 - q otherwise), and
// the exponential exp(-z^2 / 2) is shared with the pdf.
__device__ __forceinline__ void gauss_cdf_pdf(float z, float &cdf, float &pdf) {
  const float ax = fabsf(z) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float q = 0.5f * t * poly * e;
  cdf = z < 0.f ? q : 1.f - q;
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_fwd(float z) {
  float cdf, pdf;
  gauss_cdf_pdf(z, cdf, pdf);
  return z * cdf;
}
__device__ __forceinline__ float gelu_bwd(float z) {
  float cdf, pdf;
  gauss_cdf_pdf(z, cdf, pdf);
  return fmaf(z, pdf, cdf);
}

// merge two (max, sum-of-exp) pairs; (-inf, 0) is the identity
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
  float mn = fmaxf(m, m2);
  if (mn == -INFINITY) {
    m = mn;
    s = 0.f;
    return;
  }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}
__device__ __forceinline__ void lse_merge_d(double &m, double &s, double m2, double s2) {
  double mn = fmax(m, m2);
  if (mn == -INFINITY) {
    m = mn;
    s = 0.0;
    return;
  }
  s = s * exp(m - mn) + s2 * exp(m2 - mn);
  m = mn;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// k-NN candidate (fp32 filter distance, key index) and the tensor-core filter stage (knn_tc.cu)
struct KnnCand {
  float d;
  int idx;
};
struct KnnTcPlan {
  int splits, tiles_per_split, n_lists, list_len;
  size_t off_absmax, off_qn, off_q_hi, off_q_lo, off_k_hi, off_k_lo, off_pub, off_tscale, off_qscale, bytes;
};
bool knn_tc_supported(int n_keys, int n_queries, int width, int list_len);
KnnTcPlan knn_tc_plan(int n_keys, int n_queries, int width, int list_len);
// prepared keys of the tensor-core filter: fp16 hi / lo planes [n_keys, 128] and 1 / scale per 128-key tile
struct KnnTcKeys {
  void *hi, *lo;
  float *tile_inv_scale;
};
KnnTcKeys knn_tc_keys_in_workspace(const KnnTcPlan &plan, unsigned char *ws);
// per query and list: list_len candidates sorted by (d, idx), padded with (+inf, -1); lists = splits * 2
// one pass over the keys: squared row norms -> key_norms, fp16 hi / lo split with a scale per 128-key tile -> out
int knn_tc_prepare_keys(const float *keys, int n_keys, int width, const KnnTcKeys &out, float *key_norms, cudaStream_t st);
// after knn_tc_prepare_keys (and after excluded rows got +inf norms)
int knn_filter_tc(const KnnTcKeys &keys, const float *key_norms, int n_keys, int width, const float *queries,
                  int n_queries, const KnnTcPlan &plan, unsigned char *ws, KnnCand *cand, cudaStream_t st,
                  bool q_ready = false);
// queries = rows `ids` of the key pool: gather + squared norms + per-tile scale + fp16 split in one launch (then
// knn_filter_tc with q_ready)
int knn_tc_gather_queries(const float *keys, int width, const int64_t *ids, int n_queries, const KnnTcPlan &plan,
                          unsigned char *ws, float *q_out, cudaStream_t st);

// width-1 pools on the kd_tree route (knn_1d.cu): sort + per-query two-sided walk
bool knn1d_supported(int width, int exact_form);
size_t knn1d_workspace_bytes(int n_keys);
int knn1d_search(const float *keys, int n_keys, int64_t key_offset, const float *queries, int n_queries,
                 const int64_t *excluded, int n_excluded, int k, int64_t *nbr_orig, double *nbr_dist, unsigned char *ws,
                 cudaStream_t st);

// Partial-statistics layout shared by every score sweep (FFMA and tcgen05):
// part[(split * n_own + row) * 3 + {0,1,2}] = {max, sum, softplus-sum}
int combine_row_stats(const float *part, int n_splits, int n_own, float *row_max, float *row_sum, float *row_sp,
                      cudaStream_t st);

// implemented in sep_tc.cu; returns 0 and sets *handled = 1 when it ran
int sep_row_stats_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                     int flags, float *row_max, float *row_sum, float *row_sp, void *ws, size_t ws_bytes,
                     cudaStream_t st, const float *row_lse = nullptr, const float *row_sigma = nullptr);
int sep_weighted_sum_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                        int family, int include_diag, const float *shift, int shift_by_swept, const float *coef,
                        const float *dcoef, float *out, void *ws, size_t ws_bytes, cudaStream_t st,
                        float *row_sum = nullptr);
int sep_online_forward_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                          int include_diag, float *row_ref, float *wsum, float *row_sum, void *ws, size_t ws_bytes,
                          cudaStream_t st);
size_t sep_tc_workspace_bytes(int n_own, int n_all, int embed);
bool sep_tc_supported(int n_own, int n_all, int embed);

}  // namespace mimrl
