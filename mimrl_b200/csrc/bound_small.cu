// Bound values and backward coefficients from per-row sweep statistics, plus
// the materialised-score entry points and the library's error plumbing.
//
// Closed forms: SURVEY.md Appendix A, each checked against the reference's
// autograd through oracle/vmi_oracle.py.  Reductions over rows run in fp64.
#include <stdarg.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace mimrl {

static std::mutex g_err_mu;
static std::string g_err = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = buf;
}

namespace {

constexpr int kFinThreads = 1024;

// result[] slots
enum { R_MI = 0, R_LOSS = 1, R_L = 2, R_MARG = 3, R_MA = 4, R_N = 5, R_GMAX = 6 };

__device__ __forceinline__ double block_sum(double v, double *sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) sh[0] = v;
  }
  __syncthreads();
  return sh[0];
}

__device__ __forceinline__ void block_lse(double &m, double &s, double *shm, double *shs) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    double m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, s, off);
    lse_merge_d(m, s, m2, s2);
  }
  __syncthreads();
  if (lane == 0) shm[w] = m, shs[w] = s;
  __syncthreads();
  if (w == 0) {
    m = lane < (blockDim.x >> 5) ? shm[lane] : -INFINITY;
    s = lane < (blockDim.x >> 5) ? shs[lane] : 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      double m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, s, off);
      lse_merge_d(m, s, m2, s2);
    }
    if (lane == 0) shm[0] = m, shs[0] = s;
  }
  __syncthreads();
  m = shm[0];
  s = shs[0];
}

__device__ __forceinline__ double softplus_d(double z) { return fmax(z, 0.0) + log1p(exp(-fabs(z))); }
__device__ __forceinline__ double logaddexp_d(double a, double b) {
  double mx = fmax(a, b);
  if (mx == -INFINITY) return mx;
  return mx + log1p(exp(-fabs(a - b)));
}

__global__ void __launch_bounds__(kFinThreads)
bound_finalize_kernel(int bound, const float *__restrict__ row_max, const float *__restrict__ row_sum,
                      const float *__restrict__ row_sp, const float *__restrict__ diag,
                      const float *__restrict__ base, int n, float *__restrict__ result) {
  __shared__ double sh[32], sh2[32];
  double sum_d = 0, sum_a = 0, sum_nce = 0, sum_sp = 0, sum_spneg = 0;
  double gm = -INFINITY, gs = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // a log-baseline only exists for TUBA (VMI.py:148-154); it is ignored for every other bound
    const double d = diag[i], a = (base && bound == MIMRL_BOUND_TUBA) ? (double)base[i] : 0.0;
    const double m = row_max[i], s = row_sum[i];
    sum_d += d;
    sum_a += a;
    if (bound == MIMRL_BOUND_INFONCE) {
      const double rl = s > 0.0 ? m + log(s) : -INFINITY;
      sum_nce += d - logaddexp_d(rl, d);
    } else {
      lse_merge_d(gm, gs, m - a, s);
    }
    if (bound == MIMRL_BOUND_JS_FGAN) {
      sum_sp += row_sp[i];
      sum_spneg += softplus_d(-d);
    }
  }
  sum_d = block_sum(sum_d, sh);
  sum_a = block_sum(sum_a, sh);
  sum_nce = block_sum(sum_nce, sh);
  sum_sp = block_sum(sum_sp, sh);
  sum_spneg = block_sum(sum_spneg, sh);
  block_lse(gm, gs, sh, sh2);
  if (threadIdx.x != 0) return;
  const double nn = n, n_off = nn * (nn - 1.0);
  const double L = gs > 0.0 ? gm + log(gs) : -INFINITY;
  double mi = 0, loss = 0, marg = 0, ma = 1.0;
  switch (bound) {
    case MIMRL_BOUND_DV:
    case MIMRL_BOUND_SMILE:
      mi = sum_d / nn - (L - log(n_off));
      loss = -mi;
      break;
    case MIMRL_BOUND_MINE: {
      mi = sum_d / nn - (L - log(n_off));
      const double mean_et = exp(L) / (nn * nn);
      ma = 0.99 + 0.01 * mean_et;
      loss = sum_d / nn - mean_et / ma;
      break;
    }
    case MIMRL_BOUND_TUBA:
      marg = exp(L - log(n_off));
      mi = 1.0 + (sum_d - sum_a) / nn - marg;
      loss = -mi;
      break;
    case MIMRL_BOUND_NWJ:
    case MIMRL_BOUND_JS:
      marg = exp(L - 1.0 - log(n_off));
      mi = 1.0 + (sum_d / nn - 1.0) - marg;
      loss = -mi;
      break;
    case MIMRL_BOUND_INFONCE:
      mi = log(nn) + sum_nce / nn;
      loss = -mi;
      break;
    case MIMRL_BOUND_JS_FGAN:
      mi = -sum_spneg / nn - sum_sp / n_off;
      loss = -mi;
      break;
    default:
      mi = loss = NAN;
  }
  result[R_MI] = (float)mi;
  result[R_LOSS] = (float)loss;
  result[R_L] = (float)L;
  result[R_MARG] = (float)marg;
  result[R_MA] = (float)ma;
  result[R_N] = (float)nn;
  result[R_GMAX] = (float)gm;   // max over rows of (row_max - baseline): the weight reference of the backward sweeps
}

__global__ void bound_backward_coef_kernel(int bound, const float *__restrict__ result,
                                           const float *__restrict__ grad, const float *__restrict__ row_max,
                                           const float *__restrict__ row_sum, const float *__restrict__ diag,
                                           const float *__restrict__ base, int n, float *__restrict__ coef,
                                           float *__restrict__ shift, float *__restrict__ dcoef,
                                           float *__restrict__ dbase) {
  const double nn = n, n_off = nn * (nn - 1.0);
  const double g_mi = grad[0], g_loss = grad[1];
  const double g = g_mi - g_loss;  // mi_loss = -mi for every bound but MINE
  const double L = result[R_L];
  double c = 0.0;
  switch (bound) {
    case MIMRL_BOUND_INFONCE: c = -g / nn; break;
    case MIMRL_BOUND_DV: c = -g; break;
    case MIMRL_BOUND_MINE: c = -g_mi - g_loss * exp(L) / (nn * nn * (double)result[R_MA]); break;
    case MIMRL_BOUND_TUBA: c = -g * exp(L - log(n_off)); break;
    case MIMRL_BOUND_NWJ: c = -g * exp(L - 1.0 - log(n_off)); break;
    default: c = -g / n_off; break;  // sigmoid family: js_fgan, js, smile
  }
  // Global-LSE bounds (dv, mine, tuba, nwj): the pair weights are exp(S_ij - a_i - L) ~ 1 / (n (n - 1)).  The tensor-core
  // sweeps carry a weight as w * 2^14 in an fp16 hi/lo pair, so at n = 65536 such a weight would be fp16-subnormal
  // (about 6 significant bits).  Reference the weights to the global maximum G >= every S_ij - a_i instead
  // (w' = exp(S_ij - a_i - G) in (0, 1], the dominant pairs keep full precision) and move exp(G - L) into the fp32
  // coefficient: the product coef' * w' is unchanged.
  const bool global_lse = bound == MIMRL_BOUND_DV || bound == MIMRL_BOUND_MINE || bound == MIMRL_BOUND_TUBA ||
                          bound == MIMRL_BOUND_NWJ;
  const double gmax = result[R_GMAX];
  // row_max may be the online forward's reference point, which can sit up to 4.5 below the true row maximum
  // (kOnlineTau in sep_tc.cu): the margin keeps every weight <= 1 either way
  const double ref = (global_lse && isfinite(gmax) && isfinite(L)) ? gmax + 4.5 : L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) coef[0] = (float)(global_lse ? c * exp(ref - L) : c);
  if (i >= n) return;
  const double d = diag[i], a = (base && bound == MIMRL_BOUND_TUBA) ? (double)base[i] : 0.0;
  double sh = ref, dc = g / nn;
  switch (bound) {
    case MIMRL_BOUND_INFONCE: {
      const double s = row_sum[i], m = row_max[i];
      sh = logaddexp_d(s > 0.0 ? m + log(s) : -INFINITY, d);
      break;
    }
    case MIMRL_BOUND_MINE: dc = (g_mi + g_loss) / nn; break;
    case MIMRL_BOUND_TUBA: sh = a + ref; break;
    case MIMRL_BOUND_JS_FGAN:
    case MIMRL_BOUND_JS:
    case MIMRL_BOUND_SMILE: dc = g / (nn * (1.0 + exp(d))); break;  // g * sigmoid(-d) / n
    default: break;
  }
  shift[i] = (float)sh;
  dcoef[i] = (float)dc;
  if (dbase && base) {
    // d/da_i = -sum_j G_ij = -(g/n + coef * sum_{j != i} exp(S_ij - a_i - L))
    const double rs = (double)row_sum[i] * exp((double)row_max[i] - a - L);
    dbase[i] = bound == MIMRL_BOUND_TUBA ? (float)(-(g / nn + c * rs)) : 0.f;
  }
}

// ---- materialised score matrix ---------------------------------------------
__global__ void __launch_bounds__(256)
scores_row_stats_kernel(const float *__restrict__ S, int n_rows, int n_cols, int own_offset, int flags,
                        float *__restrict__ row_max, float *__restrict__ row_sum, float *__restrict__ row_sp,
                        float *__restrict__ diag) {
  __shared__ float shm[8], shs[8], shp[8];
  const int r = blockIdx.x;
  const float *row = S + (size_t)r * n_cols;
  const int dcol = own_offset + r;
  const bool clamp = flags & MIMRL_STAT_CLAMP, want_sp = flags & MIMRL_STAT_SOFTPLUS;
  float m = -INFINITY, s = 0.f, sp = 0.f;
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    if (c == dcol) continue;
    const float z = row[c];
    const float v = clamp ? fminf(fmaxf(z, -1.f), 1.f) : z;
    lse_merge(m, s, v, 1.f);
    if (want_sp) sp += softplusf(z);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, s, off);
    lse_merge(m, s, m2, s2);
    sp += __shfl_xor_sync(0xffffffffu, sp, off);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) shm[w] = m, shs[w] = s, shp[w] = sp;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      lse_merge(m, s, shm[k], shs[k]);
      sp += shp[k];
    }
    row_max[r] = m;
    row_sum[r] = s;
    if (row_sp) row_sp[r] = sp;
    if (diag) diag[r] = row[dcol];
  }
}

__global__ void scores_grad_kernel(const float *__restrict__ S, int n_rows, int n_cols, int own_offset,
                                   int family, int include_diag, const float *__restrict__ shift,
                                   const float *__restrict__ coef, const float *__restrict__ dcoef,
                                   float *__restrict__ G) {
  const size_t total = (size_t)n_rows * n_cols;
  const float c = coef[0];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / n_cols), col = (int)(idx - (size_t)r * n_cols);
    const bool on_diag = col == own_offset + r;
    float g = 0.f;
    if (!on_diag || include_diag)
      g = c * (family == MIMRL_WEIGHT_EXP ? __expf(S[idx] - shift[r]) : sigmoidf(S[idx]));
    if (on_diag) g += dcoef[r];
    G[idx] = g;
  }
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" int mimrl_version(void) { return MIMRL_ABI_VERSION; }

extern "C" const char *mimrl_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_err;
  return copy.c_str();
}

extern "C" uint64_t mimrl_launch_count(void) { return g_launches.load(); }

extern "C" int mimrl_bound_weight_family(int bound, int *weight_family, int *include_diag, int *stat_flags) {
  int fam = MIMRL_WEIGHT_EXP, inc = 0, fl = 0;
  switch (bound) {
    case MIMRL_BOUND_INFONCE: inc = 1; break;
    case MIMRL_BOUND_DV:
    case MIMRL_BOUND_MINE:
    case MIMRL_BOUND_TUBA:
    case MIMRL_BOUND_NWJ: break;
    case MIMRL_BOUND_JS_FGAN: fam = MIMRL_WEIGHT_SIGMOID; fl = MIMRL_STAT_SOFTPLUS; break;
    case MIMRL_BOUND_JS: fam = MIMRL_WEIGHT_SIGMOID; break;
    case MIMRL_BOUND_SMILE: fam = MIMRL_WEIGHT_SIGMOID; fl = MIMRL_STAT_CLAMP; break;
    default:
      set_error("bound %d has no single-sweep form (interpolate runs on materialised scores)", bound);
      return 3;
  }
  if (weight_family) *weight_family = fam;
  if (include_diag) *include_diag = inc;
  if (stat_flags) *stat_flags = fl;
  return 0;
}

extern "C" int mimrl_bound_finalize(int bound, const float *row_max, const float *row_sum, const float *row_sp,
                                    const float *diag, const float *log_baseline, int n, float *result,
                                    void *stream) {
  MIMRL_REQUIRE(n > 0, "bound_finalize: n=%d", n);
  MIMRL_REQUIRE(bound >= 0 && bound < MIMRL_BOUND_INTERPOLATE, "bound_finalize: bound %d not supported here", bound);
  MIMRL_REQUIRE(bound != MIMRL_BOUND_JS_FGAN || row_sp, "bound_finalize: js_fgan needs row_sp");
  bound_finalize_kernel<<<1, kFinThreads, 0, (cudaStream_t)stream>>>(bound, row_max, row_sum, row_sp, diag,
                                                                    log_baseline, n, result);
  return check_launch("bound_finalize");
}

extern "C" int mimrl_bound_backward_coef(int bound, const float *result, const float *grad, const float *row_max,
                                         const float *row_sum, const float *diag, const float *log_baseline, int n,
                                         float *coef, float *shift, float *dcoef, float *dbaseline, void *stream) {
  MIMRL_REQUIRE(n > 0, "bound_backward_coef: n=%d", n);
  MIMRL_REQUIRE(bound >= 0 && bound < MIMRL_BOUND_INTERPOLATE, "bound_backward_coef: bound %d not supported here", bound);
  bound_backward_coef_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      bound, result, grad, row_max, row_sum, diag, log_baseline, n, coef, shift, dcoef, dbaseline);
  return check_launch("bound_backward_coef");
}

extern "C" int mimrl_scores_row_stats(const float *scores, int n_rows, int n_cols, int own_offset, int flags,
                                      float *row_max, float *row_sum, float *row_sp, float *diag, void *stream) {
  MIMRL_REQUIRE(n_rows > 0 && n_cols > 0, "scores_row_stats: empty score matrix");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_rows <= n_cols, "scores_row_stats: row block outside the matrix");
  scores_row_stats_kernel<<<n_rows, 256, 0, (cudaStream_t)stream>>>(scores, n_rows, n_cols, own_offset, flags, row_max,
                                                                   row_sum, row_sp, diag);
  return check_launch("scores_row_stats");
}

extern "C" int mimrl_scores_grad(const float *scores, int n_rows, int n_cols, int own_offset, int weight_family,
                                 int include_diag, const float *shift, const float *coef, const float *dcoef,
                                 float *grad_scores, void *stream) {
  MIMRL_REQUIRE(n_rows > 0 && n_cols > 0, "scores_grad: empty score matrix");
  const size_t total = (size_t)n_rows * n_cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scores_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scores, n_rows, n_cols, own_offset, weight_family,
                                                              include_diag, shift, coef, dcoef, grad_scores);
  return check_launch("scores_grad");
}
