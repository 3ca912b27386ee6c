// Separable-critic score sweeps on the fp32 CUDA cores.
//
// General-width path (embed <= 256) of the two sweeps declared in
// include/mimrl_b200.h (mimrl_sep_row_stats / mimrl_sep_weighted_sum); the
// tcgen05 path in sep_tc.cu serves embed <= 128 (the only width Model.py:285
// ever builds).  Both replace VMI.py:57 + the bound reductions of
// VMI.py:136-198 without ever writing the B x B score matrix.
//
// Tiling: a CTA of 256 threads owns 64 rows (full embed width resident in
// shared memory, transposed so the inner product reads are conflict free) and
// streams 64-column tiles of the swept operand.  Thread (ty,tx) of the 16x16
// grid computes the 4x4 scores S[ty+16i][tx+16j].
#include "common.cuh"

namespace mimrl {

namespace {

constexpr int kTile = 64;
constexpr int kThreads = 256;
constexpr int kPad = 65;

template <int EP>
__device__ __forceinline__ void load_tile_transposed(float *dst, const float *__restrict__ src, int row0, int n,
                                                     int embed, int tid) {
  for (int idx = tid; idx < kTile * EP; idx += kThreads) {
    int r = idx / EP, e = idx - r * EP;
    int gr = row0 + r;
    float v = (gr < n && e < embed) ? __ldg(src + (size_t)gr * embed + e) : 0.f;
    dst[e * kPad + r] = v;
  }
}

template <int EP>
__device__ __forceinline__ void score_tile(const float *As, const float *Xs, int ty, int tx, float (&s)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
  for (int e = 0; e < EP; ++e) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[e * kPad + ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Xs[e * kPad + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
  }
}

template <int EP>
__global__ void __launch_bounds__(kThreads)
sep_row_stats_ffma_kernel(const float *__restrict__ own, const float *__restrict__ all, int n_own, int n_all,
                          int embed, int own_offset, int flags, float *__restrict__ part, int tiles_per_split) {
  extern __shared__ float smem[];
  float *As = smem;
  float *Xs = smem + EP * kPad;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * kTile;
  const int split = blockIdx.y;
  const bool clamp = flags & MIMRL_STAT_CLAMP, want_sp = flags & MIMRL_STAT_SOFTPLUS;

  load_tile_transposed<EP>(As, own, row0, n_own, embed, tid);
  float m[4], s[4], sp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = -INFINITY, s[i] = 0.f, sp[i] = 0.f;

  const int n_tiles = (n_all + kTile - 1) / kTile;
  const int t0 = split * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
  for (int t = t0; t < t1; ++t) {
    __syncthreads();
    load_tile_transposed<EP>(Xs, all, t * kTile, n_all, embed, tid);
    __syncthreads();
    float sc[4][4];
    score_tile<EP>(As, Xs, ty, tx, sc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = own_offset + row0 + ty + 16 * i;
      float v[4], tmax = -INFINITY;
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = t * kTile + tx + 16 * j;
        ok[j] = gc < n_all && gc != gr;
        const float z = sc[i][j];
        v[j] = clamp ? fminf(fmaxf(z, -1.f), 1.f) : z;
        if (ok[j]) {
          tmax = fmaxf(tmax, v[j]);
          if (want_sp) sp[i] += softplusf(z);
        }
      }
      if (tmax > -INFINITY) {
        const float mn = fmaxf(m[i], tmax);
        float acc = s[i] * __expf(m[i] - mn);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) acc += __expf(v[j] - mn);
        m[i] = mn;
        s[i] = acc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m[i], off);
      float s2 = __shfl_xor_sync(0xffffffffu, s[i], off);
      lse_merge(m[i], s[i], m2, s2);
      sp[i] += __shfl_xor_sync(0xffffffffu, sp[i], off);
    }
    const int r = row0 + ty + 16 * i;
    if (tx == 0 && r < n_own) {
      float *p = part + ((size_t)split * n_own + r) * 3;
      p[0] = m[i];
      p[1] = s[i];
      p[2] = sp[i];
    }
  }
}

template <int EP>
__global__ void __launch_bounds__(kThreads)
sep_weighted_sum_ffma_kernel(const float *__restrict__ own, const float *__restrict__ all, int n_own, int n_all,
                             int embed, int own_offset, int family, int include_diag,
                             const float *__restrict__ shift, int shift_by_swept, float *__restrict__ part,
                             int tiles_per_split) {
  extern __shared__ float smem[];
  float *As = smem;
  float *Xs = smem + EP * kPad;
  float *Ws = smem + 2 * EP * kPad;
  constexpr int NK = EP / 16;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * kTile;
  const int split = blockIdx.y;

  load_tile_transposed<EP>(As, own, row0, n_own, embed, tid);
  float acc[4][NK];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[i][k] = 0.f;
  float rshift[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty + 16 * i;
    rshift[i] = (!shift_by_swept && family == MIMRL_WEIGHT_EXP && r < n_own) ? shift[r] : 0.f;
  }

  const int n_tiles = (n_all + kTile - 1) / kTile;
  const int t0 = split * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
  for (int t = t0; t < t1; ++t) {
    __syncthreads();
    load_tile_transposed<EP>(Xs, all, t * kTile, n_all, embed, tid);
    __syncthreads();
    float sc[4][4];
    score_tile<EP>(As, Xs, ty, tx, sc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = t * kTile + tx + 16 * j;
      const float cshift = (shift_by_swept && family == MIMRL_WEIGHT_EXP && gc < n_all) ? shift[gc] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int lr = row0 + ty + 16 * i;
        const int gr = own_offset + lr;
        const bool ok = gc < n_all && lr < n_own && (include_diag || gc != gr);
        float w = 0.f;
        if (ok) w = family == MIMRL_WEIGHT_EXP ? __expf(sc[i][j] - (shift_by_swept ? cshift : rshift[i]))
                                               : sigmoidf(sc[i][j]);
        Ws[(tx + 16 * j) * kPad + ty + 16 * i] = w;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < kTile; ++c) {
      float wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = Ws[c * kPad + ty + 16 * i];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const float xv = Xs[(tx + 16 * k) * kPad + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][k] = fmaf(wv[i], xv, acc[i][k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty + 16 * i;
    if (r >= n_own) continue;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int e = tx + 16 * k;
      if (e < embed) part[((size_t)split * n_own + r) * embed + e] = acc[i][k];
    }
  }
}

// out[r][e] = coef * sum_splits part[split][r][e] + dcoef[r] * all[own_offset + r][e]
__global__ void weighted_sum_reduce_kernel(const float *__restrict__ part, int n_splits, int n_own, int embed,
                                           const float *__restrict__ all, int own_offset,
                                           const float *__restrict__ coef, const float *__restrict__ dcoef,
                                           float *__restrict__ out) {
  const size_t total = (size_t)n_own * embed;
  const float c = coef[0];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / embed);
    float a = 0.f;
    for (int s = 0; s < n_splits; ++s) a += part[(size_t)s * total + idx];
    const float d = dcoef ? dcoef[r] * all[(size_t)(own_offset + r) * embed + (idx - (size_t)r * embed)] : 0.f;
    out[idx] = fmaf(c, a, d);
  }
}

__global__ void combine_row_stats_kernel(const float *__restrict__ part, int n_splits, int n_own,
                                         float *__restrict__ row_max, float *__restrict__ row_sum,
                                         float *__restrict__ row_sp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_own) return;
  float m = -INFINITY, s = 0.f, sp = 0.f;
  for (int k = 0; k < n_splits; ++k) {
    const float *p = part + ((size_t)k * n_own + r) * 3;
    lse_merge(m, s, p[0], p[1]);
    sp += p[2];
  }
  row_max[r] = m;
  row_sum[r] = s;
  if (row_sp) row_sp[r] = sp;
}

// diag[i] = own_i . all_{own_offset + i}; one warp per row
__global__ void sep_diag_kernel(const float *__restrict__ own, const float *__restrict__ all, int n_own, int n_all,
                                int embed, int own_offset, float *__restrict__ diag) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_own) return;
  const float *a = own + (size_t)w * embed;
  const float *b = all + (size_t)(own_offset + w) * embed;
  float acc = 0.f;
  for (int e = lane; e < embed; e += 32) acc = fmaf(a[e], b[e], acc);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) diag[w] = acc;
}

int pick_splits(int n_own, int n_all) {
  const int row_tiles = ceil_div(n_own, kTile), col_tiles = ceil_div(n_all, kTile);
  int splits = (2 * 148 + row_tiles - 1) / row_tiles;
  if (splits > col_tiles) splits = col_tiles;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  return splits;
}

template <int EP>
int launch_row_stats(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                     int flags, float *part, int splits, cudaStream_t st) {
  const size_t smem = (size_t)2 * EP * kPad * sizeof(float);
  cudaFuncSetAttribute(sep_row_stats_ffma_kernel<EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int col_tiles = ceil_div(n_all, kTile);
  dim3 grid(ceil_div(n_own, kTile), splits);
  sep_row_stats_ffma_kernel<EP><<<grid, kThreads, smem, st>>>(own, all, n_own, n_all, embed, own_offset, flags, part,
                                                             ceil_div(col_tiles, splits));
  return check_launch("sep_row_stats_ffma");
}

template <int EP>
int launch_weighted_sum(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                        int family, int include_diag, const float *shift, int shift_by_swept, float *part,
                        int splits, cudaStream_t st) {
  const size_t smem = (size_t)(2 * EP + kTile) * kPad * sizeof(float);
  cudaFuncSetAttribute(sep_weighted_sum_ffma_kernel<EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int col_tiles = ceil_div(n_all, kTile);
  dim3 grid(ceil_div(n_own, kTile), splits);
  sep_weighted_sum_ffma_kernel<EP><<<grid, kThreads, smem, st>>>(own, all, n_own, n_all, embed, own_offset, family,
                                                                include_diag, shift, shift_by_swept, part,
                                                                ceil_div(col_tiles, splits));
  return check_launch("sep_weighted_sum_ffma");
}

size_t ffma_workspace_bytes(int n_own, int n_all, int embed) {
  const int splits = pick_splits(n_own, n_all);
  size_t a = (size_t)splits * n_own * 3 * sizeof(float);
  size_t b = (size_t)splits * n_own * embed * sizeof(float);
  return (a > b ? a : b) + 256;
}

}  // namespace

int combine_row_stats(const float *part, int n_splits, int n_own, float *row_max, float *row_sum, float *row_sp,
                      cudaStream_t st) {
  combine_row_stats_kernel<<<ceil_div(n_own, 256), 256, 0, st>>>(part, n_splits, n_own, row_max, row_sum, row_sp);
  return check_launch("combine_row_stats");
}

}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_sep_workspace_bytes(int n_own, int n_all, int embed) {
  size_t a = ffma_workspace_bytes(n_own, n_all, embed);
  size_t b = sep_tc_workspace_bytes(n_own, n_all, embed);
  return a > b ? a : b;
}

static bool use_tc(int impl, int n_own, int n_all, int embed) {
  if (impl == MIMRL_IMPL_FFMA) return false;
  return sep_tc_supported(n_own, n_all, embed);
}

extern "C" int mimrl_sep_selected_impl(int n_own, int n_all, int embed, int impl) {
  return use_tc(impl, n_own, n_all, embed) ? MIMRL_IMPL_TCGEN05 : MIMRL_IMPL_FFMA;
}

extern "C" int mimrl_sep_row_stats(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                                   int own_offset, int flags, int impl, float *row_max, float *row_sum,
                                   float *row_sp, float *diag, void *workspace, size_t workspace_bytes,
                                   void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && embed > 0, "sep_row_stats: empty input (n_own=%d n_all=%d embed=%d)", n_own,
                n_all, embed);
  MIMRL_REQUIRE(embed <= 256, "sep_row_stats: embed=%d > 256 is not supported", embed);
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "sep_row_stats: row block [%d,%d) outside [0,%d)",
                own_offset, own_offset + n_own, n_all);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_sep_workspace_bytes(n_own, n_all, embed), "sep_row_stats: workspace too small");
  MIMRL_REQUIRE(impl != MIMRL_IMPL_TCGEN05 || sep_tc_supported(n_own, n_all, embed),
                "sep_row_stats: tcgen05 path needs embed <= 128 (got %d)", embed);
  MIMRL_REQUIRE(!(flags & MIMRL_STAT_MAXONLY) || use_tc(impl, n_own, n_all, embed),
                "sep_row_stats: MIMRL_STAT_MAXONLY exists on the tcgen05 path only");
  if (diag) {
    sep_diag_kernel<<<ceil_div(n_own * 32, 256), 256, 0, st>>>(own_emb, all_emb, n_own, n_all, embed, own_offset, diag);
    if (check_launch("sep_diag")) return 1;
  }
  if (use_tc(impl, n_own, n_all, embed))
    return sep_row_stats_tc(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, row_max, row_sum, row_sp,
                            workspace, workspace_bytes, st);
  const int splits = pick_splits(n_own, n_all);
  float *part = (float *)workspace;
  int rc;
  if (embed <= 16) rc = launch_row_stats<16>(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, part, splits, st);
  else if (embed <= 32) rc = launch_row_stats<32>(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, part, splits, st);
  else if (embed <= 64) rc = launch_row_stats<64>(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, part, splits, st);
  else if (embed <= 128) rc = launch_row_stats<128>(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, part, splits, st);
  else rc = launch_row_stats<256>(own_emb, all_emb, n_own, n_all, embed, own_offset, flags, part, splits, st);
  if (rc) return rc;
  return combine_row_stats(part, splits, n_own, row_max, row_sum, row_sp, st);
}

// Fused forward sweep (tcgen05 path only): one pass over the own x swept score tiles gives BOTH the row statistic
// row_sum[i] = sum_{j != i} exp(S_ij - shift[i]) and the weighted sum wsum[i] = sum_j w_ij all_j (w includes the
// diagonal when include_diag).  With shift = an approximate row maximum (MIMRL_STAT_MAXONLY pre-pass) this replaces the
// exact statistics sweep AND the owned-row gradient sweep of the exp-family bounds: 4 1/3 tensor-core units instead of 5.
extern "C" int mimrl_sep_fused_forward(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                                       int own_offset, int include_diag, const float *shift, float *wsum, float *row_sum,
                                       void *workspace, size_t workspace_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && embed > 0 && shift && wsum && row_sum, "sep_fused_forward: bad arguments");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "sep_fused_forward: row block outside the batch");
  MIMRL_REQUIRE(sep_tc_supported(n_own, n_all, embed), "sep_fused_forward: tcgen05 path needs embed <= 128 (got %d)", embed);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_sep_workspace_bytes(n_own, n_all, embed), "sep_fused_forward: workspace too small");
  return sep_weighted_sum_tc(own_emb, all_emb, n_own, n_all, embed, own_offset, MIMRL_WEIGHT_EXP, include_diag, shift, 0,
                             /*coef = 1*/ nullptr, nullptr, wsum, workspace, workspace_bytes, st, row_sum);
}

// Online-softmax forward sweep (tcgen05 path only): like mimrl_sep_fused_forward, but the reference point is found
// by the sweep itself (running row maximum, lazily rescaled accumulators) and returned in row_ref.  No pre-pass.
extern "C" int mimrl_sep_online_forward(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                                        int own_offset, int include_diag, float *row_ref, float *wsum, float *row_sum,
                                        float *diag, void *workspace, size_t workspace_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && embed > 0 && row_ref && wsum && row_sum, "sep_online_forward: bad arguments");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "sep_online_forward: row block outside the batch");
  MIMRL_REQUIRE(sep_tc_supported(n_own, n_all, embed), "sep_online_forward: tcgen05 path needs embed <= 128 (got %d)", embed);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_sep_workspace_bytes(n_own, n_all, embed), "sep_online_forward: workspace too small");
  if (diag) {
    sep_diag_kernel<<<ceil_div(n_own * 32, 256), 256, 0, st>>>(own_emb, all_emb, n_own, n_all, embed, own_offset, diag);
    if (check_launch("sep_diag")) return 1;
  }
  return sep_online_forward_tc(own_emb, all_emb, n_own, n_all, embed, own_offset, include_diag, row_ref, wsum, row_sum,
                               workspace, workspace_bytes, st);
}

// Second forward sweep of the interpolated bound (VMI.py:201-250, tcgen05 path only): with p_ij = exp(S_ij - row_lse[i]),
// over the off-diagonal columns  row_q[i] = sum_j p_ij / (1 - row_sigma[i] p_ij),  row_t[i] = sum_j log(1 - row_sigma[i] p_ij).
extern "C" int mimrl_sep_interp_stats(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                                      int own_offset, const float *row_lse, const float *row_sigma, float *row_q,
                                      float *row_t, void *workspace, size_t workspace_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && embed > 0 && row_lse && row_sigma && row_q && row_t, "sep_interp_stats: bad arguments");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "sep_interp_stats: row block outside the batch");
  MIMRL_REQUIRE(sep_tc_supported(n_own, n_all, embed), "sep_interp_stats: tcgen05 path needs embed <= 128 (got %d)", embed);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_sep_workspace_bytes(n_own, n_all, embed), "sep_interp_stats: workspace too small");
  // row_q doubles as the (all-zero) row_max output of the shared combine step, which writes row_max before row_sum
  return sep_row_stats_tc(own_emb, all_emb, n_own, n_all, embed, own_offset, 0, row_q, row_q, row_t, workspace,
                          workspace_bytes, st, row_lse, row_sigma);
}

extern "C" int mimrl_sep_weighted_sum(const float *own_emb, const float *all_emb, int n_own, int n_all, int embed,
                                      int own_offset, int weight_family, int include_diag, const float *shift,
                                      int shift_by_swept, const float *coef, const float *dcoef, int impl,
                                      float *out, void *workspace, size_t workspace_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && embed > 0, "sep_weighted_sum: empty input");
  MIMRL_REQUIRE(embed <= 256, "sep_weighted_sum: embed=%d > 256 is not supported", embed);
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "sep_weighted_sum: row block outside the batch");
  MIMRL_REQUIRE(weight_family == MIMRL_WEIGHT_EXP || weight_family == MIMRL_WEIGHT_SIGMOID ||
                    weight_family == MIMRL_WEIGHT_INTERP,
                "sep_weighted_sum: unknown weight family %d", weight_family);
  MIMRL_REQUIRE(weight_family == MIMRL_WEIGHT_SIGMOID || shift, "sep_weighted_sum: this family needs its shift / parameter vectors");
  MIMRL_REQUIRE(weight_family != MIMRL_WEIGHT_INTERP || use_tc(impl, n_own, n_all, embed),
                "sep_weighted_sum: MIMRL_WEIGHT_INTERP exists on the tcgen05 path only (embed <= 128)");
  MIMRL_REQUIRE(workspace_bytes >= mimrl_sep_workspace_bytes(n_own, n_all, embed), "sep_weighted_sum: workspace too small");
  MIMRL_REQUIRE(impl != MIMRL_IMPL_TCGEN05 || sep_tc_supported(n_own, n_all, embed),
                "sep_weighted_sum: tcgen05 path needs embed <= 128 (got %d)", embed);
  if (use_tc(impl, n_own, n_all, embed))
    return sep_weighted_sum_tc(own_emb, all_emb, n_own, n_all, embed, own_offset, weight_family, include_diag, shift,
                               shift_by_swept, coef, dcoef, out, workspace, workspace_bytes, st);
  const int splits = pick_splits(n_own, n_all);
  float *part = (float *)workspace;
  int rc;
#define WS_ARGS own_emb, all_emb, n_own, n_all, embed, own_offset, weight_family, include_diag, shift, shift_by_swept, part, splits, st
  if (embed <= 16) rc = launch_weighted_sum<16>(WS_ARGS);
  else if (embed <= 32) rc = launch_weighted_sum<32>(WS_ARGS);
  else if (embed <= 64) rc = launch_weighted_sum<64>(WS_ARGS);
  else if (embed <= 128) rc = launch_weighted_sum<128>(WS_ARGS);
  else rc = launch_weighted_sum<256>(WS_ARGS);
#undef WS_ARGS
  if (rc) return rc;
  const size_t total = (size_t)n_own * embed;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  weighted_sum_reduce_kernel<<<blocks, 256, 0, st>>>(part, splits, n_own, embed, all_emb, own_offset, coef, dcoef, out);
  return check_launch("weighted_sum_reduce");
}
