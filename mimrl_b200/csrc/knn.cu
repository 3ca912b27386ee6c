// k-NN search of the conditional-MI sampler (reference Model.py:82-86, i.e.
// scikit-learn NearestNeighbors(metric='euclidean').kneighbors on the pool with
// the query rows removed).
//
// Two stages so that the result is bit-exact against scikit-learn's float64
// arithmetic while the O(m*N*width) work stays in fp32 tiles:
//   1. knn_filter_kernel: tiled fp32 distances ||q||^2 + ||z||^2 - 2 q.z, each
//      query keeps its K' = k + slack best keys per key-split in a sorted
//      shared-memory list (threshold test per pair, rare insertions).
//   2. knn_rerank_kernel: per query, merge the split lists, recompute the K'
//      survivors in float64 with scikit-learn's formula, order by
//      (distance, index) and emit the k nearest (ties -> lowest index).
// Excluded keys (the rows drawn as queries, Model.py:83-84) get norm = +inf.
#include <stdlib.h>

#include "common.cuh"

namespace mimrl {
namespace {

constexpr int kQT = 64;        // queries per CTA
constexpr int kKT = 128;       // keys per tile
constexpr int kWC = 32;        // width chunk
constexpr int kQS = 68;        // padded strides (multiples of 4 floats for LDS.128)
constexpr int kKS = 132;
constexpr int kThreads = 256;
constexpr int kMaxList = 64;   // K' upper bound
constexpr int kSlack = 4;    // fp32-class filter distances: the float64 re-rank only has to fix the ORDER of near-ties

using Cand = KnnCand;

__device__ __forceinline__ bool cand_less(float d1, int i1, float d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

__global__ void key_norms_kernel(const float *__restrict__ keys, int n, int width, float *__restrict__ kn) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const float *r = keys + (size_t)w * width;
  float acc = 0.f;
  for (int e = lane; e < width; e += 32) acc = fmaf(r[e], r[e], acc);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) kn[w] = acc;
}

__global__ void mark_excluded_kernel(const int64_t *__restrict__ ids, int n_ids, int64_t key_offset, int n_keys,
                                     float *__restrict__ kn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ids) return;
  const int64_t l = ids[i] - key_offset;
  if (l >= 0 && l < n_keys) kn[l] = INFINITY;
}

__global__ void gather_rows_kernel(const float *__restrict__ src, int width, const int64_t *__restrict__ idx,
                                   int n_idx, int repeat, int out_width, float *__restrict__ out) {
  const size_t total = (size_t)n_idx * repeat * out_width;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / out_width;
    const int c = (int)(t - r * out_width);
    out[t] = src[(size_t)idx[r / repeat] * width + (c % width)];
  }
}

// grid: (query tiles, key splits).  DIRECT: accumulate (q - z)^2 instead of q.z -- for narrow keys (the label
// pools, width 1) neighbours are so close that the GEMM form |q|^2 + |z|^2 - 2 q.z cancels to noise in fp32.
template <bool DIRECT>
__global__ void __launch_bounds__(kThreads)
knn_filter_kernel(const float *__restrict__ keys, const float *__restrict__ kn, int n_keys, int width,
                  const float *__restrict__ queries, int n_queries, int list_len, int tiles_per_split,
                  Cand *__restrict__ cand) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *Qs = reinterpret_cast<float *>(smem_raw);                 // [kWC][kQS]
  float *Ks = Qs + kWC * kQS;                                      // [kWC][kKS]
  float *qn = Ks + kWC * kKS;                                      // [kQT]
  float *tau_d = qn + kQT;                                         // [kQT]
  int *tau_i = reinterpret_cast<int *>(tau_d + kQT);               // [kQT]
  int *cnt = tau_i + kQT;                                          // [kQT]
  int *len = cnt + kQT;                                            // [kQT]
  int *any = len + kQT;                                            // [4]
  Cand *lists = reinterpret_cast<Cand *>(any + 4);                 // [kQT][kMaxList]
  Cand *buf = lists + kQT * kMaxList;                              // [kQT][kKT]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * kQT;
  const int split = blockIdx.y, n_splits = gridDim.y;
  const int wpad = (width + 3) & ~3;

  if (tid < kQT) {
    tau_d[tid] = INFINITY;
    tau_i[tid] = 0x7fffffff;
    cnt[tid] = 0;
    len[tid] = 0;
    // query norms in fp32
    float a = 0.f;
    if (q0 + tid < n_queries) {
      const float *r = queries + (size_t)(q0 + tid) * width;
      for (int e = 0; e < width; ++e) a = fmaf(r[e], r[e], a);
    }
    qn[tid] = a;
  }
  if (tid == 0) any[0] = 0;

  const int n_tiles = (n_keys + kKT - 1) / kKT;
  const int t0 = split * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
  for (int t = t0; t < t1; ++t) {
    const int k0 = t * kKT;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int w0 = 0; w0 < wpad; w0 += kWC) {
      const int wc = min(kWC, wpad - w0);
      __syncthreads();
      for (int idx = tid; idx < kQT * wc; idx += kThreads) {
        const int r = idx / wc, e = idx - r * wc;
        const int gq = q0 + r, ge = w0 + e;
        Qs[e * kQS + r] = (gq < n_queries && ge < width) ? __ldg(queries + (size_t)gq * width + ge) : 0.f;
      }
      for (int idx = tid; idx < kKT * wc; idx += kThreads) {
        const int r = idx / wc, e = idx - r * wc;
        const int gk = k0 + r, ge = w0 + e;
        Ks[e * kKS + r] = (gk < n_keys && ge < width) ? __ldg(keys + (size_t)gk * width + ge) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int e = 0; e < wc; ++e) {
        const float4 a = *reinterpret_cast<const float4 *>(Qs + e * kQS + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(Ks + e * kKS + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4 *>(Ks + e * kKS + 64 + tx * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (DIRECT) {
              const float df = av[i] - bv[j];
              acc[i][j] = fmaf(df, df, acc[i][j]);
            } else {
              acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
          }
      }
    }
    // threshold test: rows ty*4+i, keys tx*4+j (j<4) and 64+tx*4+j-4
    bool pushed = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kl = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      const int gk = k0 + kl;
      const float knv = gk < n_keys ? __ldg(kn + gk) : INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ql = ty * 4 + i;
        const float d = DIRECT ? acc[i][j] : fmaxf(qn[ql] + knv - 2.f * acc[i][j], 0.f);
        if (d <= tau_d[ql] && knv < INFINITY && q0 + ql < n_queries) {
          const int pos = atomicAdd(&cnt[ql], 1);
          buf[ql * kKT + pos] = Cand{d, gk};
          pushed = true;
        }
      }
    }
    if (pushed) any[0] = 1;
    __syncthreads();
    if (any[0]) {
      if (tid < kQT) {
        const int c = cnt[tid];
        Cand *lst = lists + tid * kMaxList;
        int n = len[tid];
        for (int u = 0; u < c; ++u) {
          const Cand x = buf[tid * kKT + u];
          if (n == list_len) {
            if (!cand_less(x.d, x.idx, lst[n - 1].d, lst[n - 1].idx)) continue;
            --n;
          }
          int p = n;
          while (p > 0 && cand_less(x.d, x.idx, lst[p - 1].d, lst[p - 1].idx)) {
            lst[p] = lst[p - 1];
            --p;
          }
          lst[p] = x;
          ++n;
        }
        len[tid] = n;
        cnt[tid] = 0;
        if (n == list_len) {
          tau_d[tid] = lst[n - 1].d;
          tau_i[tid] = lst[n - 1].idx;
        }
      }
      __syncthreads();
      if (tid == 0) any[0] = 0;
    }
  }
  __syncthreads();
  // emit lists (padded with +inf)
  for (int idx = tid; idx < kQT * list_len; idx += kThreads) {
    const int ql = idx / list_len, u = idx - ql * list_len;
    if (q0 + ql >= n_queries) continue;
    Cand c = u < len[ql] ? lists[ql * kMaxList + u] : Cand{INFINITY, -1};
    cand[((size_t)(q0 + ql) * n_splits + split) * list_len + u] = c;
  }
}

// one CTA (128 threads) per query.  cand holds n_lists sorted lists of list_len candidates for this query.
__global__ void __launch_bounds__(128)
knn_rerank_kernel(const float *__restrict__ keys, int n_keys, int width, int64_t key_offset,
                  const float *__restrict__ queries, int n_queries, const Cand *__restrict__ cand, int n_lists,
                  int list_len, int k, int exact_form, int64_t *__restrict__ nbr_orig,
                  double *__restrict__ nbr_dist, const int64_t *__restrict__ comp_ids, int n_comp_ids,
                  int64_t *__restrict__ nbr_comp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_cand = n_lists * list_len;
  Cand *cs = reinterpret_cast<Cand *>(smem_raw);                   // [n_cand]
  int *sel = reinterpret_cast<int *>(cs + n_cand);                 // [list_len]
  double *dd = reinterpret_cast<double *>(sel + ((list_len + 1) & ~1));  // [list_len]
  __shared__ float w_d[4];
  __shared__ int w_i[4], w_t[4];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_cand; i += blockDim.x) cs[i] = cand[(size_t)q * n_cand + i];
  __syncthreads();
  // 1. the list_len best by (fp32 distance, index): multi-way merge, thread t walks list t
  int head = 0;
  for (int it = 0; it < list_len; ++it) {
    float d = INFINITY;
    int idx = -1, t = tid;
    if (tid < n_lists && head < list_len) {
      const Cand c = cs[tid * list_len + head];
      if (c.idx >= 0) d = c.d, idx = c.idx;
    }
    if (idx < 0) idx = 0x7fffffff;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float d2 = __shfl_xor_sync(0xffffffffu, d, off);
      const int i2 = __shfl_xor_sync(0xffffffffu, idx, off), t2 = __shfl_xor_sync(0xffffffffu, t, off);
      if (cand_less(d2, i2, d, idx)) d = d2, idx = i2, t = t2;
    }
    if (lane == 0) w_d[warp] = d, w_i[warp] = idx, w_t[warp] = t;
    __syncthreads();
    d = w_d[0], idx = w_i[0], t = w_t[0];
#pragma unroll
    for (int w = 1; w < 4; ++w)
      if (cand_less(w_d[w], w_i[w], d, idx)) d = w_d[w], idx = w_i[w], t = w_t[w];
    if (tid == 0) sel[it] = idx == 0x7fffffff ? -1 : idx;
    if (tid == t && idx != 0x7fffffff) ++head;
    __syncthreads();
  }
  // 2. float64 distances of the survivors, one warp per candidate
  const float *qr = queries + (size_t)q * width;
  for (int c = warp; c < list_len; c += 4) {
    const int idx = sel[c];
    double dist = INFINITY;
    if (idx >= 0) {
      const float *kr = keys + (size_t)idx * width;
      if (exact_form) {
        double qn = 0, kn = 0, dot = 0;
        for (int e = lane; e < width; e += 32) {
          const double a = qr[e], b = kr[e];
          qn += a * a;
          kn += b * b;
          dot += a * b;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          qn += __shfl_xor_sync(0xffffffffu, qn, off);
          kn += __shfl_xor_sync(0xffffffffu, kn, off);
          dot += __shfl_xor_sync(0xffffffffu, dot, off);
        }
        dist = (qn + (-2.0 * dot)) + kn;
        dist = dist < 0.0 ? 0.0 : dist;
      } else {
        double acc = 0;
        for (int e = lane; e < width; e += 32) {
          const double t = (double)qr[e] - (double)kr[e];
          acc += t * t;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        dist = acc;
      }
    }
    if (lane == 0) dd[c] = dist;
  }
  __syncthreads();
  // 3. order by (float64 distance, index)
  for (int c = tid; c < list_len; c += blockDim.x) {
    const int idx = sel[c];
    if (idx < 0) continue;
    const double d = dd[c];
    int rank = 0;
    for (int j = 0; j < list_len; ++j) {
      const int ij = sel[j];
      if (ij < 0) continue;
      const double dj = dd[j];
      rank += (dj < d || (dj == d && ij < idx)) ? 1 : 0;
    }
    if (rank < k) {
      nbr_orig[(size_t)q * k + rank] = (int64_t)idx + key_offset;
      if (nbr_dist) nbr_dist[(size_t)q * k + rank] = d;
    }
  }
  // 4. nbr_comp = index with the query rows removed (what sklearn returns in the reference): nbr_orig minus the number
  // of removed ids below it, counted by the whole block for each of its k outputs
  if (nbr_comp) {
    __syncthreads();          // this block's nbr_orig entries are visible to all its threads
    for (int o = 0; o < k; ++o) {
      const int64_t v = nbr_orig[(size_t)q * k + o];
      int below = 0;
      for (int i = tid; i < n_comp_ids; i += blockDim.x) below += __ldg(comp_ids + i) < v ? 1 : 0;
#pragma unroll
      for (int off = 16; off; off >>= 1) below += __shfl_xor_sync(0xffffffffu, below, off);
      if (lane == 0) w_i[warp] = below;
      __syncthreads();
      if (tid == 0) nbr_comp[(size_t)q * k + o] = v - (w_i[0] + w_i[1] + w_i[2] + w_i[3]);
      __syncthreads();
    }
  }
}

// nbr_comp = nbr_orig - #(excluded ids < nbr_orig); one warp per output, the ids split over its lanes
__global__ void compact_index_kernel(const int64_t *__restrict__ nbr_orig, size_t n_out,
                                     const int64_t *__restrict__ ids, int n_ids, int64_t *__restrict__ nbr_comp) {
  const size_t o = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= n_out) return;
  const int64_t v = nbr_orig[o];
  int below = 0;
  for (int i = lane; i < n_ids; i += 32) below += __ldg(ids + i) < v ? 1 : 0;
#pragma unroll
  for (int off = 16; off; off >>= 1) below += __shfl_xor_sync(0xffffffffu, below, off);
  if (lane == 0) nbr_comp[o] = v - below;
}

struct Plan {
  int list_len, splits, tiles_per_split, n_lists;
  bool use_tc;
  KnnTcPlan tc;
  size_t off_kn, off_q, off_cand, off_tc, off_1d, total;
};

Plan make_plan(int n_keys, int n_queries, int width, int k) {
  Plan p;
  p.list_len = k + kSlack;
  if (p.list_len > kMaxList) p.list_len = kMaxList;
  // narrow keys (sklearn's kd_tree range, width <= 15) take the direct-difference CUDA-core filter
  p.use_tc = width > 15 && knn_tc_supported(n_keys, n_queries, width, p.list_len) && !getenv("MIMRL_KNN_FFMA");
  const int q_tiles = ceil_div(n_queries, kQT), k_tiles = ceil_div(n_keys, kKT);
  int splits = ceil_div(2 * 148, q_tiles);
  if (splits > 32) splits = 32;
  if (splits > k_tiles) splits = k_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = ceil_div(k_tiles, splits);
  p.splits = ceil_div(k_tiles, p.tiles_per_split);
  p.n_lists = p.splits;
  if (p.use_tc) {
    p.tc = knn_tc_plan(n_keys, n_queries, width, p.list_len);
    p.n_lists = p.tc.n_lists;
  }
  size_t o = 0;
  p.off_kn = o;
  o += ((((size_t)n_keys + 127) & ~(size_t)127) * sizeof(float) + 255) & ~(size_t)255;      // whole 128-key tiles (16-byte reads)
  p.off_q = o;
  o += ((size_t)n_queries * width * sizeof(float) + 255) & ~(size_t)255;
  p.off_cand = o;
  o += ((size_t)n_queries * p.n_lists * p.list_len * sizeof(Cand) + 255) & ~(size_t)255;
  p.off_tc = o;
  o += p.use_tc ? p.tc.bytes : 0;
  p.off_1d = o;
  o += width == 1 ? knn1d_workspace_bytes(n_keys) : 0;          // sorted pool of the width-1 route (knn_1d.cu)
  p.total = o;
  return p;
}

constexpr size_t kFilterSmem = (size_t)(kWC * kQS + kWC * kKS + 2 * kQT) * 4 + (size_t)(3 * kQT + 4) * 4 +
                               (size_t)kQT * kMaxList * sizeof(Cand) + (size_t)kQT * kKT * sizeof(Cand);

// Fitted pool (mimrl_knn_fit): what a search computes from the keys alone -- clean squared norms, per-tile scales, fp16
// hi / lo planes -- so that repeated searches of one pool (Model.py:323,329 search T_F_all twice per step, and the pools
// are constant over an epoch) skip the preparation pass.
struct FitLayout {
  size_t off_kn, off_tscale, off_hi, off_lo, total;
};
FitLayout fit_layout(int n_keys) {
  FitLayout f;
  size_t o = 0;
  f.off_kn = o, o += ((((size_t)n_keys + 127) & ~(size_t)127) * sizeof(float) + 255) & ~(size_t)255;
  f.off_tscale = o, o += ((size_t)ceil_div(n_keys, 128) * 4 + 255) & ~(size_t)255;
  f.off_hi = o, o += ((size_t)n_keys * 128 * 2 + 255) & ~(size_t)255;
  f.off_lo = o, o += ((size_t)n_keys * 128 * 2 + 255) & ~(size_t)255;
  f.total = o;
  return f;
}
bool fit_applies(int n_keys, int width) { return width > 15 && knn_tc_supported(n_keys, 1, width, 1) && !getenv("MIMRL_KNN_FFMA"); }
KnnTcKeys fitted_keys(const FitLayout &f, unsigned char *fitted) {
  KnnTcKeys k;
  k.hi = fitted + f.off_hi, k.lo = fitted + f.off_lo, k.tile_inv_scale = reinterpret_cast<float *>(fitted + f.off_tscale);
  return k;
}

int knn_core(const float *keys, int n_keys, int width, int64_t key_offset, const float *queries, int n_queries,
             const int64_t *excluded, int n_excluded, int k, int exact_form, int64_t *nbr_orig, double *nbr_dist,
             const Plan &p, unsigned char *ws, cudaStream_t st, const unsigned char *fitted = nullptr,
             int64_t *nbr_comp = nullptr, bool *comp_done = nullptr, bool q_ready = false) {
  if (knn1d_supported(width, exact_form))          // label pools: sort once, walk per query
    return knn1d_search(keys, n_keys, key_offset, queries, n_queries, excluded, n_excluded, k, nbr_orig, nbr_dist,
                        ws + p.off_1d, st);
  float *kn = reinterpret_cast<float *>(ws + p.off_kn);
  Cand *cand = reinterpret_cast<Cand *>(ws + p.off_cand);
  KnnTcKeys tck{};
  if (p.use_tc && fitted) {          // fitted pool: only the norms are copied (excluded rows are marked in the copy)
    const FitLayout f = fit_layout(n_keys);
    tck = fitted_keys(f, const_cast<unsigned char *>(fitted));
    cudaMemcpyAsync(kn, fitted + f.off_kn, (size_t)n_keys * sizeof(float), cudaMemcpyDeviceToDevice, st);
  } else if (p.use_tc) {          // norms, fp16 hi / lo split and per-tile scales in one pass over the keys
    tck = knn_tc_keys_in_workspace(p.tc, ws + p.off_tc);
    if (knn_tc_prepare_keys(keys, n_keys, width, tck, kn, st)) return 1;
  } else {
    key_norms_kernel<<<ceil_div(n_keys * 32, 256), 256, 0, st>>>(keys, n_keys, width, kn);
    if (check_launch("knn key_norms")) return 1;
  }
  if (n_excluded > 0) {
    mark_excluded_kernel<<<ceil_div(n_excluded, 256), 256, 0, st>>>(excluded, n_excluded, key_offset, n_keys, kn);
    if (check_launch("knn mark_excluded")) return 1;
  }
  if (p.use_tc) {
    if (knn_filter_tc(tck, kn, n_keys, width, queries, n_queries, p.tc, ws + p.off_tc, cand, st, q_ready)) return 1;
  } else {
    dim3 grid(ceil_div(n_queries, kQT), p.splits);
    if (width <= 15) {
      cudaFuncSetAttribute(knn_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFilterSmem);
      knn_filter_kernel<true><<<grid, kThreads, kFilterSmem, st>>>(keys, kn, n_keys, width, queries, n_queries,
                                                                  p.list_len, p.tiles_per_split, cand);
    } else {
      cudaFuncSetAttribute(knn_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFilterSmem);
      knn_filter_kernel<false><<<grid, kThreads, kFilterSmem, st>>>(keys, kn, n_keys, width, queries, n_queries,
                                                                   p.list_len, p.tiles_per_split, cand);
    }
    if (check_launch("knn_filter")) return 1;
  }
  const size_t rsmem = (size_t)p.n_lists * p.list_len * sizeof(Cand) + (size_t)((p.list_len + 1) & ~1) * 4 +
                       (size_t)p.list_len * 8;
  cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
  // (the compacted indices come out of the same kernel when the caller wants them: excluded = the removed query rows)
  knn_rerank_kernel<<<n_queries, 128, rsmem, st>>>(keys, n_keys, width, key_offset, queries, n_queries, cand, p.n_lists,
                                                  p.list_len, k, exact_form, nbr_orig, nbr_dist, excluded, n_excluded, nbr_comp);
  if (comp_done) *comp_done = nbr_comp != nullptr;
  return check_launch("knn_rerank");
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_knn_workspace_bytes(int n_keys, int n_queries, int width, int k) {
  if (n_keys <= 0 || n_queries <= 0 || width <= 0 || k <= 0) return 0;
  return make_plan(n_keys, n_queries, width, k).total + 256;
}

extern "C" int mimrl_gather_rows(const float *src, int n_src, int width, const int64_t *idx, int n_idx, int repeat,
                                 int out_width, float *out, void *stream) {
  MIMRL_REQUIRE(n_idx > 0 && repeat > 0 && width > 0 && out_width >= width, "gather_rows: bad sizes");
  const size_t total = (size_t)n_idx * repeat * out_width;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, width, idx, n_idx, repeat, out_width, out);
  return check_launch("gather_rows");
}

// whole_pool: the keys are the complete pool, so sklearn's n_neighbors <= n_samples_fit applies here
// (sklearn/neighbors/_base.py:840-851).  A key SHARD may hold fewer than k keys (other shards fill in): the global
// check is the caller's (model.knn_search_sharded), unfilled slots come back as (-1, +huge).
static int knn_check(int n_keys, int width, int n_queries, int k, int n_excluded, bool whole_pool) {
  MIMRL_REQUIRE(n_keys > 0 && width > 0 && n_queries > 0 && k > 0, "knn_search: empty input");
  MIMRL_REQUIRE(k <= kMaxList, "knn_search: k=%d too large (max %d)", k, kMaxList);
  if (whole_pool)
    MIMRL_REQUIRE(k <= n_keys - n_excluded,
                  "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d", k, n_keys - n_excluded);
  return 0;
}

extern "C" int mimrl_knn_search_rows(const float *keys, int n_keys, int width, int64_t key_index_offset,
                                     const float *queries, int n_queries, const int64_t *excluded_sorted,
                                     int n_excluded, int k, int exact_form, int64_t *nbr_orig, double *nbr_dist,
                                     void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = knn_check(n_keys, width, n_queries, k, 0, false)) return rc;
  const Plan p = make_plan(n_keys, n_queries, width, k);
  MIMRL_REQUIRE(workspace_bytes >= p.total, "knn_search_rows: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // fewer than k reachable keys on this shard is legal (other shards fill in): pre-fill with "none"
  cudaMemsetAsync(nbr_orig, 0xff, (size_t)n_queries * k * sizeof(int64_t), st);
  if (nbr_dist) cudaMemsetAsync(nbr_dist, 0x7f, (size_t)n_queries * k * sizeof(double), st);  // huge positive
  return knn_core(keys, n_keys, width, key_index_offset, queries, n_queries, excluded_sorted, n_excluded, k, exact_form,
                  nbr_orig, nbr_dist, p, (unsigned char *)workspace, st);
}

extern "C" size_t mimrl_knn_fit_bytes(int n_keys, int width) {
  if (n_keys <= 0 || width <= 0 || !fit_applies(n_keys, width)) return 0;
  return fit_layout(n_keys).total;
}

extern "C" int mimrl_knn_fit(const float *keys, int n_keys, int width, void *fitted, size_t fitted_bytes, void *stream) {
  MIMRL_REQUIRE(keys && n_keys > 0 && width > 0 && fit_applies(n_keys, width),
                "knn_fit: nothing to fit for a %d x %d pool (mimrl_knn_fit_bytes returns 0)", n_keys, width);
  const FitLayout f = fit_layout(n_keys);
  MIMRL_REQUIRE(fitted && fitted_bytes >= f.total, "knn_fit: buffer too small");
  unsigned char *fb = static_cast<unsigned char *>(fitted);
  return knn_tc_prepare_keys(keys, n_keys, width, fitted_keys(f, fb), reinterpret_cast<float *>(fb + f.off_kn),
                             (cudaStream_t)stream);
}

static int knn_search_impl(const float *keys, int n_keys, int width, const unsigned char *fitted, const int64_t *query_ids,
                           int n_queries, int k, int exact_form, int64_t *nbr_orig, int64_t *nbr_comp, double *nbr_dist,
                           void *workspace, size_t workspace_bytes, void *stream);

extern "C" int mimrl_knn_search_fitted(const float *keys, int n_keys, int width, const void *fitted, size_t fitted_bytes,
                                       const int64_t *query_ids, int n_queries, int k, float radius, int exact_form,
                                       int64_t *nbr_orig, int64_t *nbr_comp, double *nbr_dist, void *workspace,
                                       size_t workspace_bytes, void *stream) {
  (void)radius;
  MIMRL_REQUIRE(fitted && n_keys > 0 && width > 0 && fit_applies(n_keys, width) && fitted_bytes >= fit_layout(n_keys).total,
                "knn_search_fitted: not a buffer filled by mimrl_knn_fit for a %d x %d pool", n_keys, width);
  return knn_search_impl(keys, n_keys, width, static_cast<const unsigned char *>(fitted), query_ids, n_queries, k, exact_form,
                         nbr_orig, nbr_comp, nbr_dist, workspace, workspace_bytes, stream);
}

extern "C" int mimrl_knn_search(const float *keys, int n_keys, int width, const int64_t *query_ids, int n_queries,
                                int k, float radius, int exact_form, int64_t *nbr_orig, int64_t *nbr_comp,
                                double *nbr_dist, void *workspace, size_t workspace_bytes, void *stream) {
  (void)radius;  // Model.py:82 passes it to the constructor; kneighbors() never reads it
  return knn_search_impl(keys, n_keys, width, nullptr, query_ids, n_queries, k, exact_form, nbr_orig, nbr_comp, nbr_dist,
                         workspace, workspace_bytes, stream);
}

static int knn_search_impl(const float *keys, int n_keys, int width, const unsigned char *fitted, const int64_t *query_ids,
                           int n_queries, int k, int exact_form, int64_t *nbr_orig, int64_t *nbr_comp, double *nbr_dist,
                           void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = knn_check(n_keys, width, n_queries, k, n_queries, true)) return rc;
  const Plan p = make_plan(n_keys, n_queries, width, k);
  MIMRL_REQUIRE(workspace_bytes >= p.total, "knn_search: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  float *q = reinterpret_cast<float *>(ws + p.off_q);
  static const bool qprep_off = getenv("MIMRL_KNN_QPREP_OFF") != nullptr;
  const bool tc_route = p.use_tc && !knn1d_supported(width, exact_form) && !qprep_off;
  if (tc_route) {          // gather, norms, per-tile scale and fp16 split of the queries in one launch
    if (int rc = knn_tc_gather_queries(keys, width, query_ids, n_queries, p.tc, ws + p.off_tc, q, st)) return rc;
  } else if (int rc = mimrl_gather_rows(keys, n_keys, width, query_ids, n_queries, 1, width, q, stream)) {
    return rc;
  }
  bool comp_done = false;
  if (int rc = knn_core(keys, n_keys, width, 0, q, n_queries, query_ids, n_queries, k, exact_form, nbr_orig, nbr_dist,
                        p, ws, st, fitted, nbr_comp, &comp_done, tc_route))
    return rc;
  if (nbr_comp && !comp_done) {          // (the sorted width-1 route has no re-rank kernel)
    const size_t n_out = (size_t)n_queries * k;
    compact_index_kernel<<<(int)((n_out * 32 + 255) / 256), 256, 0, st>>>(nbr_orig, n_out, query_ids, n_queries, nbr_comp);
    return check_launch("knn compact_index");
  }
  return 0;
}
