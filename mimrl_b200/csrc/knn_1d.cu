// k-NN over a width-1 key pool (the label pools C_F_all of Model.py:326,332,366,372: two of the six searches of a
// step) on scikit-learn's kd_tree route, where the distance is the true squared difference in float64
// (oracle/knn_oracle.c, gemm_form = 0).  In one dimension fl((q - z)^2) is monotone in the position of z in the
// sorted pool on either side of q, so the search is a sort of the pool (radix sort, stable: equal values keep ascending
// row numbers) and, per query, a binary search plus a two-sided walk -- instead of the m x N sweep of the general
// filter (16 ms at 4096 x 1M).  The (distance, row number) total order of the oracle is kept exactly:
//   * D = the k-th smallest distance, from the k nearest positions on either side;
//   * every key with distance < D is a neighbour (fewer than k of them, all adjacent to q's position);
//   * the remaining slots go to the LOWEST row numbers among the keys with distance == D.  Those keys form one run of
//     positions per side; a run is a sequence of sub-runs of bit-equal values, each with ascending row numbers, so the
//     first `need` entries of every sub-run are the only candidates (label pools: one or two sub-runs of thousands of
//     duplicates; distinct values whose float64 distances collide are handled by the same loop).
// Excluded rows (the rows drawn as queries, Model.py:83-84) are given the value +inf before the sort and land behind
// the valid keys.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mimrl {
namespace {

constexpr int kMaxK = 64;
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void knn1d_keys_kernel(const float *__restrict__ keys, int n, float *__restrict__ vals, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = keys[i] + 0.f;          // -0.0 -> +0.0: same distances, one bit pattern per value
  idx[i] = i;
}

__global__ void knn1d_exclude_kernel(const int64_t *__restrict__ ids, int n_ids, int64_t key_offset, int n,
                                     float *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ids) return;
  const int64_t l = ids[i] - key_offset;
  if (l >= 0 && l < n) vals[l] = INFINITY;
}

// first position in [lo, hi) with v[pos] >= x
__device__ __forceinline__ int lower_bound(const float *__restrict__ v, int lo, int hi, float x) {
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (__ldg(v + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first position in [lo, hi) with v[pos] > x
__device__ __forceinline__ int upper_bound(const float *__restrict__ v, int lo, int hi, float x) {
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (__ldg(v + mid) <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(64)
knn1d_query_kernel(const float *__restrict__ sv, const int *__restrict__ si, int n, const float *__restrict__ queries,
                   int n_queries, int k, int64_t key_offset, int64_t *__restrict__ nbr, double *__restrict__ nbr_dist) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n_queries) return;
  const float qf = queries[qi];
  const double q = (double)qf;
  const int nv = lower_bound(sv, 0, n, INFINITY);          // valid (not excluded) keys
  const int kk = k < nv ? k : nv;
  if (kk == 0) return;
  const int p = lower_bound(sv, 0, nv, qf + 0.f);
  auto dist = [&](int pos) {
    const double t = q - (double)__ldg(sv + pos);
    return t * t;
  };
  // D: the kk-th smallest distance
  double D = 0.0;
  {
    int l = p - 1, r = p;
    double dl = l >= 0 ? dist(l) : (double)INFINITY, dr = r < nv ? dist(r) : (double)INFINITY;
    for (int c = 0; c < kk; ++c) {
      if (dl <= dr) {
        D = dl, --l;
        dl = l >= 0 ? dist(l) : (double)INFINITY;
      } else {
        D = dr, ++r;
        dr = r < nv ? dist(r) : (double)INFINITY;
      }
    }
  }
  // keys closer than D, ordered by (distance, row)
  double rd[kMaxK];
  int ri[kMaxK];
  int c = 0;
  auto push = [&](double d, int id) {
    int j = c++;
    while (j > 0 && (rd[j - 1] > d || (rd[j - 1] == d && ri[j - 1] > id))) rd[j] = rd[j - 1], ri[j] = ri[j - 1], --j;
    rd[j] = d, ri[j] = id;
  };
  int a = p - 1, b = p;
  for (; a >= 0; --a) {
    const double d = dist(a);
    if (!(d < D)) break;
    push(d, __ldg(si + a));
  }
  for (; b < nv; ++b) {
    const double d = dist(b);
    if (!(d < D)) break;
    push(d, __ldg(si + b));
  }
  // the lowest row numbers among the keys at distance D
  const int need = kk - c;
  int ti[kMaxK];
  int nt = 0;
  auto offer = [&](int id) {          // false: the list is full and id cannot enter (nor can a larger one)
    if (nt == need && id >= ti[nt - 1]) return false;
    int j = nt < need ? nt++ : nt - 1;
    while (j > 0 && ti[j - 1] > id) ti[j] = ti[j - 1], --j;
    ti[j] = id;
    return true;
  };
  while (a >= 0 && dist(a) == D) {
    const int s = lower_bound(sv, 0, a + 1, __ldg(sv + a));          // the sub-run of this value is [s, a]
    for (int j = s; j <= a && j < s + need; ++j)
      if (!offer(__ldg(si + j))) break;
    a = s - 1;
  }
  while (b < nv && dist(b) == D) {
    const int e = upper_bound(sv, b, nv, __ldg(sv + b));             // [b, e)
    for (int j = b; j < e && j < b + need; ++j)
      if (!offer(__ldg(si + j))) break;
    b = e;
  }
  int64_t *out = nbr + (size_t)qi * k;
  double *outd = nbr_dist ? nbr_dist + (size_t)qi * k : nullptr;
  for (int j = 0; j < c; ++j) {
    out[j] = (int64_t)ri[j] + key_offset;
    if (outd) outd[j] = rd[j];
  }
  for (int j = 0; j < nt; ++j) {
    out[c + j] = (int64_t)ti[j] + key_offset;
    if (outd) outd[c + j] = D;
  }
}

struct Layout1d {
  size_t off_vals, off_idx, off_svals, off_sidx, off_temp, temp_bytes, total;
};

Layout1d layout_1d(int n_keys) {
  Layout1d l;
  const size_t arr = align256((size_t)n_keys * 4);
  l.off_vals = 0, l.off_idx = arr, l.off_svals = 2 * arr, l.off_sidx = 3 * arr, l.off_temp = 4 * arr;
  l.temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, l.temp_bytes, (const float *)nullptr, (float *)nullptr, (const int *)nullptr,
                                  (int *)nullptr, n_keys);
  l.total = l.off_temp + align256(l.temp_bytes);
  return l;
}

}  // namespace

bool knn1d_supported(int width, int exact_form) {
  static const bool off = getenv("MIMRL_KNN1D_OFF") != nullptr;
  return width == 1 && !exact_form && !off;
}

size_t knn1d_workspace_bytes(int n_keys) { return layout_1d(n_keys).total; }

int knn1d_search(const float *keys, int n_keys, int64_t key_offset, const float *queries, int n_queries,
                 const int64_t *excluded, int n_excluded, int k, int64_t *nbr_orig, double *nbr_dist, unsigned char *ws,
                 cudaStream_t st) {
  MIMRL_REQUIRE(k <= kMaxK, "knn_search: k=%d too large (max %d)", k, kMaxK);
  const Layout1d l = layout_1d(n_keys);
  float *vals = reinterpret_cast<float *>(ws + l.off_vals), *svals = reinterpret_cast<float *>(ws + l.off_svals);
  int *idx = reinterpret_cast<int *>(ws + l.off_idx), *sidx = reinterpret_cast<int *>(ws + l.off_sidx);
  knn1d_keys_kernel<<<(n_keys + 255) / 256, 256, 0, st>>>(keys, n_keys, vals, idx);
  if (check_launch("knn 1-d keys")) return 1;
  if (n_excluded > 0) {
    knn1d_exclude_kernel<<<(n_excluded + 255) / 256, 256, 0, st>>>(excluded, n_excluded, key_offset, n_keys, vals);
    if (check_launch("knn 1-d exclude")) return 1;
  }
  size_t temp = l.temp_bytes;
  const cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + l.off_temp, temp, (const float *)vals, svals, (const int *)idx,
                                                        sidx, n_keys, 0, 32, st);
  MIMRL_REQUIRE(e == cudaSuccess, "knn 1-d sort: %s", cudaGetErrorString(e));
  knn1d_query_kernel<<<(n_queries + 63) / 64, 64, 0, st>>>(svals, sidx, n_keys, queries, n_queries, k, key_offset,
                                                           nbr_orig, nbr_dist);
  return check_launch("knn 1-d query");
}

}  // namespace mimrl
