// Tensor-core distance filter of the k-NN sampler (first stage of mimrl_knn_search,
// reference Model.py:82-86): d(q, z) = |q|^2 + |z|^2 - 2 q.z with the q.z tile on
// tcgen05 (fp16 hi/lo split, three products -> fp32-class), 128 queries parked in
// TMEM per CTA, keys streamed by TMA through a 4-deep ring, and the selection fused
// into the epilogue: one thread per query row compares the 128 distances of a tile
// against its current K'-th best and inserts the rare survivors into a sorted list
// in shared memory.  The m x N distance matrix is never written.  Exactness comes
// from the float64 re-rank stage (knn.cu), which only needs the true neighbours to
// be inside the K' = k + 4 candidates kept here.
#include "tc_common.cuh"

namespace mimrl {
namespace {

constexpr int kKnnThreads = 320;
constexpr int kKnnStages = 4;
constexpr uint32_t kKUnit = 32768;
constexpr uint32_t kKTile16 = 128 * 128;
constexpr int kListMax = 32;
constexpr uint32_t kKnnSmem = kKnnStages * kKUnit + 1024 /*bars*/ + 2 * 128 * 4 /*key norms*/ +
                              2 * 128 * kListMax * 8 /*lists*/ + 1024;

__global__ void knn_absmax_kernel(const float *__restrict__ a, size_t na, const float *__restrict__ b, size_t nb,
                                  unsigned *__restrict__ out) {
  const float *src = blockIdx.y == 0 ? a : b;
  const size_t n = blockIdx.y == 0 ? na : nb;
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(src + i)));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out + blockIdx.y, __float_as_uint(m));
}

// Keys: ONE pass.  A CTA takes a tile of 128 keys (the tile the filter later streams as one TMA box): squared row
// norms, the tile's absmax -> a power-of-two scale of ITS OWN (the filter folds 1 / scale into the per-tile factor of
// the distance, so nothing is paid per element), fp16 hi / lo split.  Replaces a norms pass, an absmax pass and a split pass over the keys.
__global__ void __launch_bounds__(256) knn_prep_keys_kernel(const float *__restrict__ keys, int n_keys, int width,
                                                            float *__restrict__ kn, float *__restrict__ tile_inv_scale,
                                                            __half *__restrict__ hi, __half *__restrict__ lo) {
  __shared__ float s_max[8];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  // warp w owns rows w, w + 8, ... of the tile; a lane owns columns 4 lane .. 4 lane + 3 of each (one 512-byte row per
  // warp-wide 16-byte load, one 256-byte run per plane and store)
  float4 v[16];
  const bool vec = width == 128 && (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const size_t row = (size_t)blockIdx.x * 128 + w + 8 * j;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < (size_t)n_keys) {
      if (vec) {
        q = __ldg(reinterpret_cast<const float4 *>(keys + row * 128) + lane);
      } else {
        const float *src = keys + row * width;
        const int c = 4 * lane;
        q.x = c < width ? __ldg(src + c) : 0.f, q.y = c + 1 < width ? __ldg(src + c + 1) : 0.f;
        q.z = c + 2 < width ? __ldg(src + c + 2) : 0.f, q.w = c + 3 < width ? __ldg(src + c + 3) : 0.f;
      }
    }
    v[j] = q;
  }
  // (a second loop: the shuffles below would otherwise sit between the loads and serialise them)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const size_t row = (size_t)blockIdx.x * 128 + w + 8 * j;
    const float4 q = v[j];
    float nrm = fmaf(q.w, q.w, fmaf(q.z, q.z, fmaf(q.y, q.y, q.x * q.x)));
#pragma unroll
    for (int o = 16; o; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0 && row < (size_t)n_keys) kn[row] = nrm;
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[w] = mx;
  __syncthreads();
  mx = s_max[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, s_max[k]);
  const float sc = scale_from_absmax(__float_as_uint(mx));
  if (t == 0) tile_inv_scale[blockIdx.x] = 1.f / sc;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const size_t row = (size_t)blockIdx.x * 128 + w + 8 * j;
    if (row >= (size_t)n_keys) continue;                  // rows past the end are zero-filled by the TMA box
    const float a0 = v[j].x * sc, a1 = v[j].y * sc, a2 = v[j].z * sc, a3 = v[j].w * sc;
    const __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(a0 - f0.x, a1 - f0.y), l1 = __floats2half2_rn(a2 - f1.x, a3 - f1.y);
    reinterpret_cast<uint2 *>(hi + row * 128)[lane] =
        make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
    reinterpret_cast<uint2 *>(lo + row * 128)[lane] =
        make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
  }
}

// [n, width] fp32 -> hi, lo [n, 128] fp16 (K zero-padded), scaled by 2^k
__global__ void knn_split_kernel(const float *__restrict__ src, size_t n, int width, const unsigned *__restrict__ absmax,
                                 int which, __half *__restrict__ hi, __half *__restrict__ lo) {
  const float sc = scale_from_absmax(absmax[which]);
  const size_t total = n * 64;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx >> 6;
    const int e = (int)(idx & 63) * 2;
    const float v0 = e < width ? __ldg(src + r * width + e) * sc : 0.f;
    const float v1 = e + 1 < width ? __ldg(src + r * width + e + 1) * sc : 0.f;
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__half2 *>(hi + r * 128 + e) = h;
    *reinterpret_cast<__half2 *>(lo + r * 128 + e) = l;
  }
}

__global__ void knn_row_norms_kernel(const float *__restrict__ x, int n, int width, float *__restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const float *r = x + (size_t)w * width;
  float acc = 0.f;
  for (int e = lane; e < width; e += 32) acc = fmaf(r[e], r[e], acc);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[w] = acc;
}

// Queries of the sampler are rows of the key pool.  One CTA per 128-query tile (the tile a filter CTA parks in TMEM):
// gather the rows, squared norms (same summation order as knn_row_norms_kernel), the tile's max|q| -> a power-of-two
// scale of its own, fp16 hi / lo split (K zero-padded to 128).  One launch instead of gather + absmax + norms + split.
__global__ void __launch_bounds__(256) knn_prep_queries_kernel(const float *__restrict__ keys, int width,
                                                               const int64_t *__restrict__ ids, int n, float *__restrict__ q,
                                                               float *__restrict__ qn, unsigned *__restrict__ tile_absmax,
                                                               __half *__restrict__ hi, __half *__restrict__ lo) {
  __shared__ float s_max[8];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  // warp w owns rows w, w + 8, ... of the tile; a lane owns columns lane, lane + 32, lane + 64, lane + 96
  float v[16][4];
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int row = blockIdx.x * 128 + w + 8 * j;
    const float *src = row < n ? keys + (size_t)__ldg(ids + row) * width : nullptr;
#pragma unroll
    for (int c = 0; c < 4; ++c) v[j][c] = (src && lane + 32 * c < width) ? __ldg(src + lane + 32 * c) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int row = blockIdx.x * 128 + w + 8 * j;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (lane + 32 * c < width) acc = fmaf(v[j][c], v[j][c], acc);
      mx = fmaxf(mx, fabsf(v[j][c]));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < n) {
      if (lane == 0) qn[row] = acc;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (lane + 32 * c < width) q[(size_t)row * width + lane + 32 * c] = v[j][c];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[w] = mx;
  __syncthreads();
  mx = s_max[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, s_max[k]);
  if (t == 0) tile_absmax[blockIdx.x] = __float_as_uint(mx);
  const float sc = scale_from_absmax(__float_as_uint(mx));
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int row = blockIdx.x * 128 + w + 8 * j;
    if (row >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float a = v[j][c] * sc;
      const __half h = __float2half_rn(a);
      hi[(size_t)row * 128 + lane + 32 * c] = h;
      lo[(size_t)row * 128 + lane + 32 * c] = __float2half_rn(a - __half2float(h));
    }
  }
}

struct KnnTcParams {
  int n_keys, n_queries, tiles_per_split, list_len, n_lists, jth, splits, q_tiles, q_group, refresh_mask;
  float *pub;                      // [n_queries][n_lists]: jth-smallest distance each (split, warpgroup) part has seen
  const unsigned *absmax;          // max|q|: [0] for all queries (absmax_stride 0) or one per 128-query tile (stride 1)
  int absmax_stride;
  const float *tile_inv_scale;     // 1 / (power-of-two scale) of every 128-key tile (knn_prep_keys_kernel)
  const __half *q_hi, *q_lo;
  const float *qn, *kn;
  KnnCand *cand;
};

// Candidate list of one (query row, warpgroup): a binary MAX-HEAP by (d, idx) once it is full (plain array while it
// fills; heapified at the K'-th entry), so the threshold is the root and an insertion is a sift-down of log2 K' levels
// -- two shared-memory reads per level -- instead of a rescan of all K' entries (insertions run under divergence, one
// lane's latency stalls the warp: at k = 16 the rescan was 40 % of the kernel's samples).  Sorted once at the end.
__device__ __forceinline__ bool knn_less(float d1, int i1, float d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

__device__ __forceinline__ void knn_sift_down(KnnCand *lst, int n, int i, KnnCand x) {
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    KnnCand a = lst[c];
    if (c + 1 < n) {
      const KnnCand b2 = lst[c + 1];
      if (knn_less(a.d, a.idx, b2.d, b2.idx)) a = b2, ++c;
    }
    if (!knn_less(x.d, x.idx, a.d, a.idx)) break;
    lst[i] = a;
    i = c;
  }
  lst[i] = x;
}

__device__ __noinline__ void knn_insert(KnnCand *lst, int &len, int cap, float &tau_d, int &tau_i, int &worst, float d,
                                        int idx) {
  (void)worst;
  if (len < cap) {
    lst[len] = KnnCand{d, idx};
    ++len;
    if (len < cap) return;                       // threshold stays +inf until the list is full
    for (int s2 = cap / 2 - 1; s2 >= 0; --s2) knn_sift_down(lst, cap, s2, lst[s2]);
  } else {
    if (!knn_less(d, idx, tau_d, tau_i)) return;
    knn_sift_down(lst, cap, 0, KnnCand{d, idx});
  }
  const KnnCand root = lst[0];
  tau_d = root.d, tau_i = root.idx;
}

__global__ void __launch_bounds__(kKnnThreads, 1)
knn_filter_tc_kernel(const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                     const KnnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kKnnStages * kKUnit;
  const uint32_t bFull = bars, bEmpty = bars + 64, bTFull = bars + 128, bTEmpty = bars + 144;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kKnnStages * kKUnit + 256);
  float *knbuf = reinterpret_cast<float *>(gen + kKnnStages * kKUnit + 1024);                 // [2][128]
  KnnCand *lists = reinterpret_cast<KnnCand *>(gen + kKnnStages * kKUnit + 1024 + 1024);      // [2][128][kListMax]

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // 1-D grid over (query tile, key split).  Query tiles are taken in groups of q_group; inside a group the query
  // tile is the fast index, so q_group CTAs stream the same key range at the same time (L2 reuse of the keys) while
  // all splits of a query tile still run in the same window (they tighten each other's thresholds through `pub`).
  int split, qt;
  {
    const int per_group = p.splits * p.q_group;
    const int grp = blockIdx.x / per_group, rr = blockIdx.x % per_group;
    const int gsz = min(p.q_group, p.q_tiles - grp * p.q_group);
    split = rr / gsz;
    qt = grp * p.q_group + rr % gsz;
    if (split >= p.splits) return;             // tail of a short last group
  }
  const int row0 = qt * 128;
  const int n_tiles = (p.n_keys + 127) / 128;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(n_tiles, t0 + p.tiles_per_split);
  const int T = t1 - t0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kKnnStages; ++i) {
      mbar_init(bFull + 8 * i, 1);
      mbar_init(bEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kKnnStages * kKUnit + 256), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // park the 128 query rows in TMEM columns [0,128): hi kb0, hi kb1, lo kb0, lo kb1 (one column = two K values)
  if (warp >= 2 && warp < 6) {
    const int r = (warp & 3) * 32 + lane;
    const bool ok = row0 + r < p.n_queries;
#pragma unroll 1
    for (int part = 0; part < 4; ++part) {
      const __half *src = (part < 2 ? p.q_hi : p.q_lo) + (size_t)(row0 + r) * 128 + (part & 1) * 64;
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 t = ok ? __ldg(reinterpret_cast<const uint4 *>(src) + j) : make_uint4(0, 0, 0, 0);
        v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
      }
      tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + part * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      prefetch_tmap(&map_k_hi);
      prefetch_tmap(&map_k_lo);
      for (int u = 0; u < 2 * T; ++u) {
        const int stage = u % kKnnStages;
        const int col = (t0 + (u >> 1)) * 128;
        mbar_wait(bEmpty + 8 * stage, ((u / kKnnStages) & 1) ^ 1);
        mbar_expect_tx(bFull + 8 * stage, kKUnit);
        const CUtensorMap *m = (u & 1) ? &map_k_lo : &map_k_hi;
        tma_load_2d(base + stage * kKUnit, m, bFull + 8 * stage, 0, col);
        tma_load_2d(base + stage * kKUnit + kKTile16, m, bFull + 8 * stage, 64, col);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = instr_desc_f16(128, 128);
    for (int i = 0; i < T; ++i) {
      const int buf = i & 1;
      const uint32_t d = tmem_base + 128 + buf * 128;
      mbar_wait(bTEmpty + 8 * buf, ((i >> 1) & 1) ^ 1);
      int u = 2 * i, stage = u % kKnnStages;
      mbar_wait(bFull + 8 * stage, (u / kKnnStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t b0 = base + stage * kKUnit;
#pragma unroll
        for (int a_sel = 0; a_sel < 4; a_sel += 2)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(d, tmem_base + (a_sel + kb) * 32 + k * 8, smem_desc_sw128(b0 + kb * kKTile16 + k * 32), idesc,
                          (a_sel | kb | k) ? 1u : 0u);
        umma_commit(bEmpty + 8 * stage);
      }
      __syncwarp();
      u = 2 * i + 1, stage = u % kKnnStages;
      mbar_wait(bFull + 8 * stage, (u / kKnnStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t b0 = base + stage * kKUnit;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(d, tmem_base + kb * 32 + k * 8, smem_desc_sw128(b0 + kb * kKTile16 + k * 32), idesc, 1u);
        umma_commit(bEmpty + 8 * stage);
        umma_commit(bTFull + 8 * buf);
      }
      __syncwarp();
    }
  } else {
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool row_ok = row0 + r < p.n_queries;
    const float m2q = -2.f / scale_from_absmax(p.absmax[(row0 >> 7) * p.absmax_stride]);        // times 1 / (scale of the key tile), per tile
    const float qn = row_ok ? p.qn[row0 + r] : 0.f;
    KnnCand *lst = lists + (size_t)(wg * 128 + r) * kListMax;
    float *kn_s = knbuf + wg * 128;
    int len = 0, tau_i = 0x7fffffff, worst = 0;
    float tau_d = INFINITY;          // effective threshold = min(local K'-th best, shared bound)
    float tau_local = INFINITY, tau_shared = INFINITY, published = INFINITY;
    // Shared bound across the key splits of this query: every part publishes the jth-smallest distance it has seen,
    // jth = ceil(K' / parts).  If every part holds >= jth keys within t = max over parts, the union holds >= K'
    // keys within t, so nothing beyond t can be among the K' best: the splits tighten each other's thresholds as
    // if they were one sweep, and the (warp-divergent) insertion path stays rare.
    const int part = split * 2 + wg;
    // pub[part][query]: the 32 rows of a warp read / write one 128-byte line per part
    float *pub_q = p.pub + (row0 + r);
    const size_t pub_ld = (size_t)p.q_tiles * 128;
    const int buf = wg;
    float kn_next = INFINITY;
    if (wg < T) {
      const int gk = (t0 + wg) * 128 + r;
      kn_next = gk < p.n_keys ? __ldg(p.kn + gk) : INFINITY;
    }
    for (int i = wg; i < T; i += 2) {
      const int col0 = (t0 + i) * 128;
      // stage |z|^2 of this tile's 128 keys (+inf = excluded / out of range); prefetch the next tile's
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");      // previous tile's readers are done
      kn_s[r] = kn_next;
      if (i + 2 < T) {
        const int gk = col0 + 256 + r;
        kn_next = gk < p.n_keys ? __ldg(p.kn + gk) : INFINITY;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
      const float m2inv = m2q * __ldg(p.tile_inv_scale + t0 + i);
      mbar_wait(bTFull + 8 * buf, (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + buf * 128 + ch * 32, v);
        tmem_ld_wait();
        const float4 *k4 = reinterpret_cast<const float4 *>(kn_s + ch * 32);
        float dmin = INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 kk = k4[j >> 2];
          const float d0 = fmaxf(fmaf(__uint_as_float(v[j]), m2inv, qn + kk.x), 0.f);
          const float d1 = fmaxf(fmaf(__uint_as_float(v[j + 1]), m2inv, qn + kk.y), 0.f);
          const float d2 = fmaxf(fmaf(__uint_as_float(v[j + 2]), m2inv, qn + kk.z), 0.f);
          const float d3 = fmaxf(fmaf(__uint_as_float(v[j + 3]), m2inv, qn + kk.w), 0.f);
          v[j] = __float_as_uint(d0), v[j + 1] = __float_as_uint(d1), v[j + 2] = __float_as_uint(d2),
          v[j + 3] = __float_as_uint(d3);
          dmin = fminf(dmin, fminf(fminf(d0, d1), fminf(d2, d3)));
        }
        // Survivors.  A chunk that holds one for ANY of the warp's 32 rows sends the whole warp down this path, so its
        // cost is what matters early in a sweep (the chance of a survivor decays like K'/keys seen).  Every lane first
        // builds the bit mask of its own survivors without divergence and parks its 32 distances in local memory; the
        // lanes then walk their masks together, so the serialised insertions are max-per-lane (typically one), not one
        // per distinct column position in the warp.
        if (__any_sync(0xffffffffu, row_ok && dmin <= tau_d)) {
          uint32_t mask = 0;
          float loc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dj = __uint_as_float(v[j]);
            loc[j] = dj;
            mask |= (dj <= tau_d && dj < INFINITY) ? (1u << j) : 0u;
          }
          if (!row_ok) mask = 0;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float dj = loc[j];
            if (dj <= tau_d) {
              knn_insert(lst, len, p.list_len, tau_local, tau_i, worst, dj, col0 + ch * 32 + j);
              tau_d = fminf(tau_local, tau_shared);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTEmpty + 8 * buf);
      if (row_ok) {
        const int it = i >> 1;               // tiles this warpgroup has finished
        // refresh points: tiles 1, 2, 4, 8 of this warpgroup while the lists warm up, then every 16th
        const bool refresh = it == 0 || it == 1 || it == 3 || (it & p.refresh_mask) == p.refresh_mask;
        if (refresh && len >= p.jth) {
          // jth-smallest distance of this part (jth is 1 or 2 in practice): partial selection over the unsorted list
          float lo_d = -1.f, mine = INFINITY;
          int lo_i = -1;
          for (int jj = 0; jj < p.jth; ++jj) {
            float bd = INFINITY;
            int bi = 0x7fffffff;
            for (int u = 0; u < len; ++u) {
              const KnnCand c = lst[u];
              if (knn_less(lo_d, lo_i, c.d, c.idx) && knn_less(c.d, c.idx, bd, bi)) bd = c.d, bi = c.idx;
            }
            lo_d = bd, lo_i = bi, mine = bd;
          }
          if (mine < published) {
            published = mine;
            __stcg(pub_q + part * pub_ld, mine);
          }
        }
        if (refresh) {
          // n_lists is even; 8 independent L2 loads in flight per round trip.
          // Fewer parts than K' (jth > 1): t = max over parts of their jth-smallest.  At least K' parts (jth == 1):
          // the parts are taken in K' groups and t = max over groups of (min over the group's parts) --
          // the K' group minima are K' distinct keys within t, so the K'-th best of the union is <= t.  With 74 parts
          // and K' = 20 that bound sits at the ~0.9/n quantile of the distances seen instead of ~4.7/n for the plain
          // maximum (n = keys seen per part), i.e. five times fewer survivors.
          float t = 0.f;
          int pp = 0;
          if (p.jth == 1 && p.n_lists > p.list_len) {
            // group g = parts g, g + K', g + 2 K', ...: interleaved, so that the few splits of a query tile that run in
            // the same wave already populate every group (consecutive groups would stay empty until the last wave)
            const int G = p.list_len;
            for (int g0 = 0; g0 < G; g0 += 2) {
              float m0 = INFINITY, m1 = INFINITY;
              const bool two = g0 + 1 < G;
#pragma unroll 4
              for (int e = g0; e < p.n_lists; e += G) {
                m0 = fminf(m0, __ldcg(pub_q + e * pub_ld));
                if (two && e + 1 < p.n_lists) m1 = fminf(m1, __ldcg(pub_q + (e + 1) * pub_ld));
              }
              t = fmaxf(t, two ? fmaxf(m0, m1) : m0);
            }
          } else {
            for (; pp + 8 <= p.n_lists; pp += 8) {
              const float a0 = __ldcg(pub_q + pp * pub_ld), a1 = __ldcg(pub_q + (pp + 1) * pub_ld), a2 = __ldcg(pub_q + (pp + 2) * pub_ld),
                          a3 = __ldcg(pub_q + (pp + 3) * pub_ld), a4 = __ldcg(pub_q + (pp + 4) * pub_ld), a5 = __ldcg(pub_q + (pp + 5) * pub_ld),
                          a6 = __ldcg(pub_q + (pp + 6) * pub_ld), a7 = __ldcg(pub_q + (pp + 7) * pub_ld);
              t = fmaxf(t, fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)), fmaxf(fmaxf(a4, a5), fmaxf(a6, a7))));
            }
            for (; pp < p.n_lists; pp += 2) t = fmaxf(t, fmaxf(__ldcg(pub_q + pp * pub_ld), __ldcg(pub_q + (pp + 1) * pub_ld)));
          }
          tau_shared = fminf(tau_shared, t);
          tau_d = fminf(tau_local, tau_shared);
        }
      }
    }
    if (row_ok) {
      // sort ascending by (d, idx): the re-rank merges sorted lists.  A full list is a max-heap: pop the root to the
      // end, K' - 1 times (heapsort); a list that never filled is a short plain array: insertion sort.
      if (len == p.list_len) {
        for (int u = len - 1; u > 0; --u) {
          const KnnCand top = lst[0], last = lst[u];
          lst[u] = top;
          knn_sift_down(lst, u, 0, last);
        }
      } else {
        for (int u = 1; u < len; ++u) {
          const KnnCand x = lst[u];
          int v = u;
          while (v > 0 && knn_less(x.d, x.idx, lst[v - 1].d, lst[v - 1].idx)) {
            lst[v] = lst[v - 1];
            --v;
          }
          lst[v] = x;
        }
      }
      KnnCand *o = p.cand + ((size_t)(row0 + r) * p.n_lists + (split * 2 + wg)) * p.list_len;
      for (int u = 0; u < p.list_len; ++u) o[u] = u < len ? lst[u] : KnnCand{INFINITY, -1};
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool knn_tc_supported(int n_keys, int n_queries, int width, int list_len) {
  return width <= 128 && list_len <= kListMax && n_keys >= 2048 && n_queries >= 1;
}

KnnTcPlan knn_tc_plan(int n_keys, int n_queries, int width, int list_len) {
  (void)width;
  KnnTcPlan p;
  p.list_len = list_len;
  const int q_tiles = ceil_div(n_queries, 128), k_tiles = ceil_div(n_keys, 128);
  // key splits: as few idle waves over the 148 SMs as possible (one CTA per SM).  Every split pays a warm-up while its
  // candidate lists fill and its threshold is loose (survivor path taken by every chunk).  Measured on the B200 at
  // 4096 x 1M x 128 (A/B of the constant): with the grouped shared bound of the filter 6 -> 3.71 ms, 30 -> 3.30 ms,
  // 100 -> 3.35 ms, 300 -> 3.38 ms (before it: 5.15 / 3.9 / 3.9 / 3.53 ms).  n_lists = 2 * splits <= 128 (one merge
  // thread per list in the re-rank).
  int splits = 1;
  double best = 1e300;
  for (int s = 1; s <= 64 && s <= k_tiles; ++s) {
    const int per = ceil_div(k_tiles, s), eff = ceil_div(k_tiles, per);
    static const double warm = getenv("MIMRL_KNN_WARM") ? atof(getenv("MIMRL_KNN_WARM")) : 30.0;
    const double cost = (double)ceil_div(q_tiles * eff, 148) * (per + warm);
    if (cost < best - 1e-9) best = cost, splits = eff;
  }
  p.tiles_per_split = ceil_div(k_tiles, splits);
  p.splits = ceil_div(k_tiles, p.tiles_per_split);
  p.n_lists = 2 * p.splits;
  size_t o = 0;
  p.off_absmax = o, o += 256;
  p.off_qn = o, o += align256((size_t)n_queries * 4);
  p.off_q_hi = o, o += align256((size_t)n_queries * 128 * 2);
  p.off_q_lo = o, o += align256((size_t)n_queries * 128 * 2);
  p.off_k_hi = o, o += align256((size_t)n_keys * 128 * 2);
  p.off_k_lo = o, o += align256((size_t)n_keys * 128 * 2);
  p.off_pub = o, o += align256((size_t)q_tiles * 128 * p.n_lists * 4);
  p.off_tscale = o, o += align256((size_t)k_tiles * 4);
  p.off_qscale = o, o += align256((size_t)q_tiles * 4);
  p.bytes = o;
  return p;
}

KnnTcKeys knn_tc_keys_in_workspace(const KnnTcPlan &plan, unsigned char *ws) {
  KnnTcKeys k;
  k.hi = reinterpret_cast<__half *>(ws + plan.off_k_hi), k.lo = reinterpret_cast<__half *>(ws + plan.off_k_lo);
  k.tile_inv_scale = reinterpret_cast<float *>(ws + plan.off_tscale);
  return k;
}

int knn_tc_prepare_keys(const float *keys, int n_keys, int width, const KnnTcKeys &out, float *key_norms, cudaStream_t st) {
  knn_prep_keys_kernel<<<ceil_div(n_keys, 128), 256, 0, st>>>(keys, n_keys, width, key_norms, out.tile_inv_scale,
                                                             static_cast<__half *>(out.hi), static_cast<__half *>(out.lo));
  return check_launch("knn prepare keys");
}

int knn_tc_gather_queries(const float *keys, int width, const int64_t *ids, int n_queries, const KnnTcPlan &plan,
                          unsigned char *ws, float *q_out, cudaStream_t st) {
  knn_prep_queries_kernel<<<ceil_div(n_queries, 128), 256, 0, st>>>(
      keys, width, ids, n_queries, q_out, reinterpret_cast<float *>(ws + plan.off_qn),
      reinterpret_cast<unsigned *>(ws + plan.off_qscale), reinterpret_cast<__half *>(ws + plan.off_q_hi),
      reinterpret_cast<__half *>(ws + plan.off_q_lo));
  return check_launch("knn prepare queries");
}

int knn_filter_tc(const KnnTcKeys &pk, const float *key_norms, int n_keys, int width, const float *queries,
                  int n_queries, const KnnTcPlan &plan, unsigned char *ws, KnnCand *cand, cudaStream_t st, bool q_ready) {
  unsigned *absmax = reinterpret_cast<unsigned *>(ws + plan.off_absmax);
  float *qn = reinterpret_cast<float *>(ws + plan.off_qn);
  __half *q_hi = reinterpret_cast<__half *>(ws + plan.off_q_hi), *q_lo = reinterpret_cast<__half *>(ws + plan.off_q_lo);
  __half *k_hi = static_cast<__half *>(pk.hi), *k_lo = static_cast<__half *>(pk.lo);
  if (!q_ready) {          // (knn_tc_gather_queries has produced norms, per-tile max|q| and the fp16 planes otherwise)
    cudaMemsetAsync(absmax, 0, 8, st);
    const size_t nq = (size_t)n_queries * width;
    int blocks = (int)((nq + 2047) / 2048);
    blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
    knn_absmax_kernel<<<dim3(blocks, 1), 256, 0, st>>>(queries, nq, nullptr, 0, absmax);
    if (check_launch("knn absmax")) return 1;
    knn_row_norms_kernel<<<ceil_div(n_queries * 32, 256), 256, 0, st>>>(queries, n_queries, width, qn);
    if (check_launch("knn query norms")) return 1;
    knn_split_kernel<<<ceil_div(n_queries * 64, 256), 256, 0, st>>>(queries, n_queries, width, absmax, 0, q_hi, q_lo);
    if (check_launch("knn split q")) return 1;
  }
  CUtensorMap mh, ml;
  if (make_map(&mh, k_hi, 128, n_keys, 128, 128)) return 1;
  if (make_map(&ml, k_lo, 128, n_keys, 128, 128)) return 1;
  KnnTcParams p;
  p.n_keys = n_keys, p.n_queries = n_queries, p.tiles_per_split = plan.tiles_per_split, p.list_len = plan.list_len;
  p.n_lists = plan.n_lists;
  p.jth = ceil_div(plan.list_len, plan.n_lists);
  p.pub = reinterpret_cast<float *>(ws + plan.off_pub);
  cudaMemsetAsync(p.pub, 0x7f, (size_t)ceil_div(n_queries, 128) * 128 * plan.n_lists * 4, st);     // 3.39e38: "nothing seen yet"
  p.absmax = q_ready ? reinterpret_cast<const unsigned *>(ws + plan.off_qscale) : absmax, p.absmax_stride = q_ready ? 1 : 0;
  p.q_hi = q_hi, p.q_lo = q_lo, p.qn = qn, p.kn = key_norms, p.cand = cand;
  p.tile_inv_scale = pk.tile_inv_scale;
  cudaFuncSetAttribute(knn_filter_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKnnSmem);
  p.splits = plan.splits;
  p.q_tiles = ceil_div(n_queries, 128);
  p.refresh_mask = getenv("MIMRL_KNN_REFRESH") ? atoi(getenv("MIMRL_KNN_REFRESH")) : 15;
  p.q_group = getenv("MIMRL_KNN_QGROUP") ? atoi(getenv("MIMRL_KNN_QGROUP")) : 32;
  if (p.q_group < 1) p.q_group = 1;
  const int groups = ceil_div(p.q_tiles, p.q_group);
  knn_filter_tc_kernel<<<groups * plan.splits * p.q_group, kKnnThreads, kKnnSmem, st>>>(mh, ml, p);
  return check_launch("knn_filter_tc");
}

}  // namespace mimrl
