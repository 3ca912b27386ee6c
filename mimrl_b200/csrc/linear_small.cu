// Linear layers outside the tensor-core envelope: small batches (the reference trains at bs = 128, README.md:16-26)
// and the 1- / 2-wide heads of the baseline and the CMI classifier (VMI.py:13-22 `mlps(..., out=1)`,
// Model.py:52-57).  Plain fp32 FFMA, exact fp32 products and fp32 accumulation -- at these sizes a layer is
// launch-latency bound, the point is that no library GEMM is left on the path.
//
//   mode 0: C[M,N] = A[M,K] . B[N,K]^T (+ bias[N], relu)    Linear forward
//   mode 1: C[M,N] = A[M,K] . B[K,N]                         input gradient  dz W
//   mode 2: C[M,N] = A[K,M]^T . B[K,N]                       weight gradient dz^T x
// (the modes of mimrl_gemm_f32x3).  a_mask (nullable, same shape as A): A is multiplied by (a_mask > 0) first (ReLU
// backward with the saved layer output as mask); colsum (nullable, mode 2 only): colsum[m] += sum_k A'[k, m], the bias
// gradient of the same masked matrix, accumulated by the CTAs of the first column tile.
//
// One kernel, generic element strides: C[i,j] = sum_p A(i,p) B(p,j).  16 x 32 output tile, 64-deep k slices staged in
// shared memory with the next slice prefetched into registers, 128 threads x (1 x 4) outputs.  Long contractions (weight
// gradients over a large batch) are split over blockIdx.z; partial tiles are summed in a fixed order by a second kernel,
// so results are deterministic.
#include "common.cuh"

namespace mimrl {
namespace {

constexpr int kBM = 16, kBN = 32, kBK = 64, kThreads = 128;      // small tiles: 64+ CTAs already at 128 x 256 outputs
constexpr int kLA = kBM * kBK / kThreads, kLB = kBN * kBK / kThreads;   // elements of a k-slice each thread loads

struct SmallParams {
  const float *A, *mask, *B, *bias;
  float *C, *colsum;
  int M, N, K;
  long long sa_i, sa_p, sb_p, sb_j;       // element strides
  int relu, k_per_split, splits;
};

// One k-slice ahead in registers (global -> registers while the previous slice is multiplied out of shared memory): at
// these sizes a launch is a single wave of a few dozen CTAs and its duration is the serial latency of the k loop.
__global__ void __launch_bounds__(kThreads) linear_small_kernel(const SmallParams p) {
  __shared__ float As[kBK][kBM + 1];
  __shared__ float Bs[kBK][kBN + 4];
  const int i0 = blockIdx.x * kBM, j0 = blockIdx.y * kBN;          // row tiles on x: no 65535 limit on the batch
  const int k_lo = blockIdx.z * p.k_per_split, k_hi = min(p.K, k_lo + p.k_per_split);
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;          // outputs (row ty, columns tx*4 .. +3)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float csum = 0.f;                                                // column sum of A' for row (of C) i0 + threadIdx.x
  // loader mapping: which index runs fastest in memory decides which one the lanes walk
  const bool a_i_fast = p.sa_i == 1, b_j_fast = p.sb_j == 1;
  float ra[kLA], rb[kLB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < kLA; ++t) {
      const int e = t * kThreads + threadIdx.x;
      const int ii = a_i_fast ? (e % kBM) : (e / kBK), pp = a_i_fast ? (e / kBM) : (e % kBK);
      const int gi = i0 + ii, gp = k0 + pp;
      float v = 0.f;
      if (gi < p.M && gp < k_hi) {
        const long long off = gi * p.sa_i + gp * p.sa_p;
        v = __ldg(p.A + off);
        if (p.mask && !(__ldg(p.mask + off) > 0.f)) v = 0.f;
      }
      ra[t] = v;
    }
#pragma unroll
    for (int t = 0; t < kLB; ++t) {
      const int e = t * kThreads + threadIdx.x;
      const int jj = b_j_fast ? (e % kBN) : (e / kBK), pp = b_j_fast ? (e / kBN) : (e % kBK);
      const int gj = j0 + jj, gp = k0 + pp;
      rb[t] = (gj < p.N && gp < k_hi) ? __ldg(p.B + gp * p.sb_p + gj * p.sb_j) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int t = 0; t < kLA; ++t) {
      const int e = t * kThreads + threadIdx.x;
      const int ii = a_i_fast ? (e % kBM) : (e / kBK), pp = a_i_fast ? (e / kBM) : (e % kBK);
      As[pp][ii] = ra[t];
    }
#pragma unroll
    for (int t = 0; t < kLB; ++t) {
      const int e = t * kThreads + threadIdx.x;
      const int jj = b_j_fast ? (e % kBN) : (e / kBK), pp = b_j_fast ? (e / kBN) : (e % kBK);
      Bs[pp][jj] = rb[t];
    }
  };
  if (k_lo < k_hi) fetch(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += kBK) {
    stage();
    __syncthreads();
    if (k0 + kBK < k_hi) fetch(k0 + kBK);                          // in flight while this slice is multiplied
#pragma unroll
    for (int pp = 0; pp < kBK; ++pp) {
      const float a = As[pp][ty];
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[pp][tx * 4]);
      acc[0] = fmaf(a, b.x, acc[0]), acc[1] = fmaf(a, b.y, acc[1]), acc[2] = fmaf(a, b.z, acc[2]), acc[3] = fmaf(a, b.w, acc[3]);
    }
    if (p.colsum && blockIdx.y == 0 && threadIdx.x < kBM) {
#pragma unroll
      for (int pp = 0; pp < kBK; ++pp) csum += As[pp][threadIdx.x];
    }
    __syncthreads();
  }
  const int gi = i0 + ty;
  if (gi < p.M) {
    float *dst = p.splits == 1 ? p.C : p.C + (size_t)blockIdx.z * p.M * p.N;      // split mode: C is the partial buffer
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int gj = j0 + tx * 4 + w;
      if (gj >= p.N) continue;
      float v = acc[w];
      if (p.splits == 1) {
        v += p.bias ? p.bias[gj] : 0.f;
        if (p.relu) v = fmaxf(v, 0.f);
      }
      dst[(size_t)gi * p.N + gj] = v;
    }
  }
  if (p.colsum && blockIdx.y == 0 && threadIdx.x < kBM && i0 + threadIdx.x < p.M) {
    if (p.splits == 1) p.colsum[i0 + threadIdx.x] += csum;
    else p.colsum[(size_t)(1 + blockIdx.z) * p.M + i0 + threadIdx.x] = csum;     // partials behind the output vector
  }
}

__global__ void linear_small_reduce_kernel(const float *__restrict__ part, int splits, size_t total,
                                           const float *__restrict__ bias, int N, int relu, float *__restrict__ C) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += part[(size_t)s * total + idx];
    if (bias) a += bias[idx % N];
    C[idx] = relu ? fmaxf(a, 0.f) : a;
  }
}
__global__ void linear_small_colsum_kernel(const float *__restrict__ part, int splits, int M, float *__restrict__ colsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float a = 0.f;
  for (int s = 0; s < splits; ++s) a += part[(size_t)s * M + i];
  colsum[i] += a;
}

int pick_k_splits(int M, int N, int K) {
  const int tiles = ceil_div(M, kBM) * ceil_div(N, kBN);
  // a short contraction is not worth a second (reduce) launch: at the reference's batch of 128 every layer is
  // launch-latency bound, and one launch per product is what counts
  if (tiles >= 148 || K < 2048) return 1;
  int s = ceil_div(2 * 148, tiles);
  const int max_s = ceil_div(K, 4 * kBK);
  s = s > max_s ? max_s : s;
  return s > 64 ? 64 : (s < 1 ? 1 : s);
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_linear_small_workspace_bytes(int mode, int M, int N, int K) {
  (void)mode;
  const int s = pick_k_splits(M, N, K);
  return s == 1 ? 16 : ((size_t)s * M * N + (size_t)(s + 1) * M) * sizeof(float) + 16;
}

extern "C" int mimrl_linear_small(int mode, const float *A, const float *a_mask, const float *B, int M, int N, int K,
                                  const float *bias, int relu, float *C, float *colsum, void *workspace,
                                  size_t workspace_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MIMRL_REQUIRE(mode >= 0 && mode <= 2, "linear_small: unknown mode %d", mode);
  MIMRL_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "linear_small: empty input (M=%d N=%d K=%d)", M, N, K);
  MIMRL_REQUIRE(!colsum || mode == 2, "linear_small: colsum exists for the weight-gradient mode only");
  MIMRL_REQUIRE(workspace_bytes >= mimrl_linear_small_workspace_bytes(mode, M, N, K), "linear_small: workspace too small");
  SmallParams p;
  p.A = A, p.mask = a_mask, p.B = B, p.bias = bias, p.C = C, p.colsum = colsum;
  p.M = M, p.N = N, p.K = K, p.relu = relu;
  if (mode == 0) p.sa_i = K, p.sa_p = 1, p.sb_p = 1, p.sb_j = K;           // A[M,K], B[N,K]
  else if (mode == 1) p.sa_i = K, p.sa_p = 1, p.sb_p = N, p.sb_j = 1;      // A[M,K], B[K,N]
  else p.sa_i = 1, p.sa_p = M, p.sb_p = N, p.sb_j = 1;                     // A[K,M], B[K,N]
  const int splits = pick_k_splits(M, N, K);
  p.splits = splits;
  p.k_per_split = ceil_div(ceil_div(K, splits), kBK) * kBK;
  dim3 grid(ceil_div(M, kBM), ceil_div(N, kBN), splits);
  if (splits == 1) {
    linear_small_kernel<<<grid, kThreads, 0, st>>>(p);
    return check_launch("linear_small");
  }
  float *part = (float *)workspace;
  float *cs_part = part + (size_t)splits * M * N;            // [1 + splits][M]; slot 0 unused (keeps the indexing simple)
  p.C = part;
  p.colsum = colsum ? cs_part : nullptr;
  linear_small_kernel<<<grid, kThreads, 0, st>>>(p);
  if (check_launch("linear_small")) return 1;
  const size_t total = (size_t)M * N;
  int blocks = (int)((total + 255) / 256);
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  linear_small_reduce_kernel<<<blocks, 256, 0, st>>>(part, splits, total, bias, N, relu, C);
  if (check_launch("linear_small_reduce")) return 1;
  if (colsum) {
    linear_small_colsum_kernel<<<ceil_div(M, 256), 256, 0, st>>>(cs_part + M, splits, M, colsum);
    return check_launch("linear_small_colsum");
  }
  return 0;
}
