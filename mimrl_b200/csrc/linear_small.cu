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

// ---------------------------------------------------------------------------------------------------------------
// The whole relu MLP of VMI.py:13-22 / Model.py:52-57 (Linear+ReLU x3, Linear; hidden 256) for SMALL batches in three
// launches instead of sixteen: forward (all four layers, activations of a row group kept in shared memory), data
// gradient (all four layers backwards), and one grouped launch for the four weight + bias gradients.  fp32 FFMA.  A CTA
// owns kTR rows; in the forward thread t owns output feature t (its weight row streams through L1 with 16-byte loads),
// in the backward thread t owns INPUT feature t (weight columns: coalesced across the warp).
namespace mimrl {
namespace {

constexpr int kTR = 4;            // rows per CTA
constexpr int kHidW = 256;        // hidden width (= threads per CTA)
constexpr int kMaxIn = 384;       // widest input (the CMI classifier's [x, y, z] concatenation)

struct Mlp4SmallParams {
  const float *x, *w[4], *b[4];
  float *h[3], *y;                // saved post-ReLU activations [M, 256] x3, output [M, d_out]
  int M, d_in, d_out;
};

__global__ void __launch_bounds__(kHidW) mlp4_small_fwd_kernel(const Mlp4SmallParams p) {
  __shared__ __align__(16) float act[2][kTR][kMaxIn];
  const int r0 = blockIdx.x * kTR, t = threadIdx.x;
  for (int e = t; e < kTR * p.d_in; e += kHidW) {
    const int r = e / p.d_in, k = e - r * p.d_in;
    act[0][r][k] = r0 + r < p.M ? p.x[(size_t)(r0 + r) * p.d_in + k] : 0.f;
  }
  __syncthreads();
  int cur = 0;
#pragma unroll 1
  for (int l = 0; l < 4; ++l) {
    const int K = l == 0 ? p.d_in : kHidW, N = l == 3 ? p.d_out : kHidW;
    if (t < N) {
      float acc[kTR];
      const float bias = p.b[l] ? p.b[l][t] : 0.f;
#pragma unroll
      for (int r = 0; r < kTR; ++r) acc[r] = bias;
      const float *wrow = p.w[l] + (size_t)t * K;
      if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(wrow) & 15) == 0) {
        int k = 0;
        for (; k + 32 <= K; k += 32) {                     // eight 16-byte weight loads in flight
          float4 w4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w4[j] = __ldg(reinterpret_cast<const float4 *>(wrow + k + 4 * j));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int r = 0; r < kTR; ++r) {
              const float4 a = *reinterpret_cast<const float4 *>(&act[cur][r][k + 4 * j]);
              acc[r] = fmaf(w4[j].w, a.w, fmaf(w4[j].z, a.z, fmaf(w4[j].y, a.y, fmaf(w4[j].x, a.x, acc[r]))));
            }
          }
        }
        for (; k < K; k += 4) {
          const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wrow + k));
#pragma unroll
          for (int r = 0; r < kTR; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(&act[cur][r][k]);
            acc[r] = fmaf(w4.w, a.w, fmaf(w4.z, a.z, fmaf(w4.y, a.y, fmaf(w4.x, a.x, acc[r]))));
          }
        }
      } else {
        for (int k = 0; k < K; ++k) {
          const float wv = __ldg(wrow + k);
#pragma unroll
          for (int r = 0; r < kTR; ++r) acc[r] = fmaf(wv, act[cur][r][k], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < kTR; ++r) {
        const float v = l < 3 ? fmaxf(acc[r], 0.f) : acc[r];
        act[cur ^ 1][r][t] = v;
        if (r0 + r < p.M) {
          if (l < 3) p.h[l][(size_t)(r0 + r) * kHidW + t] = v;
          else p.y[(size_t)(r0 + r) * p.d_out + t] = v;
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

struct Mlp4SmallBwdParams {
  const float *gy, *w[4], *h[3];
  float *dz[4], *gx;              // dz_l = dL/d(pre-activation of layer l) [M, N_l]; gx [M, d_in] (nullable)
  int M, d_in, d_out;
};

__global__ void __launch_bounds__(kHidW) mlp4_small_bwd_kernel(const Mlp4SmallBwdParams p) {
  __shared__ __align__(16) float g[2][kTR][kHidW];       // dz of the current layer for the CTA's rows
  const int r0 = blockIdx.x * kTR, t = threadIdx.x;
  for (int e = t; e < kTR * p.d_out; e += kHidW) {
    const int r = e / p.d_out, o = e - r * p.d_out;
    const float v = r0 + r < p.M ? p.gy[(size_t)(r0 + r) * p.d_out + o] : 0.f;
    g[0][r][o] = v;
    if (r0 + r < p.M) p.dz[3][(size_t)(r0 + r) * p.d_out + o] = v;
  }
  __syncthreads();
  int cur = 0;
#pragma unroll 1
  for (int l = 3; l >= 0; --l) {
    const int K = l == 0 ? p.d_in : kHidW, N = l == 3 ? p.d_out : kHidW;       // layer l: [N outputs] x [K inputs]
    if (l == 0 && !p.gx) break;
    for (int k = t; k < K; k += kHidW) {                                          // d_in may exceed 256
      float acc[kTR];
#pragma unroll
      for (int r = 0; r < kTR; ++r) acc[r] = 0.f;
      const float *wcol = p.w[l] + k;
      // sixteen weight loads in flight per thread: the kernel is a chain of L2 round trips otherwise (32 CTAs, every one
      // streams the whole matrix)
      int o = 0;
      for (; o + 16 <= N; o += 16) {
        float wv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = __ldg(wcol + (size_t)(o + j) * K);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
#pragma unroll
          for (int r = 0; r < kTR; ++r) {
            const float4 gv = *reinterpret_cast<const float4 *>(&g[cur][r][o + j]);
            acc[r] = fmaf(wv[j + 3], gv.w, fmaf(wv[j + 2], gv.z, fmaf(wv[j + 1], gv.y, fmaf(wv[j], gv.x, acc[r]))));
          }
        }
      }
      for (; o < N; ++o) {
        const float wv = __ldg(wcol + (size_t)o * K);
#pragma unroll
        for (int r = 0; r < kTR; ++r) acc[r] = fmaf(wv, g[cur][r][o], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kTR; ++r) {
        if (r0 + r >= p.M) continue;
        if (l > 0) {
          const float v = p.h[l - 1][(size_t)(r0 + r) * kHidW + k] > 0.f ? acc[r] : 0.f;       // ReLU mask of layer l-1
          g[cur ^ 1][r][k] = v;
          p.dz[l - 1][(size_t)(r0 + r) * kHidW + k] = v;
        } else {
          p.gx[(size_t)(r0 + r) * p.d_in + k] = acc[r];
        }
      }
      if (l > 0) {
#pragma unroll
        for (int r = 0; r < kTR; ++r)
          if (r0 + r >= p.M) g[cur ^ 1][r][k] = 0.f;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

// Grouped weight gradients: problem q computes C_q[M_q, N_q] = A_q[K, M_q]^T . B_q[K, N_q] and colsum_q[m] = sum_k A_q[k, m]
// (blockIdx.z = problem).  Same tile loop as linear_small_kernel mode 2, no split (K = the small batch).
struct GroupedParams {
  const float *A[4], *B[4];
  float *C[4], *colsum[4];
  int M[4], N[4], K;
};

__global__ void __launch_bounds__(kThreads) linear_small_grouped_kernel(const GroupedParams gp) {
  const int q = blockIdx.z;
  const int M = gp.M[q], N = gp.N[q];
  const int i0 = blockIdx.x * kBM, j0 = blockIdx.y * kBN;
  if (i0 >= M || j0 >= N) return;
  __shared__ float As[kBK][kBM + 1];
  __shared__ float Bs[kBK][kBN + 4];
  const float *A = gp.A[q], *B = gp.B[q];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float csum = 0.f;
  for (int k0 = 0; k0 < gp.K; k0 += kBK) {
#pragma unroll
    for (int t = 0; t < kLA; ++t) {
      const int e = t * kThreads + threadIdx.x, ii = e % kBM, pp = e / kBM;
      As[pp][ii] = (i0 + ii < M && k0 + pp < gp.K) ? __ldg(A + (size_t)(k0 + pp) * M + i0 + ii) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < kLB; ++t) {
      const int e = t * kThreads + threadIdx.x, jj = e % kBN, pp = e / kBN;
      Bs[pp][jj] = (j0 + jj < N && k0 + pp < gp.K) ? __ldg(B + (size_t)(k0 + pp) * N + j0 + jj) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kBK; ++pp) {
      const float a = As[pp][ty];
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[pp][tx * 4]);
      acc[0] = fmaf(a, b.x, acc[0]), acc[1] = fmaf(a, b.y, acc[1]), acc[2] = fmaf(a, b.z, acc[2]), acc[3] = fmaf(a, b.w, acc[3]);
    }
    if (gp.colsum[q] && blockIdx.y == 0 && threadIdx.x < kBM) {
#pragma unroll
      for (int pp = 0; pp < kBK; ++pp) csum += As[pp][threadIdx.x];
    }
    __syncthreads();
  }
  if (i0 + ty < M) {
#pragma unroll
    for (int w = 0; w < 4; ++w)
      if (j0 + tx * 4 + w < N) gp.C[q][(size_t)(i0 + ty) * N + j0 + tx * 4 + w] = acc[w];
  }
  if (gp.colsum[q] && blockIdx.y == 0 && threadIdx.x < kBM && i0 + threadIdx.x < M) gp.colsum[q][i0 + threadIdx.x] = csum;
}

}  // namespace
}  // namespace mimrl

extern "C" int mimrl_mlp4_small_supported(int d_in, int hidden, int d_out) {
  return hidden == kHidW && d_in >= 1 && d_in <= kMaxIn && d_out >= 1 && d_out <= kHidW;
}

// y = W4 relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3) + b4 for x [M, d_in]; h1..h3 [M, 256] are saved for the backward.
extern "C" int mimrl_mlp4_small_fwd(const float *x, int M, int d_in, const float *w1, const float *b1, const float *w2,
                                    const float *b2, const float *w3, const float *b3, const float *w4, const float *b4,
                                    int d_out, float *h1, float *h2, float *h3, float *y, void *stream) {
  MIMRL_REQUIRE(mimrl_mlp4_small_supported(d_in, kHidW, d_out), "mlp4_small_fwd: d_in=%d d_out=%d not supported", d_in, d_out);
  MIMRL_REQUIRE(M > 0 && x && w1 && w2 && w3 && w4 && h1 && h2 && h3 && y, "mlp4_small_fwd: bad arguments");
  Mlp4SmallParams p;
  p.x = x, p.w[0] = w1, p.w[1] = w2, p.w[2] = w3, p.w[3] = w4, p.b[0] = b1, p.b[1] = b2, p.b[2] = b3, p.b[3] = b4;
  p.h[0] = h1, p.h[1] = h2, p.h[2] = h3, p.y = y, p.M = M, p.d_in = d_in, p.d_out = d_out;
  mlp4_small_fwd_kernel<<<ceil_div(M, kTR), kHidW, 0, (cudaStream_t)stream>>>(p);
  return check_launch("mlp4_small_fwd");
}

// Backward: dz1..dz3 [M, 256], dz4 [M, d_out] (scratch, caller-allocated), gx [M, d_in] (nullable), then the weight
// gradients gw_l = dz_l^T input_l and bias gradients gb_l = column sums of dz_l (nullable) in one grouped launch.
extern "C" int mimrl_mlp4_small_bwd(const float *gy, const float *x, int M, int d_in, int d_out, const float *w1,
                                    const float *w2, const float *w3, const float *w4, const float *h1, const float *h2,
                                    const float *h3, float *dz1, float *dz2, float *dz3, float *dz4, float *gx, float *gw1,
                                    float *gb1, float *gw2, float *gb2, float *gw3, float *gb3, float *gw4, float *gb4,
                                    void *stream) {
  MIMRL_REQUIRE(mimrl_mlp4_small_supported(d_in, kHidW, d_out), "mlp4_small_bwd: d_in=%d d_out=%d not supported", d_in, d_out);
  MIMRL_REQUIRE(M > 0 && gy && x && w1 && w2 && w3 && w4 && h1 && h2 && h3 && dz1 && dz2 && dz3 && dz4, "mlp4_small_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Mlp4SmallBwdParams p;
  p.gy = gy, p.w[0] = w1, p.w[1] = w2, p.w[2] = w3, p.w[3] = w4, p.h[0] = h1, p.h[1] = h2, p.h[2] = h3;
  p.dz[0] = dz1, p.dz[1] = dz2, p.dz[2] = dz3, p.dz[3] = dz4, p.gx = gx, p.M = M, p.d_in = d_in, p.d_out = d_out;
  mlp4_small_bwd_kernel<<<ceil_div(M, kTR), kHidW, 0, st>>>(p);
  if (check_launch("mlp4_small_bwd")) return 1;
  if (!gw1 && !gw2 && !gw3 && !gw4) return 0;
  GroupedParams gp;
  const float *dz[4] = {dz1, dz2, dz3, dz4}, *in[4] = {x, h1, h2, h3};
  float *gw[4] = {gw1, gw2, gw3, gw4}, *gb[4] = {gb1, gb2, gb3, gb4};
  int max_m = 0, max_n = 0, n = 0;
  for (int l = 0; l < 4; ++l) {
    if (!gw[l]) continue;
    gp.A[n] = dz[l], gp.B[n] = in[l], gp.C[n] = gw[l], gp.colsum[n] = gb[l];
    gp.M[n] = l == 3 ? d_out : kHidW, gp.N[n] = l == 0 ? d_in : kHidW;
    max_m = gp.M[n] > max_m ? gp.M[n] : max_m, max_n = gp.N[n] > max_n ? gp.N[n] : max_n;
    ++n;
  }
  for (int q = n; q < 4; ++q) gp.A[q] = gp.B[q] = nullptr, gp.C[q] = gp.colsum[q] = nullptr, gp.M[q] = gp.N[q] = 0;
  gp.K = M;
  linear_small_grouped_kernel<<<dim3(ceil_div(max_m, kBM), ceil_div(max_n, kBN), n), kThreads, 0, st>>>(gp);
  return check_launch("linear_small_grouped");
}
