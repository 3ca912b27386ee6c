// fp32-class GEMM on tcgen05 for the critic / baseline / classifier MLPs
// (reference VMI.py:13-22 `mlps`, Model.py:52-57): the three contractions of a
// Linear layer's forward and backward, with bias + ReLU (forward) and the ReLU
// mask (backward) fused in.
//
//   mode 0 (NT)  C[M,N] = A[M,K] . B[N,K]^T   Linear forward      y  = x W^T + b
//   mode 1 (NN)  C[M,N] = A[M,K] . B[K,N]     input gradient      dx = dz W
//   mode 2 (TN)  C[M,N] = A[K,M]^T . B[K,N]   weight gradient     dW = dz^T x   (split over K = batch rows)
//
// Same arithmetic as the score sweeps (sep_tc.cu): each fp32 operand is scaled
// by a per-tensor power of two and split into fp16 hi + lo; every product runs
// hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM.  Operands keep their
// natural row-major layout: a matrix whose contraction index is the slow one is
// read as an MN-major UMMA operand, so no transposed copies are made.
// CTA: warp 0 TMA producer (3-stage ring of 64 KB), warp 1 MMA issuer, warps 2-5
// epilogue (thread = output row).
#include "tc_common.cuh"

namespace mimrl {
namespace {

constexpr int kGemmThreads = 192;
constexpr int kGStages = 3;
constexpr uint32_t kOp16 = 16384;                 // one operand half (hi or lo) of a stage: 16 KB
constexpr uint32_t kGStage = 4 * kOp16;           // A hi, A lo, B hi, B lo
constexpr uint32_t kGemmSmem = kGStages * kGStage + 512 + 1024;

// absmax[0] = max |A (masked)|, absmax[1] = max |B|.  float4 grid-stride loads, 4 independent chains per thread.
__global__ void __launch_bounds__(256)
gemm_absmax_kernel(const float *__restrict__ a, const float *__restrict__ mask, size_t na,
                   const float *__restrict__ b, size_t nb, unsigned *__restrict__ out) {
  const bool second = blockIdx.y == 1;
  const float *src = second ? b : a;
  const float *msk = second ? nullptr : mask;
  const size_t n = second ? nb : na;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if ((((size_t)src | (size_t)msk) & 15) == 0) {
    const size_t n4 = n >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    const float4 *k4 = reinterpret_cast<const float4 *>(msk);
    for (size_t i = tid; i < n4; i += stride) {
      float4 v = __ldg(s4 + i);
      if (msk) {
        const float4 k = __ldg(k4 + i);
        v.x = k.x > 0.f ? v.x : 0.f, v.y = k.y > 0.f ? v.y : 0.f, v.z = k.z > 0.f ? v.z : 0.f, v.w = k.w > 0.f ? v.w : 0.f;
      }
      m0 = fmaxf(m0, fabsf(v.x)), m1 = fmaxf(m1, fabsf(v.y)), m2 = fmaxf(m2, fabsf(v.z)), m3 = fmaxf(m3, fabsf(v.w));
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) {
      float v = src[i];
      if (msk && !(msk[i] > 0.f)) v = 0.f;
      m0 = fmaxf(m0, fabsf(v));
    }
  } else {
    for (size_t i = tid; i < n; i += stride) {
      float v = src[i];
      if (msk && !(msk[i] > 0.f)) v = 0.f;
      m0 = fmaxf(m0, fabsf(v));
    }
  }
  float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out + blockIdx.y, __float_as_uint(m));
}

// src [rows, cols] fp32 (optionally masked) -> hi, lo [rows, ld] fp16, zero padded, scaled by 2^k
// colsum (nullable): column sums of the masked source are ACCUMULATED there (bias gradient); the launch makes the
// grid stride a multiple of ld/2 so that every thread stays on one column pair.
// hmask (nullable, [rows, ld] fp16): alternative mask source -- the hi half of an activation split by mlp_tc.cu
__global__ void gemm_split_kernel(const float *__restrict__ src, const float *__restrict__ mask, int rows, int cols,
                                  int ld, const unsigned *__restrict__ absmax, int which, __half *__restrict__ hi,
                                  __half *__restrict__ lo, float *__restrict__ colsum,
                                  const __half *__restrict__ hmask = nullptr) {
  const float sc = scale_from_absmax(absmax[which]);
  const int half_ld = ld >> 1;
  const size_t total = (size_t)rows * half_ld;
  float cs0 = 0.f, cs1 = 0.f;
  int my_c = -1;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / half_ld;
    const int c = (int)(idx - r * half_ld) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (c < cols) {
      v0 = src[r * cols + c];
      if (mask && !(mask[r * cols + c] > 0.f)) v0 = 0.f;
      if (hmask && !(__half2float(hmask[r * ld + c]) > 0.f)) v0 = 0.f;
    }
    if (c + 1 < cols) {
      v1 = src[r * cols + c + 1];
      if (mask && !(mask[r * cols + c + 1] > 0.f)) v1 = 0.f;
      if (hmask && !(__half2float(hmask[r * ld + c + 1]) > 0.f)) v1 = 0.f;
    }
    cs0 += v0, cs1 += v1, my_c = c;
    v0 *= sc, v1 *= sc;
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__half2 *>(hi + r * ld + c) = h;
    *reinterpret_cast<__half2 *>(lo + r * ld + c) = l;
  }
  if (colsum && my_c >= 0) {
    if (my_c < cols) atomicAdd(colsum + my_c, cs0);
    if (my_c + 1 < cols) atomicAdd(colsum + my_c + 1, cs1);
  }
}

struct GemmParams {
  int M, N, K, kblocks_per_split, relu, splits;
  const unsigned *absmax;     // [0] A, [1] B -- or, when absmax_b is set, absmax[0] = A and absmax_b[0] = B
  const unsigned *absmax_b;
  const float *bias;
  float *C;       // [M, N] when splits == 1, else partials [splits][M][N]
  int atomic_out; // split-K partial sums are ADDED into C [M, N] with red.global (no partial buffer, no reduce pass)
  // narrow operands (weight gradients of the CubeMLP mixes): MMA N (multiple of 16), how many 64-wide boxes of each
  // MN-major operand exist (the second one is not loaded when the operand has <= 64 features), output transposed
  int n_mma, a_boxes, b_boxes, trans_out, ldc;
  int b_rows;     // blocked-K layout: box height of the B operand (its rows rounded up to 16)
  // blocked-K layout: ring geometry chosen by the host from the box heights (narrow operands -> small stages -> a deep
  // ring, so that enough bytes are in flight per SM); 0 stages = the fixed 3 x 64 KB ring
  int n_stages, kbs;          // kbs: k-blocks per stage (one TMA box per plane carries kbs consecutive k-tiles)
  uint32_t stage_bytes, a_plane, b_plane;          // plane = kbs x (box rows x 128 bytes)
};

// A_MN / B_MN: operand stored with the contraction index as the SLOW one (read MN-major)
// BLK: both operands K-major in the blocked-K layout of make_map_blocked (3-D tensor maps)
template <bool A_MN, bool B_MN, bool BLK = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kGStages * kGStage;
  const uint32_t bFull = bars, bEmpty = bars + 128;          // up to 16 stages
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kGStages * kGStage + 384);
  const bool ring = BLK && p.n_stages > 0;
  const int n_st = ring ? p.n_stages : kGStages;
  const uint32_t st_bytes = ring ? p.stage_bytes : kGStage;
  const uint32_t o_a1 = ring ? p.a_plane : kOp16, o_b0 = ring ? 2u * p.a_plane : 2u * kOp16,
                 o_b1 = ring ? 2u * p.a_plane + p.b_plane : 3u * kOp16;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // Persistent over work items (m tile, n tile, k split): the ring keeps streaming across items and the two TMEM
  // accumulators alternate, so the epilogue of one item overlaps the loads and MMAs of the next.
  const int n_kb = (p.K + 63) / 64;
  const int tiles_m = (p.M + 127) / 128, tiles_n = (p.N + 127) / 128;
  const long long n_items = (long long)tiles_m * tiles_n * p.splits;
  const uint32_t bAccFull = bars + 256, bAccEmpty = bars + 272;
  auto item_of = [&](long long it, int &m0, int &n0, int &split, int &kb0, int &T) {
    split = (int)(it / ((long long)tiles_m * tiles_n));
    const int rem = (int)(it - (long long)split * tiles_m * tiles_n);
    m0 = (rem / tiles_n) * 128, n0 = (rem % tiles_n) * 128;
    kb0 = split * p.kblocks_per_split;
    const int kb1 = min(n_kb, kb0 + p.kblocks_per_split);
    T = kb1 > kb0 ? kb1 - kb0 : 0;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_st; ++i) {
      mbar_init(bFull + 8 * i, 1);
      mbar_init(bEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bAccFull + 8 * i, 1);
      mbar_init(bAccEmpty + 8 * i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kGStages * kGStage + 384), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      prefetch_tmap(&map_a_hi);
      prefetch_tmap(&map_b_hi);
      uint32_t n = 0;
      for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
        int m0, n0, split, kb0, T;
        item_of(it, m0, n0, split, kb0, T);
        const int kstep = ring ? p.kbs : 1;
        for (int i = 0; i < T; i += kstep, ++n) {
          const int stage = n % n_st, k0 = (kb0 + i) * 64;
          mbar_wait(bEmpty + 8 * stage, ((n / n_st) & 1) ^ 1);
          const uint32_t fb = bFull + 8 * stage;
          mbar_expect_tx(fb, (A_MN && B_MN) ? (uint32_t)(p.a_boxes + p.b_boxes) * 2u * 8192u
                                 : ring ? 2u * p.a_plane + 2u * p.b_plane
                                 : BLK ? 2u * kOp16 + 2u * (uint32_t)p.b_rows * 128u : kGStage);
          const uint32_t dst = base + stage * st_bytes;
          if (A_MN) {   // rows = contraction index, 64 per box; two boxes cover 128 M
            tma_load_2d(dst + 0 * kOp16, &map_a_hi, fb, m0, k0);
            tma_load_2d(dst + 1 * kOp16, &map_a_lo, fb, m0, k0);
            if (!B_MN || p.a_boxes == 2) {
              tma_load_2d(dst + 0 * kOp16 + 8192, &map_a_hi, fb, m0 + 64, k0);
              tma_load_2d(dst + 1 * kOp16 + 8192, &map_a_lo, fb, m0 + 64, k0);
            }
          } else if (BLK) {
            tma_load_3d(dst, &map_a_hi, fb, 0, m0, kb0 + i);
            tma_load_3d(dst + o_a1, &map_a_lo, fb, 0, m0, kb0 + i);
          } else {
            tma_load_2d(dst + 0 * kOp16, &map_a_hi, fb, k0, m0);
            tma_load_2d(dst + 1 * kOp16, &map_a_lo, fb, k0, m0);
          }
          if (B_MN) {
            tma_load_2d(dst + 2 * kOp16, &map_b_hi, fb, n0, k0);
            tma_load_2d(dst + 3 * kOp16, &map_b_lo, fb, n0, k0);
            if (!A_MN || p.b_boxes == 2) {
              tma_load_2d(dst + 2 * kOp16 + 8192, &map_b_hi, fb, n0 + 64, k0);
              tma_load_2d(dst + 3 * kOp16 + 8192, &map_b_lo, fb, n0 + 64, k0);
            }
          } else if (BLK) {
            tma_load_3d(dst + o_b0, &map_b_hi, fb, 0, n0, kb0 + i);
            tma_load_3d(dst + o_b1, &map_b_lo, fb, 0, n0, kb0 + i);
          } else {
            tma_load_2d(dst + 2 * kOp16, &map_b_hi, fb, k0, n0);
            tma_load_2d(dst + 3 * kOp16, &map_b_lo, fb, k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = instr_desc_f16(128, p.n_mma) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
    uint32_t n = 0, j = 0;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      int m0, n0, split, kb0, T;
      item_of(it, m0, n0, split, kb0, T);
      const uint32_t acc = j & 1, tacc = tmem_base + acc * 128;
      mbar_wait(bAccEmpty + 8 * acc, ((j >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
      tc_fence_after();
      const int kstep = ring ? p.kbs : 1;
      const uint32_t a_sub = ring ? p.a_plane / (uint32_t)p.kbs : 0u, b_sub = ring ? p.b_plane / (uint32_t)p.kbs : 0u;
      for (int i = 0; i < T; i += kstep, ++n) {
        const int stage = n % n_st;
        mbar_wait(bFull + 8 * stage, (n / n_st) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t s0 = base + stage * st_bytes;
          const int subs = min(kstep, T - i);          // (a box past the end of this split is loaded but not used)
          for (int sub = 0; sub < subs; ++sub) {
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
              const uint32_t a_base = s0 + (prod == 2 ? o_a1 : 0u) + sub * a_sub;               // A hi, hi, lo
              const uint32_t b_base = s0 + (prod == 1 ? o_b1 : o_b0) + sub * b_sub;             // B hi, lo, hi
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = A_MN ? smem_desc_sw128_mn(a_base + k * 2048, 8192, 1024) : smem_desc_sw128(a_base + k * 32);
                const uint64_t bd = B_MN ? smem_desc_sw128_mn(b_base + k * 2048, 8192, 1024) : smem_desc_sw128(b_base + k * 32);
                umma_f16(tacc, ad, bd, idesc, (i | sub | prod | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(bEmpty + 8 * stage);
        }
        __syncwarp();
      }
      if (leader) umma_commit(bAccFull + 8 * acc);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const float inv = 1.f / (scale_from_absmax(p.absmax[0]) * scale_from_absmax(p.absmax_b ? p.absmax_b[0] : p.absmax[1]));
    const bool partial = p.splits > 1;
    uint32_t j = 0;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      int m0, n0, split, kb0, T;
      item_of(it, m0, n0, split, kb0, T);
      const uint32_t acc = j & 1;
      const int row = m0 + r;
      float *out = p.C + ((size_t)(p.atomic_out ? 0 : split) * p.M + row) * (p.atomic_out ? p.ldc : p.N);
      mbar_wait(bAccFull + 8 * acc, (j >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        if (T > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + ch * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = 0u;
        }
        if (row >= p.M) continue;
        const int c0 = n0 + ch * 32;
        if (c0 >= p.N) continue;
        if (p.atomic_out) {
          if (T > 0) {
            if (p.trans_out) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (c0 + jj < p.N) atomicAdd(p.C + (size_t)(c0 + jj) * p.ldc + row, __uint_as_float(v[jj]) * inv);
            } else if (c0 + 32 <= p.N && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0) {
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4)          // one 16-byte reduction per four outputs
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c0 + jj),
                             "f"(__uint_as_float(v[jj]) * inv), "f"(__uint_as_float(v[jj + 1]) * inv),
                             "f"(__uint_as_float(v[jj + 2]) * inv), "f"(__uint_as_float(v[jj + 3]) * inv)
                             : "memory");
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (c0 + jj < p.N) atomicAdd(out + c0 + jj, __uint_as_float(v[jj]) * inv);
            }
          }
        } else if (c0 + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float x = __uint_as_float(v[jj + u]) * inv;
              if (!partial) {
                if (p.bias) x += __ldg(p.bias + c0 + jj + u);
                if (p.relu) x = fmaxf(x, 0.f);
              }
              o[u] = x;
            }
            *reinterpret_cast<float4 *>(out + c0 + jj) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            if (c0 + jj < p.N) {
              float x = __uint_as_float(v[jj]) * inv;
              if (!partial) {
                if (p.bias) x += __ldg(p.bias + c0 + jj);
                if (p.relu) x = fmaxf(x, 0.f);
              }
              out[c0 + jj] = x;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bAccEmpty + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

__global__ void gemm_reduce_kernel(const float *__restrict__ part, int splits, size_t mn, int N, const float *__restrict__ bias,
                                   int relu, float *__restrict__ C) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      a0 += part[(size_t)s * mn + i];
      a1 += part[(size_t)(s + 1) * mn + i];
      a2 += part[(size_t)(s + 2) * mn + i];
      a3 += part[(size_t)(s + 3) * mn + i];
    }
    for (; s < splits; ++s) a0 += part[(size_t)s * mn + i];
    float r = (a0 + a1) + (a2 + a3);
    if (bias) r += bias[i % (size_t)N];
    C[i] = relu ? fmaxf(r, 0.f) : r;
  }
}

struct GemmLayout {
  int a_rows, a_cols, b_rows, b_cols, lda, ldb, splits;
  size_t off_absmax, off_a_hi, off_a_lo, off_b_hi, off_b_lo, off_part, total;
};

GemmLayout gemm_layout(int mode, int M, int N, int K) {
  GemmLayout g;
  // stored shapes of the operands (rows x cols, row-major)
  g.a_rows = mode == 2 ? K : M;
  g.a_cols = mode == 2 ? M : K;
  g.b_rows = mode == 0 ? N : K;
  g.b_cols = mode == 0 ? K : N;
  g.lda = (g.a_cols + 63) & ~63;
  g.ldb = (g.b_cols + 63) & ~63;
  const int n_kb = (K + 63) / 64;
  const int tiles = ceil_div(M, 128) * ceil_div(N, 128);
  int splits = 1;
  if (mode == 2) {   // few output tiles, long contraction: spread the k-blocks over the SMs
    splits = ceil_div(2 * 148, tiles);
    if (splits > n_kb) splits = n_kb;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    // The tensor core adds into its fp32 accumulator with truncation, so the error of one accumulator grows
    // linearly with the number of k-steps (measured: 1e-4 relative after 3072 of them).  Keep a split to 32
    // k-blocks (384 accumulations, ~1e-5) and let the round-to-nearest reduction add the partial sums.
  }
  constexpr int kMaxKbPerSplit = 32;
  if (n_kb > (mode == 2 ? kMaxKbPerSplit : 4 * kMaxKbPerSplit) && ceil_div(n_kb, splits) > kMaxKbPerSplit)
    splits = ceil_div(n_kb, kMaxKbPerSplit);
  if (splits > 1) {
    const int per = ceil_div(n_kb, splits);
    splits = ceil_div(n_kb, per);
  }
  g.splits = splits;
  size_t o = 0;
  g.off_absmax = o, o += 256;
  g.off_a_hi = o, o += align256((size_t)g.a_rows * g.lda * 2);
  g.off_a_lo = o, o += align256((size_t)g.a_rows * g.lda * 2);
  g.off_b_hi = o, o += align256((size_t)g.b_rows * g.ldb * 2);
  g.off_b_lo = o, o += align256((size_t)g.b_rows * g.ldb * 2);
  g.off_part = o, o += splits > 1 ? align256((size_t)splits * M * N * sizeof(float)) : 0;
  g.total = o;
  return g;
}

template <bool A_MN, bool B_MN, bool BLK = false>
int launch_gemm(const CUtensorMap &ah, const CUtensorMap &al, const CUtensorMap &bh, const CUtensorMap &bl,
                const GemmParams &p, dim3 grid, cudaStream_t st) {
  cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
  const long long items = (long long)grid.x * grid.y * grid.z;          // persistent: one CTA per SM walks the items
  gemm_tc_kernel<A_MN, B_MN, BLK><<<(int)(items < 148 ? items : 148), kGemmThreads, kGemmSmem, st>>>(ah, al, bh, bl, p);
  return check_launch("gemm_tc");
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

namespace {
struct SplitLayout {
  int ld;
  size_t off_hi, off_lo, total;
};
SplitLayout split_layout(int rows, int cols) {
  SplitLayout L;
  L.ld = (cols + 63) & ~63;
  L.off_hi = 256;
  L.off_lo = 256 + align256((size_t)rows * L.ld * 2);
  L.total = L.off_lo + align256((size_t)rows * L.ld * 2);
  return L;
}
}  // namespace

extern "C" size_t mimrl_split_bytes(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return split_layout(rows, cols).total;
}

static int split_f32_impl(const float *src, const float *mask, const __half *hmask, int rows, int cols, void *out,
                          float *colsum, void *stream) {
  MIMRL_REQUIRE(rows > 0 && cols > 0 && src && out, "split_f32: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  const SplitLayout L = split_layout(rows, cols);
  unsigned char *o = (unsigned char *)out;
  unsigned *absmax = reinterpret_cast<unsigned *>(o);
  cudaMemsetAsync(absmax, 0, 8, st);
  const size_t n = (size_t)rows * cols;
  int blocks = (int)((n + 4095) / 4096);
  blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
  gemm_absmax_kernel<<<dim3(blocks, 1), 256, 0, st>>>(src, mask, n, nullptr, 0, absmax);
  if (check_launch("split absmax")) return 1;
  // grid stride (blocks * 256 threads) must be a multiple of ld/2 (64, 128 or 192 for the widths in use): 3k blocks
  const size_t total = (size_t)rows * (L.ld / 2);
  int b = (int)((total + 255) / 256);
  b = b > 148 * 6 ? 148 * 6 : b;
  if (colsum) {
    const int unit = (L.ld / 2) / 64 > 0 ? 3 * (L.ld / 2) : 3;     // generous multiple; exact check below
    (void)unit;
    while (b > 1 && ((size_t)b * 256) % (size_t)(L.ld / 2) != 0) --b;
    MIMRL_REQUIRE(((size_t)b * 256) % (size_t)(L.ld / 2) == 0, "split_f32: column sums need ld/2 to divide the grid stride");
  }
  gemm_split_kernel<<<b, 256, 0, st>>>(src, mask, rows, cols, L.ld, absmax, 0, reinterpret_cast<__half *>(o + L.off_hi),
                                     reinterpret_cast<__half *>(o + L.off_lo), colsum, hmask);
  return check_launch("split_f32");
}

extern "C" int mimrl_split_f32(const float *src, const float *mask, int rows, int cols, void *out, float *colsum,
                               void *stream) {
  return split_f32_impl(src, mask, nullptr, rows, cols, out, colsum, stream);
}

// Same, with the mask taken from another split buffer of the same shape: entries whose hi half is not positive are
// zeroed (ReLU backward against an activation that only exists as the fp16 operand written by mimrl_mlp4_fwd).
extern "C" int mimrl_split_f32_hmask(const float *src, const void *mask_split, int rows, int cols, void *out, float *colsum,
                                     void *stream) {
  MIMRL_REQUIRE(mask_split, "split_f32_hmask: no mask operand");
  const SplitLayout L = split_layout(rows, cols);
  return split_f32_impl(src, nullptr, reinterpret_cast<const __half *>((const unsigned char *)mask_split + L.off_hi), rows, cols,
                        out, colsum, stream);
}

// GEMM on operands already split by mimrl_split_f32.  Stored shapes: mode 0: A [M,K], B [N,K]; mode 1: A [M,K],
// B [K,N]; mode 2: A [K,M], B [K,N].
extern "C" size_t mimrl_gemm_split_workspace_bytes(int mode, int M, int N, int K) {
  if (mode < 0 || mode > 2 || M <= 0 || N <= 0 || K <= 0) return 0;
  const GemmLayout g = gemm_layout(mode, M, N, K);
  return (g.splits > 1 ? align256((size_t)g.splits * M * N * sizeof(float)) : 0) + 256;
}

extern "C" int mimrl_gemm_split(int mode, const void *a_split, const void *b_split, int M, int N, int K,
                                const float *bias, int relu, float *C, void *workspace, size_t workspace_bytes,
                                void *stream) {
  MIMRL_REQUIRE(mode >= 0 && mode <= 2, "gemm_split: unknown mode %d", mode);
  MIMRL_REQUIRE(M > 0 && N > 0 && K > 0 && a_split && b_split, "gemm_split: empty problem");
  const GemmLayout g = gemm_layout(mode, M, N, K);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_gemm_split_workspace_bytes(mode, M, N, K), "gemm_split: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const SplitLayout la = split_layout(g.a_rows, g.a_cols), lb = split_layout(g.b_rows, g.b_cols);
  const unsigned char *pa = (const unsigned char *)a_split, *pb = (const unsigned char *)b_split;
  // the kernel reads absmax[0] (A) and absmax[1] (B) from one array: gather the two headers
  unsigned *absmax = reinterpret_cast<unsigned *>((unsigned char *)workspace + workspace_bytes - 256);
  cudaMemcpyAsync(absmax, pa, 4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(absmax + 1, pb, 4, cudaMemcpyDeviceToDevice, st);
  const bool a_mn = mode == 2, b_mn = mode != 0;
  CUtensorMap ah, al, bh, bl;
  if (make_map(&ah, pa + la.off_hi, g.a_cols, g.a_rows, la.ld, a_mn ? 64 : 128)) return 1;
  if (make_map(&al, pa + la.off_lo, g.a_cols, g.a_rows, la.ld, a_mn ? 64 : 128)) return 1;
  if (make_map(&bh, pb + lb.off_hi, g.b_cols, g.b_rows, lb.ld, b_mn ? 64 : 128)) return 1;
  if (make_map(&bl, pb + lb.off_lo, g.b_cols, g.b_rows, lb.ld, b_mn ? 64 : 128)) return 1;
  GemmParams p{};
  p.M = M, p.N = N, p.K = K;
  p.kblocks_per_split = ceil_div(ceil_div(K, 64), g.splits);
  p.splits = g.splits;
  p.relu = g.splits > 1 ? 0 : relu;          // split-K: bias and ReLU are applied by the reduction
  p.absmax = absmax;
  p.absmax_b = nullptr, p.atomic_out = 0;
  p.n_mma = 128, p.a_boxes = 2, p.b_boxes = 2, p.trans_out = 0, p.ldc = N, p.b_rows = 128;
  p.bias = g.splits > 1 ? nullptr : bias;
  p.C = g.splits > 1 ? reinterpret_cast<float *>(workspace) : C;
  dim3 grid(ceil_div(M, 128), ceil_div(N, 128), g.splits);
  int rc;
  if (mode == 0) rc = launch_gemm<false, false>(ah, al, bh, bl, p, grid, st);
  else if (mode == 1) rc = launch_gemm<false, true>(ah, al, bh, bl, p, grid, st);
  else rc = launch_gemm<true, true>(ah, al, bh, bl, p, grid, st);
  if (rc) return rc;
  if (g.splits > 1) {
    const size_t mn = (size_t)M * N;
    int b = (int)((mn + 255) / 256);
    b = b > 148 * 8 ? 148 * 8 : b;
    gemm_reduce_kernel<<<b, 256, 0, st>>>(p.C, g.splits, mn, N, bias, relu, C);
    return check_launch("gemm reduce");
  }
  return 0;
}

// C[M,N] = A[M,K] . B[N,K]^T with both operands in the blocked-K layout (tiles of 64 consecutive k, each [rows][64]
// contiguous; same header / hi / lo placement and total size as mimrl_split_f32 of a [rows, K] matrix, K % 64 == 0).
// For contractions over millions of entries (weight gradients over pairs or fibres): every TMA box is one
// contiguous run in memory.  Workspace: mimrl_gemm_split_workspace_bytes(0, M, N, K).
extern "C" int mimrl_gemm_split_blocked(const void *a_split, const void *b_split, int M, int N, int K, float *C,
                                        void *workspace, size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(M > 0 && N > 0 && K > 0 && (K % 64) == 0 && a_split && b_split && C, "gemm_split_blocked: bad arguments");
  const GemmLayout g = gemm_layout(0, M, N, K);
  MIMRL_REQUIRE(workspace_bytes >= mimrl_gemm_split_workspace_bytes(0, M, N, K), "gemm_split_blocked: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const SplitLayout la = split_layout(M, K), lb = split_layout(N, K);
  const unsigned char *pa = (const unsigned char *)a_split, *pb = (const unsigned char *)b_split;
  unsigned *absmax = reinterpret_cast<unsigned *>((unsigned char *)workspace + workspace_bytes - 256);
  cudaMemcpyAsync(absmax, pa, 4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(absmax + 1, pb, 4, cudaMemcpyDeviceToDevice, st);
  CUtensorMap ah, al, bh, bl;
  const uint64_t kt = (uint64_t)K / 64;
  if (make_map_blocked(&ah, pa + la.off_hi, M, kt, 128) || make_map_blocked(&al, pa + la.off_lo, M, kt, 128) ||
      make_map_blocked(&bh, pb + lb.off_hi, N, kt, 128) || make_map_blocked(&bl, pb + lb.off_lo, N, kt, 128))
    return 1;
  GemmParams p{};
  p.M = M, p.N = N, p.K = K;
  p.kblocks_per_split = ceil_div(ceil_div(K, 64), g.splits);
  p.splits = g.splits;
  p.relu = 0;
  p.absmax = absmax;
  p.absmax_b = nullptr, p.atomic_out = 0;
  p.n_mma = 128, p.a_boxes = 2, p.b_boxes = 2, p.trans_out = 0, p.ldc = N, p.b_rows = 128;
  p.bias = nullptr;
  p.C = g.splits > 1 ? reinterpret_cast<float *>(workspace) : C;
  dim3 grid(ceil_div(M, 128), ceil_div(N, 128), g.splits);
  if (int rc = launch_gemm<false, false, true>(ah, al, bh, bl, p, grid, st)) return rc;
  if (g.splits > 1) {
    const size_t mn = (size_t)M * N;
    int b = (int)((mn + 255) / 256);
    b = b > 148 * 8 ? 148 * 8 : b;
    gemm_reduce_kernel<<<b, 256, 0, st>>>(p.C, g.splits, mn, N, nullptr, 0, C);
    return check_launch("gemm reduce");
  }
  return 0;
}

// Same contraction, ADDED into C [M, N] (the caller zero-fills or passes a running sum): every split-K work item adds
// its partial sum with red.global.add.f32 -- no partial buffer, no reduction launch, no header copies (the kernel reads
// the two absmax headers in place).  The order of the additions is not fixed, so the last bits can differ from run to
// run; the CubeMLP weight gradients use it, the critic MLPs keep the deterministic two-stage sum.
extern "C" int mimrl_gemm_split_blocked_acc(const void *a_split, const void *b_split, int M, int N, int K, float *C,
                                            void *stream) {
  MIMRL_REQUIRE(M > 0 && N > 0 && M <= 128 && N <= 128 && K > 0 && (K % 64) == 0 && a_split && b_split && C,
                "gemm_split_blocked_acc: bad arguments (M, N <= 128, K %% 64 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  // the wider operand takes the 128 MMA rows, the narrower one the MMA columns rounded up to 16 (a 10 x 50 gradient
  // is a 128 x 16 MMA, not 128 x 128) and a TMA box of that height; the output is written transposed when swapped
  const bool swap = N > M;
  const unsigned char *pm = (const unsigned char *)(swap ? b_split : a_split), *pn = (const unsigned char *)(swap ? a_split : b_split);
  const int Mm = swap ? N : M, Nn = swap ? M : N;
  const int n_pad = (Nn + 15) & ~15;
  const SplitLayout lm = split_layout(Mm, K), ln = split_layout(Nn, K);
  CUtensorMap mh, ml, nh, nl;
  const uint64_t kt = (uint64_t)K / 64;
  // box heights = the operands' own rows (rounded to the 8-row swizzle atom / to the MMA N): a stage is as small as the
  // data it carries and the ring as deep as fits (<= 16 stages, 176 KB -- the MMA reads 128 rows from the A plane, so
  // the last 16 KB of the ring stay unused; rows past the box are other planes' finite or not, they only reach
  // accumulator rows >= M that are never written out)
  static const bool deep = !getenv("MIMRL_GEMM_RING_OFF");
  const int m_box = deep ? (Mm + 7) & ~7 : 128;
  // several consecutive k-tiles per TMA box when the operands are narrow (a 10 x 10 gradient over 393216 fibres was
  // bound by the number of 2 KB boxes the TMA unit issues, not by bytes): ~40 KB per stage
  int kbs = 1;
  if (deep) {
    kbs = (int)(40960u / (2u * (uint32_t)(m_box + n_pad) * 128u));
    kbs = kbs < 1 ? 1 : (kbs > 8 ? 8 : kbs);
    if (getenv("MIMRL_GEMM_KBS")) kbs = atoi(getenv("MIMRL_GEMM_KBS"));
  }
  if (make_map_blocked(&mh, pm + lm.off_hi, Mm, kt, m_box, kbs) || make_map_blocked(&ml, pm + lm.off_lo, Mm, kt, m_box, kbs) ||
      make_map_blocked(&nh, pn + ln.off_hi, Nn, kt, n_pad, kbs) || make_map_blocked(&nl, pn + ln.off_lo, Nn, kt, n_pad, kbs))
    return 1;
  // enough pieces of the contraction for every SM, at most 32 k-blocks per accumulator (tensor-core adds truncate)
  const int n_kb = K / 64;
  const int want = ceil_div(n_kb, 32);
  int splits = want <= 148 ? 148 : 148 * ceil_div(want, 148);
  if (splits > n_kb) splits = n_kb;
  GemmParams p{};
  p.M = Mm, p.N = Nn, p.K = K;
  p.kblocks_per_split = ceil_div(n_kb, splits);
  p.splits = ceil_div(n_kb, p.kblocks_per_split);
  p.relu = 0;
  p.absmax = reinterpret_cast<const unsigned *>(pm);
  p.absmax_b = reinterpret_cast<const unsigned *>(pn);
  p.atomic_out = 1;
  p.n_mma = n_pad, p.a_boxes = 2, p.b_boxes = 2, p.trans_out = swap ? 1 : 0, p.ldc = N, p.b_rows = n_pad;
  if (deep) {
    p.kbs = kbs;
    p.a_plane = (uint32_t)kbs * m_box * 128u, p.b_plane = (uint32_t)kbs * n_pad * 128u;
    p.stage_bytes = 2u * p.a_plane + 2u * p.b_plane;          // multiples of 1024: the swizzle atoms stay aligned
    const uint32_t room = kGStages * kGStage - (m_box == 128 ? 0u : kOp16);          // (a full-height A plane needs no slack)
    p.n_stages = (int)(room / p.stage_bytes);
    if (p.n_stages > 16) p.n_stages = 16;
  }
  p.bias = nullptr;
  p.C = C;
  dim3 grid(1, 1, p.splits);
  return launch_gemm<false, false, true>(mh, ml, nh, nl, p, grid, st);
}

extern "C" size_t mimrl_gemm_workspace_bytes(int mode, int M, int N, int K) {
  if (mode < 0 || mode > 2 || M <= 0 || N <= 0 || K <= 0) return 0;
  return gemm_layout(mode, M, N, K).total + 256;
}

extern "C" int mimrl_gemm_f32x3(int mode, const float *A, const float *a_mask, const float *B, int M, int N, int K,
                                const float *bias, int relu, float *C, void *workspace, size_t workspace_bytes,
                                void *stream) {
  MIMRL_REQUIRE(mode >= 0 && mode <= 2, "gemm_f32x3: unknown mode %d", mode);
  MIMRL_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f32x3: empty problem %dx%dx%d", M, N, K);
  const GemmLayout g = gemm_layout(mode, M, N, K);
  MIMRL_REQUIRE(workspace_bytes >= g.total, "gemm_f32x3: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  unsigned *absmax = reinterpret_cast<unsigned *>(ws + g.off_absmax);
  __half *a_hi = reinterpret_cast<__half *>(ws + g.off_a_hi), *a_lo = reinterpret_cast<__half *>(ws + g.off_a_lo);
  __half *b_hi = reinterpret_cast<__half *>(ws + g.off_b_hi), *b_lo = reinterpret_cast<__half *>(ws + g.off_b_lo);
  cudaMemsetAsync(absmax, 0, 8, st);
  const size_t na = (size_t)g.a_rows * g.a_cols, nb = (size_t)g.b_rows * g.b_cols;
  int blocks = (int)(((na > nb ? na : nb) + 4095) / 4096);
  blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
  gemm_absmax_kernel<<<dim3(blocks, 2), 256, 0, st>>>(A, a_mask, na, B, nb, absmax);
  if (check_launch("gemm absmax")) return 1;
  auto split = [&](const float *src, const float *mask, int rows, int cols, int ld, int which, __half *hi, __half *lo) {
    const size_t total = (size_t)rows * (ld / 2);
    int b = (int)((total + 255) / 256);
    b = b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b);
    gemm_split_kernel<<<b, 256, 0, st>>>(src, mask, rows, cols, ld, absmax, which, hi, lo, nullptr);
    return check_launch("gemm split");
  };
  if (split(A, a_mask, g.a_rows, g.a_cols, g.lda, 0, a_hi, a_lo)) return 1;
  if (split(B, nullptr, g.b_rows, g.b_cols, g.ldb, 1, b_hi, b_lo)) return 1;
  // tensor maps: K-major operand -> box [128 rows x 64 k]; MN-major operand -> box [64 k-rows x 64 mn]
  const bool a_mn = mode == 2, b_mn = mode != 0;
  CUtensorMap ah, al, bh, bl;
  if (make_map(&ah, a_hi, g.a_cols, g.a_rows, g.lda, a_mn ? 64 : 128)) return 1;
  if (make_map(&al, a_lo, g.a_cols, g.a_rows, g.lda, a_mn ? 64 : 128)) return 1;
  if (make_map(&bh, b_hi, g.b_cols, g.b_rows, g.ldb, b_mn ? 64 : 128)) return 1;
  if (make_map(&bl, b_lo, g.b_cols, g.b_rows, g.ldb, b_mn ? 64 : 128)) return 1;
  GemmParams p{};
  p.M = M, p.N = N, p.K = K;
  p.kblocks_per_split = ceil_div(ceil_div(K, 64), g.splits);
  p.splits = g.splits;
  p.relu = g.splits > 1 ? 0 : relu;          // split-K: bias and ReLU are applied by the reduction
  p.absmax = absmax;
  p.absmax_b = nullptr, p.atomic_out = 0;
  p.n_mma = 128, p.a_boxes = 2, p.b_boxes = 2, p.trans_out = 0, p.ldc = N, p.b_rows = 128;
  p.bias = g.splits > 1 ? nullptr : bias;
  p.C = g.splits > 1 ? reinterpret_cast<float *>(ws + g.off_part) : C;
  dim3 grid(ceil_div(M, 128), ceil_div(N, 128), g.splits);
  int rc;
  if (mode == 0) rc = launch_gemm<false, false>(ah, al, bh, bl, p, grid, st);
  else if (mode == 1) rc = launch_gemm<false, true>(ah, al, bh, bl, p, grid, st);
  else rc = launch_gemm<true, true>(ah, al, bh, bl, p, grid, st);
  if (rc) return rc;
  if (g.splits > 1) {
    const size_t mn = (size_t)M * N;
    int b = (int)((mn + 255) / 256);
    b = b > 148 * 8 ? 148 * 8 : b;
    gemm_reduce_kernel<<<b, 256, 0, st>>>(p.C, g.splits, mn, N, bias, relu, C);
    return check_launch("gemm reduce");
  }
  return 0;
}

