// Concat critic (reference VMI.py:58-65): scores[i, j] = f([x_i, y_j]) for every pair, f the 4-layer MLP
// `mlps(dx + dy, 256, 1, layers = 2)` (VMI.py:13-22), on the tensor cores without the B^2 x 256 pair matrix.
//
// The first layer is linear in the concatenation, so it factorises: W1 [x_i ; y_j] + b1 = u_i + v_j with
// u = x W1x^T + b1 and v = y W1y^T (two B x 256 GEMMs done by the caller).  A CTA scores tiles of 128 pairs
// (4 rows i x 32 columns j; a pair is one TMEM lane):
//
//   h1 = relu(u_i + v_j)                      built by the epilogue warps, fp16 hi/lo, parked in TMEM (A operand)
//   D2 = h1 W2^T   (tcgen05, N = 256)         W2 streamed through a TMA ring (hi and lo halves)
//   h2 = relu(D2 + b2)                        rewritten in place over D2 as the next A operand
//   D3 = h2 W3^T                              accumulates where h1 was
//   s  = w4 . relu(D3 + b3) + b4              thread-local dot product, one float per pair to HBM
//
// Three products (hi.hi + hi.lo + lo.hi) per contraction keep fp32-class accuracy.  Operand scales are powers
// of two from analytic bounds (|h1| <= max|u| + max|v|, |h2| <= |h1|max max_n sum_k |W2[n,k]| + max|b2|), so no
// pass over the data is needed to find them; fp16's exponent range leaves > 2^10 of slack for a loose bound.
#include "tc_common.cuh"

namespace mimrl {
namespace {

constexpr int kFwdWG = 4;                           // epilogue warpgroups of the forward kernel
constexpr int kFwdCh = 8 / kFwdWG;                  // 32-feature chunks per epilogue thread
constexpr int kCcThreads = 64 + 128 * kFwdWG;       // warp 0 TMA, warp 1 MMA, then the epilogue warpgroups
constexpr int kHid = 256;
constexpr uint32_t kCcUnit = 256 * 128;            // ring unit: 64 k of a 256 x 256 weight, hi OR lo half, 32 KB
constexpr int kCcStages = 5;                       // units in flight (hi and lo of a k-block are separate units)
constexpr uint32_t kCcRing = kCcStages * kCcUnit;
constexpr uint32_t kCcVecOff = kCcRing + 256;                       // b2 | b3 | w4 (3 x 256 floats)
constexpr uint32_t kCcPartOff = kCcVecOff + 3 * kHid * 4;           // 128 partial dot products
constexpr uint32_t kCcSmem = kCcPartOff + 4 * 128 * 4 + 1024;
constexpr uint32_t kR0 = 0, kR1 = 256;             // TMEM regions

struct ConcatParams {
  const float *u;        // [n_own, 256]
  const float *vt;       // [256, ldv], column j = y_j W1y^T
  const float *b2, *b3, *w4, *b4;
  float *scores;         // [n_own, n_all]
  const float *scales;   // [0] scale of h1, [1] scale of h2 (powers of two)
  const unsigned *sc_w2, *sc_w3;
  int n_own, n_all, ldv;
  long long n_tiles;
  int n_iq;              // row quads
  // fused bound (no score matrix): row statistics in the forward, gradient weights formed in the backward
  float *part;           // STATS: [split][n_own][3] = {max, sum exp(. - max), softplus sum} over the off-diagonal columns
  int own_offset;        // global row index of local row 0 (the diagonal pair of row i is column own_offset + i)
  int stat_flags;        // MIMRL_STAT_CLAMP | MIMRL_STAT_SOFTPLUS
  int n_jb, q_per, jb_per;   // STATS tiling: column blocks, row quads per CTA (grid.x), column blocks per split (grid.y)
};

// A-operand layout inside a 256-column TMEM region: features [32c, 32c+32) keep their own 32 columns,
// hi halves in the first 16, lo halves in the last 16 -- so an accumulator chunk can be replaced in place.
__device__ __forceinline__ uint32_t a_col(int k16, int lo) { return 32u * (k16 >> 1) + 8u * (k16 & 1) + (lo ? 16u : 0u); }

__device__ __forceinline__ void split32h(const float (&v)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const __half2 h = __floats2half2_rn(v[j], v[j + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
    hi[j >> 1] = *reinterpret_cast<const uint32_t *>(&h);
    lo[j >> 1] = *reinterpret_cast<const uint32_t *>(&l);
  }
}

// STATS = false: scores[i, j] are written (mimrl_concat_scores, the materialising API).
// STATS = true:  nothing B x B leaves the SM.  Tiles run row-quad major (a warp keeps its row i over consecutive column
//                blocks), every lane keeps an online (max, sum exp, softplus sum) of its own columns of the row, and the
//                32 lanes are merged once per row: the off-diagonal row statistics every bound of VMI.py:136-198 needs.
template <bool STATS>
__global__ void __launch_bounds__(kCcThreads, 1)
concat_fwd_kernel(const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                  const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo,
                  const ConcatParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kCcRing;
  // barriers: full[3] | empty[3] | h1 ready | d2 full | h2 ready | d3 full
  const uint32_t bFull = bars, bEmpty = bars + 40, bH1 = bars + 80, bD2 = bars + 88, bH2 = bars + 96, bD3 = bars + 104;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kCcRing + 160);
  float *s_b2 = reinterpret_cast<float *>(gen + kCcVecOff), *s_b3 = s_b2 + kHid, *s_w4 = s_b2 + 2 * kHid;
  float *s_part = reinterpret_cast<float *>(gen + kCcPartOff);
  for (int t = threadIdx.x; t < kHid; t += blockDim.x) {
    s_b2[t] = p.b2 ? p.b2[t] : 0.f;
    s_b3[t] = p.b3 ? p.b3[t] : 0.f;
    s_w4[t] = p.w4[t];
  }

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // scores: contiguous tile range per CTA, tiles ordered column block major so a CTA keeps its v columns hot.
  // STATS: CTA (x, y) owns row quads [q0, q1) x column blocks [jb0, jb1), row-quad major.
  const long long per = (p.n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = STATS ? 0 : per * blockIdx.x;
  const int q0 = STATS ? p.q_per * blockIdx.x : 0, q1 = STATS ? min(p.n_iq, q0 + p.q_per) : 0;
  const int jb0 = STATS ? p.jb_per * blockIdx.y : 0, jb1 = STATS ? min(p.n_jb, jb0 + p.jb_per) : 0;
  const int njl = STATS ? jb1 - jb0 : 1;
  const long long t_end = STATS ? (long long)max(q1 - q0, 0) * max(njl, 0)
                                : ((t_begin + per < p.n_tiles) ? t_begin + per : p.n_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCcStages; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    mbar_init(bH1, 4 * kFwdWG);
    mbar_init(bD2, 1);
    mbar_init(bH2, 4 * kFwdWG);
    mbar_init(bD3, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kCcRing + 160), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    uint32_t n = 0;
    for (long long tile = t_begin; tile < t_end; ++tile) {
      for (int layer = 0; layer < 2; ++layer) {
        for (int un = 0; un < 8; ++un, ++n) {        // unit = (k-block, hi | lo)
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kCcStages, round = n / kCcStages;
          if (round > 0) mbar_wait(bEmpty + 8 * s, (round - 1) & 1);
          if (leader) {
            mbar_expect_tx(bFull + 8 * s, kCcUnit);
            tma_load_2d(base + s * kCcUnit, layer ? (lo ? &map_w3_lo : &map_w3_hi) : (lo ? &map_w2_lo : &map_w2_hi),
                        bFull + 8 * s, kb * 64, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = instr_desc_f16(128, 256);
    uint32_t n = 0, it = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const uint32_t ph = it & 1;
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(layer ? bH2 : bH1, ph);
        tc_fence_after();
        const uint32_t ra = tmem_base + (layer ? kR1 : kR0), rd = tmem_base + (layer ? kR0 : kR1);
        for (int un = 0; un < 8; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kCcStages, round = n / kCcStages;
          mbar_wait(bFull + 8 * s, round & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t b0 = base + s * kCcUnit;
            // the hi unit feeds a_hi.b_hi and a_lo.b_hi, the lo unit a_hi.b_lo
            for (int a_lo = 0; a_lo < (lo ? 1 : 2); ++a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(rd, ra + a_col(kb * 4 + k, a_lo), smem_desc_sw128(b0 + k * 32), idesc, (un | a_lo | k) ? 1u : 0u);
            }
            umma_commit(bEmpty + 8 * s);
          }
          __syncwarp();
        }
        if (leader) umma_commit(layer ? bD3 : bD2);
        __syncwarp();
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int g = e >> 2;                   // feature half handled by this warpgroup
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float s1 = p.scales[0], s2 = p.scales[1];
    const float inv12 = 1.f / (s1 * scale_from_absmax(p.sc_w2[0])), inv23 = 1.f / (s2 * scale_from_absmax(p.sc_w3[0]));
    const float b4 = p.b4 ? p.b4[0] : 0.f;
    uint32_t it = 0;
    float rm = -INFINITY, rs = 0.f, rsp = 0.f;      // STATS: this lane's running statistics of the current row
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const uint32_t ph = it & 1;
      const long long jb = STATS ? jb0 + (tile % njl) : tile / p.n_iq;
      const int iq = STATS ? q0 + (int)(tile / njl) : (int)(tile - jb * p.n_iq);
      const int i = iq * 4 + q, j = (int)jb * 32 + lane;
      const bool ok = i < p.n_own && j < p.n_all;
      const float *ui = p.u + (size_t)(i < p.n_own ? i : 0) * kHid;
      const float *vj = p.vt + (j < p.n_all ? j : 0);
      // ---- h1 = relu(u_i + v_j) -> TMEM region 0
      for (int c = kFwdCh * g; c < kFwdCh * g + kFwdCh; ++c) {
        float v[32];
#pragma unroll
        for (int t4 = 0; t4 < 8; ++t4) {
          const float4 uu = __ldg(reinterpret_cast<const float4 *>(ui + 32 * c) + t4);
          const float uv[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int f = 32 * c + 4 * t4 + w;
            v[4 * t4 + w] = ok ? fmaxf(uv[w] + __ldg(vj + (size_t)f * p.ldv), 0.f) * s1 : 0.f;
          }
        }
        uint32_t hi[16], lo[16];
        split32h(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kR0 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kR0 + 32 * c + 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bH1);
      // ---- h2 = relu(D2 + b2), in place over D2 (region 1)
      mbar_wait(bD2, ph);
      tc_fence_after();
      for (int c = kFwdCh * g; c < kFwdCh * g + kFwdCh; ++c) {
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR1 + 32 * c, d);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = fmaxf(fmaf(__uint_as_float(d[t]), inv12, s_b2[32 * c + t]), 0.f) * s2;
        uint32_t hi[16], lo[16];
        split32h(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c + 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bH2);
      // ---- s = w4 . relu(D3 + b3) + b4
      mbar_wait(bD3, ph);
      tc_fence_after();
      float dot = 0.f;
      for (int c = kFwdCh * g; c < kFwdCh * g + kFwdCh; ++c) {
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR0 + 32 * c, d);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t)
          dot = fmaf(fmaxf(fmaf(__uint_as_float(d[t]), inv23, s_b3[32 * c + t]), 0.f), s_w4[32 * c + t], dot);
      }
      tc_fence_before();
      if (g > 0) s_part[(g - 1) * 128 + q * 32 + lane] = dot;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kFwdWG) : "memory");
      if (g == 0) {
#pragma unroll
        for (int w = 0; w < kFwdWG - 1; ++w) dot += s_part[w * 128 + q * 32 + lane];
        const float sc = dot + b4;
        if (!STATS) {
          if (ok) p.scores[(size_t)i * p.n_all + j] = sc;
        } else {
          if (ok && j != p.own_offset + i) {
            if (p.stat_flags & MIMRL_STAT_SOFTPLUS) rsp += softplusf(sc);
            const float z = (p.stat_flags & MIMRL_STAT_CLAMP) ? fminf(fmaxf(sc, -1.f), 1.f) : sc;
            if (z > rm) {
              rs = rs * __expf(rm - z) + 1.f;        // rm = -inf: rs = 0 * 0 + 1
              rm = z;
            } else {
              rs += __expf(z - rm);
            }
          }
          if ((tile % njl) == njl - 1) {             // last column block of this row in this CTA: merge the 32 lanes
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const float m2 = __shfl_xor_sync(0xffffffffu, rm, off), s2v = __shfl_xor_sync(0xffffffffu, rs, off);
              lse_merge(rm, rs, m2, s2v);
              rsp += __shfl_xor_sync(0xffffffffu, rsp, off);
            }
            if (lane == 0 && i < p.n_own) {
              float *o = p.part + ((size_t)blockIdx.y * p.n_own + i) * 3;
              o[0] = rm, o[1] = rs, o[2] = rsp;
            }
            rm = -INFINITY, rs = 0.f, rsp = 0.f;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- backward ------------------------------------------------------------------------------------------
// Given G = dL/dscores, one pass per tile recomputes the forward (h1, h2, the layer-3 mask) and runs the two
// data-gradient contractions on the tensor cores with the SAME weight tiles read MN-major:
//
//   g3 = G w4 [pre3 > 0]           (A operand, region 0)      g_h2 = g3 W3  -> region 1
//   g2 = g_h2 [h2 > 0]             (in place, region 1)       g_h1 = g2 W2  -> region 0
//   g1 = g_h1 [h1 > 0]             -> g_u[i] (sum over the warp's 32 columns, butterfly) and g_v[j] (sum over rows:
//                                     shared-memory accumulator kept across the CTA's tiles of one column block)
//
// The weight gradients contract over PAIRS (the TMEM lane axis), so they cannot share this kernel's
// accumulators: h1, h2, g2, g3 are written once to HBM as fp16 hi/lo operands (4 KB per pair, feature-major) and
// two split-K GEMMs (gemm_tc.cu, blocked-K operands) finish gW2 = g2 h1^T and gW3 = g3 h2^T.  Bias and w4 gradients are column sums over
// pairs: 32x32 butterfly transposes leave lane t with feature 32c+t, accumulated in registers over the tiles.
constexpr int kBwdWG = 4;                           // epilogue warpgroups of the backward kernel (2 chunks of 32 features each)
constexpr int kBwdCh = 8 / kBwdWG;
constexpr int kBwdThreads = 64 + 128 * kBwdWG;
constexpr uint32_t kCbGvOff = kCcPartOff + 128 * 4;                // g_v accumulator [256 f][32 j] fp32
constexpr uint32_t kCbDotOff = kCbGvOff + kHid * 32 * 4;           // FUSED: partial score dot products [2][kBwdWG][128]
constexpr uint32_t kCbSmem = kCbDotOff + 2 * kBwdWG * 128 * 4 + 1024;

struct ConcatBwdParams {
  ConcatParams f;
  const float *g;            // [n_own, n_all] dL/dscores
  float *g_u;                // [n_own, 256]   (+=)
  float *g_vt;               // [256, ldv]     (+=)
  float *g_b2, *g_b3, *g_w4; // [256] each     (+=)
  __half *op[4][2];          // h1, h2, g2, g3 operands: hi / lo, [256, n_tiles * 128]
  // FUSED (no dL/dscores matrix): g_ij = coef[0] * w(s_ij) off the diagonal, 0 on it, formed from the recomputed score
  const float *coef;         // [1]
  const float *shift;        // [n_own] exp family: w = exp(s_ij - shift[i]); sigmoid family: unused
  const float *b4;           // [1] (nullable) last-layer bias, needed to rebuild s_ij
  int family;                // MIMRL_WEIGHT_EXP | MIMRL_WEIGHT_SIGMOID
  float *g_b4;               // [1] (+=) sum of g_ij
};

// sum over the 32 lanes of v[t] for every t; lane t returns the total of entry t
__device__ __forceinline__ float lane_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool upper = lane & s;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float keep = upper ? v[k + n / 2] : v[k];
      const float send = upper ? v[k] : v[k + n / 2];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Operands are stored feature-major in the blocked-K layout of make_map_blocked: tiles of 64 consecutive pairs, each
// [256 features][64 pairs] contiguous.  For one feature the 32 lanes of a warp (32 consecutive pairs) write 64
// contiguous bytes (one LSU wavefront per store instead of 32 with a pair-major layout), and the weight-gradient GEMM
// reads every [128 features x 64 pairs] box as one contiguous 16 KB run.
__device__ __forceinline__ void store_op32(__half *base_hi, __half *base_lo, size_t row, int f0, const uint32_t (&hi)[16],
                                           const uint32_t (&lo)[16]) {
  const size_t off = ((row >> 6) * kHid + f0) * 64 + (row & 63);
  unsigned short *h = reinterpret_cast<unsigned short *>(base_hi) + off, *l = reinterpret_cast<unsigned short *>(base_lo) + off;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    h[(2 * t) * 64] = (unsigned short)(hi[t] & 0xffffu);
    h[(2 * t + 1) * 64] = (unsigned short)(hi[t] >> 16);
    l[(2 * t) * 64] = (unsigned short)(lo[t] & 0xffffu);
    l[(2 * t + 1) * 64] = (unsigned short)(lo[t] >> 16);
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(kBwdThreads, 1)
concat_bwd_kernel(const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                  const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo,
                  const __grid_constant__ CUtensorMap mn_w2_hi, const __grid_constant__ CUtensorMap mn_w2_lo,
                  const __grid_constant__ CUtensorMap mn_w3_hi, const __grid_constant__ CUtensorMap mn_w3_lo,
                  const ConcatBwdParams bp) {
  const ConcatParams &p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kCcRing;
  // barriers: full[3] | empty[3] | operand ready[4] | accumulator full[4]
  const uint32_t bFull = bars, bEmpty = bars + 40, bReady = bars + 80, bAcc = bars + 112;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kCcRing + 160);
  float *s_b2 = reinterpret_cast<float *>(gen + kCcVecOff), *s_b3 = s_b2 + kHid, *s_w4 = s_b2 + 2 * kHid;
  float *s_gv = reinterpret_cast<float *>(gen + kCbGvOff);
  float *s_dot = reinterpret_cast<float *>(gen + kCbDotOff);
  for (int t = threadIdx.x; t < kHid; t += blockDim.x) {
    s_b2[t] = p.b2 ? p.b2[t] : 0.f;
    s_b3[t] = p.b3 ? p.b3[t] : 0.f;
    s_w4[t] = p.w4[t];
  }
  for (int t = threadIdx.x; t < kHid * 32; t += blockDim.x) s_gv[t] = 0.f;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const long long per = (p.n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = per * blockIdx.x, t_end = (t_begin + per < p.n_tiles) ? t_begin + per : p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCcStages; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bReady + 8 * s, 4 * kBwdWG);
      mbar_init(bAcc + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kCcRing + 160), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // phases: 0: D2 = h1 W2^T (K-major W2)   1: D3 = h2 W3^T (K-major W3)
  //         2: g_h2 = g3 W3 (MN-major W3)  3: g_h1 = g2 W2 (MN-major W2)
  if (warp == 0) {
    const uint32_t leader = elect_one();
    uint32_t n = 0;
    for (long long tile = t_begin; tile < t_end; ++tile) {
      for (int ph = 0; ph < 4; ++ph) {
        for (int un = 0; un < 8; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kCcStages, round = n / kCcStages;
          if (round > 0) mbar_wait(bEmpty + 8 * s, (round - 1) & 1);
          if (leader) {
            const uint32_t dst = base + s * kCcUnit, fb = bFull + 8 * s;
            mbar_expect_tx(fb, kCcUnit);
            if (ph < 2) {
              tma_load_2d(dst, ph ? (lo ? &map_w3_lo : &map_w3_hi) : (lo ? &map_w2_lo : &map_w2_hi), fb, kb * 64, 0);
            } else {          // rows [64 kb, 64 kb + 64) of W, four 64-column blocks
              const CUtensorMap *m = ph == 2 ? (lo ? &mn_w3_lo : &mn_w3_hi) : (lo ? &mn_w2_lo : &mn_w2_hi);
#pragma unroll
              for (int nb = 0; nb < 4; ++nb) tma_load_2d(dst + nb * 8192, m, fb, nb * 64, kb * 64);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc_k = instr_desc_f16(128, 256), idesc_mn = instr_desc_f16_bmn(128, 256);
    uint32_t n = 0, it = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const uint32_t par = it & 1;
      for (int ph = 0; ph < 4; ++ph) {
        mbar_wait(bReady + 8 * ph, par);
        tc_fence_after();
        const uint32_t ra = tmem_base + ((ph & 1) ? kR1 : kR0), rd = tmem_base + ((ph & 1) ? kR0 : kR1);
        for (int un = 0; un < 8; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kCcStages, round = n / kCcStages;
          mbar_wait(bFull + 8 * s, round & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t b0 = base + s * kCcUnit;
            for (int a_lo = 0; a_lo < (lo ? 1 : 2); ++a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t bdesc = ph < 2 ? smem_desc_sw128(b0 + k * 32) : smem_desc_sw128_mn(b0 + k * 2048, 8192, 1024);
                umma_f16_ts(rd, ra + a_col(kb * 4 + k, a_lo), bdesc, ph < 2 ? idesc_k : idesc_mn, (un | a_lo | k) ? 1u : 0u);
              }
            }
            umma_commit(bEmpty + 8 * s);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bAcc + 8 * ph);
        __syncwarp();
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;
    const int g = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float s1 = p.scales[0], s2 = p.scales[1], sg3 = p.scales[4], sg2 = p.scales[5];
    const float sw2 = scale_from_absmax(p.sc_w2[0]), sw3 = scale_from_absmax(p.sc_w3[0]);
    const float inv12 = 1.f / (s1 * sw2), inv23 = 1.f / (s2 * sw3), inv_c = 1.f / (sg3 * sw3), inv_d = 1.f / (sg2 * sw2);
    float acc_b2[kBwdCh] = {}, acc_b3[kBwdCh] = {}, acc_w4[kBwdCh] = {};
    float acc_gb4 = 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(128 * kBwdWG) : "memory");      // s_gv zeroed (the block-wide barrier above already ordered it; cheap)
    long long cur_jb = -1;
    uint32_t it = 0;
    auto flush_gv = [&](long long jb) {
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kBwdWG) : "memory");
      for (int t = (e * 32 + lane); t < kHid * 32; t += 128 * kBwdWG) {
        const int f = t >> 5, jj = t & 31;
        const long long j = jb * 32 + jj;
        const float val = s_gv[t];
        if (j < p.n_all && val != 0.f) atomicAdd(bp.g_vt + (size_t)f * p.ldv + j, val);
        s_gv[t] = 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kBwdWG) : "memory");
    };
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const uint32_t par = it & 1;
      const long long jb = tile / p.n_iq;
      if (jb != cur_jb) {
        if (cur_jb >= 0) flush_gv(cur_jb);
        cur_jb = jb;
      }
      const int iq = (int)(tile - jb * p.n_iq);
      const int i = iq * 4 + q, j = (int)jb * 32 + lane;
      const bool ok = i < p.n_own && j < p.n_all;
      const float *ui = p.u + (size_t)(i < p.n_own ? i : 0) * kHid;
      const float *vj = p.vt + (j < p.n_all ? j : 0);
      float gp = (!FUSED && ok) ? __ldg(bp.g + (size_t)i * p.n_all + j) : 0.f;
      const size_t row = (size_t)(tile - 0) * 128 + q * 32 + lane;        // operand row of this pair
      uint32_t mask1[kBwdCh], mask2[kBwdCh];
      // ---- phase 0 operand: h1
      for (int cc = 0; cc < kBwdCh; ++cc) {
        const int c = kBwdCh * g + cc;
        float v[32];
        uint32_t m = 0;
#pragma unroll
        for (int t4 = 0; t4 < 8; ++t4) {
          const float4 uu = __ldg(reinterpret_cast<const float4 *>(ui + 32 * c) + t4);
          const float uv[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int f = 32 * c + 4 * t4 + w;
            const float z = ok ? uv[w] + __ldg(vj + (size_t)f * p.ldv) : 0.f;
            m |= (z > 0.f ? 1u : 0u) << (4 * t4 + w);
            v[4 * t4 + w] = fmaxf(z, 0.f) * s1;
          }
        }
        mask1[cc] = m;
        uint32_t hi[16], lo[16];
        split32h(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kR0 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kR0 + 32 * c + 16, lo);
        store_op32(bp.op[0][0], bp.op[0][1], row, 32 * c, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 0);
      // ---- phase 1 operand: h2 in place over D2
      mbar_wait(bAcc + 0, par);
      tc_fence_after();
      for (int cc = 0; cc < kBwdCh; ++cc) {
        const int c = kBwdCh * g + cc;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR1 + 32 * c, d);
        tmem_ld_wait();
        float v[32];
        uint32_t m = 0;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float z = fmaf(__uint_as_float(d[t]), inv12, s_b2[32 * c + t]);
          m |= (z > 0.f ? 1u : 0u) << t;
          v[t] = fmaxf(z, 0.f) * s2;
        }
        mask2[cc] = m;
        uint32_t hi[16], lo[16];
        split32h(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c + 16, lo);
        store_op32(bp.op[1][0], bp.op[1][1], row, 32 * c, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 8);
      // ---- phase 2 operand: g3 = G w4 [pre3 > 0] in place over D3; w4 and b3 gradients
      mbar_wait(bAcc + 8, par);
      tc_fence_after();
      if (FUSED) {
        // the pair's gradient weight from its recomputed score: s = w4 . relu(D3 + b3) + b4, every warpgroup sums its
        // own features, the four partial sums meet in shared memory (double-buffered by tile parity)
        float dot = 0.f;
        for (int cc = 0; cc < kBwdCh; ++cc) {
          const int c = kBwdCh * g + cc;
          uint32_t d[32];
          tmem_ld32(tmem_base + lane_off + kR0 + 32 * c, d);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t)
            dot = fmaf(fmaxf(fmaf(__uint_as_float(d[t]), inv23, s_b3[32 * c + t]), 0.f), s_w4[32 * c + t], dot);
        }
        float *sd = s_dot + par * (kBwdWG * 128);
        sd[g * 128 + q * 32 + lane] = dot;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * kBwdWG) : "memory");
        float sc = bp.b4 ? bp.b4[0] : 0.f;
#pragma unroll
        for (int w = 0; w < kBwdWG; ++w) sc += sd[w * 128 + q * 32 + lane];
        gp = 0.f;
        if (ok && j != p.own_offset + i) {
          const float wgt = bp.family == MIMRL_WEIGHT_EXP ? __expf(sc - __ldg(bp.shift + i)) : sigmoidf(sc);
          gp = bp.coef[0] * wgt;
        }
        if (g == 0) acc_gb4 += gp;
      }
      for (int cc = 0; cc < kBwdCh; ++cc) {
        const int c = kBwdCh * g + cc;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR0 + 32 * c, d);
        tmem_ld_wait();
        float gb[32];
        {
          float gh[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float z = fmaf(__uint_as_float(d[t]), inv23, s_b3[32 * c + t]);
            gh[t] = gp * fmaxf(z, 0.f);
            gb[t] = z > 0.f ? gp * s_w4[32 * c + t] : 0.f;
          }
          acc_w4[cc] += lane_transpose_sum(gh, lane);
        }
        {
          float v[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = gb[t] * sg3;
          uint32_t hi[16], lo[16];
          split32h(v, hi, lo);
          tmem_st16(tmem_base + lane_off + kR0 + 32 * c, hi);
          tmem_st16(tmem_base + lane_off + kR0 + 32 * c + 16, lo);
          store_op32(bp.op[3][0], bp.op[3][1], row, 32 * c, hi, lo);
        }
        acc_b3[cc] += lane_transpose_sum(gb, lane);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 16);
      // ---- phase 3 operand: g2 = g_h2 [h2 > 0] in place (region 1); b2 gradient
      mbar_wait(bAcc + 16, par);
      tc_fence_after();
      for (int cc = 0; cc < kBwdCh; ++cc) {
        const int c = kBwdCh * g + cc;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR1 + 32 * c, d);
        tmem_ld_wait();
        float v[32], gb[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float g2 = ((mask2[cc] >> t) & 1u) ? __uint_as_float(d[t]) * inv_c : 0.f;
          gb[t] = g2;
          v[t] = g2 * sg2;
        }
        uint32_t hi[16], lo[16];
        split32h(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kR1 + 32 * c + 16, lo);
        store_op32(bp.op[2][0], bp.op[2][1], row, 32 * c, hi, lo);
        acc_b2[cc] += lane_transpose_sum(gb, lane);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 24);
      // ---- g1 = g_h1 [h1 > 0]: row sums to g_u, column sums to the shared g_v accumulator
      mbar_wait(bAcc + 24, par);
      tc_fence_after();
      for (int cc = 0; cc < kBwdCh; ++cc) {
        const int c = kBwdCh * g + cc;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kR0 + 32 * c, d);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          v[t] = ((mask1[cc] >> t) & 1u) ? __uint_as_float(d[t]) * inv_d : 0.f;
          if (v[t] != 0.f) atomicAdd(s_gv + (32 * c + t) * 32 + lane, v[t]);
        }
        const float su = lane_transpose_sum(v, lane);
        if (i < p.n_own && su != 0.f) atomicAdd(bp.g_u + (size_t)i * kHid + 32 * c + lane, su);
      }
      tc_fence_before();
    }
    if (cur_jb >= 0) flush_gv(cur_jb);
#pragma unroll
    for (int cc = 0; cc < kBwdCh; ++cc) {
      const int f = 32 * (kBwdCh * g + cc) + lane;
      atomicAdd(bp.g_b2 + f, acc_b2[cc]);
      atomicAdd(bp.g_b3 + f, acc_b3[cc]);
      atomicAdd(bp.g_w4 + f, acc_w4[cc]);
    }
    if (FUSED && g == 0) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, off);
      if (lane == 0 && bp.g_b4) atomicAdd(bp.g_b4, acc_gb4);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- operand scales from analytic bounds ------------------------------------------------------------------
__global__ void concat_absmax_kernel(const float *a, size_t na, const float *b, size_t nb, unsigned *out) {
  float m0 = 0.f, m1 = 0.f;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < na; t += (size_t)gridDim.x * blockDim.x)
    m0 = fmaxf(m0, fabsf(a[t]));
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nb; t += (size_t)gridDim.x * blockDim.x)
    m1 = fmaxf(m1, fabsf(b[t]));
  for (int o = 16; o; o >>= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, __float_as_uint(m0));
    atomicMax(out + 1, __float_as_uint(m1));
  }
}

__device__ __forceinline__ float pow2_for(float bound) {      // bound * scale in [2^13, 2^14)
  if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
  int e;
  frexpf(bound, &e);
  int sh = 14 - e;
  sh = sh < -60 ? -60 : (sh > 60 ? 60 : sh);
  return ldexpf(1.f, sh);
}

// one block of 256 threads: analytic bounds -> power-of-two operand scales; also stamps the bounds into the
// headers of the materialised operands so the split-K GEMMs recover the same scales
//   absmax: [0] u  [1] vt  [2] G  [3] w4
//   scales: [0] h1 [1] h2 [2..3] the bounds  [4] g3 [5] g2
__global__ void concat_scales_kernel(const unsigned *absmax, const float *w2, const float *b2, const float *w3,
                                     float *scales, unsigned *hdr_h1, unsigned *hdr_h2, unsigned *hdr_g2, unsigned *hdr_g3,
                                     const float *g_bound = nullptr) {
  __shared__ float red[256];
  const int n = threadIdx.x;
  auto block_max = [&](float v) {
    __syncthreads();
    red[n] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if (n < o) red[n] = fmaxf(red[n], red[n + o]);
      __syncthreads();
    }
    return red[0];
  };
  float row2 = 0.f, col3 = 0.f;
  for (int k = 0; k < kHid; ++k) {
    row2 += fabsf(w2[(size_t)n * kHid + k]);         // sum_k |W2[n, k]|
    col3 += fabsf(w3[(size_t)k * kHid + n]);         // sum_m |W3[m, n]|
  }
  const float row2_max = block_max(row2), col3_max = block_max(col3), b2_max = block_max(b2 ? fabsf(b2[n]) : 0.f);
  if (n == 0) {
    const float m1 = __uint_as_float(absmax[0]) + __uint_as_float(absmax[1]);
    const float m2 = m1 * row2_max + b2_max;
    // fused bound: |g_ij| = |coef| w_ij <= |coef| (the weights are referred to a shift that keeps them <= 1)
    const float mg3 = (g_bound ? fabsf(g_bound[0]) : __uint_as_float(absmax[2])) * __uint_as_float(absmax[3]);
    const float mg2 = mg3 * col3_max;
    scales[0] = pow2_for(m1), scales[1] = pow2_for(m2), scales[2] = m1, scales[3] = m2;
    scales[4] = pow2_for(mg3), scales[5] = pow2_for(mg2);
    if (hdr_h1) *hdr_h1 = __float_as_uint(m1);
    if (hdr_h2) *hdr_h2 = __float_as_uint(m2);
    if (hdr_g2) *hdr_g2 = __float_as_uint(mg2);
    if (hdr_g3) *hdr_g3 = __float_as_uint(mg3);
  }
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_split_bytes(int rows, int cols);
extern "C" int mimrl_split_f32(const float *src, const float *mask, int rows, int cols, void *out, float *colsum,
                               void *stream);

extern "C" int mimrl_concat_tc_supported(int hidden, int layers) { return hidden == kHid && layers == 2; }

extern "C" size_t mimrl_concat_workspace_bytes(int hidden) {
  if (hidden != kHid) return 0;
  return 256 + 2 * mimrl_split_bytes(kHid, kHid);
}

// rows of the materialised weight-gradient operands for an [n_own, n_all] block of pairs (whole 128-pair tiles)
extern "C" long long mimrl_concat_pair_rows(int n_own, int n_all) {
  if (n_own <= 0 || n_all <= 0) return 0;
  return (long long)((n_own + 3) / 4) * ((n_all + 31) / 32) * 128;
}

namespace {
struct ConcatHost {
  unsigned *absmax;
  float *scales;
  unsigned char *s2, *s3;
  CUtensorMap k2h, k2l, k3h, k3l, m2h, m2l, m3h, m3l;
};

// weights -> fp16 hi/lo, tensor maps (K-major boxes of 256 rows; MN-major boxes of 64 rows), absmax of u / vt / G / w4
int concat_prepare(ConcatHost &h, const float *u, const float *vt, int n_own, int ldv, const float *w2, const float *w3,
                   const float *g, size_t n_g, const float *w4, void *workspace, bool mn_maps, cudaStream_t st) {
  unsigned char *ws = (unsigned char *)workspace;
  h.absmax = reinterpret_cast<unsigned *>(ws);
  h.scales = reinterpret_cast<float *>(ws + 64);
  h.s2 = ws + 256, h.s3 = h.s2 + mimrl_split_bytes(kHid, kHid);
  cudaMemsetAsync(h.absmax, 0, 16, st);
  concat_absmax_kernel<<<148, 256, 0, st>>>(u, (size_t)n_own * kHid, vt, (size_t)kHid * ldv, h.absmax);
  if (check_launch("concat absmax")) return 1;
  if (g) {
    concat_absmax_kernel<<<148, 256, 0, st>>>(g, n_g, w4, (size_t)kHid, h.absmax + 2);
    if (check_launch("concat absmax (G)")) return 1;
  }
  if (int rc = mimrl_split_f32(w2, nullptr, kHid, kHid, h.s2, nullptr, (void *)st)) return rc;
  if (int rc = mimrl_split_f32(w3, nullptr, kHid, kHid, h.s3, nullptr, (void *)st)) return rc;
  const size_t off_lo = 256 + align256((size_t)kHid * kHid * 2);
  if (make_map(&h.k2h, h.s2 + 256, kHid, kHid, kHid, 256) || make_map(&h.k2l, h.s2 + off_lo, kHid, kHid, kHid, 256) ||
      make_map(&h.k3h, h.s3 + 256, kHid, kHid, kHid, 256) || make_map(&h.k3l, h.s3 + off_lo, kHid, kHid, kHid, 256))
    return 1;
  if (mn_maps &&
      (make_map(&h.m2h, h.s2 + 256, kHid, kHid, kHid, 64) || make_map(&h.m2l, h.s2 + off_lo, kHid, kHid, kHid, 64) ||
       make_map(&h.m3h, h.s3 + 256, kHid, kHid, kHid, 64) || make_map(&h.m3l, h.s3 + off_lo, kHid, kHid, kHid, 64)))
    return 1;
  return 0;
}

void concat_fill(ConcatParams &p, const ConcatHost &h, const float *u, const float *vt, int n_own, int n_all, int ldv,
                 const float *b2, const float *b3, const float *w4, const float *b4, float *scores) {
  p.u = u, p.vt = vt, p.b2 = b2, p.b3 = b3, p.w4 = w4, p.b4 = b4, p.scores = scores, p.scales = h.scales;
  p.sc_w2 = reinterpret_cast<const unsigned *>(h.s2), p.sc_w3 = reinterpret_cast<const unsigned *>(h.s3);
  p.n_own = n_own, p.n_all = n_all, p.ldv = ldv;
  p.n_iq = (n_own + 3) / 4;
  p.n_tiles = (long long)p.n_iq * ((n_all + 31) / 32);
  p.part = nullptr, p.own_offset = 0, p.stat_flags = 0, p.n_jb = (n_all + 31) / 32, p.q_per = 0, p.jb_per = 0;
}
}  // namespace

// scores[i, j] = w4 . relu(W3 relu(W2 relu(u_i + v_j) + b2) + b3) + b4
extern "C" int mimrl_concat_scores(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden,
                                   const float *w2, const float *b2, const float *w3, const float *b3, const float *w4,
                                   const float *b4, float *scores, void *workspace, size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(hidden == kHid, "concat_scores: hidden width %d not supported (256 only)", hidden);
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && ldv >= n_all && u && vt && w2 && w3 && w4 && scores, "concat_scores: bad arguments");
  MIMRL_REQUIRE(workspace && workspace_bytes >= mimrl_concat_workspace_bytes(hidden), "concat_scores: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  ConcatHost h;
  if (int rc = concat_prepare(h, u, vt, n_own, ldv, w2, w3, nullptr, 0, w4, workspace, false, st)) return rc;
  concat_scales_kernel<<<1, 256, 0, st>>>(h.absmax, w2, b2, w3, h.scales, nullptr, nullptr, nullptr, nullptr);
  if (check_launch("concat scales")) return 1;
  ConcatParams p;
  concat_fill(p, h, u, vt, n_own, n_all, ldv, b2, b3, w4, b4, scores);
  cudaFuncSetAttribute(concat_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCcSmem);
  const int blocks = (int)(p.n_tiles < 148 ? p.n_tiles : 148);
  concat_fwd_kernel<false><<<blocks, kCcThreads, kCcSmem, st>>>(h.k2h, h.k2l, h.k3h, h.k3l, p);
  return check_launch("concat_fwd");
}

// Backward of mimrl_concat_scores for g = dL/dscores [n_own, n_all].  Accumulates (+=) into g_u [n_own, 256],
// g_vt [256, ldv], g_b2, g_b3, g_w4 [256]; writes the four weight-gradient operands (mimrl_split_f32 format of a
// [256, mimrl_concat_pair_rows(n_own, n_all)] matrix): gW2 = op_g2 op_h1^T, gW3 = op_g3 op_h2^T via mimrl_gemm_split mode 0.
extern "C" int mimrl_concat_grad(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden,
                                 const float *w2, const float *b2, const float *w3, const float *b3, const float *w4,
                                 const float *g, float *g_u, float *g_vt, float *g_b2, float *g_b3, float *g_w4,
                                 void *op_h1, void *op_h2, void *op_g2, void *op_g3, void *workspace,
                                 size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(hidden == kHid, "concat_grad: hidden width %d not supported (256 only)", hidden);
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && ldv >= n_all && u && vt && w2 && w3 && w4 && g && g_u && g_vt && g_b2 && g_b3 &&
                    g_w4 && op_h1 && op_h2 && op_g2 && op_g3,
                "concat_grad: bad arguments");
  MIMRL_REQUIRE(workspace && workspace_bytes >= mimrl_concat_workspace_bytes(hidden), "concat_grad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  ConcatHost h;
  if (int rc = concat_prepare(h, u, vt, n_own, ldv, w2, w3, g, (size_t)n_own * n_all, w4, workspace, true, st)) return rc;
  concat_scales_kernel<<<1, 256, 0, st>>>(h.absmax, w2, b2, w3, h.scales, (unsigned *)op_h1, (unsigned *)op_h2,
                                          (unsigned *)op_g2, (unsigned *)op_g3);
  if (check_launch("concat scales")) return 1;
  ConcatBwdParams bp;
  concat_fill(bp.f, h, u, vt, n_own, n_all, ldv, b2, b3, w4, nullptr, nullptr);
  bp.g = g, bp.g_u = g_u, bp.g_vt = g_vt, bp.g_b2 = g_b2, bp.g_b3 = g_b3, bp.g_w4 = g_w4;
  const size_t rows = (size_t)bp.f.n_tiles * 128;
  const size_t off_lo = 256 + align256((size_t)kHid * rows * 2);
  void *ops[4] = {op_h1, op_h2, op_g2, op_g3};
  for (int t = 0; t < 4; ++t) {
    bp.op[t][0] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256);
    bp.op[t][1] = reinterpret_cast<__half *>((unsigned char *)ops[t] + off_lo);
  }
  bp.coef = bp.shift = bp.b4 = nullptr, bp.family = 0, bp.g_b4 = nullptr;
  cudaFuncSetAttribute(concat_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCbSmem);
  const int blocks = (int)(bp.f.n_tiles < 148 ? bp.f.n_tiles : 148);
  concat_bwd_kernel<false><<<blocks, kBwdThreads, kCbSmem, st>>>(h.k2h, h.k2l, h.k3h, h.k3l, h.m2h, h.m2l, h.m3h, h.m3l, bp);
  return check_launch("concat_bwd");
}

// ---- fused bound: no score matrix, no dL/dscores matrix --------------------------------------------------------
namespace {
struct StatsGrid {
  int gx, gy, q_per, jb_per;
};
StatsGrid concat_stats_grid(int n_own, int n_all) {
  const int n_iq = (n_own + 3) / 4, n_jb = (n_all + 31) / 32;
  StatsGrid g;
  g.q_per = (n_iq + 147) / 148;
  g.gx = (n_iq + g.q_per - 1) / g.q_per;
  int gy = g.gx >= 148 ? 1 : (148 + g.gx - 1) / g.gx;
  gy = gy > n_jb ? n_jb : (gy > 32 ? 32 : gy);
  g.jb_per = (n_jb + gy - 1) / gy;
  g.gy = (n_jb + g.jb_per - 1) / g.jb_per;
  return g;
}
}  // namespace

extern "C" size_t mimrl_concat_stats_workspace_bytes(int hidden, int n_own, int n_all) {
  if (hidden != kHid || n_own <= 0 || n_all <= 0) return 0;
  const StatsGrid g = concat_stats_grid(n_own, n_all);
  return mimrl_concat_workspace_bytes(hidden) + 256 + (size_t)g.gy * n_own * 3 * sizeof(float);
}

// Off-diagonal row statistics of the all-pairs scores, the scores never leave the SM:
// row_max[i] = max_{j != o+i} t(s_ij), row_sum[i] = sum exp(t(s_ij) - row_max[i]), row_sp[i] = sum softplus(s_ij)
extern "C" int mimrl_concat_row_stats(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden,
                                      int own_offset, int flags, const float *w2, const float *b2, const float *w3,
                                      const float *b3, const float *w4, const float *b4, float *row_max, float *row_sum,
                                      float *row_sp, void *workspace, size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(hidden == kHid, "concat_row_stats: hidden width %d not supported (256 only)", hidden);
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && ldv >= n_all && u && vt && w2 && w3 && w4 && row_max && row_sum,
                "concat_row_stats: bad arguments");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "concat_row_stats: row block outside the batch");
  MIMRL_REQUIRE(workspace && workspace_bytes >= mimrl_concat_stats_workspace_bytes(hidden, n_own, n_all),
                "concat_row_stats: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  ConcatHost h;
  if (int rc = concat_prepare(h, u, vt, n_own, ldv, w2, w3, nullptr, 0, w4, workspace, false, st)) return rc;
  concat_scales_kernel<<<1, 256, 0, st>>>(h.absmax, w2, b2, w3, h.scales, nullptr, nullptr, nullptr, nullptr);
  if (check_launch("concat scales")) return 1;
  ConcatParams p;
  concat_fill(p, h, u, vt, n_own, n_all, ldv, b2, b3, w4, b4, nullptr);
  const StatsGrid g = concat_stats_grid(n_own, n_all);
  p.part = reinterpret_cast<float *>((unsigned char *)workspace + align256(mimrl_concat_workspace_bytes(hidden)));
  p.own_offset = own_offset, p.stat_flags = flags, p.n_jb = (n_all + 31) / 32, p.q_per = g.q_per, p.jb_per = g.jb_per;
  cudaFuncSetAttribute(concat_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCcSmem);
  concat_fwd_kernel<true><<<dim3(g.gx, g.gy), kCcThreads, kCcSmem, st>>>(h.k2h, h.k2l, h.k3h, h.k3l, p);
  if (check_launch("concat_fwd(stats)")) return 1;
  return combine_row_stats(p.part, g.gy, n_own, row_max, row_sum, row_sp, st);
}

// Backward of the fused bound: g_ij = coef[0] * w(s_ij) for j != own_offset + i (exp family: w = exp(s_ij - shift[i]);
// sigmoid family: w = sigmoid(s_ij)), 0 on the diagonal, formed inside the kernel from the recomputed score.
// Outputs as mimrl_concat_grad, plus g_b4[0] += sum_ij g_ij.
extern "C" int mimrl_concat_grad_fused(const float *u, const float *vt, int n_own, int n_all, int ldv, int hidden,
                                       int own_offset, const float *w2, const float *b2, const float *w3,
                                       const float *b3, const float *w4, const float *b4, int family, const float *coef,
                                       const float *shift, float *g_u, float *g_vt, float *g_b2, float *g_b3,
                                       float *g_w4, float *g_b4, void *op_h1, void *op_h2, void *op_g2, void *op_g3,
                                       void *workspace, size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(hidden == kHid, "concat_grad_fused: hidden width %d not supported (256 only)", hidden);
  MIMRL_REQUIRE(n_own > 0 && n_all > 0 && ldv >= n_all && u && vt && w2 && w3 && w4 && coef && g_u && g_vt && g_b2 && g_b3 &&
                    g_w4 && op_h1 && op_h2 && op_g2 && op_g3,
                "concat_grad_fused: bad arguments");
  MIMRL_REQUIRE(family == MIMRL_WEIGHT_EXP || family == MIMRL_WEIGHT_SIGMOID, "concat_grad_fused: unknown weight family %d", family);
  MIMRL_REQUIRE(family != MIMRL_WEIGHT_EXP || shift, "concat_grad_fused: exp family needs a shift vector");
  MIMRL_REQUIRE(own_offset >= 0 && own_offset + n_own <= n_all, "concat_grad_fused: row block outside the batch");
  MIMRL_REQUIRE(workspace && workspace_bytes >= mimrl_concat_workspace_bytes(hidden), "concat_grad_fused: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  ConcatHost h;
  // absmax[2..3]: only max |w4| is needed here (the bound of |G| is |coef|)
  if (int rc = concat_prepare(h, u, vt, n_own, ldv, w2, w3, w4, (size_t)kHid, w4, workspace, true, st)) return rc;
  concat_scales_kernel<<<1, 256, 0, st>>>(h.absmax, w2, b2, w3, h.scales, (unsigned *)op_h1, (unsigned *)op_h2,
                                          (unsigned *)op_g2, (unsigned *)op_g3, coef);
  if (check_launch("concat scales")) return 1;
  ConcatBwdParams bp;
  concat_fill(bp.f, h, u, vt, n_own, n_all, ldv, b2, b3, w4, nullptr, nullptr);
  bp.f.own_offset = own_offset;
  bp.g = nullptr, bp.g_u = g_u, bp.g_vt = g_vt, bp.g_b2 = g_b2, bp.g_b3 = g_b3, bp.g_w4 = g_w4;
  bp.coef = coef, bp.shift = shift, bp.b4 = b4, bp.family = family, bp.g_b4 = g_b4;
  const size_t rows = (size_t)bp.f.n_tiles * 128;
  const size_t off_lo = 256 + align256((size_t)kHid * rows * 2);
  void *ops[4] = {op_h1, op_h2, op_g2, op_g3};
  for (int t = 0; t < 4; ++t) {
    bp.op[t][0] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256);
    bp.op[t][1] = reinterpret_cast<__half *>((unsigned char *)ops[t] + off_lo);
  }
  cudaFuncSetAttribute(concat_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCbSmem);
  const int blocks = (int)(bp.f.n_tiles < 148 ? bp.f.n_tiles : 148);
  concat_bwd_kernel<true><<<blocks, kBwdThreads, kCbSmem, st>>>(h.k2h, h.k2l, h.k3h, h.k3l, h.m2h, h.m2l, h.m3h, h.m3l, bp);
  return check_launch("concat_bwd(fused)");
}
