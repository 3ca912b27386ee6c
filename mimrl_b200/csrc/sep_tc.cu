// Separable-critic score sweeps on the 5th-generation tensor cores (sm_100a).
//
//   S = own . all^T            tcgen05.mma kind::f16, operands staged by TMA
//                              (128B swizzle), accumulators in TMEM
//   fp32-class accuracy:       every fp32 operand is pre-scaled by a power of two
//                              and split v = hi + lo (two fp16 values); each
//                              contraction runs hi.hi + hi.lo + lo.hi with fp32
//                              accumulation (dropped lo.lo term ~2^-22 relative).
//                              fp16 runs at twice the tf32 rate, so this costs
//                              the same tensor time as 1.5 tf32 passes while
//                              meeting the 1e-4 parity bound.
//   operand residency:         the owned 128-row block is parked in TMEM once per CTA
//                              (A operand of every score MMA); only the swept
//                              operand streams through a 6-deep ring of 32 KB
//                              shared-memory units, so shared-memory bandwidth is
//                              spent on one operand and TMA runs tiles ahead.
//   row_stats kernel:          warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 =
//                              two epilogue warpgroups (thread = score row) that
//                              ping-pong over double-buffered TMEM accumulators:
//                              online (max, sum exp) straight out of TMEM.
//   weighted_sum kernel:       per 128x64 score tile the epilogue turns S into the
//                              weight tile W (exp / sigmoid family) and writes W as
//                              an fp16 hi/lo pair back INTO the TMEM columns S came
//                              from; a second MMA (A = W from TMEM, B = the same
//                              shared-memory tile read MN-major) accumulates
//                              O += W . all into a resident 128x128 TMEM accumulator
//                              (flash-attention shaped, operands loaded once).
// The B x B matrix only ever exists as 128x128 / 128x64 tiles in TMEM.
#include "tc_common.cuh"

namespace mimrl {
namespace {

// --------------------------------------------------------------- prepass ----
// absmax[which] = max |v| over the tensor (float bits compare as unsigned for v >= 0); float4 grid-stride loads
__global__ void __launch_bounds__(256)
absmax2_kernel(const float *__restrict__ a, size_t na, const float *__restrict__ b, size_t nb, unsigned *__restrict__ out) {
  const float *src = blockIdx.y == 0 ? a : b;
  const size_t n = blockIdx.y == 0 ? na : nb;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (((size_t)src & 15) == 0) {
    const size_t n4 = n >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    for (size_t i = tid; i < n4; i += stride) {
      const float4 v = __ldg(s4 + i);
      m0 = fmaxf(m0, fabsf(v.x)), m1 = fmaxf(m1, fabsf(v.y)), m2 = fmaxf(m2, fabsf(v.z)), m3 = fmaxf(m3, fabsf(v.w));
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) m0 = fmaxf(m0, fabsf(src[i]));
  } else {
    for (size_t i = tid; i < n; i += stride) m0 = fmaxf(m0, fabsf(src[i]));
  }
  float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out + blockIdx.y, __float_as_uint(m));
}

// [n, embed] fp32 -> hi, lo [n, 128] fp16 (K padded with zeros), scaled by 2^k
__global__ void split_rows_kernel(const float *__restrict__ a, int na, const float *__restrict__ b, int nb, int embed,
                                  const unsigned *__restrict__ absmax, __half *__restrict__ a_hi,
                                  __half *__restrict__ a_lo, __half *__restrict__ b_hi, __half *__restrict__ b_lo) {
  const int which = blockIdx.y;
  const float *src = which == 0 ? a : b;
  const int n = which == 0 ? na : nb;
  __half *hi = which == 0 ? a_hi : b_hi, *lo = which == 0 ? a_lo : b_lo;
  const float sc = scale_from_absmax(absmax[which]);
  const size_t total = (size_t)n * 64;  // two elements per thread
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx >> 6;
    const int e = (int)(idx & 63) * 2;
    const float v0 = e < embed ? src[r * embed + e] * sc : 0.f;
    const float v1 = e + 1 < embed ? src[r * embed + e + 1] * sc : 0.f;
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__half2 *>(hi + r * 128 + e) = h;
    *reinterpret_cast<__half2 *>(lo + r * 128 + e) = l;
  }
}


// ------------------------------------------------------------ common bits ----
constexpr int kTcThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2-9 = two epilogue warpgroups
constexpr int kStages = 6;               // ring depth
constexpr uint32_t kUnit = 32768;        // ring unit: 32 KB of the swept operand
constexpr uint32_t kTile16 = 128 * 128;  // 128 rows x 128 B
constexpr uint32_t kXTile = 64 * 128;    // 64 rows x 128 B
constexpr uint32_t kAux = 1024;          // barriers, tmem slot, staged shifts
constexpr uint32_t kRingSmem = kStages * kUnit + kAux + 1024;
constexpr uint32_t kWsumSmem = kRingSmem + 2048;   // + the online mode's reference / row-sum exchange ([2][2][128] floats)
// TMEM columns: [0,128) owned block (hi kb0, hi kb1, lo kb0, lo kb1: 32 columns each, one column = two K values)
constexpr uint32_t kTmemA = 0;

// Park the owned 128-row block (fp16 hi/lo, K = 128) in TMEM: lane = row, column c of a 32-column part holds
// K elements (2c, 2c+1).  Executed by epilogue warpgroup 0 (warps 2-5).
__device__ __forceinline__ void park_own_block(uint32_t tmem_a, const __half *own_hi, const __half *own_lo, int row0,
                                               int n_own, int warp, int lane) {
  const int r = (warp & 3) * 32 + lane;
  const bool ok = row0 + r < n_own;
#pragma unroll 1
  for (int part = 0; part < 4; ++part) {
    const __half *src = (part < 2 ? own_hi : own_lo) + (size_t)(row0 + r) * 128 + (part & 1) * 64;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 t = ok ? __ldg(reinterpret_cast<const uint4 *>(src) + j) : make_uint4(0, 0, 0, 0);
      v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
    }
    tmem_st32(tmem_a + ((uint32_t)((warp & 3) * 32) << 16) + part * 32, v);
  }
  tmem_st_wait();
}

// -------------------------------------------------------------- row stats ----
// TMEM: [0,128) owned block, [128,256) accumulator 0, [256,384) accumulator 1.
// Ring unit u = 2*tile + h: the hi (h=0) or lo (h=1) half of a 128-column tile, [kb0 | kb1] x (128 rows x 128 B).
struct StatsParams {
  int n_own, n_all, own_offset, tiles_per_split;
  const unsigned *absmax;
  const __half *own_hi, *own_lo;
  float *part;
  const float *row_lse, *row_sigma;   // kStatInterp: per owned row, LSE_i of the whole row and sigma_i (see below)
};

// FLAGS == kStatInterp: the second forward sweep of the interpolated bound (VMI.py:201-250).  With p_ij =
// exp(S_ij - LSE_i) and sigma_i from the first sweep, per owned row over the off-diagonal columns:
//   part[.][1] = Q_i = sum_j p_ij / (1 - sigma_i p_ij),   part[.][2] = T_i = sum_j log(1 - sigma_i p_ij)
// (part[.][0] = 0, so that the (max, sum) merge of combine_row_stats degenerates to a plain sum).
constexpr int kStatInterp = 8;

// log(1 - x) for x in [0, 1): series below 2^-6 (the common case, p ~ 1/n), libm above
__device__ __forceinline__ float log1m(float x) {
  if (x < 0.015625f) return -x * fmaf(x, fmaf(x, fmaf(x, 0.25f, 0.33333334f), 0.5f), 1.f);
  return log1pf(-x);
}

template <int FLAGS>
__global__ void __launch_bounds__(kTcThreads, 1)
sep_stats_tc_kernel(const __grid_constant__ CUtensorMap map_all_hi, const __grid_constant__ CUtensorMap map_all_lo,
                    const StatsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kStages * kUnit;
  const uint32_t bFull = bars, bEmpty = bars + 64, bTFull = bars + 128, bTEmpty = bars + 144;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kStages * kUnit + 256);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;
  const int split = blockIdx.y;
  const int n_tiles = (p.n_all + 127) / 128;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(n_tiles, t0 + p.tiles_per_split);
  const int T = t1 - t0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bFull + 8 * i, 1);
      mbar_init(bEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 4);          // one elected arrive per warp of the owning epilogue warpgroup
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kStages * kUnit + 256), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tmem_a = tmem_base + kTmemA;
  if (warp >= 2 && warp < 6) park_own_block(tmem_a, p.own_hi, p.own_lo, row0, p.n_own, warp, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // FLAGS & 4 (MIMRL_STAT_MAXONLY): approximate row maxima only -- one product (own_hi . x_hi, 11-bit operands) instead
  // of three and no exponentials: the reference point of the fused forward sweep, which does not need to be exact.
  constexpr bool kMaxOnly = (FLAGS & 4) != 0;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      prefetch_tmap(&map_all_hi);
      prefetch_tmap(&map_all_lo);
      for (int u = 0; u < (kMaxOnly ? T : 2 * T); ++u) {
        const int stage = u % kStages;
        const int col = (t0 + (kMaxOnly ? u : (u >> 1))) * 128;
        mbar_wait(bEmpty + 8 * stage, ((u / kStages) & 1) ^ 1);
        mbar_expect_tx(bFull + 8 * stage, kUnit);
        const CUtensorMap *m = (!kMaxOnly && (u & 1)) ? &map_all_lo : &map_all_hi;
        tma_load_2d(base + stage * kUnit, m, bFull + 8 * stage, 0, col);
        tma_load_2d(base + stage * kUnit + kTile16, m, bFull + 8 * stage, 64, col);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = instr_desc_f16(128, 128);
    for (int i = 0; i < T; ++i) {
      const int buf = i & 1;
      const uint32_t d = tmem_base + 128 + buf * 128;
      mbar_wait(bTEmpty + 8 * buf, ((i >> 1) & 1) ^ 1);
      if (kMaxOnly) {
        const int stage = i % kStages;
        mbar_wait(bFull + 8 * stage, (i / kStages) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t b0 = base + stage * kUnit;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(d, tmem_a + kb * 32 + k * 8, smem_desc_sw128(b0 + kb * kTile16 + k * 32), idesc, (kb | k) ? 1u : 0u);
          umma_commit(bEmpty + 8 * stage);
          umma_commit(bTFull + 8 * buf);
        }
        __syncwarp();
        continue;
      }
      // hi half of the swept tile: own_hi . x_hi and own_lo . x_hi
      int u = 2 * i, stage = u % kStages;
      mbar_wait(bFull + 8 * stage, (u / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t b0 = base + stage * kUnit;
#pragma unroll
        for (int a_sel = 0; a_sel < 4; a_sel += 2)          // own hi (parts 0,1) then own lo (parts 2,3)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(d, tmem_a + (a_sel + kb) * 32 + k * 8, smem_desc_sw128(b0 + kb * kTile16 + k * 32), idesc,
                          (a_sel | kb | k) ? 1u : 0u);
        umma_commit(bEmpty + 8 * stage);
      }
      __syncwarp();
      // lo half: own_hi . x_lo
      u = 2 * i + 1, stage = u % kStages;
      mbar_wait(bFull + 8 * stage, (u / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t b0 = base + stage * kUnit;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(d, tmem_a + kb * 32 + k * 8, smem_desc_sw128(b0 + kb * kTile16 + k * 32), idesc, 1u);
        umma_commit(bEmpty + 8 * stage);
        umma_commit(bTFull + 8 * buf);
      }
      __syncwarp();
    }
  } else {
    // Two epilogue warpgroups ping-pong over the tiles: warpgroup g owns accumulator g and the tiles with
    // (i & 1) == g, so every SM sub-partition has two epilogue warps in flight to hide TMEM/MUFU latency.
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int gr = p.own_offset + row0 + r;
    const float inv = 1.f / (scale_from_absmax(p.absmax[0]) * scale_from_absmax(p.absmax[1]));
    const float c2 = inv * kLog2e;
    // running max in RAW accumulator units: (acc - max) is formed first, then scaled to log2 units, so the
    // dominant terms keep full fp32 accuracy; inv is a power of two, so max * inv is exact
    float m2 = -INFINITY, s = 0.f, sp = 0.f;
    const int grmin = p.own_offset + row0, grmax = grmin + 127;
    const int buf = wg;
    constexpr bool kInterp = FLAGS == kStatInterp;
    float i_lse = 0.f, i_sig = 0.f;
    if (kInterp) {
      m2 = 0.f;
      if (row0 + r < p.n_own) i_lse = p.row_lse[row0 + r], i_sig = p.row_sigma[row0 + r];
    }
    for (int i = wg; i < T; i += 2) {
      const int col0 = (t0 + i) * 128;
      const bool clean = (FLAGS & 3) == 0 && col0 + 128 <= p.n_all && (col0 + 128 <= grmin || col0 > grmax);
      mbar_wait(bTFull + 8 * buf, (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + buf * 128 + ch * 32, v);
        tmem_ld_wait();
        if (kInterp) {
          float qa[2] = {0.f, 0.f}, ta[2] = {0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gc = col0 + ch * 32 + j;
            const bool ok = clean || (gc < p.n_all && gc != gr);
            // exp(S - LSE): the difference is formed at score magnitude, then scaled to log2 units
            const float pij = ok ? ex2(fmaf(__uint_as_float(v[j]), inv, -i_lse) * kLog2e) : 0.f;
            const float x = fminf(i_sig * pij, 0.99999994f);
            qa[j & 1] += __fdividef(pij, 1.f - x);
            ta[j & 1] += log1m(x);
          }
          s += qa[0] + qa[1];
          sp += ta[0] + ta[1];
          continue;
        }
        if (!clean) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gc = col0 + ch * 32 + j;
            const bool ok = gc < p.n_all && gc != gr;
            const float z = __uint_as_float(v[j]) * inv;
            if ((FLAGS & MIMRL_STAT_SOFTPLUS) && ok)
              sp += fmaxf(z, 0.f) + kLn2 * lg2(1.f + ex2(-fabsf(z) * kLog2e));
            // clamp acts on S in natural units; the value goes back to raw units (t(S) / inv)
            const float val = (FLAGS & MIMRL_STAT_CLAMP) ? fminf(fmaxf(z, -1.f), 1.f) * (1.f / inv) : __uint_as_float(v[j]);
            v[j] = __float_as_uint(ok ? val : -INFINITY);
          }
        }
        float mx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = __uint_as_float(v[u]);
#pragma unroll
        for (int j = 4; j < 32; j += 4)
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(v[j + u]));
        const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (kMaxOnly) {
          m2 = fmaxf(m2, cmax);
          continue;
        }
        if (cmax > m2) {                      // also false when the whole chunk is masked (-inf)
          s *= ex2((m2 - cmax) * c2);
          m2 = cmax;
        }
        if (m2 > -INFINITY) {
          float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; j += 4)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += ex2((__uint_as_float(v[j + u]) - m2) * c2);
          s += (a[0] + a[1]) + (a[2] + a[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTEmpty + 8 * buf);
    }
    if (row0 + r < p.n_own) {
      float *o = p.part + ((size_t)(split * 2 + wg) * p.n_own + row0 + r) * 3;
      o[0] = m2 * inv;
      o[1] = s;
      o[2] = sp;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ----------------------------------------------------------- weighted sum ----
// TMEM: [0,128) owned block, [128,192) S/W buffer 0, [192,256) S/W buffer 1, [256,384) output accumulator
// (ONLINE: [256,384) accumulator of epilogue warpgroup 0's columns, [384,512) of warpgroup 1's).
// Ring unit = one 64-row tile of the swept operand: [hi kb0 | hi kb1 | lo kb0 | lo kb1] x (64 rows x 128 B).
// It is read K-major by the score MMAs (rows = N) and MN-major by the W.X MMAs (rows = K).
//
// ONLINE (the forward of the exp-family bounds, mimrl_sep_online_forward): no reference point is given.  Every
// epilogue thread keeps a running reference ref for its row (online softmax): weights are exp(S - ref), and when a
// tile's maximum exceeds ref by more than kOnlineTau the accumulated sums are rescaled by exp(ref - ref') -- lazily,
// so that after the first tiles of a sweep almost no tile pays for it.  The two epilogue warpgroups own disjoint
// column halves of every score tile; giving each its OWN output accumulator (k-steps 0,1 -> accumulator 0, k-steps
// 2,3 -> accumulator 1) makes their references independent, so no cross-warpgroup exchange happens per tile; the two
// are merged once at the end.  A rescale of accumulator g must not overlap an MMA that adds into it: the epilogue
// warp waits for the W.X MMAs of the previous tile (bODone) before it touches the accumulator, and the MMAs of the
// current tile cannot start before this warp has delivered its weights (bWFull).
constexpr float kOnlineWExp = 9.f;       // online weights are carried as w * 2^9: w <= 2^kOnlineTau stays below fp16 max
constexpr float kOnlineTau = 4.5f;       // rescale when a score exceeds the reference by more than this (natural units)

// One thread's 32 scores of a tile -> weights (fp16 hi/lo pairs) + the row-sum statistic.  EDGE = the tile holds the
// row's diagonal or runs past the batch: invalid entries get weight 0 and the statistic leaves the diagonal out.  It
// is a separate instantiation so that the common tile carries none of the per-element index tests.
// MIMRL_WEIGHT_INTERP (backward of the interpolated bound): w = p (a1 + a2 / (1 - sg p)), p = exp(S - lse), with the
// four parameters of the score's ROW (lse, a1, a2, sg), scaled by the caller so that |w| <= 1.  They belong to the
// owned row (scalars) or, when the operands are swapped, to the swept row (staged in shared memory as [4][64]).
struct InterpPar {
  float lse, a1, a2, sg;
  const float *sm;          // by_swept: this thread's 32 columns of the staged [4][64] block
};

template <int FAMILY, bool ONLINE, bool EDGE>
__device__ __forceinline__ void weights_of_tile(const uint32_t (&v)[32], uint32_t (&hi)[16], uint32_t (&lo)[16],
                                                float2 (&rs2)[2], float inv, float c2, float wexp, float shift_row,
                                                const float4 *sh4, bool by_swept, bool want_rsum, int cbase, int gr,
                                                int n_all, bool include_diag, const InterpPar &ip) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float shv[4] = {shift_row, shift_row, shift_row, shift_row};
    if (!ONLINE && by_swept) {
      const float4 t4 = sh4[j >> 2];
      shv[0] = t4.x, shv[1] = t4.y, shv[2] = t4.z, shv[3] = t4.w;
    }
    float w[4];
    if (FAMILY == MIMRL_WEIGHT_EXP) {
      // two packed FFMA2 per pair: (acc * inv - shift) at score magnitude, then * log2e + wexp
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const float2 a2 = make_float2(__uint_as_float(v[j + u]), __uint_as_float(v[j + u + 1]));
        const float2 t2 = ffma2(a2, make_float2(inv, inv), make_float2(-shv[u], -shv[u + 1]));
        const float2 e2 = ffma2(t2, make_float2(kLog2e, kLog2e), make_float2(wexp, wexp));
        // ONLINE: every entry is <= ref + tau by construction, no cap needed
        w[u] = ex2(ONLINE ? e2.x : fminf(e2.x, 15.9f));
        w[u + 1] = ex2(ONLINE ? e2.y : fminf(e2.y, 15.9f));
      }
    } else if (FAMILY == MIMRL_WEIGHT_INTERP) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float lse = by_swept ? ip.sm[j + u] : ip.lse, a1 = by_swept ? ip.sm[64 + j + u] : ip.a1;
        const float a2 = by_swept ? ip.sm[128 + j + u] : ip.a2, sg = by_swept ? ip.sm[192 + j + u] : ip.sg;
        const float pij = ex2(fmaf(__uint_as_float(v[j + u]), inv, -lse) * kLog2e);
        w[u] = 16384.f * pij * fmaf(a2, __fdividef(1.f, 1.f - fminf(sg * pij, 0.99999994f)), a1);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = __fdividef(16384.f, 1.f + ex2(-__uint_as_float(v[j + u]) * c2));
    }
    float a[4] = {w[0], w[1], w[2], w[3]};
    if (EDGE) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gc = cbase + j + u;
        if (!(gc < n_all && (include_diag || gc != gr))) w[u] = 0.f;
        if (gc == gr || gc >= n_all) a[u] = 0.f;           // the statistic always leaves the diagonal out
      }
    }
    if (want_rsum) {
      rs2[0] = fadd2(rs2[0], make_float2(a[0], a[1]));
      rs2[1] = fadd2(rs2[1], make_float2(a[2], a[3]));
    }
    // hi = w truncated to 11 significant bits (exactly representable in fp16), lo = w - hi
    const float2 wh01 = make_float2(__uint_as_float(__float_as_uint(w[0]) & 0xFFFFE000u),
                                    __uint_as_float(__float_as_uint(w[1]) & 0xFFFFE000u));
    const float2 wh23 = make_float2(__uint_as_float(__float_as_uint(w[2]) & 0xFFFFE000u),
                                    __uint_as_float(__float_as_uint(w[3]) & 0xFFFFE000u));
    const float2 wl01 = fsub2(make_float2(w[0], w[1]), wh01), wl23 = fsub2(make_float2(w[2], w[3]), wh23);
    const __half2 h0 = __floats2half2_rn(wh01.x, wh01.y), h1 = __floats2half2_rn(wh23.x, wh23.y);
    const __half2 l0 = __floats2half2_rn(wl01.x, wl01.y), l1 = __floats2half2_rn(wl23.x, wl23.y);
    hi[j >> 1] = *reinterpret_cast<const uint32_t *>(&h0);
    hi[(j >> 1) + 1] = *reinterpret_cast<const uint32_t *>(&h1);
    lo[j >> 1] = *reinterpret_cast<const uint32_t *>(&l0);
    lo[(j >> 1) + 1] = *reinterpret_cast<const uint32_t *>(&l1);
  }
}

struct WsumParams {
  int n_own, n_all, own_offset, tiles_per_split, include_diag, shift_by_swept;
  const unsigned *absmax;
  const __half *own_hi, *own_lo;
  const float *shift;
  float *part;  // [split][n_own][128]
  float *rsum_part;   // nullable: [split * 2 + warpgroup][n_own] sum of the off-diagonal weights of the row
                      // (ONLINE: [split][n_own], referred to ref_part)
  float *ref_part;    // ONLINE: [split][n_own] final reference point of the row in this split (-inf: nothing seen)
  int n_par;          // MIMRL_WEIGHT_INTERP: shift holds four vectors [4][n_par] (lse, a1, a2, sg of the score's row)
};

template <int FAMILY, bool ONLINE>
__global__ void __launch_bounds__(kTcThreads, 1)
sep_wsum_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                   const WsumParams p) {
  static_assert(!ONLINE || FAMILY == MIMRL_WEIGHT_EXP, "online mode is for the exp family");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kStages * kUnit;
  const uint32_t bXFull = bars, bXEmpty = bars + 64, bSFull = bars + 128, bWFull = bars + 144, bOFull = bars + 160;
  const uint32_t bODone = bars + 168;                                            // [2], ONLINE only
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kStages * kUnit + 256);
  float *sh_smem = reinterpret_cast<float *>(gen + kStages * kUnit + 512);      // [2][64] staged column shifts
  float *sh_x = reinterpret_cast<float *>(gen + kStages * kUnit + kAux);        // ONLINE: [2][2][128] ref / rsum exchange

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;
  const int split = blockIdx.y;
  const int n_tiles = (p.n_all + 63) / 64;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(n_tiles, t0 + p.tiles_per_split);
  const int T = t1 - t0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bXFull + 8 * i, 1);
      mbar_init(bXEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bSFull + 8 * i, 1);
      mbar_init(bWFull + 8 * i, 8);           // one elected arrive per epilogue warp
      mbar_init(bODone + 8 * i, 1);
    }
    mbar_init(bOFull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kStages * kUnit + 256), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tmem_a = tmem_base + kTmemA;
  const uint32_t tmem_s = tmem_base + 128;
  const uint32_t tmem_o = tmem_base + 256;
  if (warp >= 2 && warp < 6) park_own_block(tmem_a, p.own_hi, p.own_lo, row0, p.n_own, warp, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      prefetch_tmap(&map_x_hi);
      prefetch_tmap(&map_x_lo);
      for (int i = 0; i < T; ++i) {
        const int stage = i % kStages;
        const int col = (t0 + i) * 64;
        mbar_wait(bXEmpty + 8 * stage, ((i / kStages) & 1) ^ 1);
        const uint32_t fb = bXFull + 8 * stage;
        mbar_expect_tx(fb, kUnit);
        const uint32_t dst = base + stage * kUnit;
        tma_load_2d(dst + 0 * kXTile, &map_x_hi, fb, 0, col);
        tma_load_2d(dst + 1 * kXTile, &map_x_hi, fb, 64, col);
        tma_load_2d(dst + 2 * kXTile, &map_x_lo, fb, 0, col);
        tma_load_2d(dst + 3 * kXTile, &map_x_lo, fb, 64, col);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc1 = instr_desc_f16(128, 64);
    constexpr uint32_t idesc2 = instr_desc_f16_bmn(128, 128);
    // S(i) = own . x_i^T : own from TMEM, x tile K-major
    auto issue_scores = [&](int i) {
      const int stage = i % kStages, buf = i & 1;
      mbar_wait(bXFull + 8 * stage, (i / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t d = tmem_s + buf * 64;
        const uint32_t x0 = base + stage * kUnit;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t a_sel = prod == 2 ? 2 : 0;   // own hi, hi, lo
          const uint32_t b_sel = prod == 1 ? 2 : 0;   // x   hi, lo, hi
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(d, tmem_a + (a_sel + kb) * 32 + k * 8, smem_desc_sw128(x0 + (b_sel + kb) * kXTile + k * 32),
                          idesc1, (prod | kb | k) ? 1u : 0u);
        }
        umma_commit(bSFull + 8 * buf);
      }
      __syncwarp();
    };
    if (T > 0) {
      issue_scores(0);
      if (T > 1) issue_scores(1);
      for (int i = 0; i < T; ++i) {
        const int stage = i % kStages, buf = i & 1;
        mbar_wait(bWFull + 8 * buf, (i >> 1) & 1);
        tc_fence_after();
        if (leader) {
          // O += W(i) . x_i : W from the TMEM columns the scores came from (hi | lo per warpgroup half), the x
          // tile read MN-major (row = K index, 64 contiguous N per 128-byte row, next 64 N one kb block further).
          // k-steps 0,1 are the columns of epilogue warpgroup 0, k-steps 2,3 those of warpgroup 1.
          const uint32_t x0 = base + stage * kUnit;
          const uint32_t w0 = tmem_s + buf * 64;
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {
            const uint32_t w_lo = prod == 2 ? 16 : 0;    // W hi, hi, lo
            const uint32_t x_sel = prod == 1 ? 2 : 0;    // x hi, lo, hi
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t d = ONLINE ? tmem_o + (k >> 1) * 128 : tmem_o;
              const uint32_t acc = ONLINE ? ((i > 0 || prod > 0 || (k & 1)) ? 1u : 0u)
                                          : ((i > 0 || prod > 0 || k > 0) ? 1u : 0u);
              umma_f16_ts(d, w0 + 32 * (k >> 1) + 8 * (k & 1) + w_lo,
                          smem_desc_sw128_mn(x0 + x_sel * kXTile + k * 2048, kXTile, 1024), idesc2, acc);
            }
          }
          umma_commit(bXEmpty + 8 * stage);
          if (ONLINE) umma_commit(bODone + 8 * buf);
        }
        __syncwarp();
        if (i + 2 < T) issue_scores(i + 2);      // overwrites S/W buffer (i & 1): ordered after the MMAs above
      }
      if (leader) umma_commit(bOFull);
      __syncwarp();
    }
  } else {
    // Two epilogue warpgroups split every 128x64 score tile by columns: warpgroup g turns columns
    // [32g, 32g+32) into weights and writes them back over those same TMEM columns: hi in the first 16,
    // lo in the last 16 (one 32-bit column = two consecutive K values).
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int et = wg * 128 + r;              // index among the 256 epilogue threads
    const int gr = p.own_offset + row0 + r;
    const bool row_ok = row0 + r < p.n_own;
    const float s_all = scale_from_absmax(p.absmax[1]);
    const float inv = 1.f / (scale_from_absmax(p.absmax[0]) * s_all);
    const float c2 = inv * kLog2e;
    const bool by_swept = !ONLINE && FAMILY != MIMRL_WEIGHT_SIGMOID && p.shift_by_swept;
    constexpr bool kInterp = FAMILY == MIMRL_WEIGHT_INTERP;
    InterpPar ip = {0.f, 0.f, 0.f, 0.f, nullptr};
    if (kInterp && !p.shift_by_swept && row_ok) {
      const float *q4 = p.shift + row0 + r;
      ip.lse = q4[0], ip.a1 = q4[p.n_par], ip.a2 = q4[2 * (size_t)p.n_par], ip.sg = q4[3 * (size_t)p.n_par];
    }
    // exp family: w * 2^14 = ex2((acc*inv - shift) * log2e + 14); acc*inv is exact (power of two) and the
    // subtraction happens at score magnitude, so the dominant weights keep full fp32 accuracy
    float shift_row = 0.f;
    if (!ONLINE && FAMILY == MIMRL_WEIGHT_EXP && !p.shift_by_swept) shift_row = row_ok ? p.shift[row0 + r] : 0.f;
    const int grmin = p.own_offset + row0, grmax = grmin + 127;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // column shifts of the current tile, staged once per tile (double-buffered) instead of 64 global loads
    // per thread: thread et < 64 prefetches shift[col0 + et] one tile ahead
    float sh_next = 0.f;
    if (by_swept && (kInterp || et < 64) && T > 0) {
      const int gc = t0 * 64 + (et & 63);
      sh_next = gc < p.n_all ? __ldg(p.shift + (kInterp ? (size_t)(et >> 6) * p.n_par : 0) + gc) : 0.f;
    }
    const bool want_rsum = ONLINE || p.rsum_part != nullptr;
    // weights are carried as w * 2^wexp in fp16 hi/lo.  With an exact shift w <= 1 and 2^14 uses the whole fp16 range;
    // the fused forward's reference point is approximate (w can exceed 1), so it leaves 2^6 of headroom, and the
    // exponent is capped so that even a wildly wrong reference point cannot overflow fp16.
    const float wexp = ONLINE ? kOnlineWExp : ((FAMILY == MIMRL_WEIGHT_EXP && want_rsum) ? 10.f : (float)kWExp);
    // row-sum statistic: the 32 weights of a tile are summed on their own first and only the tile total joins the
    // running sum (Kahan-compensated): a row with one dominant weight and tens of thousands of tiny ones would
    // otherwise lose the tail to the rounding of a large accumulator
    float rs_run = 0.f, rs_cmp = 0.f;
    float ref = -INFINITY;                    // ONLINE: running reference point of this row (natural score units)
    const uint32_t tmem_og = tmem_o + (ONLINE ? wg * 128 : 0);
    for (int i = 0; i < T; ++i) {
      const int buf = i & 1;
      const int col0 = (t0 + i) * 64;
      const bool clean = col0 + 64 <= p.n_all && (p.include_diag || col0 + 64 <= grmin || col0 > grmax);
      const bool clean_stat = col0 + 64 <= p.n_all && (col0 + 64 <= grmin || col0 > grmax);
      if (by_swept) {
        if (kInterp) {                        // four parameter vectors of the swept rows: [4][64] per tile, double-buffered
          sh_x[buf * 256 + et] = sh_next;
          const int gc = col0 + 64 + (et & 63);
          sh_next = (i + 1 < T && gc < p.n_all) ? __ldg(p.shift + (size_t)(et >> 6) * p.n_par + gc) : 0.f;
        } else if (et < 64) {
          sh_smem[buf * 64 + et] = sh_next;
          const int gc = col0 + 64 + et;
          sh_next = (i + 1 < T && gc < p.n_all) ? __ldg(p.shift + gc) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(bSFull + 8 * buf, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tcol = tmem_s + lane_off + buf * 64 + wg * 32;
      uint32_t v[32];
      tmem_ld32(tcol, v);
      tmem_ld_wait();
      if (ONLINE) {
        // masked entries (past the batch, or the excluded diagonal) become -inf: weight 0, ignored by the maximum
        if (!clean) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gc = col0 + wg * 32 + j;
            if (!(gc < p.n_all && (p.include_diag || gc != gr))) v[j] = __float_as_uint(-INFINITY);
          }
        }
        float mx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = __uint_as_float(v[u]);
#pragma unroll
        for (int j = 4; j < 32; j += 4)
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(v[j + u]));
        const float cand = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * inv;     // exact: inv is a power of two
        const bool need = cand > ref + kOnlineTau;            // also the first finite tile (ref = -inf); false for -inf
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? ex2((ref - cand) * kLog2e) : 1.f;                      // ref = -inf -> 0
          if (i > 0) {
            // the W.X MMAs of tile i-1 were the last to add into this accumulator; tile i's wait for our weights
            mbar_wait(bODone + 8 * (buf ^ 1), ((i - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t o[32];
              tmem_ld32(tmem_og + lane_off + ch * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
              tmem_st32(tmem_og + lane_off + ch * 32, o);
            }
            tmem_st_wait();
          }
          rs_run *= f, rs_cmp *= f;
          if (need) ref = cand;
        }
      }
      const float ref_use = ONLINE ? (ref > -INFINITY ? ref : 0.f) : shift_row;
      uint32_t hi[16], lo[16];
      float2 rs2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float4 *sh4 = reinterpret_cast<const float4 *>(sh_smem + buf * 64 + wg * 32);
      ip.sm = sh_x + buf * 256 + wg * 32;
      if (clean && clean_stat)
        weights_of_tile<FAMILY, ONLINE, false>(v, hi, lo, rs2, inv, c2, wexp, ref_use, sh4, by_swept, want_rsum,
                                               col0 + wg * 32, gr, p.n_all, p.include_diag != 0, ip);
      else
        weights_of_tile<FAMILY, ONLINE, true>(v, hi, lo, rs2, inv, c2, wexp, ref_use, sh4, by_swept, want_rsum,
                                              col0 + wg * 32, gr, p.n_all, p.include_diag != 0, ip);
      tmem_st16(tcol, hi);
      tmem_st16(tcol + 16, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bWFull + 8 * buf);
      if (want_rsum) {                        // off the critical path: the MMAs of this tile are already released
        const float yk = ((rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y)) - rs_cmp;
        const float tk = rs_run + yk;
        rs_cmp = (tk - rs_run) - yk;
        rs_run = tk;
      }
    }
    const float rs_tot = rs_run * ex2(-wexp);
    if (ONLINE) {
      // merge the two warpgroups' references: both scale their sums to R = max(ref_0, ref_1)
      sh_x[wg * 128 + r] = ref;
      sh_x[256 + wg * 128 + r] = rs_tot;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float ref0 = sh_x[r], ref1 = sh_x[128 + r];
      const float R = fmaxf(ref0, ref1);
      const float f0 = ref0 > -INFINITY ? ex2((ref0 - R) * kLog2e) : 0.f;
      const float f1 = ref1 > -INFINITY ? ex2((ref1 - R) * kLog2e) : 0.f;
      if (wg == 0 && row_ok) {
        p.rsum_part[(size_t)split * p.n_own + row0 + r] = sh_x[256 + r] * f0 + sh_x[384 + r] * f1;
        p.ref_part[(size_t)split * p.n_own + row0 + r] = R;
      }
      if (T > 0) {
        mbar_wait(bOFull, 0);
        tc_fence_after();
        const float os = ex2(-wexp) / s_all;
        const float g0 = f0 * os, g1 = f1 * os;
#pragma unroll 1
        for (int ch = wg * 2; ch < wg * 2 + 2; ++ch) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tmem_o + lane_off + ch * 32, v0);
          tmem_ld32(tmem_o + 128 + lane_off + ch * 32, v1);
          tmem_ld_wait();
          if (row_ok) {
            float4 *o = reinterpret_cast<float4 *>(p.part + ((size_t)split * p.n_own + row0 + r) * 128 + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(fmaf(__uint_as_float(v0[4 * j]), g0, __uint_as_float(v1[4 * j]) * g1),
                                 fmaf(__uint_as_float(v0[4 * j + 1]), g0, __uint_as_float(v1[4 * j + 1]) * g1),
                                 fmaf(__uint_as_float(v0[4 * j + 2]), g0, __uint_as_float(v1[4 * j + 2]) * g1),
                                 fmaf(__uint_as_float(v0[4 * j + 3]), g0, __uint_as_float(v1[4 * j + 3]) * g1));
          }
        }
      } else if (row_ok) {
        float4 *o = reinterpret_cast<float4 *>(p.part + ((size_t)split * p.n_own + row0 + r) * 128 + wg * 64);
        for (int j = 0; j < 16; ++j) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      if (want_rsum && row_ok) p.rsum_part[(size_t)(split * 2 + wg) * p.n_own + row0 + r] = rs_tot;
      if (T > 0) {
        mbar_wait(bOFull, 0);
        tc_fence_after();
        const float oscale = ex2(-wexp) / s_all;
#pragma unroll 1
        for (int ch = wg * 2; ch < wg * 2 + 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(tmem_o + lane_off + ch * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float4 *o = reinterpret_cast<float4 *>(p.part + ((size_t)split * p.n_own + row0 + r) * 128 + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(__uint_as_float(v[4 * j]) * oscale, __uint_as_float(v[4 * j + 1]) * oscale,
                                 __uint_as_float(v[4 * j + 2]) * oscale, __uint_as_float(v[4 * j + 3]) * oscale);
          }
        }
      } else if (row_ok) {
        float4 *o = reinterpret_cast<float4 *>(p.part + ((size_t)split * p.n_own + row0 + r) * 128 + wg * 64);
        for (int j = 0; j < 16; ++j) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ONLINE: merge the column splits of a row, each referred to its own reference point:
// R = max_s ref_s; out = sum_s exp(ref_s - R) part_s; row_sum likewise; row_ref = R (0 when the row saw nothing)
__global__ void online_reduce_tc_kernel(const float *__restrict__ part, const float *__restrict__ rsum_part,
                                        const float *__restrict__ ref_part, int n_splits, int n_own, int embed,
                                        float *__restrict__ out, float *__restrict__ row_sum, float *__restrict__ row_ref) {
  const size_t total = (size_t)n_own * embed;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / embed), e = (int)(idx - (size_t)r * embed);
    float R = -INFINITY;
    for (int s = 0; s < n_splits; ++s) R = fmaxf(R, ref_part[(size_t)s * n_own + r]);
    float a = 0.f, rs = 0.f;
    for (int s = 0; s < n_splits; ++s) {
      const float rf = ref_part[(size_t)s * n_own + r];
      const float f = rf > -INFINITY ? __expf(rf - R) : 0.f;
      a = fmaf(f, part[((size_t)s * n_own + r) * 128 + e], a);
      rs = fmaf(f, rsum_part[(size_t)s * n_own + r], rs);
    }
    out[idx] = a;
    if (e == 0) {
      row_sum[r] = rs;
      row_ref[r] = R > -INFINITY ? R : 0.f;
    }
  }
}

__global__ void rsum_reduce_tc_kernel(const float *__restrict__ part, int n_parts, int n_own, float *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_own) return;
  float a = 0.f;
  for (int s = 0; s < n_parts; ++s) a += part[(size_t)s * n_own + r];
  out[r] = a;
}

// out[r][e] = coef * sum_splits part[s][r][e] + dcoef[r] * all[own_offset + r][e]     (part rows are 128 wide)
__global__ void wsum_reduce_tc_kernel(const float *__restrict__ part, int n_splits, int n_own, int embed,
                                      const float *__restrict__ all, int own_offset, const float *__restrict__ coef,
                                      const float *__restrict__ dcoef, float *__restrict__ out) {
  const size_t total = (size_t)n_own * embed;
  const float c = coef ? coef[0] : 1.f;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / embed), e = (int)(idx - (size_t)r * embed);
    float a = 0.f;
    for (int s = 0; s < n_splits; ++s) a += part[((size_t)s * n_own + r) * 128 + e];
    const float d = dcoef ? dcoef[r] * all[(size_t)(own_offset + r) * embed + e] : 0.f;
    out[idx] = fmaf(c, a, d);
  }
}

// ------------------------------------------------------------------ host ----
struct TcLayout {
  size_t off_absmax, off_own_hi, off_own_lo, off_all_hi, off_all_lo, off_part, off_rsum, off_ref, total;
};


int tc_max_splits() { return 16; }

TcLayout tc_layout(int n_own, int n_all) {
  TcLayout L;
  size_t o = 0;
  L.off_absmax = o;
  o += 256;
  L.off_own_hi = o;
  o += align256((size_t)n_own * 128 * 2);
  L.off_own_lo = o;
  o += align256((size_t)n_own * 128 * 2);
  L.off_all_hi = o;
  o += align256((size_t)n_all * 128 * 2);
  L.off_all_lo = o;
  o += align256((size_t)n_all * 128 * 2);
  L.off_part = o;
  o += align256((size_t)tc_max_splits() * n_own * 128 * sizeof(float));
  L.off_rsum = o;
  o += align256((size_t)2 * tc_max_splits() * n_own * sizeof(float));
  L.off_ref = o;
  o += align256((size_t)tc_max_splits() * n_own * sizeof(float));
  L.total = o;
  return L;
}

// number of column splits: fill the 148 SMs (one CTA each) with as few idle waves as possible
int tc_pick_splits(int row_tiles, int col_tiles) {
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= tc_max_splits() && s <= col_tiles; ++s) {
    const int per = ceil_div(col_tiles, s);
    const int eff = ceil_div(col_tiles, per);
    const double waves = (double)ceil_div(row_tiles * eff, 148);
    const double cost = waves * (per + 2.0);   // +2: per-CTA prologue/epilogue in tile units
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = eff;
    }
  }
  return best;
}

int tc_prepass(const float *own, const float *all, int n_own, int n_all, int embed, const TcLayout &L,
               unsigned char *ws, cudaStream_t st) {
  unsigned *absmax = reinterpret_cast<unsigned *>(ws + L.off_absmax);
  cudaMemsetAsync(absmax, 0, 8, st);
  const size_t na = (size_t)n_own * embed, nb = (size_t)n_all * embed;
  int blocks = (int)(((na > nb ? na : nb) + 4095) / 4096);
  blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
  absmax2_kernel<<<dim3(blocks, 2), 256, 0, st>>>(own, na, all, nb, absmax);
  if (check_launch("tc absmax")) return 1;
  const size_t pairs = (size_t)(n_own > n_all ? n_own : n_all) * 64;
  blocks = (int)((pairs + 255) / 256);
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  split_rows_kernel<<<dim3(blocks, 2), 256, 0, st>>>(
      own, n_own, all, n_all, embed, absmax, reinterpret_cast<__half *>(ws + L.off_own_hi),
      reinterpret_cast<__half *>(ws + L.off_own_lo), reinterpret_cast<__half *>(ws + L.off_all_hi),
      reinterpret_cast<__half *>(ws + L.off_all_lo));
  return check_launch("tc split_rows");
}

}  // namespace

bool sep_tc_supported(int n_own, int n_all, int embed) { return embed <= 128 && n_own > 0 && n_all > 0; }

size_t sep_tc_workspace_bytes(int n_own, int n_all, int embed) {
  if (!sep_tc_supported(n_own, n_all, embed)) return 0;
  return tc_layout(n_own, n_all).total + 256;
}

int sep_row_stats_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset, int flags,
                     float *row_max, float *row_sum, float *row_sp, void *workspace, size_t ws_bytes,
                     cudaStream_t st, const float *row_lse, const float *row_sigma) {
  const TcLayout L = tc_layout(n_own, n_all);
  MIMRL_REQUIRE(ws_bytes >= L.total, "sep_row_stats(tcgen05): workspace too small");
  unsigned char *ws = (unsigned char *)workspace;
  if (tc_prepass(own, all, n_own, n_all, embed, L, ws, st)) return 1;
  CUtensorMap m_all_hi, m_all_lo;
  if (make_map(&m_all_hi, ws + L.off_all_hi, 128, n_all, 128, 128)) return 1;
  if (make_map(&m_all_lo, ws + L.off_all_lo, 128, n_all, 128, 128)) return 1;
  const int row_tiles = ceil_div(n_own, 128), col_tiles = ceil_div(n_all, 128);
  const int splits = tc_pick_splits(row_tiles, col_tiles);
  StatsParams p;
  p.n_own = n_own;
  p.n_all = n_all;
  p.own_offset = own_offset;
  p.tiles_per_split = ceil_div(col_tiles, splits);
  p.absmax = reinterpret_cast<const unsigned *>(ws + L.off_absmax);
  p.part = reinterpret_cast<float *>(ws + L.off_part);
  p.own_hi = reinterpret_cast<const __half *>(ws + L.off_own_hi);
  p.own_lo = reinterpret_cast<const __half *>(ws + L.off_own_lo);
  p.row_lse = row_lse, p.row_sigma = row_sigma;
  dim3 grid(row_tiles, splits);
#define LAUNCH_STATS(F)                                                                                          \
  do {                                                                                                           \
    cudaFuncSetAttribute(sep_stats_tc_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingSmem);    \
    sep_stats_tc_kernel<F><<<grid, kTcThreads, kRingSmem, st>>>(m_all_hi, m_all_lo, p);                           \
  } while (0)
  if (row_lse) {
    LAUNCH_STATS(8);                          // kStatInterp
  } else switch (flags & 7) {
    case 0: LAUNCH_STATS(0); break;
    case 1: LAUNCH_STATS(1); break;
    case 2: LAUNCH_STATS(2); break;
    case 3: LAUNCH_STATS(3); break;
    default: LAUNCH_STATS(4); break;          // MIMRL_STAT_MAXONLY (clamp / softplus do not apply)
  }
#undef LAUNCH_STATS
  if (check_launch("sep_stats_tc")) return 1;
  return combine_row_stats(p.part, splits * 2, n_own, row_max, row_sum, row_sp, st);
}

int sep_weighted_sum_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                        int family, int include_diag, const float *shift, int shift_by_swept, const float *coef,
                        const float *dcoef, float *out, void *workspace, size_t ws_bytes, cudaStream_t st,
                        float *row_sum) {
  const TcLayout L = tc_layout(n_own, n_all);
  MIMRL_REQUIRE(ws_bytes >= L.total, "sep_weighted_sum(tcgen05): workspace too small");
  unsigned char *ws = (unsigned char *)workspace;
  if (tc_prepass(own, all, n_own, n_all, embed, L, ws, st)) return 1;
  CUtensorMap m_x_hi, m_x_lo;
  if (make_map(&m_x_hi, ws + L.off_all_hi, 128, n_all, 128, 64)) return 1;
  if (make_map(&m_x_lo, ws + L.off_all_lo, 128, n_all, 128, 64)) return 1;
  const int row_tiles = ceil_div(n_own, 128), col_tiles = ceil_div(n_all, 64);
  const int splits = tc_pick_splits(row_tiles, col_tiles);
  WsumParams p;
  p.n_own = n_own;
  p.n_all = n_all;
  p.own_offset = own_offset;
  p.tiles_per_split = ceil_div(col_tiles, splits);
  p.include_diag = include_diag;
  p.shift_by_swept = shift_by_swept;
  p.own_hi = reinterpret_cast<const __half *>(ws + L.off_own_hi);
  p.own_lo = reinterpret_cast<const __half *>(ws + L.off_own_lo);
  p.absmax = reinterpret_cast<const unsigned *>(ws + L.off_absmax);
  p.shift = shift;
  p.part = reinterpret_cast<float *>(ws + L.off_part);
  p.rsum_part = row_sum ? reinterpret_cast<float *>(ws + L.off_rsum) : nullptr;
  dim3 grid(row_tiles, splits);
  p.ref_part = nullptr;
  p.n_par = shift_by_swept ? n_all : n_own;
  if (family == MIMRL_WEIGHT_INTERP) {
    cudaFuncSetAttribute(sep_wsum_tc_kernel<MIMRL_WEIGHT_INTERP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kWsumSmem);
    sep_wsum_tc_kernel<MIMRL_WEIGHT_INTERP, false><<<grid, kTcThreads, kWsumSmem, st>>>(m_x_hi, m_x_lo, p);
  } else if (family == MIMRL_WEIGHT_EXP) {
    cudaFuncSetAttribute(sep_wsum_tc_kernel<MIMRL_WEIGHT_EXP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kWsumSmem);
    sep_wsum_tc_kernel<MIMRL_WEIGHT_EXP, false><<<grid, kTcThreads, kWsumSmem, st>>>(m_x_hi, m_x_lo, p);
  } else {
    cudaFuncSetAttribute(sep_wsum_tc_kernel<MIMRL_WEIGHT_SIGMOID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kWsumSmem);
    sep_wsum_tc_kernel<MIMRL_WEIGHT_SIGMOID, false><<<grid, kTcThreads, kWsumSmem, st>>>(m_x_hi, m_x_lo, p);
  }
  if (check_launch("sep_wsum_tc")) return 1;
  const size_t total = (size_t)n_own * embed;
  int blocks = (int)((total + 255) / 256);
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  wsum_reduce_tc_kernel<<<blocks, 256, 0, st>>>(p.part, splits, n_own, embed, all, own_offset, coef, dcoef, out);
  if (check_launch("wsum_reduce_tc")) return 1;
  if (row_sum) {
    rsum_reduce_tc_kernel<<<ceil_div(n_own, 256), 256, 0, st>>>(p.rsum_part, 2 * splits, n_own, row_sum);
    return check_launch("rsum_reduce_tc");
  }
  return 0;
}

// Online-softmax forward (mimrl_sep_online_forward): one sweep, no reference point supplied.
int sep_online_forward_tc(const float *own, const float *all, int n_own, int n_all, int embed, int own_offset,
                          int include_diag, float *row_ref, float *wsum, float *row_sum, void *workspace, size_t ws_bytes,
                          cudaStream_t st) {
  const TcLayout L = tc_layout(n_own, n_all);
  MIMRL_REQUIRE(ws_bytes >= L.total, "sep_online_forward(tcgen05): workspace too small");
  unsigned char *ws = (unsigned char *)workspace;
  if (tc_prepass(own, all, n_own, n_all, embed, L, ws, st)) return 1;
  CUtensorMap m_x_hi, m_x_lo;
  if (make_map(&m_x_hi, ws + L.off_all_hi, 128, n_all, 128, 64)) return 1;
  if (make_map(&m_x_lo, ws + L.off_all_lo, 128, n_all, 128, 64)) return 1;
  const int row_tiles = ceil_div(n_own, 128), col_tiles = ceil_div(n_all, 64);
  const int splits = tc_pick_splits(row_tiles, col_tiles);
  WsumParams p;
  p.n_own = n_own;
  p.n_all = n_all;
  p.own_offset = own_offset;
  p.tiles_per_split = ceil_div(col_tiles, splits);
  p.include_diag = include_diag;
  p.shift_by_swept = 0;
  p.own_hi = reinterpret_cast<const __half *>(ws + L.off_own_hi);
  p.own_lo = reinterpret_cast<const __half *>(ws + L.off_own_lo);
  p.absmax = reinterpret_cast<const unsigned *>(ws + L.off_absmax);
  p.shift = nullptr;
  p.part = reinterpret_cast<float *>(ws + L.off_part);
  p.rsum_part = reinterpret_cast<float *>(ws + L.off_rsum);
  p.ref_part = reinterpret_cast<float *>(ws + L.off_ref);
  p.n_par = 0;
  dim3 grid(row_tiles, splits);
  cudaFuncSetAttribute(sep_wsum_tc_kernel<MIMRL_WEIGHT_EXP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)kWsumSmem);
  sep_wsum_tc_kernel<MIMRL_WEIGHT_EXP, true><<<grid, kTcThreads, kWsumSmem, st>>>(m_x_hi, m_x_lo, p);
  if (check_launch("sep_wsum_tc(online)")) return 1;
  const size_t total = (size_t)n_own * embed;
  int blocks = (int)((total + 255) / 256);
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  online_reduce_tc_kernel<<<blocks, 256, 0, st>>>(p.part, p.rsum_part, p.ref_part, splits, n_own, embed, wsum, row_sum,
                                                  row_ref);
  return check_launch("online_reduce_tc");
}

}  // namespace mimrl
