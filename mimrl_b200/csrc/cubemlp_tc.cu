// CubeMLP axis mix, forward, on the tensor cores (reference MLPProcess.py:94-122,
// the default ln_first = False placement):
//
//     y = LayerNorm_{A'}( W2 act(W1 x + b1) + b2 + (Wres x | x) )        per fibre x[A]
//
// A fibre (one (outer, inner) position of x [outer, A, inner]) is one TMEM lane:
// a CTA handles 128 fibres at a time.  The compute warpgroup (thread = fibre)
// reads its fibre from global memory (coalesced across lanes for inner > 1, one
// contiguous row for inner == 1), scales it by its own power of two, splits it
// into fp16 hi/lo and parks it in TMEM as the A operand.  W1, W2 and Wres
// (fp16 hi/lo, zero padded to 128x128, 128B-swizzled K-major) stay resident in
// shared memory for the whole kernel.  MMA 1 gives the pre-activations; the
// epilogue adds b1, applies the activation and writes h back to TMEM (hi/lo) as
// the A operand of MMA 2 (h W2^T) while x W_res^T accumulates next to it; the
// last epilogue adds b2, normalises over A' thread-locally (a fibre's outputs
// all sit in its own lane) and writes the fibre once.  Three products
// (hi.hi + hi.lo + lo.hi) per contraction keep fp32-class accuracy.
// No permute copies, no intermediate tensors: one read and one write of x / y.
#include "tc_common.cuh"
#include "cubemlp_tc.cuh"

namespace mimrl {
namespace {

constexpr int kCubeWG = 4;                             // compute warpgroups
constexpr int kCubeCh = 4 / kCubeWG;                   // 32-feature chunks of a fibre per compute thread
constexpr int kCubeThreads = 64 + 128 * kCubeWG;       // warp 0 TMA, warp 1 MMA, then the compute warpgroups that
                                                       // split the 32-feature chunks of a fibre (thread = fibre x chunk parity)
constexpr uint32_t kW16 = 128 * 128;                   // one 128-row x 64-K block: 16 KB
constexpr uint32_t kWMat = 4 * kW16;                   // hi kb0, hi kb1, lo kb0, lo kb1
constexpr uint32_t kCubeSmem = 3 * kWMat + 256 + 4 * 128 * 4 + kCubeWG * 128 * 4 + 1024;    // weights, barriers, b1/b2/ln_w/ln_b, LN partials, alignment
// TMEM columns
constexpr uint32_t kTX = 0, kTD1 = 128, kTH = 256, kTD2 = 384;

// 32 consecutive features [a0, a0 + 32) of one fibre (element a at base[a * stride]); entries past n are 0.  When the
// mixed axis is the innermost one (stride 1, the D mix) the fibre is a contiguous row: 16-byte accesses.
__device__ __forceinline__ void fibre_load32(const float *base, int a0, int n, size_t stride, bool vec, bool ok, float (&v)[32]) {
  if (vec) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int a = a0 + 4 * t;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && a < n) q = __ldg(reinterpret_cast<const float4 *>(base + a));          // n % 4 == 0 on this path
      v[4 * t] = q.x, v[4 * t + 1] = q.y, v[4 * t + 2] = q.z, v[4 * t + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int a = a0 + j;
      v[j] = (ok && a < n) ? __ldg(base + (size_t)a * stride) : 0.f;
    }
  }
}
__device__ __forceinline__ void fibre_store32(float *base, int a0, int n, size_t stride, bool vec, const float (&v)[32]) {
  if (vec) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int a = a0 + 4 * t;
      if (a < n) *reinterpret_cast<float4 *>(base + a) = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int a = a0 + j;
      if (a < n) base[(size_t)a * stride] = v[j];
    }
  }
}

template <int ACT>      // activation as a template parameter: only its own code is unrolled into the epilogues
__global__ void __launch_bounds__(kCubeThreads, 1)
cubemlp_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                      const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                      const __grid_constant__ CUtensorMap map_wr_hi, const __grid_constant__ CUtensorMap map_wr_lo,
                      const CubeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t sW1 = base, sW2 = base + kWMat, sWr = base + 2 * kWMat;
  const uint32_t bars = base + 3 * kWMat;
  const uint32_t bWFull = bars, bXReady = bars + 8, bD1Full = bars + 16, bHReady = bars + 24, bD2Full = bars + 32;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + 3 * kWMat + 128);
  float *s_b1 = reinterpret_cast<float *>(gen + 3 * kWMat + 256), *s_b2 = s_b1 + 128, *s_lw = s_b1 + 256, *s_lb = s_b1 + 384;
  for (int t = threadIdx.x; t < 128; t += blockDim.x) {
    s_b1[t] = (p.b1 && t < p.H) ? p.b1[t] : 0.f;
    s_b2[t] = (p.b2 && t < p.A2) ? p.b2[t] : 0.f;
    s_lw[t] = t < p.A2 ? p.ln_w[t] : 0.f;
    s_lb[t] = t < p.A2 ? p.ln_b[t] : 0.f;
  }

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const long long n_tiles = (p.n_cols + 127) / 128;
  const int ks1 = (p.A + 15) / 16, ks2 = (p.H + 15) / 16;         // k-steps of the two contractions
  const int n1 = (p.H + 15) & ~15, n2 = (p.A2 + 15) & ~15;        // MMA N (multiple of 16)

  if (threadIdx.x == 0) {
    mbar_init(bWFull, 1);
    mbar_init(bXReady, 4 * kCubeWG);
    mbar_init(bD1Full, 1);
    mbar_init(bHReady, 4 * kCubeWG);
    mbar_init(bD2Full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + 3 * kWMat + 128), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {      // the three weight matrices, once
      mbar_expect_tx(bWFull, (p.has_res ? 3 : 2) * kWMat);
      tma_load_2d(sW1 + 0 * kW16, &map_w1_hi, bWFull, 0, 0);
      tma_load_2d(sW1 + 1 * kW16, &map_w1_hi, bWFull, 64, 0);
      tma_load_2d(sW1 + 2 * kW16, &map_w1_lo, bWFull, 0, 0);
      tma_load_2d(sW1 + 3 * kW16, &map_w1_lo, bWFull, 64, 0);
      tma_load_2d(sW2 + 0 * kW16, &map_w2_hi, bWFull, 0, 0);
      tma_load_2d(sW2 + 1 * kW16, &map_w2_hi, bWFull, 64, 0);
      tma_load_2d(sW2 + 2 * kW16, &map_w2_lo, bWFull, 0, 0);
      tma_load_2d(sW2 + 3 * kW16, &map_w2_lo, bWFull, 64, 0);
      if (p.has_res) {
        tma_load_2d(sWr + 0 * kW16, &map_wr_hi, bWFull, 0, 0);
        tma_load_2d(sWr + 1 * kW16, &map_wr_hi, bWFull, 64, 0);
        tma_load_2d(sWr + 2 * kW16, &map_wr_lo, bWFull, 0, 0);
        tma_load_2d(sWr + 3 * kW16, &map_wr_lo, bWFull, 64, 0);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc1 = instr_desc_f16(128, n1), idesc2 = instr_desc_f16(128, n2);
    mbar_wait(bWFull, 0);
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      // pre = x W1^T
      mbar_wait(bXReady, ph);
      tc_fence_after();
      if (leader) {
        uint32_t acc = 0;
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t a_off = prod == 2 ? 64 : 0;        // x hi, hi, lo
          const uint32_t b_off = prod == 1 ? 2 * kW16 : 0;  // W hi, lo, hi
          for (int k = 0; k < ks1; ++k) {
            umma_f16_ts(tmem_base + kTD1, tmem_base + kTX + a_off + k * 8,
                        smem_desc_sw128(sW1 + b_off + (k >> 2) * kW16 + (k & 3) * 32), idesc1, acc);
            acc = 1;
          }
        }
        umma_commit(bD1Full);
      }
      __syncwarp();
      // o = h W2^T into D2; r = x Wres^T into the D1 columns (free once the epilogue has read pre)
      mbar_wait(bHReady, ph);
      tc_fence_after();
      if (leader) {
        uint32_t acc = 0;
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t a_off = prod == 2 ? 64 : 0;
          const uint32_t b_off = prod == 1 ? 2 * kW16 : 0;
          for (int k = 0; k < ks2; ++k) {
            umma_f16_ts(tmem_base + kTD2, tmem_base + kTH + a_off + k * 8,
                        smem_desc_sw128(sW2 + b_off + (k >> 2) * kW16 + (k & 3) * 32), idesc2, acc);
            acc = 1;
          }
        }
        if (p.has_res) {
          acc = 0;
          for (int prod = 0; prod < 3; ++prod) {
            const uint32_t a_off = prod == 2 ? 64 : 0;
            const uint32_t b_off = prod == 1 ? 2 * kW16 : 0;
            for (int k = 0; k < ks1; ++k) {
              umma_f16_ts(tmem_base + kTD1, tmem_base + kTX + a_off + k * 8,
                          smem_desc_sw128(sWr + b_off + (k >> 2) * kW16 + (k & 3) * 32), idesc2, acc);
              acc = 1;
            }
          }
        }
        umma_commit(bD2Full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float *s_part = s_lb + 128;                                  // [2 warpgroups][128 fibres]
    const float sx = p.scales[0], sh = p.scales[1];
    const float s1 = 1.f / (sx * scale_from_absmax(p.sc_w1[0])), s2 = 1.f / (sh * scale_from_absmax(p.sc_w2[0]));
    const float s3 = p.has_res ? 1.f / (sx * scale_from_absmax(p.sc_wr[0])) : 0.f;
    const bool vec_in = p.inner == 1 && (p.A & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
    const bool vec_out = p.inner == 1 && (p.A2 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;
    float rstd_max = 0.f;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long c = tile * 128 + r;
      const bool ok = c < p.n_cols;
      const long long o = ok ? c / p.inner : 0, i = ok ? c - o * p.inner : 0;
      const float *xf = p.x + (size_t)o * p.A * p.inner + (size_t)i;
      float *yf = p.y + (size_t)o * p.A2 * p.inner + (size_t)i;
      // ---- 1. fibre -> TMEM (one power-of-two scale for the tensor, fp16 hi/lo)
      for (int ch = g; ch * 32 < ks1 * 16; ch += kCubeWG) {
        float v[32];
        fibre_load32(xf, ch * 32, p.A, p.inner, vec_in, ok, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sx;
        uint32_t hi[16], lo[16];
        split32(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kTX + ch * 16, hi);
        tmem_st16(tmem_base + lane_off + kTX + 64 + ch * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bXReady);
      // ---- 2. h = act(pre + b1) (scale from the bound |act(z)| <= |z| <= max|x| max_h sum_a |W1[h,a]| + max|b1|)
      mbar_wait(bD1Full, ph);
      tc_fence_after();
      for (int ch = g; ch * 32 < ks2 * 16; ch += kCubeWG) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + kTD1 + ch * 32, v);
        tmem_ld_wait();
        float hv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = ch * 32 + j;
          hv[j] = h < p.H ? cube_act(ACT, fmaf(__uint_as_float(v[j]), s1, s_b1[h])) * sh : 0.f;
        }
        uint32_t hi[16], lo[16];
        split32(hv, hi, lo);
        tmem_st16(tmem_base + lane_off + kTH + ch * 16, hi);
        tmem_st16(tmem_base + lane_off + kTH + 64 + ch * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bHReady);
      // ---- 3. z = o + r + b2 (+ x), LayerNorm over A' (this thread's chunks stay in registers; the two
      //         warpgroups exchange their partial sums through shared memory), one write
      mbar_wait(bD2Full, ph);
      tc_fence_after();
      float z[kCubeCh][32];
      float sum = 0.f;
#pragma unroll
      for (int cc = 0; cc < kCubeCh; ++cc) {
        const int ch = g + kCubeWG * cc;
        if (ch * 32 < n2) {
          uint32_t v[32], w[32];
          tmem_ld32(tmem_base + lane_off + kTD2 + ch * 32, v);
          if (p.has_res) tmem_ld32(tmem_base + lane_off + kTD1 + ch * 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int a2 = ch * 32 + j;
            float t = 0.f;
            if (a2 < p.A2) {
              t = fmaf(__uint_as_float(v[j]), s2, s_b2[a2]);
              if (p.has_res) t = fmaf(__uint_as_float(w[j]), s3, t);
              else if (ok) t += __ldg(xf + (size_t)a2 * p.inner);
            }
            z[cc][j] = t;
            sum += t;                                            // entries past A2 are zero
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) z[cc][j] = 0.f;
        }
      }
      tc_fence_before();
      s_part[g * 128 + r] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kCubeWG) : "memory");
      float sum_all = 0.f;
#pragma unroll
      for (int w = 0; w < kCubeWG; ++w) sum_all += s_part[w * 128 + r];
      const float mean = sum_all / p.A2;
      float var = 0.f;
#pragma unroll
      for (int cc = 0; cc < kCubeCh; ++cc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dlt = ((g + kCubeWG * cc) * 32 + j < p.A2) ? z[cc][j] - mean : 0.f;
          var = fmaf(dlt, dlt, var);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kCubeWG) : "memory");             // everyone has read the sums
      s_part[g * 128 + r] = var;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kCubeWG) : "memory");
      float var_all = 0.f;
#pragma unroll
      for (int w = 0; w < kCubeWG; ++w) var_all += s_part[w * 128 + r];
      const float rstd = rsqrtf(var_all / p.A2 + 1e-6f);
      if (ok) {
#pragma unroll
        for (int cc = 0; cc < kCubeCh; ++cc) {
          float yv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int a2 = (g + kCubeWG * cc) * 32 + j;
            yv[j] = a2 < p.A2 ? (z[cc][j] - mean) * rstd * s_lw[a2] + s_lb[a2] : 0.f;
          }
          fibre_store32(yf, (g + kCubeWG * cc) * 32, p.A2, p.inner, vec_out, yv);
        }
        if (g == 0) {
          p.saved[2 * c] = mean;
          p.saved[2 * c + 1] = rstd;
          rstd_max = fmaxf(rstd_max, rstd);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kCubeWG) : "memory");             // partial-sum slots are free for the next tile
    }
    if (g == 0) {                     // the backward scales its operands with the largest rstd: no extra pass over `saved`
      for (int o = 16; o; o >>= 1) rstd_max = fmaxf(rstd_max, __shfl_xor_sync(0xffffffffu, rstd_max, o));
      if (lane == 0) atomicMax(p.absmax + 2, __float_as_uint(rstd_max));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ---- backward on the tensor cores -------------------------------------------------------------------------
// Same tile (128 fibres in the TMEM lanes, W1 / W2 / Wres resident in shared memory), two epilogue warpgroups
// that split the 32-feature chunks of a fibre between them.  One pass per tile:
//
//   x -> X            pre = X W1^T -> D1        h = act(pre + b1) -> H          o = H W2^T -> D2,  r = X Wres^T -> H
//   z = o + r + b2 (+x), zhat from the saved statistics, LayerNorm backward -> gz -> X (over x)
//   gh = GZ W2 -> D2  (the SAME shared-memory tiles, read MN-major)            gpre = gh act'(pre) -> H
//   gx = GPRE W1 -> D1  +  GZ Wres -> D2  (+ gz)                               one write of dL/dx
//
// The weight gradients contract over fibres (the lane axis): x, h, gz, gpre are written once as fp16 hi/lo
// operands, feature-major (coalesced: one feature of 32 consecutive fibres per store), and three split-K GEMMs
// (gemm_tc.cu) finish gW1 = gpre x^T, gW2 = gz h^T, gWres = gz x^T.  Bias and LayerNorm parameter gradients are sums
// over fibres: 32x32 butterfly transposes, accumulated in registers over the CTA's tiles.
// Operand scales are powers of two from bounds known before the launch (absmax of x and gy, max rstd of the saved
// statistics, L1 norms of the weights): |gz| <= rstd (2 + sqrt(A')) max|gy ln_w|.
constexpr int kCubeBwdWG = 2;                          // (four warpgroups spill: 96 registers per thread are too few here)
constexpr int kCubeBwdCh = 4 / kCubeBwdWG;
constexpr int kCubeBwdThreads = 64 + 128 * kCubeBwdWG;
constexpr uint32_t kCbVec = 3 * kWMat + 256;                   // b1 | b2 | ln_w (3 x 128 floats)
constexpr uint32_t kCbPart = kCbVec + 3 * 128 * 4;             // [warpgroups][128 fibres][2] partial LN sums
constexpr uint32_t kCubeBwdSmem = kCbPart + kCubeBwdWG * 128 * 2 * 4 + 1024;




// features [f0, f0 + 32) of one fibre -> feature-major operand in the blocked-K layout of make_map_blocked (tiles of
// 64 consecutive fibres, each [n_feat][64] contiguous); features past n_feat do not exist
__device__ __forceinline__ void cube_store_op(__half *hi_base, __half *lo_base, size_t row, int f0, int n_feat,
                                              const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
  const size_t off = ((row >> 6) * n_feat + f0) * 64 + (row & 63);
  unsigned short *h = reinterpret_cast<unsigned short *>(hi_base) + off, *l = reinterpret_cast<unsigned short *>(lo_base) + off;
  if (f0 + 32 <= n_feat) {            // whole chunk in range: no per-element predicates
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      h[(2 * t) * 64] = (unsigned short)(hi[t] & 0xffffu);
      l[(2 * t) * 64] = (unsigned short)(lo[t] & 0xffffu);
      h[(2 * t + 1) * 64] = (unsigned short)(hi[t] >> 16);
      l[(2 * t + 1) * 64] = (unsigned short)(lo[t] >> 16);
    }
    return;
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (f0 + 2 * t < n_feat) {
      h[(2 * t) * 64] = (unsigned short)(hi[t] & 0xffffu);
      l[(2 * t) * 64] = (unsigned short)(lo[t] & 0xffffu);
    }
    if (f0 + 2 * t + 1 < n_feat) {
      h[(2 * t + 1) * 64] = (unsigned short)(hi[t] >> 16);
      l[(2 * t + 1) * 64] = (unsigned short)(lo[t] >> 16);
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(kCubeBwdThreads, 1)
cubemlp_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                      const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                      const __grid_constant__ CUtensorMap map_wr_hi, const __grid_constant__ CUtensorMap map_wr_lo,
                      const CubeBwdParams bp) {
  const CubeTcParams &p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t sW1 = base, sW2 = base + kWMat, sWr = base + 2 * kWMat;
  const uint32_t bars = base + 3 * kWMat;
  // barriers: weights | x ready | pre full | h ready | o full (mid) | o,r full | gz ready | gh full | gpre ready | gx full
  const uint32_t bW = bars, bX = bars + 8, bD1 = bars + 16, bH = bars + 24, bMid = bars + 32, bD2 = bars + 40, bGz = bars + 48,
                 bD3 = bars + 56, bGp = bars + 64, bD4 = bars + 72;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + 3 * kWMat + 128);
  float *s_b1 = reinterpret_cast<float *>(gen + kCbVec), *s_b2 = s_b1 + 128, *s_lw = s_b1 + 256;
  float *s_part = reinterpret_cast<float *>(gen + kCbPart);
  for (int t = threadIdx.x; t < 128; t += blockDim.x) {
    s_b1[t] = (p.b1 && t < p.H) ? p.b1[t] : 0.f;
    s_b2[t] = (p.b2 && t < p.A2) ? p.b2[t] : 0.f;
    s_lw[t] = t < p.A2 ? p.ln_w[t] : 0.f;
  }

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const long long n_tiles = (p.n_cols + 127) / 128;
  const int ks_a = (p.A + 15) / 16, ks_h = (p.H + 15) / 16, ks_q = (p.A2 + 15) / 16;       // k-steps over A, H, A'
  const int n_a = (p.A + 15) & ~15, n_h = (p.H + 15) & ~15, n_q = (p.A2 + 15) & ~15;         // MMA N

  if (threadIdx.x == 0) {
    mbar_init(bW, 1);
    mbar_init(bX, 4 * kCubeBwdWG);
    mbar_init(bD1, 1);
    mbar_init(bH, 4 * kCubeBwdWG);
    mbar_init(bMid, 1);
    mbar_init(bD2, 1);
    mbar_init(bGz, 4 * kCubeBwdWG);
    mbar_init(bD3, 1);
    mbar_init(bGp, 4 * kCubeBwdWG);
    mbar_init(bD4, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + 3 * kWMat + 128), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx(bW, (p.has_res ? 3 : 2) * kWMat);
      const CUtensorMap *maps[6] = {&map_w1_hi, &map_w1_lo, &map_w2_hi, &map_w2_lo, &map_wr_hi, &map_wr_lo};
      for (int m = 0; m < (p.has_res ? 3 : 2); ++m)
        for (int half = 0; half < 2; ++half)
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(base + m * kWMat + (half * 2 + kb) * kW16, maps[m * 2 + half], bW, kb * 64, 0);
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    mbar_wait(bW, 0);
    // K-major product set: D (+)= A . W^T with W [n rows, k cols] as loaded
    auto mma_k = [&](uint32_t d, uint32_t a, uint32_t sw, int ksteps, uint32_t idesc) {
      uint32_t acc = 0;
      for (int prod = 0; prod < 3; ++prod) {
        const uint32_t a_off = prod == 2 ? 64 : 0, b_off = prod == 1 ? 2 * kW16 : 0;
        for (int k = 0; k < ksteps; ++k) {
          umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128(sw + b_off + (k >> 2) * kW16 + (k & 3) * 32), idesc, acc);
          acc = 1;
        }
      }
    };
    // MN-major product set: D (+)= A . W with the contraction over the ROWS of the same tiles
    auto mma_mn = [&](uint32_t d, uint32_t a, uint32_t sw, int ksteps, uint32_t idesc) {
      uint32_t acc = 0;
      for (int prod = 0; prod < 3; ++prod) {
        const uint32_t a_off = prod == 2 ? 64 : 0, b_off = prod == 1 ? 2 * kW16 : 0;
        for (int k = 0; k < ksteps; ++k) {
          umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128_mn(sw + b_off + k * 2048, kW16, 1024), idesc, acc);
          acc = 1;
        }
      }
    };
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const uint32_t tX = tmem_base + kTX, tD1 = tmem_base + kTD1, tH = tmem_base + kTH, tD2 = tmem_base + kTD2;
      mbar_wait(bX, ph);
      tc_fence_after();
      if (leader) {
        mma_k(tD1, tX, sW1, ks_a, instr_desc_f16(128, n_h));
        umma_commit(bD1);
      }
      __syncwarp();
      mbar_wait(bH, ph);
      tc_fence_after();
      if (leader) {
        mma_k(tD2, tH, sW2, ks_h, instr_desc_f16(128, n_q));
        umma_commit(bMid);
      }
      __syncwarp();
      mbar_wait(bMid, ph);               // the h operand has been consumed: its columns become the accumulator of r
      tc_fence_after();
      if (leader) {
        if (p.has_res) mma_k(tH, tX, sWr, ks_a, instr_desc_f16(128, n_q));
        umma_commit(bD2);
      }
      __syncwarp();
      mbar_wait(bGz, ph);
      tc_fence_after();
      if (leader) {
        mma_mn(tD2, tX, sW2, ks_q, instr_desc_f16_bmn(128, n_h));
        umma_commit(bD3);
      }
      __syncwarp();
      mbar_wait(bGp, ph);
      tc_fence_after();
      if (leader) {
        mma_mn(tD1, tH, sW1, ks_h, instr_desc_f16_bmn(128, n_a));
        if (p.has_res) mma_mn(tD2, tX, sWr, ks_q, instr_desc_f16_bmn(128, n_a));
        umma_commit(bD4);
      }
      __syncwarp();
    }
  } else {
    const int e = warp - 2, q = warp & 3, g = e >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tX = tmem_base + lane_off + kTX, tD1 = tmem_base + lane_off + kTD1, tH = tmem_base + lane_off + kTH,
                   tD2 = tmem_base + lane_off + kTD2;
    const float sx = bp.scales[0], sh = bp.scales[1], sgz = bp.scales[2], sgp = bp.scales[3];
    const float sw1 = scale_from_absmax(p.sc_w1[0]), sw2 = scale_from_absmax(p.sc_w2[0]);
    const float swr = p.has_res ? scale_from_absmax(p.sc_wr[0]) : 1.f;
    const float i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr), i_gh = 1.f / (sgz * sw2),
                i_gx1 = 1.f / (sgp * sw1), i_gx2 = 1.f / (sgz * swr), i_gz = 1.f / sgz;
    float acc_lnw[kCubeBwdCh] = {}, acc_lnb[kCubeBwdCh] = {}, acc_b2[kCubeBwdCh] = {}, acc_b1[kCubeBwdCh] = {};
    const bool vec_in = p.inner == 1 && (p.A & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(bp.gx)) & 15) == 0;
    const bool vec_out = p.inner == 1 && (p.A2 & 3) == 0 && (reinterpret_cast<uintptr_t>(bp.gy) & 15) == 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long c = tile * 128 + r;
      const bool ok = c < p.n_cols;
      const long long o = ok ? c / p.inner : 0, i = ok ? c - o * p.inner : 0;
      const float *xf = p.x + (size_t)o * p.A * p.inner + (size_t)i;
      const float *gyf = bp.gy + (size_t)o * p.A2 * p.inner + (size_t)i;
      float *gxf = bp.gx + (size_t)o * p.A * p.inner + (size_t)i;
      const size_t row = (size_t)c;
      const float mean = ok ? p.saved[2 * c] : 0.f, rstd = ok ? p.saved[2 * c + 1] : 0.f;
      // ---- x -> X operand (+ feature-major copy for the weight gradients)
      for (int ch = g; ch * 32 < ks_a * 16; ch += kCubeBwdWG) {
        float v[32];
        fibre_load32(xf, ch * 32, p.A, p.inner, vec_in, ok, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sx;
        uint32_t hi[16], lo[16];
        split32(v, hi, lo);
        tmem_st16(tX + ch * 16, hi);
        tmem_st16(tX + 64 + ch * 16, lo);
        cube_store_op(bp.op[0][0], bp.op[0][1], row, ch * 32, p.A, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bX);
      // ---- h = act(pre + b1) -> H operand
      mbar_wait(bD1, ph);
      tc_fence_after();
      for (int ch = g; ch * 32 < ks_h * 16; ch += kCubeBwdWG) {
        uint32_t d[32];
        tmem_ld32(tD1 + ch * 32, d);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = ch * 32 + j;
          v[j] = h < p.H ? cube_act(ACT, fmaf(__uint_as_float(d[j]), i_pre, s_b1[h])) * sh : 0.f;
        }
        uint32_t hi[16], lo[16];
        split32(v, hi, lo);
        tmem_st16(tH + ch * 16, hi);
        tmem_st16(tH + 64 + ch * 16, lo);
        cube_store_op(bp.op[1][0], bp.op[1][1], row, ch * 32, p.H, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bH);
      // ---- z, LayerNorm backward -> gz -> X operand
      mbar_wait(bD2, ph);
      tc_fence_after();
      auto z_chunk = [&](int ch, float (&z)[32]) {
        uint32_t d[32], w[32];
        tmem_ld32(tD2 + ch * 32, d);
        if (p.has_res) tmem_ld32(tH + ch * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int a2 = ch * 32 + j;
          float t = 0.f;
          if (a2 < p.A2) {
            t = fmaf(__uint_as_float(d[j]), i_o, s_b2[a2]);
            if (p.has_res) t = fmaf(__uint_as_float(w[j]), i_r, t);
            else if (ok) t += __ldg(xf + (size_t)a2 * p.inner);
          }
          z[j] = t;
        }
      };
      float sum_g = 0.f, sum_gz = 0.f;
      {
        int cc = 0;
        for (int ch = g; ch * 32 < n_q; ch += kCubeBwdWG, ++cc) {
          float z[32], t1[32], t2[32];
          z_chunk(ch, z);
          fibre_load32(gyf, ch * 32, p.A2, p.inner, vec_out, ok, t2);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int a2 = ch * 32 + j;
            const float gyv = t2[j];
            const float zh = (z[j] - mean) * rstd;
            const float gh_ = gyv * s_lw[a2 < p.A2 ? a2 : 0];
            sum_g += gh_;
            sum_gz = fmaf(gh_, zh, sum_gz);
            t1[j] = gyv * zh;
            t2[j] = gyv;
          }
          const float a1 = cube_lane_sum(t1, lane), a2s = cube_lane_sum(t2, lane);
#pragma unroll
          for (int w = 0; w < kCubeBwdCh; ++w)
            if (cc == w) acc_lnw[w] += a1, acc_lnb[w] += a2s;
        }
      }
      s_part[(g * 128 + r) * 2] = sum_g;
      s_part[(g * 128 + r) * 2 + 1] = sum_gz;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kCubeBwdWG) : "memory");
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int w = 0; w < kCubeBwdWG; ++w) m1 += s_part[(w * 128 + r) * 2], m2 += s_part[(w * 128 + r) * 2 + 1];
      m1 /= p.A2, m2 /= p.A2;
      {
        int cc = 0;
        for (int ch = g; ch * 32 < ks_q * 16; ch += kCubeBwdWG, ++cc) {
          float z[32], v[32], t1[32];
          z_chunk(ch, z);
          fibre_load32(gyf, ch * 32, p.A2, p.inner, vec_out, ok, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int a2 = ch * 32 + j;
            const float gyv = v[j];
            const float zh = (z[j] - mean) * rstd;
            const float gzv = a2 < p.A2 ? rstd * (gyv * s_lw[a2 < p.A2 ? a2 : 0] - m1 - zh * m2) : 0.f;
            t1[j] = gzv;
            v[j] = gzv * sgz;
          }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          tmem_st16(tX + ch * 16, hi);
          tmem_st16(tX + 64 + ch * 16, lo);
          cube_store_op(bp.op[2][0], bp.op[2][1], row, ch * 32, p.A2, hi, lo);
          const float a1 = cube_lane_sum(t1, lane);
#pragma unroll
          for (int w = 0; w < kCubeBwdCh; ++w)
            if (cc == w) acc_b2[w] += a1;
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bGz);
      // ---- gpre = gh act'(pre) -> H operand
      mbar_wait(bD3, ph);
      tc_fence_after();
      {
        int cc = 0;
        for (int ch = g; ch * 32 < ks_h * 16; ch += kCubeBwdWG, ++cc) {
          uint32_t d[32], w[32];
          tmem_ld32(tD2 + ch * 32, d);
          tmem_ld32(tD1 + ch * 32, w);
          tmem_ld_wait();
          float v[32], t1[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int h = ch * 32 + j;
            const float pre = fmaf(__uint_as_float(w[j]), i_pre, s_b1[h < p.H ? h : 0]);
            const float gp = h < p.H ? __uint_as_float(d[j]) * i_gh * cube_dact(ACT, pre) : 0.f;
            t1[j] = gp;
            v[j] = gp * sgp;
          }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          tmem_st16(tH + ch * 16, hi);
          tmem_st16(tH + 64 + ch * 16, lo);
          cube_store_op(bp.op[3][0], bp.op[3][1], row, ch * 32, p.H, hi, lo);
          const float a1 = cube_lane_sum(t1, lane);
#pragma unroll
          for (int w = 0; w < kCubeBwdCh; ++w)
            if (cc == w) acc_b1[w] += a1;
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bGp);
      // ---- gx = gpre W1 + gz Wres (+ gz)
      mbar_wait(bD4, ph);
      tc_fence_after();
      for (int ch = g; ch * 32 < n_a; ch += kCubeBwdWG) {
        uint32_t d[32], w[32];
        tmem_ld32(tD1 + ch * 32, d);
        if (p.has_res) tmem_ld32(tD2 + ch * 32, w);
        else {                                         // identity residual: gz itself, back from its operand columns
          uint32_t hi[16], lo[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]),
                         "=r"(hi[8]), "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(hi[12]), "=r"(hi[13]), "=r"(hi[14]), "=r"(hi[15])
                       : "r"(tX + ch * 16) : "memory");
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]),
                         "=r"(lo[8]), "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(lo[12]), "=r"(lo[13]), "=r"(lo[14]), "=r"(lo[15])
                       : "r"(tX + 64 + ch * 16) : "memory");
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&hi[t]));
            const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&lo[t]));
            w[2 * t] = __float_as_uint((a.x + b.x) * i_gz);
            w[2 * t + 1] = __float_as_uint((a.y + b.y) * i_gz);
          }
        }
        tmem_ld_wait();
        if (ok) {
          float gv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = __uint_as_float(d[j]) * i_gx1;
            t += p.has_res ? __uint_as_float(w[j]) * i_gx2 : __uint_as_float(w[j]);
            gv[j] = t;
          }
          fibre_store32(gxf, ch * 32, p.A, p.inner, vec_in, gv);
        }
      }
      tc_fence_before();
    }
#pragma unroll
    for (int cc = 0; cc < kCubeBwdCh; ++cc) {
      const int f = 32 * (g + kCubeBwdWG * cc) + lane;
      if (f < p.A2) {
        atomicAdd(bp.g_lnw + f, acc_lnw[cc]);
        atomicAdd(bp.g_lnb + f, acc_lnb[cc]);
        if (bp.g_b2) atomicAdd(bp.g_b2 + f, acc_b2[cc]);
      }
      if (f < p.H && bp.g_b1) atomicAdd(bp.g_b1 + f, acc_b1[cc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- one preparation launch per mix --------------------------------------------------------------------------
// Blocks [0, n_abs): absmax of x, gy and of the saved rstd column (grid-stride, one atomicMax per warp).
// Blocks n_abs + {0, 1, 2}: fp16 hi/lo split of W1 / W2 / Wres in the mimrl_split_f32 layout (absmax header, zero
// padded rows of ld halves); the matrices are <= 128 x 128, one block each.  The LAST block to finish (atomic ticket)
// derives the power-of-two operand scales from those maxima and from L1 norms of the weights:
//   |h| <= |pre| <= max|x| max_h sum_a |W1[h,a]| + max|b1|          (|act(z)| <= |z| for gelu, relu, tanh)
//   |gz| <= rstd (2 + sqrt(A')) max|gy ln_w|,     |gpre| <= 1.2 |gz| max_h sum_q |W2[q,h]|   (|gelu'| < 1.13)
// and, with a residual projection, ties the scales of gz and gpre together so that gpre W1 and gz Wres can share one
// accumulator (s_gpre s_W1 == s_gz s_Wres; both are powers of two and the bounds are loose, so nothing is lost).
// tail: [0] max|x|  [1] max|gy|  [2] max rstd  [3] ticket; scales at tail + 16 (floats).
struct CubePrepArgs {
  const float *x, *gy, *saved;
  size_t nx, ngy, n_cols;
  const float *w[3];                 // W1 [H, A], W2 [A2, H], Wres [A2, A] (nullable)
  unsigned char *split[3];           // split buffers; nullptr = already split (backward after forward)
  const unsigned char *hdr[3];       // where the absmax headers of the splits live (always valid)
  const float *b1, *ln_w;
  const float *prev_ln_w, *prev_ln_b;      // LayerNorm parameters of the mix that produced x (nullable), prev_n of them
  int prev_n;
  int A, H, A2, n_abs, backward;
  unsigned *tail;
  unsigned *hdr_op[4];               // backward: absmax headers of the weight-gradient operands x, h, gz, gpre
};

constexpr int kPrepThreads = 1024;
__device__ __forceinline__ void cube_prep_body(const CubePrepArgs &a, const int block, const int n_blocks) {
  __shared__ float red[kPrepThreads];
  __shared__ unsigned s_last;
  const int t = threadIdx.x;
  auto block_max = [&](float v) {
    __syncthreads();
    red[t] = v;
    __syncthreads();
    for (int o = kPrepThreads / 2; o; o >>= 1) {
      if (t < o) red[t] = fmaxf(red[t], red[t + o]);
      __syncthreads();
    }
    return red[0];
  };
  if (block < a.n_abs) {
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    const size_t t0 = (size_t)block * blockDim.x + t, st = (size_t)a.n_abs * blockDim.x;
    auto amax4 = [&](const float *p, size_t n, float m) {          // 16-byte loads where the pointer allows it
      if (!p) return m;
      size_t head = 0;
      if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        auto take = [&](const float4 v) { m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w))); };
        size_t i = t0;
        for (; i + 7 * st < n / 4; i += 8 * st) {          // eight independent 16-byte loads in flight per thread
          float4 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __ldg(p4 + i + k * st);
#pragma unroll
          for (int k = 0; k < 8; ++k) take(v[k]);
        }
        for (; i < n / 4; i += st) take(__ldg(p4 + i));
        head = (n / 4) * 4;
      }
      for (size_t i = head + t0; i < n; i += st) m = fmaxf(m, fabsf(p[i]));
      return m;
    };
    m0 = amax4(a.x, a.nx, m0);
    m1 = amax4(a.gy, a.ngy, m1);
    if (a.saved)
      for (size_t i = t0; i < a.n_cols; i += st) m2 = fmaxf(m2, fabsf(a.saved[2 * i + 1]));
    for (int o = 16; o; o >>= 1) {
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if ((t & 31) == 0) {
      if (a.x) atomicMax(a.tail, __float_as_uint(m0));
      if (a.gy) atomicMax(a.tail + 1, __float_as_uint(m1));
      if (a.saved) atomicMax(a.tail + 2, __float_as_uint(m2));
    }
  } else {
    const int m = block - a.n_abs;
    const float *w = a.w[m];
    unsigned char *out = a.split[m];
    if (w && out) {                                            // block-uniform
      const int rows = m == 0 ? a.H : a.A2, cols = m == 1 ? a.H : a.A;
      const int ld = (cols + 63) & ~63;
      float mx = 0.f;
      for (int i = t; i < rows * cols; i += kPrepThreads) mx = fmaxf(mx, fabsf(w[i]));
      const unsigned bits = __float_as_uint(block_max(mx));
      if (t == 0) *reinterpret_cast<unsigned *>(out) = bits;
      const float sc = scale_from_absmax(bits);
      __half *hi = reinterpret_cast<__half *>(out + 256);
      __half *lo = reinterpret_cast<__half *>(out + 256 + align256((size_t)rows * ld * 2));
      const int half_ld = ld >> 1;
      for (int i = t; i < rows * half_ld; i += kPrepThreads) {
        const int r = i / half_ld, c = (i - r * half_ld) * 2;
        const float v0 = c < cols ? w[r * cols + c] * sc : 0.f, v1 = c + 1 < cols ? w[r * cols + c + 1] * sc : 0.f;
        const __half2 h = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(h);
        *reinterpret_cast<__half2 *>(hi + r * ld + c) = h;
        *reinterpret_cast<__half2 *>(lo + r * ld + c) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      }
      // L1 norms the scale bounds need, while this block has the matrix in cache (tail[12]: max_h sum_a |W1[h,a]|,
      // tail[13]: max_h sum_q |W2[q,h]|); they stay in the workspace for a backward that reuses it
      if (m < 2) {
        float part = 0.f;
        if (m == 0) {                                   // row sums: eight threads per row
          const int r = t >> 3;
          if (r < rows)
            for (int c = t & 7; c < cols; c += 8) part += fabsf(w[r * cols + c]);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
        } else {                                        // column sums: eight threads per column (coalesced rows)
          const int c = t & 127;
          if (c < cols)
            for (int r = t >> 7; r < rows; r += 8) part += fabsf(w[r * cols + c]);
          __syncthreads();
          red[t] = part;
          __syncthreads();
          if (t < 128) {
            part = 0.f;
            for (int k = 0; k < 8; ++k) part += red[t + 128 * k];
          } else {
            part = 0.f;
          }
        }
        const float nm = block_max(part);
        if (t == 0) a.tail[12 + m] = __float_as_uint(nm);
      }
    }
  }
  // ---- ticket: the last block computes the scales
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(a.tail + 3, 1u) == (unsigned)(n_blocks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int A = a.A, H = a.H, A2 = a.A2;
  const float row1_max = __uint_as_float(reinterpret_cast<volatile unsigned *>(a.tail)[12]);
  const float col2_max = __uint_as_float(reinterpret_cast<volatile unsigned *>(a.tail)[13]);
  const float b1_max = block_max((a.b1 && t < H) ? fabsf(a.b1[t]) : 0.f), lw_max = block_max(t < A2 ? fabsf(a.ln_w[t]) : 0.f);
  // x is the output of a LayerNorm over prev_n features: |x| <= max|ln_w| sqrt(prev_n - 1) + max|ln_b| (a z-score of n
  // values is at most sqrt(n - 1)); loose by a few binades, which the fp16 hi/lo pair has to spare -- no pass over x
  float prev_bound = 0.f;
  if (a.prev_ln_w) {
    const float pw = block_max(t < a.prev_n ? fabsf(a.prev_ln_w[t]) : 0.f);
    const float pb = block_max((a.prev_ln_b && t < a.prev_n) ? fabsf(a.prev_ln_b[t]) : 0.f);
    prev_bound = pw * sqrtf((float)(a.prev_n > 1 ? a.prev_n - 1 : 1)) * 1.0001f + pb;
  }
  if (t == 0) {
    volatile unsigned *tl = a.tail;
    float m_x = __uint_as_float(tl[0]);
    const float m_gy = __uint_as_float(tl[1]), m_rstd = __uint_as_float(tl[2]);
    if (prev_bound > 0.f) m_x = prev_bound, tl[0] = __float_as_uint(prev_bound);
    float *scales = reinterpret_cast<float *>(a.tail + 4);
    tl[14] = __float_as_uint(b1_max);             // (with tail[12]: what a per-fibre bound of |h| needs)
    const float m_h = m_x * row1_max + b1_max;
    scales[0] = pow2_scale(m_x), scales[1] = pow2_scale(m_h);
    if (a.backward) {
      const float m_gz = m_rstd * (2.f + sqrtf((float)A2)) * m_gy * lw_max;
      const float m_gp = 1.2f * m_gz * col2_max;
      float s_gz = pow2_scale(m_gz), s_gp = pow2_scale(m_gp);
      if (a.w[2]) {                  // s_gp s_w1 == s_gz s_wr, neither operand above 2^14
        const float sw1 = scale_from_absmax(*reinterpret_cast<const volatile unsigned *>(a.hdr[0]));
        const float swr = scale_from_absmax(*reinterpret_cast<const volatile unsigned *>(a.hdr[2]));
        const float rho = swr / sw1;                       // required s_gp / s_gz
        if (s_gz * rho <= s_gp) s_gp = s_gz * rho;         // gpre gets the smaller scale
        else s_gz = s_gp / rho;                            // gz gets the smaller scale
      }
      scales[2] = s_gz, scales[3] = s_gp;
      if (a.hdr_op[0]) {
        *a.hdr_op[0] = __float_as_uint(m_x), *a.hdr_op[1] = __float_as_uint(m_h);
        // the operand headers carry the absmax the GEMM derives its scale from: invert pow2_scale (2^13.5 / scale)
        *a.hdr_op[2] = __float_as_uint(11585.2375f / s_gz), *a.hdr_op[3] = __float_as_uint(11585.2375f / s_gp);
      }
    }
    tl[3] = 0;                        // ticket ready for the next launch on this workspace
    if (!a.backward) tl[2] = 0;       // the forward kernel collects its largest rstd here
  }
}

__global__ void __launch_bounds__(kPrepThreads) cube_prep_kernel(const CubePrepArgs a) { cube_prep_body(a, blockIdx.x, gridDim.x); }

// the preparation of SEVERAL mixes in one launch (a whole encoder forward: every mix but the first is independent of the
// activations -- weights and the LayerNorm bound of its input only); mix m owns blocks [first[m], first[m + 1])
constexpr int kPrepMany = 8;
struct CubePrepBatch {
  CubePrepArgs a[kPrepMany];
  int first[kPrepMany + 1];
  int n;
};
__global__ void __launch_bounds__(kPrepThreads) cube_prep_many_kernel(const __grid_constant__ CubePrepBatch b) {
  int m = 0;
  while (m + 1 < b.n && (int)blockIdx.x >= b.first[m + 1]) ++m;
  cube_prep_body(b.a[m], blockIdx.x - b.first[m], b.first[m + 1] - b.first[m]);
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_split_bytes(int rows, int cols);

extern "C" int mimrl_cubemlp_tc_supported(int a_in, int a_hid, int a_out, int ln_first, int act) {
  const bool small = a_in <= 4 && a_hid <= 4 && a_out <= 4;      // the register-resident kernel of cubemlp.cu
  return !ln_first && !small && a_in <= 128 && a_hid <= 128 && a_out <= 128 && act >= 0 && act <= 2;
}

extern "C" size_t mimrl_cubemlp_tc_workspace_bytes(int a_in, int a_hid, int a_out) {
  return mimrl_split_bytes(a_hid, a_in) + mimrl_split_bytes(a_out, a_hid) + mimrl_split_bytes(a_out, a_in) + 256;
}

namespace {

struct CubeWs {
  unsigned char *s[3];
  unsigned *tail;
};
CubeWs cube_ws(void *workspace, int a_in, int a_hid, int a_out) {
  CubeWs w;
  w.s[0] = (unsigned char *)workspace;
  w.s[1] = w.s[0] + mimrl_split_bytes(a_hid, a_in);
  w.s[2] = w.s[1] + mimrl_split_bytes(a_out, a_hid);
  w.tail = reinterpret_cast<unsigned *>(w.s[2] + mimrl_split_bytes(a_out, a_in));           // the 256 spare bytes
  return w;
}

// maps[0..5] = W1 hi, lo | W2 hi, lo | Wres hi, lo with the given box heights
int cube_weight_maps(const CubeWs &w, int a_in, int a_hid, int a_out, bool has_res, int r1, int r2, int rr, CUtensorMap *maps) {
  auto two = [&](unsigned char *s, int rows, int cols, int box_rows, CUtensorMap *hi, CUtensorMap *lo) {
    const int ld = (cols + 63) & ~63;
    const size_t off_lo = 256 + align256((size_t)rows * ld * 2);
    if (make_map(hi, s + 256, cols, rows, ld, box_rows)) return 1;
    return make_map(lo, s + off_lo, cols, rows, ld, box_rows);
  };
  if (two(w.s[0], a_hid, a_in, r1, &maps[0], &maps[1])) return 1;
  if (two(w.s[1], a_out, a_hid, r2, &maps[2], &maps[3])) return 1;
  if (has_res) {
    if (two(w.s[2], a_out, a_in, rr, &maps[4], &maps[5])) return 1;
  } else {
    maps[4] = maps[0], maps[5] = maps[1];
  }
  return 0;
}

}  // namespace

namespace {
// The compile-time specialised sequence-mix forward (cubemlp_tc2.cu) holds a whole fibre per thread: with an input that is
// not a LayerNorm output (the first mix of an encoder) it takes the operand scales per fibre instead of from a pass over x.
bool cube_fibre_scale(const float *prev_ln_w, int prev_n, int a_in, int a_hid, int a_out, int inner, long long n_cols, int act,
                      bool has_res) {
  const bool bounded = prev_ln_w != nullptr && prev_n > 0 && prev_n <= 256;
  return !bounded && !getenv("MIMRL_CUBE_FIBRE_SCALE_OFF") && cube2_supported(a_in, a_hid, a_out, inner, n_cols, act, has_res ? 1 : 0);
}
CubePrepArgs cube_fwd_prep_args(const CubeWs &w, const float *x, long long n_cols, int a_in, int a_hid, int a_out, const float *w1,
                                const float *b1, const float *w2, const float *wres, const float *ln_w, const float *prev_ln_w,
                                const float *prev_ln_b, int prev_n, bool scale_in_kernel) {
  CubePrepArgs pa{};
  const bool bounded = prev_ln_w != nullptr && prev_n > 0 && prev_n <= 256;
  // scale_in_kernel (cube_fibre_scale): the forward kernel scales every fibre by its own max|x| -- no pass over x here
  pa.x = bounded || scale_in_kernel ? nullptr : x, pa.nx = bounded || scale_in_kernel ? 0 : (size_t)n_cols * a_in;
  pa.prev_ln_w = bounded ? prev_ln_w : nullptr, pa.prev_ln_b = bounded ? prev_ln_b : nullptr, pa.prev_n = bounded ? prev_n : 0;
  pa.w[0] = w1, pa.w[1] = w2, pa.w[2] = wres;
  for (int m = 0; m < 3; ++m) pa.split[m] = w.s[m], pa.hdr[m] = w.s[m];
  pa.b1 = b1, pa.ln_w = ln_w, pa.A = a_in, pa.H = a_hid, pa.A2 = a_out, pa.backward = 0, pa.tail = w.tail;
  const size_t want = (pa.nx + 32767) / 32768;
  pa.n_abs = (int)(want > 145 ? 145 : (want < 1 ? 1 : want));          // one wave of 1024-thread blocks (one resident per SM)
  return pa;
}
}  // namespace

// The preparation (weight split, operand scales) of n <= 8 mixes of an encoder forward in ONE launch.  Every argument is
// an array of n entries, in execution order; workspace[m] must be ZERO-FILLED (mimrl_cubemlp_tc_workspace_bytes each).
// Mix m > 0 must carry prev_ln_w / prev_ln_b / prev_n (its input is the previous mix's LayerNorm output); x / n_cols
// describe the input of mix 0 (ignored when mix 0 has prev_ln_w).  Afterwards call mimrl_cubemlp_mix_fwd_tc with
// prepared = 1 on the same workspaces.
extern "C" int mimrl_cubemlp_prep_many(int n, const float *x, const long long *n_cols, const int *a_in, const int *a_hid,
                                       const int *a_out, const int *inner, const int *act, const float *const *w1,
                                       const float *const *b1,
                                       const float *const *w2, const float *const *wres, const float *const *ln_w,
                                       const float *const *prev_ln_w, const float *const *prev_ln_b, const int *prev_n,
                                       void *const *workspace, void *stream) {
  MIMRL_REQUIRE(n >= 1 && n <= kPrepMany, "cubemlp_prep_many: 1..%d mixes", kPrepMany);
  CubePrepBatch b{};
  b.n = n;
  int first = 0;
  for (int m = 0; m < n; ++m) {
    MIMRL_REQUIRE(mimrl_cubemlp_tc_supported(a_in[m], a_hid[m], a_out[m], 0, 0) && w1[m] && w2[m] && ln_w[m] && workspace[m],
                  "cubemlp_prep_many: mix %d is outside the tensor-core kernels", m);
    MIMRL_REQUIRE(m == 0 || (prev_ln_w[m] && prev_n[m] > 0 && prev_n[m] <= 256), "cubemlp_prep_many: mix %d needs the LayerNorm bound of its input", m);
    const CubeWs w = cube_ws(workspace[m], a_in[m], a_hid[m], a_out[m]);
    const bool fs = cube_fibre_scale(prev_ln_w[m], prev_n[m], a_in[m], a_hid[m], a_out[m], inner[m], n_cols[m], act[m], wres[m] != nullptr);
    b.a[m] = cube_fwd_prep_args(w, m == 0 ? x : nullptr, n_cols[m], a_in[m], a_hid[m], a_out[m], w1[m], b1[m], w2[m], wres[m], ln_w[m],
                                prev_ln_w[m], prev_ln_b[m], prev_n[m], fs);
    MIMRL_REQUIRE(m != 0 || b.a[0].prev_ln_w || x, "cubemlp_prep_many: mix 0 needs x or a LayerNorm bound");
    // 1024-thread blocks at 56 registers: one resident block per SM -- keep the whole launch to one wave
    if (m == 0 && b.a[0].n_abs > 148 - 3 * n) b.a[0].n_abs = 148 - 3 * n;
    b.first[m] = first;
    first += b.a[m].n_abs + 3;
  }
  b.first[n] = first;
  cube_prep_many_kernel<<<first, kPrepThreads, 0, (cudaStream_t)stream>>>(b);
  return check_launch("cubemlp prep (encoder)");
}

extern "C" int mimrl_cubemlp_mix_fwd_tc(const float *x, int outer, int a_in, int inner, const float *w1, const float *b1,
                                        int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                                        const float *ln_w, const float *ln_b, int act, float *y, float *saved,
                                        void *workspace, size_t workspace_bytes, const float *prev_ln_w,
                                        const float *prev_ln_b, int prev_n, int prepared, void *stream) {
  MIMRL_REQUIRE(mimrl_cubemlp_tc_supported(a_in, a_hid, a_out, 0, act), "cubemlp_mix_fwd_tc: sizes %d/%d/%d act %d not supported",
                a_in, a_hid, a_out, act);
  MIMRL_REQUIRE(outer > 0 && inner > 0 && x && y && saved && w1 && w2 && ln_w && ln_b, "cubemlp_mix_fwd_tc: bad arguments");
  MIMRL_REQUIRE(wres || a_in == a_out, "cubemlp_mix: without res_project d_in must equal d_out (MLPProcess.py:46-48)");
  MIMRL_REQUIRE(workspace_bytes >= mimrl_cubemlp_tc_workspace_bytes(a_in, a_hid, a_out), "cubemlp_mix_fwd_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const CubeWs w = cube_ws(workspace, a_in, a_hid, a_out);
  const long long n_cols = (long long)outer * inner;
  const bool fs = cube_fibre_scale(prev_ln_w, prev_n, a_in, a_hid, a_out, inner, n_cols, act, wres != nullptr);
  if (!prepared) {
    cudaMemsetAsync(w.tail, 0, 16, st);
    const CubePrepArgs pa = cube_fwd_prep_args(w, x, n_cols, a_in, a_hid, a_out, w1, b1, w2, wres, ln_w, prev_ln_w, prev_ln_b, prev_n, fs);
    cube_prep_kernel<<<pa.n_abs + 3, kPrepThreads, 0, st>>>(pa);
    if (check_launch("cubemlp prep")) return 1;
  }
  CubeTcParams p;
  p.scales = reinterpret_cast<const float *>(w.tail + 4);
  p.absmax = w.tail;
  p.x = x, p.b1 = b1, p.b2 = b2, p.ln_w = ln_w, p.ln_b = ln_b, p.y = y, p.saved = saved;
  p.sc_w1 = reinterpret_cast<const unsigned *>(w.s[0]), p.sc_w2 = reinterpret_cast<const unsigned *>(w.s[1]);
  p.sc_wr = reinterpret_cast<const unsigned *>(w.s[2]);
  p.outer = outer, p.A = a_in, p.H = a_hid, p.A2 = a_out, p.inner = inner, p.act = act, p.has_res = wres ? 1 : 0;
  p.n_cols = n_cols;
  p.fibre_scale = fs ? 1 : 0;
  if (!getenv("MIMRL_CUBE2_PREFETCH_OFF")) p.fibre_scale |= 2;          // bit 1: L2 prefetch of the next tile's rows
  CUtensorMap maps[6];
  if (cube2_supported(a_in, a_hid, a_out, inner, n_cols, act, p.has_res)) {
    int r1, r2, rr, handled = 0;
    cube2_box_rows(a_in, a_hid, a_out, &r1, &r2, &rr);
    if (cube_weight_maps(w, a_in, a_hid, a_out, wres != nullptr, r1, r2, rr, maps)) return 1;
    if (int rc = cube2_fwd(maps, p, st, &handled)) return rc;
    if (handled) return 0;
  }
  if (cube_weight_maps(w, a_in, a_hid, a_out, wres != nullptr, 128, 128, 128, maps)) return 1;
  if (cube3_supported(a_in, a_hid, a_out, inner, n_cols, act, p.has_res) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    int handled = 0;
    if (int rc = cube3_fwd(maps, p, st, &handled)) return rc;
    if (handled) return 0;
  }
  const long long n_tiles = (p.n_cols + 127) / 128;
  const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCubeSmem);
    kern<<<blocks, kCubeThreads, kCubeSmem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  };
  if (act == 0) launch(cubemlp_tc_fwd_kernel<0>);
  else if (act == 1) launch(cubemlp_tc_fwd_kernel<1>);
  else launch(cubemlp_tc_fwd_kernel<2>);
  return check_launch("cubemlp_tc_fwd");
}

// bytes of one weight-gradient operand buffer with `features` features over R fibre rows (either layout fits)
extern "C" size_t mimrl_cubemlp_tc_op_bytes(int features, long long R) {
  if (features <= 0 || R <= 0) return 0;
  return mimrl_split_bytes(features, (int)R);
}

// rows (fibres, whole tiles) of the feature-major weight-gradient operands
extern "C" long long mimrl_cubemlp_tc_fibre_rows(int outer, int inner) {
  if (outer <= 0 || inner <= 0) return 0;
  return (((long long)outer * inner + 127) / 128) * 128;
}

// Backward of mimrl_cubemlp_mix_fwd_tc.  Writes gx; accumulates (+=) g_b1 [a_hid], g_b2 [a_out], gln_w, gln_b [a_out];
// writes op_x [a_in, R], op_h [a_hid, R], op_gz [a_out, R], op_gpre [a_hid, R] (R = mimrl_cubemlp_tc_fibre_rows,
// mimrl_split_f32 sizes, blocked-K order): gW1 = op_gpre op_x^T, gW2 = op_gz op_h^T, gWres = op_gz op_x^T via
// mimrl_gemm_split_blocked -- launched here: gw1 [a_hid, a_in], gw2 [a_out, a_hid], gwres [a_out, a_in] are ACCUMULATED
// (+=, caller zero-fills).  Operand buffers: mimrl_cubemlp_tc_op_bytes each.  ws_from_forward != 0: `workspace` is the buffer the forward of the same mix filled
// (same x and weights), so the weight split, max|x| and the rstd maximum are taken from it instead of recomputed.
extern "C" int mimrl_cubemlp_mix_bwd_tc(const float *x, const float *gy, int outer, int a_in, int inner, const float *w1,
                                        const float *b1, int a_hid, const float *w2, const float *b2, int a_out,
                                        const float *wres, const float *ln_w, const float *ln_b, int act,
                                        const float *saved, float *gx, float *g_b1, float *g_b2, float *gln_w, float *gln_b,
                                        float *gw1, float *gw2, float *gwres, void *op_x, void *op_h, void *op_gz,
                                        void *op_gpre, void *workspace, size_t workspace_bytes, int ws_from_forward,
                                        void *stream) {
  MIMRL_REQUIRE(mimrl_cubemlp_tc_supported(a_in, a_hid, a_out, 0, act), "cubemlp_mix_bwd_tc: sizes %d/%d/%d act %d not supported",
                a_in, a_hid, a_out, act);
  MIMRL_REQUIRE(outer > 0 && inner > 0 && x && gy && saved && gx && w1 && w2 && ln_w && gln_w && gln_b && op_x && op_h && op_gz &&
                    op_gpre && gw1 && gw2 && (gwres || !wres),
                "cubemlp_mix_bwd_tc: bad arguments");
  MIMRL_REQUIRE(wres || a_in == a_out, "cubemlp_mix: without res_project d_in must equal d_out (MLPProcess.py:46-48)");
  MIMRL_REQUIRE(workspace_bytes >= mimrl_cubemlp_tc_workspace_bytes(a_in, a_hid, a_out), "cubemlp_mix_bwd_tc: workspace too small");
  (void)ln_b;
  cudaStream_t st = (cudaStream_t)stream;
  const CubeWs w = cube_ws(workspace, a_in, a_hid, a_out);
  const long long n_cols = (long long)outer * inner;
  CubePrepArgs pa{};
  if (ws_from_forward) {
    cudaMemsetAsync(w.tail + 1, 0, 4, st);                      // max|gy| only; x, rstd and the ticket stay
  } else {
    cudaMemsetAsync(w.tail, 0, 16, st);
    pa.x = x, pa.nx = (size_t)n_cols * a_in, pa.saved = saved, pa.n_cols = (size_t)n_cols;
  }
  pa.gy = gy, pa.ngy = (size_t)n_cols * a_out;
  pa.w[0] = w1, pa.w[1] = w2, pa.w[2] = wres;
  for (int m = 0; m < 3; ++m) pa.split[m] = ws_from_forward ? nullptr : w.s[m], pa.hdr[m] = w.s[m];
  pa.b1 = b1, pa.ln_w = ln_w, pa.A = a_in, pa.H = a_hid, pa.A2 = a_out, pa.backward = 1, pa.tail = w.tail;
  pa.hdr_op[0] = (unsigned *)op_x, pa.hdr_op[1] = (unsigned *)op_h, pa.hdr_op[2] = (unsigned *)op_gz, pa.hdr_op[3] = (unsigned *)op_gpre;
  {
    const size_t want = ((ws_from_forward ? 0 : pa.nx) + pa.ngy + 32767) / 32768;
    pa.n_abs = (int)(want > 145 ? 145 : (want < 1 ? 1 : want));          // one wave of 1024-thread blocks (one resident per SM)
  }
  cube_prep_kernel<<<pa.n_abs + 3, kPrepThreads, 0, st>>>(pa);
  if (check_launch("cubemlp prep")) return 1;
  CubeBwdParams bp;
  CubeTcParams &p = bp.f;
  p.scales = reinterpret_cast<const float *>(w.tail + 4);
  p.absmax = w.tail;
  p.x = x, p.b1 = b1, p.b2 = b2, p.ln_w = ln_w, p.ln_b = ln_b, p.y = nullptr, p.saved = const_cast<float *>(saved);
  p.sc_w1 = reinterpret_cast<const unsigned *>(w.s[0]), p.sc_w2 = reinterpret_cast<const unsigned *>(w.s[1]);
  p.sc_wr = reinterpret_cast<const unsigned *>(w.s[2]);
  p.outer = outer, p.A = a_in, p.H = a_hid, p.A2 = a_out, p.inner = inner, p.act = act, p.has_res = wres ? 1 : 0;
  p.n_cols = n_cols;
  p.fibre_scale = 0;          // (forward-only switches; the same prefetch in the backward kernel measured no gain)
  bp.gy = gy, bp.gx = gx, bp.g_b1 = g_b1, bp.g_b2 = g_b2, bp.g_lnw = gln_w, bp.g_lnb = gln_b;
  bp.scales = p.scales;
  const long long n_tiles = (n_cols + 127) / 128;
  bp.ld = (size_t)n_tiles * 128;
  void *ops[4] = {op_x, op_h, op_gz, op_gpre};
  const int feats[4] = {a_in, a_hid, a_out, a_hid};
  const long long R = n_tiles * 128;
  CUtensorMap maps[6];
  const bool c2 = cube2_supported(a_in, a_hid, a_out, inner, n_cols, act, p.has_res);
  const bool c3 = !c2 && cube3_supported(a_in, a_hid, a_out, inner, n_cols, act, p.has_res) &&
                  ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
  // weight-gradient operands: feature-major fp16 hi/lo planes in blocked-K order (tiles of 64 fibres, each
  // [features][64] contiguous): a warp's 2-byte stores of one feature are one 64-byte run
  for (int t = 0; t < 4; ++t) {
    bp.op[t][0] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256);
    bp.op[t][1] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256 + align256((size_t)feats[t] * bp.ld * 2));
  }
  auto wgrads = [&]() -> int {
    int fused = 0;
    if (int rc = cube_wgrad_fused(op_x, op_h, op_gz, op_gpre, a_in, a_hid, a_out, R, gw1, gw2, wres ? gwres : nullptr, st, &fused)) return rc;
    if (fused) return 0;
    if (int rc = mimrl_gemm_split_blocked_acc(op_gpre, op_x, a_hid, a_in, (int)R, gw1, stream)) return rc;
    if (int rc = mimrl_gemm_split_blocked_acc(op_gz, op_h, a_out, a_hid, (int)R, gw2, stream)) return rc;
    if (wres)
      if (int rc = mimrl_gemm_split_blocked_acc(op_gz, op_x, a_out, a_in, (int)R, gwres, stream)) return rc;
    return 0;
  };
  if (c2 || c3) {
    int handled = 0;
    if (c2) {
      int r1, r2, rr;
      cube2_box_rows(a_in, a_hid, a_out, &r1, &r2, &rr);
      if (cube_weight_maps(w, a_in, a_hid, a_out, wres != nullptr, r1, r2, rr, maps)) return 1;
      if (int rc = cube2_bwd(maps, bp, st, &handled)) return rc;
    } else {
      if (cube_weight_maps(w, a_in, a_hid, a_out, wres != nullptr, 128, 128, 128, maps)) return 1;
      if (int rc = cube3_bwd(maps, bp, st, &handled)) return rc;
    }
    return wgrads();
  }
  // general kernel
  if (cube_weight_maps(w, a_in, a_hid, a_out, wres != nullptr, 128, 128, 128, maps)) return 1;
  const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCubeBwdSmem);
    kern<<<blocks, kCubeBwdThreads, kCubeBwdSmem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], bp);
  };
  if (act == 0) launch(cubemlp_tc_bwd_kernel<0>);
  else if (act == 1) launch(cubemlp_tc_bwd_kernel<1>);
  else launch(cubemlp_tc_bwd_kernel<2>);
  if (check_launch("cubemlp_tc_bwd")) return 1;
  return wgrads();
}
