// CubeMLP strided-axis mix (the sequence mix L of MLPProcess.py:94-104, x viewed as [outer, A, inner] with
// inner = K*D a multiple of 128), forward and backward, specialised at COMPILE TIME for the shapes of the reference
// configuration (README block 1: 100 -> 50 -> 50, block 2: 50 -> 10 -> 10, gelu, residual projection):
//
//     y = LayerNorm_{A'}( W2 gelu(W1 x + b1) + b2 + Wres x )              per fibre x[A]
//
// Why a second kernel family next to cubemlp_tc.cu.  The general kernel keeps one 128-fibre tile per SM in flight and
// walks it through five dependent phases (load, MMA, activation, MMA, LayerNorm); at these small widths every phase
// is a latency (global load, tcgen05 round trip, barrier), so the SM idles: 10-12 us per tile, issue slots 12-20 %
// busy (profiles/cubemlp_kernels_r2.md).  Here
//   * every size is a template constant: no index division, no per-element predicates, immediate address offsets, the
//     padding of the MMA shapes is elided at compile time;
//   * a CTA is ONE warpgroup (thread = fibre = TMEM lane; the whole fibre lives in that thread's registers, so the
//     LayerNorm and its backward are thread-local: no shared-memory exchange, no named barriers) plus one control warp
//     (weights by TMA once, then the MMA issue loop);
//   * TMEM is packed: regions are reused as soon as their contents are dead (X -> H | o -> GZ | GPRE, and
//     pre | r -> gh -> gx), the two products of dL/dx share one accumulator (the preparation kernel ties their operand
//     scales), so a block-1 tile needs 256 columns and a block-2 tile 128;
//   * weights are stored un-padded (80 KB / 12 KB instead of 192 KB);
// so 2 (block 1) or 4 (block 2) CTAs are resident per SM and hide each other's latencies.
// Arithmetic is unchanged: fp16 hi/lo split operands, three products per contraction, fp32 accumulation in TMEM.
#include "cubemlp_tc.cuh"

namespace mimrl {
namespace {

template <int A_, int H_, int Q_, int INNER_>
struct C2 {
  static constexpr int A = A_, H = H_, Q = Q_, INNER = INNER_;
  static constexpr int PA = (A + 15) & ~15, PH = (H + 15) & ~15, PQ = (Q + 15) & ~15;
  static constexpr int PS = PH > PQ ? PH : PQ;                    // slot of a hidden- or output-wide tensor
  static constexpr int RA = PA > 2 * PS ? PA : 2 * PS;            // region A: X, later H | o, later GZ | GPRE
  static constexpr int B0 = RA;                                   // region B: pre | r, later gh (over r), later gx
  static constexpr int FWD_COLS = RA + 2 * PS, BWD_COLS = 2 * RA;
  static constexpr int KB_A = (PA + 63) / 64, KB_H = (PH + 63) / 64;
  static constexpr uint32_t W1_HALF = KB_A * PH * 128, WR_HALF = KB_A * PQ * 128, W2_HALF = KB_H * PQ * 128;
  static constexpr uint32_t W_BYTES = 2 * (W1_HALF + WR_HALF + W2_HALF);
  static constexpr uint32_t OFF_WR = 2 * W1_HALF, OFF_W2 = OFF_WR + 2 * WR_HALF, OFF_BARS = W_BYTES;
  static constexpr uint32_t OFF_VEC = OFF_BARS + 256;             // b1 [PH] | b2 [PQ] | ln_w [PQ] | ln_b [PQ] | reduce [4][32]
  static constexpr uint32_t SMEM = OFF_VEC + (PH + 3 * PQ) * 4 + 4 * 32 * 4 * 8 + 1024;
  static constexpr int TPO = INNER / 128;                         // tiles per outer index
  static constexpr int pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }
  static constexpr int FWD_ALLOC = pow2(FWD_COLS), BWD_ALLOC = pow2(BWD_COLS);
  static constexpr int cap(int alloc) {
    int by_tmem = 512 / alloc, by_smem = (int)((227u * 1024u) / SMEM);
    int c = by_tmem < by_smem ? by_tmem : by_smem;
    return c > 4 ? 4 : (c < 1 ? 1 : c);
  }
  static constexpr int FWD_CTAS = cap(FWD_ALLOC), BWD_CTAS = cap(BWD_ALLOC);
  static_assert(INNER % 128 == 0 && PA <= 128 && PS <= 64, "shape outside the strided-mix specialisation");
};

constexpr int kC2Threads = 160;      // warps 0-3: compute (thread = fibre = TMEM lane), warp 4: TMA + MMA issue

// ---------------------------------------------------------------------------------------------------------- forward
template <class C>
__global__ void __launch_bounds__(kC2Threads, C::FWD_CTAS)
cube2_fwd_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                 const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3,
                 const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m5, const CubeTcParams p) {
  constexpr int A = C::A, H = C::H, Q = C::Q, INNER = C::INNER, PA = C::PA, PH = C::PH, PQ = C::PQ, PS = C::PS, B0 = C::B0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + C::OFF_BARS;
  const uint32_t bX = bars + 8, bD1 = bars + 16, bH = bars + 24, bD2 = bars + 32;
  const float *s_b1 = reinterpret_cast<const float *>(gen + C::OFF_VEC), *s_b2 = s_b1 + PH, *s_lw = s_b2 + PQ, *s_lb = s_lw + PQ;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  {
    float *vec = reinterpret_cast<float *>(gen + C::OFF_VEC);
    for (int t = threadIdx.x; t < PH + 3 * PQ; t += blockDim.x) {
      float v = 0.f;
      if (t < PH) v = (p.b1 && t < H) ? p.b1[t] : 0.f;
      else if (t < PH + PQ) v = (p.b2 && t - PH < Q) ? p.b2[t - PH] : 0.f;
      else if (t < PH + 2 * PQ) v = t - PH - PQ < Q ? p.ln_w[t - PH - PQ] : 0.f;
      else v = t - PH - 2 * PQ < Q ? p.ln_b[t - PH - 2 * PQ] : 0.f;
      vec[t] = v;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bX, 4), mbar_init(bD1, 1), mbar_init(bH, 4), mbar_init(bD2, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(gen + C::OFF_BARS + 128), C::FWD_ALLOC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(gen + C::OFF_BARS + 128), 0);
  const long long n_tiles = p.n_cols / 128;

  if (warp == 4) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx(bars, C::W_BYTES);
      for (int half = 0; half < 2; ++half)
        for (int kb = 0; kb < C::KB_A; ++kb) {
          tma_load_2d(base + half * C::W1_HALF + kb * PH * 128, half ? &m1 : &m0, bars, kb * 64, 0);
          tma_load_2d(base + C::OFF_WR + half * C::WR_HALF + kb * PQ * 128, half ? &m5 : &m4, bars, kb * 64, 0);
        }
      for (int half = 0; half < 2; ++half)
        for (int kb = 0; kb < C::KB_H; ++kb)
          tma_load_2d(base + C::OFF_W2 + half * C::W2_HALF + kb * PQ * 128, half ? &m3 : &m2, bars, kb * 64, 0);
    }
    mbar_wait(bars, 0);
    constexpr uint32_t id_h = instr_desc_f16(128, PH), id_q = instr_desc_f16(128, PQ);
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bX, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_k(tmem_base + B0, tmem_base, PA / 2, base, C::W1_HALF, PH, PA / 16, id_h, 0);                      // pre
        c2_mma_k(tmem_base + B0 + PS, tmem_base, PA / 2, base + C::OFF_WR, C::WR_HALF, PQ, PA / 16, id_q, 0);     // r
        umma_commit(bD1);
      }
      __syncwarp();
      mbar_wait(bH, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_k(tmem_base + PS, tmem_base, PH / 2, base + C::OFF_W2, C::W2_HALF, PQ, PH / 16, id_q, 0);         // o
        umma_commit(bD2);
      }
      __syncwarp();
    }
  } else {
    const int r = threadIdx.x;
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    float sx = p.scales[0], sh = p.scales[1];
    const float sw1 = scale_from_absmax(p.sc_w1[0]), sw2 = scale_from_absmax(p.sc_w2[0]), swr = scale_from_absmax(p.sc_wr[0]);
    float i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr);
    // fibre_scale: |h| <= max|x_fibre| * max_h sum_a |W1[h,a]| + max|b1| (the preparation kernel left both in the tail)
    const bool fibre_scale = (p.fibre_scale & 1) != 0;
    const float row1_max = fibre_scale ? __uint_as_float(p.absmax[12]) : 0.f, b1_max = fibre_scale ? __uint_as_float(p.absmax[14]) : 0.f;
    float rstd_max = 0.f, x_max = 0.f;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long o = tile / C::TPO;
      const int i0 = (int)(tile - o * C::TPO) * 128 + r;
      const float *xf = p.x + (size_t)o * A * INNER + i0;
      float *yf = p.y + (size_t)o * Q * INNER + i0;
      // ---- 1. fibre -> X operand (fp16 hi / lo)
      if (!fibre_scale) {
#pragma unroll
        for (int u = 0; u < PA / 16; ++u) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = u * 16 + j < A ? __ldg(xf + (size_t)(u * 16 + j) * INNER) * sx : 0.f;
          uint32_t hi[8], lo[8];
          split16(v, hi, lo);
          tmem_st8(tb + u * 8, hi);
          tmem_st8(tb + PA / 2 + u * 8, lo);
        }
      } else if constexpr (PH + PS >= PA) {
        // the scale needs the fibre's max first: the raw values are parked in the (idle) pre | r columns of this
        // thread's TMEM lane, sixteen at a time as they arrive, and read back once the scale is known
        float m = 0.f;
#pragma unroll
        for (int u = 0; u < PA / 16; ++u) {
          uint32_t raw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = u * 16 + j < A ? __ldg(xf + (size_t)(u * 16 + j) * INNER) : 0.f;
            m = fmaxf(m, fabsf(v));
            raw[j] = __float_as_uint(v);
          }
          tmem_st8(tb + B0 + u * 16, raw);
          tmem_st8(tb + B0 + u * 16 + 8, raw + 8);
        }
        tmem_st_wait();
        x_max = fmaxf(x_max, m);
        sx = pow2_scale_bits(m), sh = pow2_scale_bits(fmaf(m, row1_max, b1_max));
        i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr);
#pragma unroll
        for (int u = 0; u < PA / 16; ++u) {
          uint32_t d[16];
          tmem_ld16(tb + B0 + u * 16, d);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(d[j]) * sx;
          uint32_t hi[8], lo[8];
          split16(v, hi, lo);
          tmem_st8(tb + u * 8, hi);
          tmem_st8(tb + PA / 2 + u * 8, lo);
        }
      } else {
        float xv[PA];
#pragma unroll
        for (int j = 0; j < PA; ++j) xv[j] = j < A ? __ldg(xf + (size_t)j * INNER) : 0.f;
        float m = 0.f;
#pragma unroll
        for (int j = 0; j < A; ++j) m = fmaxf(m, fabsf(xv[j]));
        x_max = fmaxf(x_max, m);
        sx = pow2_scale_bits(m), sh = pow2_scale_bits(fmaf(m, row1_max, b1_max));
        i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr);
#pragma unroll
        for (int u = 0; u < PA / 16; ++u) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = xv[u * 16 + j] * sx;
          uint32_t hi[8], lo[8];
          split16(v, hi, lo);
          tmem_st8(tb + u * 8, hi);
          tmem_st8(tb + PA / 2 + u * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bX);
      if (p.fibre_scale & 2) {          // next tile's fibre rows towards L2 while this tile computes
        const long long nt = tile + gridDim.x;
        if (nt < n_tiles) {
          const long long no = nt / C::TPO;
          const float *nx = p.x + (size_t)no * A * INNER + (int)(nt - no * C::TPO) * 128 + r;
#pragma unroll 4
          for (int j = 0; j < A; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)j * INNER));
        }
      }
      // ---- 2. h = gelu(pre + b1) -> H operand over the dead X columns
      mbar_wait(bD1, ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < PH / 16; ++u) {
        uint32_t d[16];
        tmem_ld16(tb + B0 + u * 16, d);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int h = u * 16 + j;
          float2 g = make_float2(0.f, 0.f);
          if (h < H) {
            const float2 pre = ffma2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), make_float2(i_pre, i_pre),
                                     make_float2(s_b1[h], s_b1[h + 1]));
            g = fmul2(gelu2(pre), make_float2(sh, sh));
            if (h + 1 >= H) g.y = 0.f;
          }
          v[j] = g.x, v[j + 1] = g.y;
        }
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + u * 8, hi);
        tmem_st8(tb + PH / 2 + u * 8, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bH);
      // ---- 3. z = o + r + b2, LayerNorm over A' in this thread's registers, one write
      mbar_wait(bD2, ph);
      tc_fence_after();
      float z[PQ];
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < PQ / 16; ++u) {
        uint32_t d[16], w[16];
        tmem_ld16(tb + PS + u * 16, d);
        tmem_ld16(tb + B0 + PS + u * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int q = u * 16 + j;
          float t = 0.f;
          if (q < Q) t = fmaf(__uint_as_float(w[j]), i_r, fmaf(__uint_as_float(d[j]), i_o, s_b2[q]));
          z[q] = t;
          sum += t;
        }
      }
      tc_fence_before();
      const float mean = sum * (1.f / Q);
      float var = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float dlt = z[q] - mean;
        var = fmaf(dlt, dlt, var);
      }
      const float rstd = rsqrtf(var * (1.f / Q) + 1e-6f);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float g = rstd * s_lw[q];
        yf[(size_t)q * INNER] = fmaf(z[q], g, fmaf(-mean, g, s_lb[q]));
      }
      reinterpret_cast<float2 *>(p.saved)[tile * 128 + r] = make_float2(mean, rstd);
      rstd_max = fmaxf(rstd_max, rstd);
    }
    for (int o = 16; o; o >>= 1) {
      rstd_max = fmaxf(rstd_max, __shfl_xor_sync(0xffffffffu, rstd_max, o));
      x_max = fmaxf(x_max, __shfl_xor_sync(0xffffffffu, x_max, o));
    }
    if (lane == 0) atomicMax(p.absmax + 2, __float_as_uint(rstd_max));
    if (lane == 0 && fibre_scale) atomicMax(p.absmax, __float_as_uint(x_max));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, C::FWD_ALLOC);
}

// --------------------------------------------------------------------------------------------------------- backward
template <class C>
__global__ void __launch_bounds__(kC2Threads, C::BWD_CTAS)
cube2_bwd_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                 const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3,
                 const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m5, const CubeBwdParams bp) {
  constexpr int A = C::A, H = C::H, Q = C::Q, INNER = C::INNER, PA = C::PA, PH = C::PH, PQ = C::PQ, PS = C::PS, B0 = C::B0;
  const CubeTcParams &p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + C::OFF_BARS;
  const uint32_t bX = bars + 8, bD1 = bars + 16, bH = bars + 24, bD2 = bars + 32, bGz = bars + 40, bD3 = bars + 48, bGp = bars + 56,
                 bD4 = bars + 64;
  float *vec = reinterpret_cast<float *>(gen + C::OFF_VEC);
  const float *s_b1 = vec, *s_b2 = s_b1 + PH, *s_lw = s_b2 + PQ;
  float *s_red = vec + PH + 3 * PQ;                                // [4 warps][8 quantities][32]
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < PH + 3 * PQ; t += blockDim.x) {
    float v = 0.f;
    if (t < PH) v = (p.b1 && t < H) ? p.b1[t] : 0.f;
    else if (t < PH + PQ) v = (p.b2 && t - PH < Q) ? p.b2[t - PH] : 0.f;
    else if (t < PH + 2 * PQ) v = t - PH - PQ < Q ? p.ln_w[t - PH - PQ] : 0.f;
    vec[t] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bX, 4), mbar_init(bD1, 1), mbar_init(bH, 4), mbar_init(bD2, 1);
    mbar_init(bGz, 4), mbar_init(bD3, 1), mbar_init(bGp, 4), mbar_init(bD4, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(gen + C::OFF_BARS + 128), C::BWD_ALLOC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(gen + C::OFF_BARS + 128), 0);
  const long long n_tiles = p.n_cols / 128;

  if (warp == 4) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx(bars, C::W_BYTES);
      for (int half = 0; half < 2; ++half)
        for (int kb = 0; kb < C::KB_A; ++kb) {
          tma_load_2d(base + half * C::W1_HALF + kb * PH * 128, half ? &m1 : &m0, bars, kb * 64, 0);
          tma_load_2d(base + C::OFF_WR + half * C::WR_HALF + kb * PQ * 128, half ? &m5 : &m4, bars, kb * 64, 0);
        }
      for (int half = 0; half < 2; ++half)
        for (int kb = 0; kb < C::KB_H; ++kb)
          tma_load_2d(base + C::OFF_W2 + half * C::W2_HALF + kb * PQ * 128, half ? &m3 : &m2, bars, kb * 64, 0);
    }
    mbar_wait(bars, 0);
    constexpr uint32_t id_h = instr_desc_f16(128, PH), id_q = instr_desc_f16(128, PQ);
    constexpr uint32_t id_h_mn = instr_desc_f16_bmn(128, PH), id_a_mn = instr_desc_f16_bmn(128, PA);
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bX, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_k(tmem_base + B0, tmem_base, PA / 2, base, C::W1_HALF, PH, PA / 16, id_h, 0);                      // pre
        c2_mma_k(tmem_base + B0 + PS, tmem_base, PA / 2, base + C::OFF_WR, C::WR_HALF, PQ, PA / 16, id_q, 0);     // r
        umma_commit(bD1);
      }
      __syncwarp();
      mbar_wait(bH, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_k(tmem_base + PS, tmem_base, PH / 2, base + C::OFF_W2, C::W2_HALF, PQ, PH / 16, id_q, 0);         // o
        umma_commit(bD2);
      }
      __syncwarp();
      mbar_wait(bGz, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_mn(tmem_base + B0 + PS, tmem_base, PQ / 2, base + C::OFF_W2, C::W2_HALF, PQ, PQ / 16, id_h_mn, 0);   // gh = GZ W2
        umma_commit(bD3);
      }
      __syncwarp();
      mbar_wait(bGp, ph);
      tc_fence_after();
      if (leader) {
        c2_mma_mn(tmem_base + B0, tmem_base + PS, PH / 2, base, C::W1_HALF, PH, PH / 16, id_a_mn, 0);               // gx = GPRE W1
        c2_mma_mn(tmem_base + B0, tmem_base, PQ / 2, base + C::OFF_WR, C::WR_HALF, PQ, PQ / 16, id_a_mn, 1);        //    + GZ Wres
        umma_commit(bD4);
      }
      __syncwarp();
    }
  } else {
    const int r = threadIdx.x;
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sx = bp.scales[0], sh = bp.scales[1], sgz = bp.scales[2], sgp = bp.scales[3];
    const float sw1 = scale_from_absmax(p.sc_w1[0]), sw2 = scale_from_absmax(p.sc_w2[0]), swr = scale_from_absmax(p.sc_wr[0]);
    const float i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr), i_gh = 1.f / (sgz * sw2),
                i_gx = 1.f / (sgp * sw1);            // == 1 / (sgz * swr): the preparation kernel ties the two scales
    constexpr int NQ = (PQ + 31) / 32, NH = (PH + 31) / 32;          // 32-feature chunks of the parameter-gradient sums
    float acc_lnw[NQ] = {}, acc_lnb[NQ] = {}, acc_b2[NQ] = {}, acc_b1[NH] = {};
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long o = tile / C::TPO;
      const int i0 = (int)(tile - o * C::TPO) * 128 + r;
      const float *xf = p.x + (size_t)o * A * INNER + i0;
      const float *gyf = bp.gy + (size_t)o * Q * INNER + i0;
      float *gxf = bp.gx + (size_t)o * A * INNER + i0;
      const size_t row = (size_t)tile * 128 + r;
      // ---- 1. x -> X operand (+ the feature-major copy for the weight gradients)
#pragma unroll
      for (int u = 0; u < PA / 16; ++u) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = u * 16 + j < A ? __ldg(xf + (size_t)(u * 16 + j) * INNER) * sx : 0.f;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + u * 8, hi);
        tmem_st8(tb + PA / 2 + u * 8, lo);
        c2_store_op<A>(bp.op[0][0], bp.op[0][1], row, u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bX);
      // the incoming gradient of this fibre, fetched while the first products run
      float gyv[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) gyv[q] = __ldg(gyf + (size_t)q * INNER);
      const float2 ms = __ldg(reinterpret_cast<const float2 *>(p.saved) + row);
      const float mean = ms.x, rstd = ms.y;
      // ---- 2. h = gelu(pre + b1) -> H operand
      mbar_wait(bD1, ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < PH / 16; ++u) {
        uint32_t d[16];
        tmem_ld16(tb + B0 + u * 16, d);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int h = u * 16 + j;
          float2 g = make_float2(0.f, 0.f);
          if (h < H) {
            const float2 pre = ffma2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), make_float2(i_pre, i_pre),
                                     make_float2(s_b1[h], s_b1[h + 1]));
            g = fmul2(gelu2(pre), make_float2(sh, sh));
            if (h + 1 >= H) g.y = 0.f;
          }
          v[j] = g.x, v[j + 1] = g.y;
        }
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + u * 8, hi);
        tmem_st8(tb + PH / 2 + u * 8, lo);
        c2_store_op<H>(bp.op[1][0], bp.op[1][1], row, u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bH);
      // ---- 3. z = o + r + b2, LayerNorm backward in registers -> gz -> GZ operand
      mbar_wait(bD2, ph);
      tc_fence_after();
      float zh[PQ];
      float sum_g = 0.f, sum_gz = 0.f;
#pragma unroll
      for (int u = 0; u < PQ / 16; ++u) {
        uint32_t d[16], w[16];
        tmem_ld16(tb + PS + u * 16, d);
        tmem_ld16(tb + B0 + PS + u * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int q = u * 16 + j;
          float t = 0.f;
          if (q < Q) {
            const float z = fmaf(__uint_as_float(w[j]), i_r, fmaf(__uint_as_float(d[j]), i_o, s_b2[q]));
            t = (z - mean) * rstd;
            const float gw = gyv[q] * s_lw[q];
            sum_g += gw;
            sum_gz = fmaf(gw, t, sum_gz);
          }
          zh[q] = t;
        }
      }
      // LayerNorm parameter gradients: sums over fibres of gy zhat and gy
#pragma unroll
      for (int c = 0; c < NQ; ++c) {
        constexpr int W = PQ >= 32 ? 32 : 16;
        float t1[W], t2[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const int q = c * 32 + j;
          t1[j] = q < Q ? gyv[q < Q ? q : 0] * zh[q < PQ ? q : 0] : 0.f;
          t2[j] = q < Q ? gyv[q < Q ? q : 0] : 0.f;
        }
        acc_lnw[c] += lane_sum<W>(t1, lane);
        acc_lnb[c] += lane_sum<W>(t2, lane);
      }
      const float m1 = sum_g * (1.f / Q), m2 = sum_gz * (1.f / Q);
#pragma unroll
      for (int q = 0; q < Q; ++q) zh[q] = rstd * (fmaf(gyv[q], s_lw[q], -m1) - zh[q] * m2);       // zh now holds gz
#pragma unroll
      for (int c = 0; c < NQ; ++c) {
        constexpr int W = PQ >= 32 ? 32 : 16;
        float t1[W];
#pragma unroll
        for (int j = 0; j < W; ++j) t1[j] = c * 32 + j < Q ? zh[c * 32 + j < PQ ? c * 32 + j : 0] : 0.f;
        acc_b2[c] += lane_sum<W>(t1, lane);
      }
#pragma unroll
      for (int u = 0; u < PQ / 16; ++u) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = u * 16 + j < Q ? zh[u * 16 + j] * sgz : 0.f;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + u * 8, hi);
        tmem_st8(tb + PQ / 2 + u * 8, lo);
        c2_store_op<Q>(bp.op[2][0], bp.op[2][1], row, u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bGz);
      // ---- 4. gpre = gh gelu'(pre) -> GPRE operand
      mbar_wait(bD3, ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < PH / 16; ++u) {
        uint32_t d[16], w[16];
        tmem_ld16(tb + B0 + PS + u * 16, d);
        tmem_ld16(tb + B0 + u * 16, w);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int h = u * 16 + j;
          float2 g = make_float2(0.f, 0.f);
          if (h < H) {
            const float2 pre = ffma2(make_float2(__uint_as_float(w[j]), __uint_as_float(w[j + 1])), make_float2(i_pre, i_pre),
                                     make_float2(s_b1[h], s_b1[h + 1]));
            g = fmul2(fmul2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), make_float2(i_gh, i_gh)), gelu_bwd2(pre));
            if (h + 1 >= H) g.y = 0.f;
          }
          v[j] = g.x, v[j + 1] = g.y;
        }
        {                                         // bias gradient of the first layer
          constexpr int W = 16;
          float t1[W];
#pragma unroll
          for (int j = 0; j < W; ++j) t1[j] = v[j];
          const float s = lane_sum<W>(t1, lane);   // lanes t and t + 16 hold feature u * 16 + t % 16
          if (((u & 1) != 0) == (lane >= 16)) acc_b1[u >> 1] += s;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= sgp;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + PS + u * 8, hi);
        tmem_st8(tb + PS + PH / 2 + u * 8, lo);
        c2_store_op<H>(bp.op[3][0], bp.op[3][1], row, u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bGp);
      // ---- 5. gx = gpre W1 + gz Wres
      mbar_wait(bD4, ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < PA / 16; ++u) {
        uint32_t d[16];
        tmem_ld16(tb + B0 + u * 16, d);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (u * 16 + j < A) gxf[(size_t)(u * 16 + j) * INNER] = __uint_as_float(d[j]) * i_gx;
      }
      tc_fence_before();
    }
    // parameter-gradient sums: lanes -> shared memory -> one atomic per CTA and feature
    float *mine = s_red + warp * 8 * 32;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      mine[(0 + c) * 32 + lane] = c < NQ ? acc_lnw[c < NQ ? c : 0] : 0.f;
      mine[(2 + c) * 32 + lane] = c < NQ ? acc_lnb[c < NQ ? c : 0] : 0.f;
      mine[(4 + c) * 32 + lane] = c < NQ ? acc_b2[c < NQ ? c : 0] : 0.f;
      mine[(6 + c) * 32 + lane] = c < NH ? acc_b1[c < NH ? c : 0] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      const int slot = threadIdx.x;                                 // 128 threads: 4 quantities x 32 lanes, two chunks each
      const int quant = slot >> 5, ln = slot & 31;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += s_red[w * 8 * 32 + (quant * 2 + c) * 32 + ln];
        // N = 16 butterflies leave feature t in lanes t and t + 16: take lanes < 16 (b1 uses both halves, see above)
        int f;
        bool ok;
        if (quant == 3) {
          f = c * 32 + ln;
          ok = c < NH && f < H;
        } else {
          f = c * 32 + ln;
          ok = c < NQ && f < Q && (PQ >= 32 || ln < 16);
        }
        if (ok) {
          float *dst = quant == 0 ? bp.g_lnw : quant == 1 ? bp.g_lnb : quant == 2 ? bp.g_b2 : bp.g_b1;
          if (dst) atomicAdd(dst + f, s);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, C::BWD_ALLOC);
}

using CfgL1 = C2<100, 50, 50, 384>;
using CfgL2 = C2<50, 10, 10, 384>;

template <class C>
int launch_fwd(const CUtensorMap *m, const CubeTcParams &p, cudaStream_t st) {
  const long long n_tiles = p.n_cols / 128;
  const int cap = 148 * C::FWD_CTAS;
  const int blocks = (int)(n_tiles < cap ? n_tiles : cap);
  cudaFuncSetAttribute(cube2_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  cube2_fwd_kernel<C><<<blocks, kC2Threads, C::SMEM, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], p);
  return check_launch("cube2_fwd");
}
template <class C>
int launch_bwd(const CUtensorMap *m, const CubeBwdParams &bp, cudaStream_t st) {
  const long long n_tiles = bp.f.n_cols / 128;
  const int cap = 148 * C::BWD_CTAS;
  const int blocks = (int)(n_tiles < cap ? n_tiles : cap);
  cudaFuncSetAttribute(cube2_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  cube2_bwd_kernel<C><<<blocks, kC2Threads, C::SMEM, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], bp);
  return check_launch("cube2_bwd");
}

int c2_which(int a_in, int a_hid, int a_out, int inner) {
  if (inner != 384) return 0;
  if (a_in == 100 && a_hid == 50 && a_out == 50) return 1;
  if (a_in == 50 && a_hid == 10 && a_out == 10) return 2;
  return 0;
}

}  // namespace

bool cube2_supported(int a_in, int a_hid, int a_out, int inner, long long n_cols, int act, int has_res) {
  if (getenv("MIMRL_CUBE2_OFF")) return false;
  return c2_which(a_in, a_hid, a_out, inner) != 0 && n_cols % 128 == 0 && act == 0 && has_res;
}

void cube2_box_rows(int a_in, int a_hid, int a_out, int *rows_w1, int *rows_w2, int *rows_wr) {
  const int w = c2_which(a_in, a_hid, a_out, 384);
  *rows_w1 = w == 1 ? CfgL1::PH : w == 2 ? CfgL2::PH : 0;
  *rows_w2 = *rows_wr = w == 1 ? CfgL1::PQ : w == 2 ? CfgL2::PQ : 0;
}

int cube2_fwd(const CUtensorMap *maps, const CubeTcParams &p, cudaStream_t st, int *handled) {
  const int w = c2_which(p.A, p.H, p.A2, p.inner);
  *handled = w != 0;
  if (w == 1) return launch_fwd<CfgL1>(maps, p, st);
  if (w == 2) return launch_fwd<CfgL2>(maps, p, st);
  return 0;
}

int cube2_bwd(const CUtensorMap *maps, const CubeBwdParams &bp, cudaStream_t st, int *handled) {
  const int w = c2_which(bp.f.A, bp.f.H, bp.f.A2, bp.f.inner);
  *handled = w != 0;
  if (w == 1) return launch_bwd<CfgL1>(maps, bp, st);
  if (w == 2) return launch_bwd<CfgL2>(maps, bp, st);
  return 0;
}

}  // namespace mimrl
