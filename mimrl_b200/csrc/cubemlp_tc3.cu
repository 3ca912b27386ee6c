// CubeMLP channel mix D of the reference configuration (MLPProcess.py:114-120 with d = 128: x [rows, 128] ->
// [rows, 128], hidden 128, gelu, residual projection), forward and backward, specialised at compile time:
//
//     y = LayerNorm_128( W2 gelu(W1 x + b1) + b2 + Wres x )              per row
//
// 128-wide operands leave room for ONE 128-row tile per SM (three resident 128 x 128 fp16 hi/lo weights are 192 KB of
// shared memory, the tile's tensors fill the 512 TMEM columns), so the latencies are hidden INSIDE the tile instead of
// across CTAs (cubemlp_tc2.cu does the latter for the narrow sequence mix).  Two warpgroups own the feature halves
// [0, 64) and [64, 128) of every row (thread = (row, half)), and every contraction is issued in the matching halves:
//   pre = X W1^T as two N = 64 products with their own commits -- warpgroup 0 starts its gelu while the tensor core
//       still works on the second half and on r = X Wres^T;
//   o = H W2^T as two K = 64 accumulation steps, each issued as soon as that warpgroup has delivered its half of H;
//   backward: gh = GZ W2 the same way (K halves follow the gz halves).
// TMEM plan (columns): [0,128) X -> o -> gh -> gx | [128,256) H -> GZ | [256,384) pre | [384,512) r -> GPRE; an MMA
// that overwrites a region is issued behind the MMA that last read it (the tensor pipe executes in issue order) and
// behind a barrier all readers of that region have passed.  The two products of dL/dx share one accumulator (operand
// scales tied by the preparation kernel).  LayerNorm statistics: partial sums of the two halves meet in shared memory.
// Everything else (fp16 hi/lo operands, three products per contraction, weight-gradient operands for
// mimrl_gemm_split_blocked, parameter-gradient sums by lane butterflies) is as in cubemlp_tc.cu.
#include "cubemlp_tc.cuh"

namespace mimrl {
namespace {

// G warpgroups share a row: thread = (row, feature slice of 128 / G); warp 4 G: TMA + MMA issue
template <int G> struct C3 {
  static constexpr int FT = 128 / G, NU = FT / 16, NC = FT / 32, THREADS = 128 * G + 32;
  static_assert(G == 2 || G == 4, "two or four warpgroups");
};
constexpr uint32_t kC3Blk = 128 * 128;                 // one 64-wide K block of a weight: 128 rows x 128 B
constexpr uint32_t kC3Half = 2 * kC3Blk;               // hi or lo
constexpr uint32_t kC3W = 2 * kC3Half;                 // one weight
constexpr uint32_t kC3Bars = 3 * kC3W;
constexpr uint32_t kC3Vec = kC3Bars + 256;             // b1 | b2 | ln_w | ln_b (4 x 128 floats)
constexpr uint32_t kC3Part = kC3Vec + 4 * 128 * 4;     // [2 rounds][G slices][128 rows] float2
constexpr uint32_t kC3Red = kC3Part + 2 * 4 * 128 * 8; // [4 G warps][8][32] floats
constexpr uint32_t kC3Smem = kC3Red + 16 * 8 * 32 * 4 + 1024;
constexpr uint32_t kTX3 = 0, kTH3 = 128, kTPre3 = 256, kTR3 = 384;      // o, gh, gx reuse kTX3; GZ kTH3; GPRE kTR3

// barrier indices
enum { B_W = 0, B_X, B_D1A, B_D1B, B_HA, B_HB, B_D2, B_GZA, B_GZB, B_D3, B_GP, B_D4, B_COUNT };

template <int G>
__device__ __forceinline__ uint32_t c3_setup(uint8_t *gen, uint32_t base, const CubeTcParams &p, int warp) {
  const uint32_t bars = base + kC3Bars;
  float *vec = reinterpret_cast<float *>(gen + kC3Vec);
  for (int t = threadIdx.x; t < 512; t += blockDim.x) {
    const int f = t & 127, w = t >> 7;
    vec[t] = w == 0 ? (p.b1 ? p.b1[f] : 0.f) : w == 1 ? (p.b2 ? p.b2[f] : 0.f) : w == 2 ? p.ln_w[f] : (p.ln_b ? p.ln_b[f] : 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(bars + 8 * B_W, 1);
    mbar_init(bars + 8 * B_X, 4 * G);
    mbar_init(bars + 8 * B_D1A, 1), mbar_init(bars + 8 * B_D1B, 1);
    mbar_init(bars + 8 * B_HA, 2 * G), mbar_init(bars + 8 * B_HB, 2 * G);
    mbar_init(bars + 8 * B_D2, 1);
    mbar_init(bars + 8 * B_GZA, 2 * G), mbar_init(bars + 8 * B_GZB, 2 * G);
    mbar_init(bars + 8 * B_D3, 1);
    mbar_init(bars + 8 * B_GP, 4 * G);
    mbar_init(bars + 8 * B_D4, 1);
    fence_barrier_init();
  }
  if (warp == 4 * G) {
    tmem_alloc(smem_u32(gen + kC3Bars + 128), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(gen + kC3Bars + 128), 0);
}

// three-product sets of the issue warp ------------------------------------------------------------------------
// D[128 x n] (+)= A[:, k0 .. k0 + 16 ks) . W[row0 .. row0 + n, same k]^T   (K-major weight tile)
__device__ __forceinline__ void c3_mma_k(uint32_t d, uint32_t a, int k0, int ks, uint32_t sw, int row0, uint32_t idesc, uint32_t acc) {
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_off = prod == 2 ? 64 : 0, b_off = prod == 1 ? kC3Half : 0;
    for (int k = k0; k < k0 + ks; ++k) {
      umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128(sw + b_off + (k >> 2) * kC3Blk + row0 * 128 + (k & 3) * 32), idesc, acc);
      acc = 1;
    }
  }
}
// D[128 x 128] (+)= A[:, k0 .. k0 + 16 ks) . W[rows k0*16 .., :]   (contraction over the ROWS of the same tile)
__device__ __forceinline__ void c3_mma_mn(uint32_t d, uint32_t a, int k0, int ks, uint32_t sw, uint32_t idesc, uint32_t acc) {
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_off = prod == 2 ? 64 : 0, b_off = prod == 1 ? kC3Half : 0;
    for (int k = k0; k < k0 + ks; ++k) {
      umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128_mn(sw + b_off + k * 2048, kC3Blk, 1024), idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ void c3_load_weights(uint32_t base, uint32_t bar, const CUtensorMap *w1h, const CUtensorMap *w1l,
                                                const CUtensorMap *w2h, const CUtensorMap *w2l, const CUtensorMap *wrh,
                                                const CUtensorMap *wrl) {
  mbar_expect_tx(bar, 3 * kC3W);
  const CUtensorMap *maps[6] = {w1h, w1l, wrh, wrl, w2h, w2l};            // shared-memory order: W1 | Wres | W2
  for (int m = 0; m < 3; ++m)
    for (int half = 0; half < 2; ++half)
      for (int kb = 0; kb < 2; ++kb)
        tma_load_2d(base + m * kC3W + half * kC3Half + kb * kC3Blk, maps[m * 2 + half], bar, kb * 64, 0);
}

// 64 features of a row: 16 x 16-byte loads
__device__ __forceinline__ void c3_load16(const float *src, bool ok, float (&v)[16]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float4 q = ok ? __ldg(reinterpret_cast<const float4 *>(src) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[4 * t] = q.x, v[4 * t + 1] = q.y, v[4 * t + 2] = q.z, v[4 * t + 3] = q.w;
  }
}

// gelu(pre + b1) of one 16-feature unit, scaled and split, into the H operand
__device__ __forceinline__ void c3_hidden_unit(uint32_t tb, int fb, int u, const float *s_b1, float i_pre, float sh, uint32_t (&hi)[8],
                                               uint32_t (&lo)[8]) {
  uint32_t d[16];
  tmem_ld16(tb + kTPre3 + fb + u * 16, d);
  tmem_ld_wait();
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const int h = fb + u * 16 + j;
    const float2 pre = ffma2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), make_float2(i_pre, i_pre),
                             make_float2(s_b1[h], s_b1[h + 1]));
    const float2 gl = fmul2(gelu2(pre), make_float2(sh, sh));
    v[j] = gl.x, v[j + 1] = gl.y;
  }
  split16(v, hi, lo);
  tmem_st8(tb + kTH3 + fb / 2 + u * 8, hi);
  tmem_st8(tb + kTH3 + 64 + fb / 2 + u * 8, lo);
}

// ---------------------------------------------------------------------------------------------------------- forward
template <int G>
__global__ void __launch_bounds__(C3<G>::THREADS, 1)
cube3_fwd_kernel(const __grid_constant__ CUtensorMap w1h, const __grid_constant__ CUtensorMap w1l,
                 const __grid_constant__ CUtensorMap w2h, const __grid_constant__ CUtensorMap w2l,
                 const __grid_constant__ CUtensorMap wrh, const __grid_constant__ CUtensorMap wrl, const CubeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kC3Bars;
  const uint32_t sW1 = base, sWr = base + kC3W, sW2 = base + 2 * kC3W;
  const float *s_b1 = reinterpret_cast<const float *>(gen + kC3Vec), *s_b2 = s_b1 + 128, *s_lw = s_b1 + 256, *s_lb = s_b1 + 384;
  float *s_part = reinterpret_cast<float *>(gen + kC3Part);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  constexpr int FT = C3<G>::FT, NU = C3<G>::NU;
  const uint32_t tmem_base = c3_setup<G>(gen, base, p, warp);
  const long long n_tiles = (p.n_cols + 127) / 128;

  if (warp == 4 * G) {
    const uint32_t leader = elect_one();
    if (leader) c3_load_weights(base, bars + 8 * B_W, &w1h, &w1l, &w2h, &w2l, &wrh, &wrl);
    mbar_wait(bars + 8 * B_W, 0);
    constexpr uint32_t id64 = instr_desc_f16(128, 64), id128 = instr_desc_f16(128, 128);
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bars + 8 * B_X, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_k(tmem_base + kTPre3, tmem_base + kTX3, 0, 8, sW1, 0, id64, 0);
        umma_commit(bars + 8 * B_D1A);
        c3_mma_k(tmem_base + kTPre3 + 64, tmem_base + kTX3, 0, 8, sW1, 64, id64, 0);
        umma_commit(bars + 8 * B_D1B);
        c3_mma_k(tmem_base + kTR3, tmem_base + kTX3, 0, 8, sWr, 0, id128, 0);
      }
      __syncwarp();
      mbar_wait(bars + 8 * B_HA, ph);
      tc_fence_after();
      if (leader) c3_mma_k(tmem_base + kTX3, tmem_base + kTH3, 0, 4, sW2, 0, id128, 0);         // o over the dead X columns
      __syncwarp();
      mbar_wait(bars + 8 * B_HB, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_k(tmem_base + kTX3, tmem_base + kTH3, 4, 4, sW2, 0, id128, 1);
        umma_commit(bars + 8 * B_D2);
      }
      __syncwarp();
    }
  } else {
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const int fb = FT * g, hf = (2 * g) / G;            // first feature of this thread's slice; its half of the row
    const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sx = p.scales[0], sh = p.scales[1];
    const float i_pre = 1.f / (sx * scale_from_absmax(p.sc_w1[0])), i_o = 1.f / (sh * scale_from_absmax(p.sc_w2[0]));
    const float i_r = 1.f / (sx * scale_from_absmax(p.sc_wr[0]));
    float rstd_max = 0.f;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long c = tile * 128 + r;
      const bool ok = c < p.n_cols;
      const float *xr = p.x + (size_t)(ok ? c : 0) * 128 + fb;
      // ---- 1. my half of the row -> X operand
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float v[16];
        c3_load16(xr + u * 16, ok, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= sx;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + kTX3 + fb / 2 + u * 8, hi);
        tmem_st8(tb + kTX3 + 64 + fb / 2 + u * 8, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_X);
      // ---- 2. h = gelu(pre + b1), my half, as soon as its product is done
      mbar_wait(bars + 8 * (hf ? B_D1B : B_D1A), ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t hi[8], lo[8];
        c3_hidden_unit(tb, fb, u, s_b1, i_pre, sh, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (hf ? B_HB : B_HA));
      // ---- 3. z = o + r + b2; LayerNorm over the 128 features (two halves meet in shared memory)
      mbar_wait(bars + 8 * B_D2, ph);
      tc_fence_after();
      float z[FT];
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t d[16], w[16];
        tmem_ld16(tb + kTX3 + fb + u * 16, d);
        tmem_ld16(tb + kTR3 + fb + u * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = fmaf(__uint_as_float(w[j]), i_r, fmaf(__uint_as_float(d[j]), i_o, s_b2[fb + u * 16 + j]));
          z[u * 16 + j] = t;
          sum += t;
        }
      }
      tc_fence_before();
      s_part[g * 128 + r] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < G; ++w) tot += s_part[w * 128 + r];
      const float mean = tot * (1.f / 128.f);
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        const float dlt = z[j] - mean;
        var = fmaf(dlt, dlt, var);
      }
      s_part[512 + g * 128 + r] = var;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");
      float vtot = 0.f;
#pragma unroll
      for (int w = 0; w < G; ++w) vtot += s_part[512 + w * 128 + r];
      const float rstd = rsqrtf(vtot * (1.f / 128.f) + 1e-6f);
      if (ok) {
        float4 *yr = reinterpret_cast<float4 *>(p.y + (size_t)c * 128 + fb);
#pragma unroll
        for (int t = 0; t < FT / 4; ++t) {
          float o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int f = fb + 4 * t + e;
            const float gg = rstd * s_lw[f];
            o4[e] = fmaf(z[4 * t + e], gg, fmaf(-mean, gg, s_lb[f]));
          }
          yr[t] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        if (g == 0) {
          reinterpret_cast<float2 *>(p.saved)[c] = make_float2(mean, rstd);
          rstd_max = fmaxf(rstd_max, rstd);
        }
      }
    }
    if (g == 0) {
      for (int o = 16; o; o >>= 1) rstd_max = fmaxf(rstd_max, __shfl_xor_sync(0xffffffffu, rstd_max, o));
      if (lane == 0) atomicMax(p.absmax + 2, __float_as_uint(rstd_max));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * G) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------------------------------------- backward
template <int G>
__global__ void __launch_bounds__(C3<G>::THREADS, 1)
cube3_bwd_kernel(const __grid_constant__ CUtensorMap w1h, const __grid_constant__ CUtensorMap w1l,
                 const __grid_constant__ CUtensorMap w2h, const __grid_constant__ CUtensorMap w2l,
                 const __grid_constant__ CUtensorMap wrh, const __grid_constant__ CUtensorMap wrl, const CubeBwdParams bp) {
  const CubeTcParams &p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kC3Bars;
  const uint32_t sW1 = base, sWr = base + kC3W, sW2 = base + 2 * kC3W;
  const float *s_b1 = reinterpret_cast<const float *>(gen + kC3Vec), *s_b2 = s_b1 + 128, *s_lw = s_b1 + 256;
  float2 *s_part = reinterpret_cast<float2 *>(gen + kC3Part);
  float *s_red = reinterpret_cast<float *>(gen + kC3Red);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  constexpr int FT = C3<G>::FT, NU = C3<G>::NU, NC = C3<G>::NC;
  const uint32_t tmem_base = c3_setup<G>(gen, base, p, warp);
  const long long n_tiles = (p.n_cols + 127) / 128;

  if (warp == 4 * G) {
    const uint32_t leader = elect_one();
    if (leader) c3_load_weights(base, bars + 8 * B_W, &w1h, &w1l, &w2h, &w2l, &wrh, &wrl);
    mbar_wait(bars + 8 * B_W, 0);
    constexpr uint32_t id64 = instr_desc_f16(128, 64), id128 = instr_desc_f16(128, 128), id128mn = instr_desc_f16_bmn(128, 128);
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bars + 8 * B_X, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_k(tmem_base + kTPre3, tmem_base + kTX3, 0, 8, sW1, 0, id64, 0);
        umma_commit(bars + 8 * B_D1A);
        c3_mma_k(tmem_base + kTPre3 + 64, tmem_base + kTX3, 0, 8, sW1, 64, id64, 0);
        umma_commit(bars + 8 * B_D1B);
        c3_mma_k(tmem_base + kTR3, tmem_base + kTX3, 0, 8, sWr, 0, id128, 0);
      }
      __syncwarp();
      mbar_wait(bars + 8 * B_HA, ph);
      tc_fence_after();
      if (leader) c3_mma_k(tmem_base + kTX3, tmem_base + kTH3, 0, 4, sW2, 0, id128, 0);          // o
      __syncwarp();
      mbar_wait(bars + 8 * B_HB, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_k(tmem_base + kTX3, tmem_base + kTH3, 4, 4, sW2, 0, id128, 1);
        umma_commit(bars + 8 * B_D2);
      }
      __syncwarp();
      mbar_wait(bars + 8 * B_GZA, ph);                 // every thread has passed its reads of o (it needs both halves' sums)
      tc_fence_after();
      if (leader) c3_mma_mn(tmem_base + kTX3, tmem_base + kTH3, 0, 4, sW2, id128mn, 0);           // gh = GZ W2
      __syncwarp();
      mbar_wait(bars + 8 * B_GZB, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_mn(tmem_base + kTX3, tmem_base + kTH3, 4, 4, sW2, id128mn, 1);
        umma_commit(bars + 8 * B_D3);
      }
      __syncwarp();
      mbar_wait(bars + 8 * B_GP, ph);
      tc_fence_after();
      if (leader) {
        c3_mma_mn(tmem_base + kTX3, tmem_base + kTR3, 0, 8, sW1, id128mn, 0);                     // gx = GPRE W1
        c3_mma_mn(tmem_base + kTX3, tmem_base + kTH3, 0, 8, sWr, id128mn, 1);                     //    + GZ Wres
        umma_commit(bars + 8 * B_D4);
      }
      __syncwarp();
    }
  } else {
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const int fb = FT * g, hf = (2 * g) / G;            // first feature of this thread's slice; its half of the row
    const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sx = bp.scales[0], sh = bp.scales[1], sgz = bp.scales[2], sgp = bp.scales[3];
    const float sw1 = scale_from_absmax(p.sc_w1[0]), sw2 = scale_from_absmax(p.sc_w2[0]), swr = scale_from_absmax(p.sc_wr[0]);
    const float i_pre = 1.f / (sx * sw1), i_o = 1.f / (sh * sw2), i_r = 1.f / (sx * swr), i_gh = 1.f / (sgz * sw2),
                i_gx = 1.f / (sgp * sw1);            // == 1 / (sgz * swr)
    float acc_lnw[NC] = {}, acc_lnb[NC] = {}, acc_b2[NC] = {}, acc_b1[NC] = {};
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long c = tile * 128 + r;
      const bool ok = c < p.n_cols;
      const size_t row = (size_t)c;
      const float *xr = p.x + (size_t)(ok ? c : 0) * 128 + fb;
      const float *gyr = bp.gy + (size_t)(ok ? c : 0) * 128 + fb;
      // ---- 1. x -> X operand (+ weight-gradient operand)
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float v[16];
        c3_load16(xr + u * 16, ok, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= sx;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + kTX3 + fb / 2 + u * 8, hi);
        tmem_st8(tb + kTX3 + 64 + fb / 2 + u * 8, lo);
        c2_store_op<128>(bp.op[0][0], bp.op[0][1], row, fb + u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_X);
      float gyv[FT];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float v[16];
        c3_load16(gyr + u * 16, ok, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) gyv[u * 16 + j] = v[j];
      }
      const float2 ms = ok ? __ldg(reinterpret_cast<const float2 *>(p.saved) + c) : make_float2(0.f, 0.f);
      const float mean = ms.x, rstd = ms.y;
      // ---- 2. h
      mbar_wait(bars + 8 * (hf ? B_D1B : B_D1A), ph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t hi[8], lo[8];
        c3_hidden_unit(tb, fb, u, s_b1, i_pre, sh, hi, lo);
        c2_store_op<128>(bp.op[1][0], bp.op[1][1], row, fb + u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (hf ? B_HB : B_HA));
      // ---- 3. LayerNorm backward -> gz
      mbar_wait(bars + 8 * B_D2, ph);
      tc_fence_after();
      float zh[FT];
      float sum_g = 0.f, sum_gz = 0.f;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t d[16], w[16];
        tmem_ld16(tb + kTX3 + fb + u * 16, d);
        tmem_ld16(tb + kTR3 + fb + u * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int f = fb + u * 16 + j;
          const float z = fmaf(__uint_as_float(w[j]), i_r, fmaf(__uint_as_float(d[j]), i_o, s_b2[f]));
          const float t = (z - mean) * rstd;
          const float gw = gyv[u * 16 + j] * s_lw[f];
          sum_g += gw;
          sum_gz = fmaf(gw, t, sum_gz);
          zh[u * 16 + j] = t;
        }
      }
      tc_fence_before();
      s_part[g * 128 + r] = make_float2(sum_g, sum_gz);
      asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int w = 0; w < G; ++w) m1 += s_part[w * 128 + r].x, m2 += s_part[w * 128 + r].y;
      m1 *= (1.f / 128.f), m2 *= (1.f / 128.f);
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) {
        float t1[32], t2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t1[j] = gyv[c2 * 32 + j] * zh[c2 * 32 + j], t2[j] = gyv[c2 * 32 + j];
        acc_lnw[c2] += lane_sum<32>(t1, lane);
        acc_lnb[c2] += lane_sum<32>(t2, lane);
      }
#pragma unroll
      for (int j = 0; j < FT; ++j) zh[j] = rstd * (fmaf(gyv[j], s_lw[fb + j], -m1) - zh[j] * m2);         // gz
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) {
        float t1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t1[j] = zh[c2 * 32 + j];
        acc_b2[c2] += lane_sum<32>(t1, lane);
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = zh[u * 16 + j] * sgz;
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        tmem_st8(tb + kTH3 + fb / 2 + u * 8, hi);
        tmem_st8(tb + kTH3 + 64 + fb / 2 + u * 8, lo);
        c2_store_op<128>(bp.op[2][0], bp.op[2][1], row, fb + u * 16, hi, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (hf ? B_GZB : B_GZA));
      // ---- 4. gpre = gh gelu'(pre)
      mbar_wait(bars + 8 * B_D3, ph);
      tc_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) {
        float gp[32];
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const int u = c2 * 2 + uu;
          uint32_t d[16], w[16];
          tmem_ld16(tb + kTX3 + fb + u * 16, d);
          tmem_ld16(tb + kTPre3 + fb + u * 16, w);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 pre = ffma2(make_float2(__uint_as_float(w[j]), __uint_as_float(w[j + 1])), make_float2(i_pre, i_pre),
                                     make_float2(s_b1[fb + u * 16 + j], s_b1[fb + u * 16 + j + 1]));
            const float2 gpv = fmul2(fmul2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), make_float2(i_gh, i_gh)),
                                     gelu_bwd2(pre));
            gp[uu * 16 + j] = gpv.x, gp[uu * 16 + j + 1] = gpv.y;
            v[j] = gpv.x * sgp, v[j + 1] = gpv.y * sgp;
          }
          uint32_t hi[8], lo[8];
          split16(v, hi, lo);
          tmem_st8(tb + kTR3 + fb / 2 + u * 8, hi);
          tmem_st8(tb + kTR3 + 64 + fb / 2 + u * 8, lo);
          c2_store_op<128>(bp.op[3][0], bp.op[3][1], row, fb + u * 16, hi, lo);
        }
        acc_b1[c2] += lane_sum<32>(gp, lane);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_GP);
      // ---- 5. gx
      mbar_wait(bars + 8 * B_D4, ph);
      tc_fence_after();
      float4 *gxr = reinterpret_cast<float4 *>(bp.gx + (size_t)(ok ? c : 0) * 128 + fb);
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t d[16];
        tmem_ld16(tb + kTX3 + fb + u * 16, d);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            gxr[u * 4 + t] = make_float4(__uint_as_float(d[4 * t]) * i_gx, __uint_as_float(d[4 * t + 1]) * i_gx,
                                         __uint_as_float(d[4 * t + 2]) * i_gx, __uint_as_float(d[4 * t + 3]) * i_gx);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");      // the other half has read its gx columns: X of the next tile may land
    }
    // parameter-gradient sums: lane t of a warp holds feature 64 g + 32 c + t
    float *mine = s_red + warp * 8 * 32;
#pragma unroll
    for (int c2 = 0; c2 < NC; ++c2) {
      mine[(0 + c2) * 32 + lane] = acc_lnw[c2];
      mine[(2 + c2) * 32 + lane] = acc_lnb[c2];
      mine[(4 + c2) * 32 + lane] = acc_b2[c2];
      mine[(6 + c2) * 32 + lane] = acc_b1[c2];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");
    {
      const int quant = q;                                             // one quantity per warp of the half
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += s_red[(g * 4 + w) * 8 * 32 + (quant * 2 + c2) * 32 + lane];
        float *dst = quant == 0 ? bp.g_lnw : quant == 1 ? bp.g_lnb : quant == 2 ? bp.g_b2 : bp.g_b1;
        if (dst) atomicAdd(dst + fb + 32 * c2 + lane, s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * G) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static int c3_groups() {
  static const int g = getenv("MIMRL_CUBE3_G") ? atoi(getenv("MIMRL_CUBE3_G")) : 4;
  return g == 2 ? 2 : 4;
}

bool cube3_supported(int a_in, int a_hid, int a_out, int inner, long long n_cols, int act, int has_res) {
  if (getenv("MIMRL_CUBE3_OFF")) return false;
  return a_in == 128 && a_hid == 128 && a_out == 128 && inner == 1 && n_cols >= 128 && act == 0 && has_res;
}

int cube3_fwd(const CUtensorMap *m, const CubeTcParams &p, cudaStream_t st, int *handled) {
  *handled = 1;
  const long long n_tiles = (p.n_cols + 127) / 128;
  const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);
  if (c3_groups() == 4) {
    cudaFuncSetAttribute(cube3_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
    cube3_fwd_kernel<4><<<blocks, C3<4>::THREADS, kC3Smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], p);
  } else {
    cudaFuncSetAttribute(cube3_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
    cube3_fwd_kernel<2><<<blocks, C3<2>::THREADS, kC3Smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], p);
  }
  return check_launch("cube3_fwd");
}

int cube3_bwd(const CUtensorMap *m, const CubeBwdParams &bp, cudaStream_t st, int *handled) {
  *handled = 1;
  const long long n_tiles = (bp.f.n_cols + 127) / 128;
  const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);
  if (c3_groups() == 4) {
    cudaFuncSetAttribute(cube3_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
    cube3_bwd_kernel<4><<<blocks, C3<4>::THREADS, kC3Smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], bp);
  } else {
    cudaFuncSetAttribute(cube3_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
    cube3_bwd_kernel<2><<<blocks, C3<2>::THREADS, kC3Smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], bp);
  }
  return check_launch("cube3_bwd");
}

}  // namespace mimrl
