// Shared declarations of the tensor-core CubeMLP axis-mix kernels (cubemlp_tc.cu: general shapes, one CTA per SM;
// cubemlp_tc2.cu: the strided-axis mixes of the reference configuration specialised at compile time, several CTAs
// per SM).
#pragma once

#include "tc_common.cuh"

namespace mimrl {

struct CubeTcParams {
  const float *x, *b1, *b2, *ln_w, *ln_b;
  float *y, *saved;
  const unsigned *sc_w1, *sc_w2, *sc_wr;               // absmax headers of the split weights
  const float *scales;                                 // forward: [0] scale of x, [1] scale of h (powers of two)
  unsigned *absmax;                                    // [0] max|x| [1] max|gy| [2] max rstd (the forward adds its rstd here)
  int outer, A, H, A2, inner, act, has_res;
  long long n_cols;
  int fibre_scale;          // cube2 forward: operand scales per fibre from the fibre's own max|x| (no pass over x before
                            // the kernel); the kernel publishes max|x| to absmax[0] for a backward that reuses the workspace
};

struct CubeBwdParams {
  CubeTcParams f;
  const float *gy;
  float *gx, *g_b1, *g_b2, *g_lnw, *g_lnb;
  __half *op[4][2];          // x, h, gz, gpre: hi / lo, [features][ld]
  size_t ld;
  const float *scales;       // [0] x [1] h [2] gz [3] gpre
};

namespace {


__device__ __forceinline__ float cube_act(int act, float z) {
  if (act == 0) return gelu_fwd(z);
  if (act == 1) return fmaxf(z, 0.f);
  return tanhf(z);
}

// the same scale from the exponent bits (per-fibre use inside a kernel): 2^(14 - e), amax = f * 2^e with f in [0.5, 1)
__device__ __forceinline__ float pow2_scale_bits(float amax) {
  const int E = (int)((__float_as_uint(amax) >> 23) & 0xffu);          // amax >= 0
  int sh = 14 - (E - 126);
  sh = sh > 60 ? 60 : (sh < -60 ? -60 : sh);
  return __uint_as_float((uint32_t)(127 + sh) << 23);
}

__device__ __forceinline__ float pow2_scale(float amax) {      // amax * scale in [2^13, 2^14)
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);
  int sh = 14 - e;
  sh = sh < -60 ? -60 : (sh > 60 ? 60 : sh);
  return ldexpf(1.f, sh);
}

// split 32 scaled values into fp16 hi / lo pairs (column c = values 2c, 2c+1)
__device__ __forceinline__ void split32(const float (&v)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const __half2 h = __floats2half2_rn(v[j], v[j + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
    hi[j >> 1] = *reinterpret_cast<const uint32_t *>(&h);
    lo[j >> 1] = *reinterpret_cast<const uint32_t *>(&l);
  }
}



__device__ __forceinline__ float cube_dact(int act, float z) {
  if (act == 0) return gelu_bwd(z);
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  const float t = tanhf(z);
  return 1.f - t * t;
}

// sum over the 32 lanes of v[t] for every t; lane t returns the total of entry t
__device__ __forceinline__ float cube_lane_sum(float (&v)[32], int lane) {
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

// ---- helpers of the compile-time specialised kernels (cubemlp_tc2.cu, cubemlp_tc3.cu) ----
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// 16 scaled values -> 8 + 8 packed fp16 hi / lo words (word c = values 2c, 2c+1)
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const __half2 h = __floats2half2_rn(v[j], v[j + 1]);
    const float2 hf = __half22float2(h);
    const float2 d = fsub2(make_float2(v[j], v[j + 1]), hf);
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi[j >> 1] = *reinterpret_cast<const uint32_t *>(&h);
    lo[j >> 1] = *reinterpret_cast<const uint32_t *>(&l);
  }
}

// exact-erf GELU of two values (Abramowitz-Stegun 7.1.26 as in common.cuh::gauss_cdf_pdf) in packed fp32x2 arithmetic.
// With q = Phi(-|z|): gelu(z) = max(z, 0) - |z| q, which needs no select on the sign of z.
__device__ __forceinline__ float2 gelu2(float2 z) {
  const float2 ax = make_float2(fabsf(z.x) * 0.70710678118654752f, fabsf(z.y) * 0.70710678118654752f);
  const float2 den = ffma2(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.f, 1.f));
  const float2 t = make_float2(__fdividef(1.f, den.x), __fdividef(1.f, den.y));
  const float2 arg = fmul2(fmul2(ax, ax), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 e = make_float2(ex2(arg.x), ex2(arg.y));
  float2 poly = ffma2(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  poly = ffma2(t, poly, make_float2(1.421413741f, 1.421413741f));
  poly = ffma2(t, poly, make_float2(-0.284496736f, -0.284496736f));
  poly = ffma2(t, poly, make_float2(0.254829592f, 0.254829592f));
  const float2 q = fmul2(fmul2(t, poly), fmul2(e, make_float2(0.5f, 0.5f)));
  return make_float2(fmaf(-fabsf(z.x), q.x, fmaxf(z.x, 0.f)), fmaf(-fabsf(z.y), q.y, fmaxf(z.y, 0.f)));
}

// gelu'(z) = Phi(z) + z phi(z) of two values, same packed evaluation
__device__ __forceinline__ float2 gelu_bwd2(float2 z) {
  const float2 ax = make_float2(fabsf(z.x) * 0.70710678118654752f, fabsf(z.y) * 0.70710678118654752f);
  const float2 den = ffma2(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.f, 1.f));
  const float2 t = make_float2(__fdividef(1.f, den.x), __fdividef(1.f, den.y));
  const float2 arg = fmul2(fmul2(ax, ax), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 e = make_float2(ex2(arg.x), ex2(arg.y));
  float2 poly = ffma2(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  poly = ffma2(t, poly, make_float2(1.421413741f, 1.421413741f));
  poly = ffma2(t, poly, make_float2(-0.284496736f, -0.284496736f));
  poly = ffma2(t, poly, make_float2(0.254829592f, 0.254829592f));
  const float2 q = fmul2(fmul2(t, poly), fmul2(e, make_float2(0.5f, 0.5f)));
  const float2 zp = fmul2(z, fmul2(e, make_float2(0.3989422804014327f, 0.3989422804014327f)));
  return make_float2(zp.x + (z.x < 0.f ? q.x : 1.f - q.x), zp.y + (z.y < 0.f ? q.y : 1.f - q.y));
}

// column sums over the 32 lanes of a warp for N (16 or 32) values per lane: lane t (and t + 16 for N = 16) returns
// the total of entry t % N
template <int N>
__device__ __forceinline__ float lane_sum(float (&v)[N], int lane) {
#pragma unroll
  for (int s = N / 2, n = N; s >= 1; s >>= 1, n >>= 1) {
    const bool upper = lane & s;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float keep = upper ? v[k + n / 2] : v[k];
      const float send = upper ? v[k] : v[k + n / 2];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  if (N == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
  return v[0];
}

// D (+)= A . W^T, W [rows x K] K-major as loaded (three products: hi.hi + hi.lo + lo.hi); a_lo = column offset of the
// lo half of the TMEM operand, w_half = byte offset of the lo half of the weight, rows = box height of the weight
__device__ __forceinline__ void c2_mma_k(uint32_t d, uint32_t a, uint32_t a_lo, uint32_t sw, uint32_t w_half, uint32_t rows,
                                         int ksteps, uint32_t idesc, uint32_t acc) {
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_off = prod == 2 ? a_lo : 0, b_off = prod == 1 ? w_half : 0;
    for (int k = 0; k < ksteps; ++k) {
      umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128(sw + b_off + (k >> 2) * rows * 128 + (k & 3) * 32), idesc, acc);
      acc = 1;
    }
  }
}
// D (+)= A . W, the contraction running over the ROWS of the same tile (MN-major view)
__device__ __forceinline__ void c2_mma_mn(uint32_t d, uint32_t a, uint32_t a_lo, uint32_t sw, uint32_t w_half, uint32_t rows,
                                          int ksteps, uint32_t idesc, uint32_t acc) {
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_off = prod == 2 ? a_lo : 0, b_off = prod == 1 ? w_half : 0;
    for (int k = 0; k < ksteps; ++k) {
      umma_f16_ts(d, a + a_off + k * 8, smem_desc_sw128_mn(sw + b_off + k * 2048, rows * 128, 1024), idesc, acc);
      acc = 1;
    }
  }
}

// features [f0, f0 + 16) of fibre `row` -> feature-major weight-gradient operand (blocked-K layout of
// make_map_blocked: tiles of 64 consecutive fibres, each [NF][64] contiguous)
template <int NF>
__device__ __forceinline__ void c2_store_op(__half *hi_base, __half *lo_base, size_t row, int f0, const uint32_t (&hi)[8],
                                            const uint32_t (&lo)[8]) {
  const size_t off = ((row >> 6) * NF) * 64 + (row & 63);
  unsigned short *h = reinterpret_cast<unsigned short *>(hi_base) + off, *l = reinterpret_cast<unsigned short *>(lo_base) + off;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int f = f0 + 2 * t;
    if (f < NF) {
      h[f * 64] = (unsigned short)(hi[t] & 0xffffu);
      l[f * 64] = (unsigned short)(lo[t] & 0xffffu);
    }
    if (f + 1 < NF) {
      h[(f + 1) * 64] = (unsigned short)(hi[t] >> 16);
      l[(f + 1) * 64] = (unsigned short)(lo[t] >> 16);
    }
  }
}

}  // namespace

// cubemlp_wgrad.cu: the three weight-gradient contractions of a mix in one kernel (every operand read once)
int cube_wgrad_fused(const void *op_x, const void *op_h, const void *op_gz, const void *op_gpre, int A, int H, int Q, long long R,
                     float *gw1, float *gw2, float *gwr, cudaStream_t st, int *handled);

// cubemlp_tc2.cu: returns 0 and sets *handled = 1 when a compile-time specialisation exists for the shape
int cube2_fwd(const CUtensorMap *maps, const CubeTcParams &p, cudaStream_t st, int *handled);
int cube2_bwd(const CUtensorMap *maps, const CubeBwdParams &bp, cudaStream_t st, int *handled);
// cubemlp_tc3.cu: the channel mix D (contiguous fibres of 128 features, 128 -> 128 -> 128)
int cube3_fwd(const CUtensorMap *maps, const CubeTcParams &p, cudaStream_t st, int *handled);
int cube3_bwd(const CUtensorMap *maps, const CubeBwdParams &bp, cudaStream_t st, int *handled);
bool cube3_supported(int a_in, int a_hid, int a_out, int inner, long long n_cols, int act, int has_res);
bool cube2_supported(int a_in, int a_hid, int a_out, int inner, long long n_cols, int act, int has_res);
// rows of the weight boxes (TMA box heights) the specialisation wants for W1 / W2 / Wres; 0 = not specialised
void cube2_box_rows(int a_in, int a_hid, int a_out, int *rows_w1, int *rows_w2, int *rows_wr);

}  // namespace mimrl
