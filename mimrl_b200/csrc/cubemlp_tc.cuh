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

}  // namespace

// cubemlp_tc2.cu: returns 0 and sets *handled = 1 when a compile-time specialisation exists for the shape
int cube2_fwd(const CUtensorMap *maps, const CubeTcParams &p, cudaStream_t st, int *handled);
int cube2_bwd(const CUtensorMap *maps, const CubeBwdParams &bp, cudaStream_t st, int *handled);
bool cube2_supported(int a_in, int a_hid, int a_out, int inner, long long n_cols, int act, int has_res);
// rows of the weight boxes (TMA box heights) the specialisation wants for W1 / W2 / Wres; 0 = not specialised
void cube2_box_rows(int a_in, int a_hid, int a_out, int *rows_w1, int *rows_w2, int *rows_wr);

}  // namespace mimrl
