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

namespace mimrl {
namespace {

constexpr int kCubeThreads = 192;                      // warp 0 TMA, warp 1 MMA, warps 2-5 compute (thread = fibre)
constexpr uint32_t kW16 = 128 * 128;                   // one 128-row x 64-K block: 16 KB
constexpr uint32_t kWMat = 4 * kW16;                   // hi kb0, hi kb1, lo kb0, lo kb1
constexpr uint32_t kCubeSmem = 3 * kWMat + 256 + 4 * 128 * 4 + 1024;    // weights, barriers, b1/b2/ln_w/ln_b, alignment
// TMEM columns
constexpr uint32_t kTX = 0, kTD1 = 128, kTH = 256, kTD2 = 384;

struct CubeTcParams {
  const float *x, *b1, *b2, *ln_w, *ln_b;
  float *y, *saved;
  const unsigned *sc_w1, *sc_w2, *sc_wr;               // absmax headers of the split weights
  int outer, A, H, A2, inner, act, has_res;
  long long n_cols;
};

__device__ __forceinline__ float cube_act(int act, float z) {
  if (act == 0) return 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
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
    mbar_init(bXReady, 4);
    mbar_init(bD1Full, 1);
    mbar_init(bHReady, 4);
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
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float inv_w1 = 1.f / scale_from_absmax(p.sc_w1[0]), inv_w2 = 1.f / scale_from_absmax(p.sc_w2[0]);
    const float inv_wr = p.has_res ? 1.f / scale_from_absmax(p.sc_wr[0]) : 0.f;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const long long c = tile * 128 + r;
      const bool ok = c < p.n_cols;
      const long long o = ok ? c / p.inner : 0, i = ok ? c - o * p.inner : 0;
      const float *xf = p.x + (size_t)o * p.A * p.inner + (size_t)i;
      float *yf = p.y + (size_t)o * p.A2 * p.inner + (size_t)i;
      // ---- 1. fibre -> TMEM (own power-of-two scale, fp16 hi/lo)
      float amax = 0.f;
      if (ok) {          // 8 independent loads in flight per round trip
        float m8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int a = 0;
        for (; a + 8 <= p.A; a += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) m8[u] = fmaxf(m8[u], fabsf(__ldg(xf + (size_t)(a + u) * p.inner)));
        }
        for (; a < p.A; ++a) m8[0] = fmaxf(m8[0], fabsf(__ldg(xf + (size_t)a * p.inner)));
        amax = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      }
      const float sx = pow2_scale(amax), inv_x = 1.f / sx;
      for (int ch = 0; ch * 32 < ks1 * 16; ++ch) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int a = ch * 32 + j;
          v[j] = (ok && a < p.A) ? __ldg(xf + (size_t)a * p.inner) * sx : 0.f;
        }
        uint32_t hi[16], lo[16];
        split32(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kTX + ch * 16, hi);
        tmem_st16(tmem_base + lane_off + kTX + 64 + ch * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bXReady);
      // ---- 2. pre-activation -> h (bound |act(z)| <= |z| gives the scale without a second activation pass)
      mbar_wait(bD1Full, ph);
      tc_fence_after();
      const float s1 = inv_x * inv_w1;
      float hmax = 0.f;
      for (int ch = 0; ch * 32 < n1; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + kTD1 + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = ch * 32 + j;
          if (h < p.H) hmax = fmaxf(hmax, fabsf(fmaf(__uint_as_float(v[j]), s1, s_b1[h])));
        }
      }
      const float sh = pow2_scale(hmax), inv_h = 1.f / sh;
      for (int ch = 0; ch * 32 < ks2 * 16; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + kTD1 + ch * 32, v);
        tmem_ld_wait();
        float hv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = ch * 32 + j;
          hv[j] = h < p.H ? cube_act(p.act, fmaf(__uint_as_float(v[j]), s1, s_b1[h])) * sh : 0.f;
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
      // ---- 3. z = o + r + b2 (+ x), LayerNorm over A' thread-locally, one write
      mbar_wait(bD2Full, ph);
      tc_fence_after();
      const float s2 = inv_h * inv_w2, s3 = inv_x * inv_wr;
      auto z_chunk = [&](int ch, float (&z)[32]) {
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
          z[j] = t;
        }
      };
      float sum = 0.f;
      for (int ch = 0; ch * 32 < n2; ++ch) {
        float z[32];
        z_chunk(ch, z);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += z[j];            // entries past A2 are zero
      }
      const float mean = sum / p.A2;
      float var = 0.f;
      for (int ch = 0; ch * 32 < n2; ++ch) {
        float z[32];
        z_chunk(ch, z);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dlt = (ch * 32 + j < p.A2) ? z[j] - mean : 0.f;
          var = fmaf(dlt, dlt, var);
        }
      }
      const float rstd = rsqrtf(var / p.A2 + 1e-6f);
      for (int ch = 0; ch * 32 < n2; ++ch) {
        float z[32];
        z_chunk(ch, z);
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int a2 = ch * 32 + j;
            if (a2 < p.A2) yf[(size_t)a2 * p.inner] = (z[j] - mean) * rstd * s_lw[a2] + s_lb[a2];
          }
        }
      }
      if (ok) {
        p.saved[2 * c] = mean;
        p.saved[2 * c + 1] = rstd;
      }
      tc_fence_before();        // TMEM reads of this tile are done before the next tile's stores / MMAs
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_split_bytes(int rows, int cols);
extern "C" int mimrl_split_f32(const float *src, const float *mask, int rows, int cols, void *out, float *colsum,
                               void *stream);

extern "C" int mimrl_cubemlp_tc_supported(int a_in, int a_hid, int a_out, int ln_first, int act) {
  const bool small = a_in <= 8 && a_hid <= 8 && a_out <= 8;
  return !ln_first && !small && a_in <= 128 && a_hid <= 128 && a_out <= 128 && act >= 0 && act <= 2;
}

extern "C" size_t mimrl_cubemlp_tc_workspace_bytes(int a_in, int a_hid, int a_out) {
  return mimrl_split_bytes(a_hid, a_in) + mimrl_split_bytes(a_out, a_hid) + mimrl_split_bytes(a_out, a_in) + 256;
}

extern "C" int mimrl_cubemlp_mix_fwd_tc(const float *x, int outer, int a_in, int inner, const float *w1, const float *b1,
                                        int a_hid, const float *w2, const float *b2, int a_out, const float *wres,
                                        const float *ln_w, const float *ln_b, int act, float *y, float *saved,
                                        void *workspace, size_t workspace_bytes, void *stream) {
  MIMRL_REQUIRE(mimrl_cubemlp_tc_supported(a_in, a_hid, a_out, 0, act), "cubemlp_mix_fwd_tc: sizes %d/%d/%d act %d not supported",
                a_in, a_hid, a_out, act);
  MIMRL_REQUIRE(outer > 0 && inner > 0 && x && y && saved && w1 && w2 && ln_w && ln_b, "cubemlp_mix_fwd_tc: bad arguments");
  MIMRL_REQUIRE(wres || a_in == a_out, "cubemlp_mix: without res_project d_in must equal d_out (MLPProcess.py:46-48)");
  MIMRL_REQUIRE(workspace_bytes >= mimrl_cubemlp_tc_workspace_bytes(a_in, a_hid, a_out), "cubemlp_mix_fwd_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  unsigned char *s1 = ws, *s2 = s1 + mimrl_split_bytes(a_hid, a_in), *s3 = s2 + mimrl_split_bytes(a_out, a_hid);
  if (int rc = mimrl_split_f32(w1, nullptr, a_hid, a_in, s1, nullptr, stream)) return rc;
  if (int rc = mimrl_split_f32(w2, nullptr, a_out, a_hid, s2, nullptr, stream)) return rc;
  if (wres)
    if (int rc = mimrl_split_f32(wres, nullptr, a_out, a_in, s3, nullptr, stream)) return rc;
  auto maps = [&](unsigned char *s, int rows, int cols, CUtensorMap *hi, CUtensorMap *lo) {
    const int ld = (cols + 63) & ~63;
    const size_t off_lo = 256 + align256((size_t)rows * ld * 2);
    if (make_map(hi, s + 256, cols, rows, ld, 128)) return 1;
    return make_map(lo, s + off_lo, cols, rows, ld, 128);
  };
  CUtensorMap m1h, m1l, m2h, m2l, mrh, mrl;
  if (maps(s1, a_hid, a_in, &m1h, &m1l)) return 1;
  if (maps(s2, a_out, a_hid, &m2h, &m2l)) return 1;
  if (wres) {
    if (maps(s3, a_out, a_in, &mrh, &mrl)) return 1;
  } else {
    mrh = m1h, mrl = m1l;
  }
  CubeTcParams p;
  p.x = x, p.b1 = b1, p.b2 = b2, p.ln_w = ln_w, p.ln_b = ln_b, p.y = y, p.saved = saved;
  p.sc_w1 = reinterpret_cast<const unsigned *>(s1), p.sc_w2 = reinterpret_cast<const unsigned *>(s2);
  p.sc_wr = reinterpret_cast<const unsigned *>(s3);
  p.outer = outer, p.A = a_in, p.H = a_hid, p.A2 = a_out, p.inner = inner, p.act = act, p.has_res = wres ? 1 : 0;
  p.n_cols = (long long)outer * inner;
  const long long n_tiles = (p.n_cols + 127) / 128;
  cudaFuncSetAttribute(cubemlp_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCubeSmem);
  const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);
  cubemlp_tc_fwd_kernel<<<blocks, kCubeThreads, kCubeSmem, st>>>(m1h, m1l, m2h, m2l, mrh, mrl, p);
  return check_launch("cubemlp_tc_fwd");
}
