// The critic MLP of the reference (`mlps(dim, 256, out, layers = 2, 'relu')`, VMI.py:13-22:
// Linear(dim,256)+ReLU, 2 x [Linear(256,256)+ReLU], Linear(256,out)) as ONE forward kernel on the tensor cores.
//
// A CTA handles tiles of 128 rows; a row is one TMEM lane.  The input row is scaled, split into fp16 hi/lo and parked
// in TMEM as the A operand; each layer's weights stream through a TMA ring (hi and lo halves of a 64-wide k-block
// are separate 32 KB units); each hidden activation relu(D + b) is rewritten IN PLACE over its accumulator chunk as
// the next layer's A operand (hi in the first 16 columns of a 32-feature chunk, lo in the last 16), so the 512 TMEM
// columns alternate as operand (256) and accumulator (256).  Three products (hi.hi + hi.lo + lo.hi) per contraction
// keep fp32-class accuracy.  Operand scales are powers of two from bounds known before the launch
// (|h_l| <= |h_{l-1}|max max_n sum_k |W_l[n,k]| + |b_l|max), so there is no absmax pass over any activation.
//
// For the backward (per-layer kernels of gemm_tc.cu) the kernel also leaves x, h1, h2, h3 behind as fp16 hi/lo
// operands in the row-major layout of mimrl_split_f32: they are the weight-gradient operands, and the sign of the hi
// half is the ReLU mask.  Compared with four Linear calls this removes 4 absmax passes, 4 split passes, 3 fp32
// activation round trips and 3 launches' worth of pipeline fill per MLP.
#include "tc_common.cuh"

namespace mimrl {
namespace {

constexpr int kMWG = 4;                              // epilogue warpgroups
constexpr int kMCh = 8 / kMWG;                       // 32-feature chunks of a 256-wide layer per epilogue thread
constexpr int kMThreads = 64 + 128 * kMWG;           // warp 0 TMA, warp 1 MMA, then the epilogue warpgroups
constexpr int kHidden = 256;
constexpr uint32_t kMUnit = 256 * 128;               // ring unit: [256 n x 64 k] fp16, hi OR lo: 32 KB
constexpr int kMStages = 5;
constexpr uint32_t kMRing = kMStages * kMUnit;
constexpr uint32_t kMVec = kMRing + 256;             // b1 | b2 | b3 (256 each) | b4 (128)
constexpr uint32_t kMSmem = kMVec + (3 * 256 + 128) * 4 + 1024;
constexpr uint32_t kMR0 = 0, kMR1 = 256;

struct Mlp4Params {
  const float *x;
  const float *b[4];
  float *y;                  // [M, OUT]
  __half *op[4][2];          // split (hi, lo) of x [M, ld0], h1, h2, h3 [M, 256], row-major
  const float *scales;       // [0] x [1] h1 [2] h2 [3] h3 (powers of two)
  const unsigned *sc_w[4];   // absmax headers of the split weights
  int M, K0, ld0, OUT;
  long long n_tiles;
};

__device__ __forceinline__ uint32_t m_a_col(int k16, int lo) { return 32u * (k16 >> 1) + 8u * (k16 & 1) + (lo ? 16u : 0u); }

__device__ __forceinline__ void m_split32(const float (&v)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const __half2 h = __floats2half2_rn(v[j], v[j + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
    hi[j >> 1] = *reinterpret_cast<const uint32_t *>(&h);
    lo[j >> 1] = *reinterpret_cast<const uint32_t *>(&l);
  }
}

__device__ __forceinline__ void m_store64(__half *dst, const uint32_t (&w)[16]) {
  uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
  for (int t = 0; t < 4; ++t) d[t] = make_uint4(w[4 * t], w[4 * t + 1], w[4 * t + 2], w[4 * t + 3]);
}

__global__ void __launch_bounds__(kMThreads, 1)
mlp4_fwd_kernel(const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo,
                const __grid_constant__ CUtensorMap map_w4_hi, const __grid_constant__ CUtensorMap map_w4_lo,
                const Mlp4Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kMRing;
  // barriers: full[5] | empty[5] | operand ready[4] | accumulator full[4]
  const uint32_t bFull = bars, bEmpty = bars + 40, bReady = bars + 80, bAcc = bars + 112;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kMRing + 160);
  float *s_b = reinterpret_cast<float *>(gen + kMVec);            // b1 | b2 | b3 | b4
  for (int t = threadIdx.x; t < 3 * 256 + 128; t += blockDim.x) {
    const int l = t < 768 ? t / 256 : 3, i = t < 768 ? t % 256 : t - 768;
    const int width = l < 3 ? kHidden : p.OUT;
    s_b[t] = (p.b[l] && i < width) ? p.b[l][i] : 0.f;
  }

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int kb0 = p.ld0 / 64;                          // k-blocks of the first layer (1 or 2)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMStages; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bReady + 8 * s, 4 * kMWG);
      mbar_init(bAcc + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kMRing + 160), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    uint32_t n = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int layer = 0; layer < 4; ++layer) {
        const int units = 2 * (layer == 0 ? kb0 : 4);
        const uint32_t bytes = layer == 3 ? kMUnit / 2 : kMUnit;
        for (int un = 0; un < units; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kMStages, round = n / kMStages;
          if (round > 0) mbar_wait(bEmpty + 8 * s, (round - 1) & 1);
          if (leader) {
            const CUtensorMap *m = layer == 0 ? (lo ? &map_w1_lo : &map_w1_hi)
                                   : layer == 1 ? (lo ? &map_w2_lo : &map_w2_hi)
                                   : layer == 2 ? (lo ? &map_w3_lo : &map_w3_hi) : (lo ? &map_w4_lo : &map_w4_hi);
            mbar_expect_tx(bFull + 8 * s, bytes);
            tma_load_2d(base + s * kMUnit, m, bFull + 8 * s, kb * 64, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc256 = instr_desc_f16(128, 256), idesc128 = instr_desc_f16(128, 128);
    uint32_t n = 0, it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      for (int layer = 0; layer < 4; ++layer) {
        mbar_wait(bReady + 8 * layer, par);
        tc_fence_after();
        const uint32_t ra = tmem_base + ((layer & 1) ? kMR1 : kMR0), rd = tmem_base + ((layer & 1) ? kMR0 : kMR1);
        const int units = 2 * (layer == 0 ? kb0 : 4);
        const uint32_t idesc = layer == 3 ? idesc128 : idesc256;
        for (int un = 0; un < units; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t s = n % kMStages, round = n / kMStages;
          mbar_wait(bFull + 8 * s, round & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t b0 = base + s * kMUnit;
            // the hi unit feeds a_hi.b_hi and a_lo.b_hi, the lo unit a_hi.b_lo
            for (int a_lo = 0; a_lo < (lo ? 1 : 2); ++a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(rd, ra + m_a_col(kb * 4 + k, a_lo), smem_desc_sw128(b0 + k * 32), idesc, (un | a_lo | k) ? 1u : 0u);
            }
            umma_commit(bEmpty + 8 * s);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bAcc + 8 * layer);
        __syncwarp();
      }
    }
  } else {
    const int e = warp - 2, q = warp & 3, g = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float inv[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) inv[l] = 1.f / (p.scales[l] * scale_from_absmax(p.sc_w[l][0]));
    const float sx = p.scales[0];
    const bool vec = (p.K0 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const long long row = tile * 128 + q * 32 + lane;
      const bool ok = row < p.M;
      const float *xr = p.x + (size_t)(ok ? row : 0) * p.K0;
      // ---- input row -> TMEM region 0 (+ its split copy for the weight gradient of layer 1)
      if (g * 32 < p.ld0) {
        const int c = g;
        float v[32];
        if (vec) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int a = 32 * c + 4 * t;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && a < p.K0) f = __ldg(reinterpret_cast<const float4 *>(xr + a));
            v[4 * t] = f.x * sx, v[4 * t + 1] = f.y * sx, v[4 * t + 2] = f.z * sx, v[4 * t + 3] = f.w * sx;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (ok && 32 * c + j < p.K0) ? __ldg(xr + 32 * c + j) * sx : 0.f;
        }
        uint32_t hi[16], lo[16];
        m_split32(v, hi, lo);
        tmem_st16(tmem_base + lane_off + kMR0 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kMR0 + 32 * c + 16, lo);
        if (ok) {
          m_store64(p.op[0][0] + (size_t)row * p.ld0 + 32 * c, hi);
          m_store64(p.op[0][1] + (size_t)row * p.ld0 + 32 * c, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 0);
      // ---- hidden layers: h = relu(D + b) in place over the accumulator, as the next A operand
      for (int layer = 0; layer < 3; ++layer) {
        mbar_wait(bAcc + 8 * layer, par);
        tc_fence_after();
        const uint32_t reg = (layer & 1) ? kMR0 : kMR1;          // accumulator of this layer
        const float iv = inv[layer], sn = p.scales[layer + 1];
        const float *bl = s_b + layer * 256;
        for (int cc = 0; cc < kMCh; ++cc) {
          const int c = kMCh * g + cc;
          uint32_t d[32];
          tmem_ld32(tmem_base + lane_off + reg + 32 * c, d);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = fmaxf(fmaf(__uint_as_float(d[t]), iv, bl[32 * c + t]), 0.f) * sn;
          uint32_t hi[16], lo[16];
          m_split32(v, hi, lo);
          tmem_st16(tmem_base + lane_off + reg + 32 * c, hi);
          tmem_st16(tmem_base + lane_off + reg + 32 * c + 16, lo);
          if (ok) {
            m_store64(p.op[layer + 1][0] + (size_t)row * kHidden + 32 * c, hi);
            m_store64(p.op[layer + 1][1] + (size_t)row * kHidden + 32 * c, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bReady + 8 * (layer + 1));
      }
      // ---- output layer: y = D + b4 (accumulator in region 0, 128 columns: one chunk per warpgroup)
      mbar_wait(bAcc + 24, par);
      tc_fence_after();
      {
        const int c = g;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kMR0 + 32 * c, d);
        tmem_ld_wait();
        if (ok) {
          float *yr = p.y + (size_t)row * p.OUT;
          if (32 * c + 32 <= p.OUT && (p.OUT & 3) == 0) {
#pragma unroll
            for (int t = 0; t < 32; t += 4)
              *reinterpret_cast<float4 *>(yr + 32 * c + t) =
                  make_float4(fmaf(__uint_as_float(d[t]), inv[3], s_b[768 + 32 * c + t]),
                              fmaf(__uint_as_float(d[t + 1]), inv[3], s_b[768 + 32 * c + t + 1]),
                              fmaf(__uint_as_float(d[t + 2]), inv[3], s_b[768 + 32 * c + t + 2]),
                              fmaf(__uint_as_float(d[t + 3]), inv[3], s_b[768 + 32 * c + t + 3]));
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (32 * c + t < p.OUT) yr[32 * c + t] = fmaf(__uint_as_float(d[t]), inv[3], s_b[768 + 32 * c + t]);
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void mlp4_absmax_kernel(const float *a, size_t na, unsigned *out) {
  float m0 = 0.f;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  size_t head = 0;
  if ((reinterpret_cast<uintptr_t>(a) & 15) == 0) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    for (size_t t = t0; t < na / 4; t += st) {
      const float4 v = __ldg(a4 + t);
      m0 = fmaxf(fmaxf(m0, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    head = (na / 4) * 4;
  }
  for (size_t t = head + t0; t < na; t += st) m0 = fmaxf(m0, fabsf(a[t]));
  for (int o = 16; o; o >>= 1) m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m0));
}

__device__ __forceinline__ float m_pow2_for(float bound) {      // bound * scale in [2^13, 2^14)
  if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
  int e;
  frexpf(bound, &e);
  int sh = 14 - e;
  sh = sh < -60 ? -60 : (sh > 60 ? 60 : sh);
  return ldexpf(1.f, sh);
}

// one block of 256 threads (thread n = output row n of the hidden layers): activation bounds -> scales and headers
__global__ void mlp4_scales_kernel(const unsigned *absmax_x, const float *w1, const float *b1, int K0, const float *w2,
                                   const float *b2, const float *w3, const float *b3, float *scales, unsigned *hdr0,
                                   unsigned *hdr1, unsigned *hdr2, unsigned *hdr3) {
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
  // row L1 norms, one warp per row (coalesced): lane sums its columns, the warp reduces, lane 0 keeps the running max
  float r1 = 0.f, r2 = 0.f, r3 = 0.f;
  {
    const int w = n >> 5, lane = n & 31;
    for (int row = w; row < kHidden; row += 8) {
      float a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int k = lane; k < K0; k += 32) a1 += fabsf(w1[(size_t)row * K0 + k]);
#pragma unroll
      for (int k = lane; k < kHidden; k += 32) {
        a2 += fabsf(w2[(size_t)row * kHidden + k]);
        a3 += fabsf(w3[(size_t)row * kHidden + k]);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        a3 += __shfl_xor_sync(0xffffffffu, a3, o);
      }
      r1 = fmaxf(r1, a1), r2 = fmaxf(r2, a2), r3 = fmaxf(r3, a3);
    }
  }
  const float r1m = block_max(r1), r2m = block_max(r2), r3m = block_max(r3);
  const float b1m = block_max(b1 ? fabsf(b1[n]) : 0.f), b2m = block_max(b2 ? fabsf(b2[n]) : 0.f),
              b3m = block_max(b3 ? fabsf(b3[n]) : 0.f);
  if (n == 0) {
    const float m0 = __uint_as_float(absmax_x[0]);
    const float m1 = m0 * r1m + b1m, m2 = m1 * r2m + b2m, m3 = m2 * r3m + b3m;
    scales[0] = m_pow2_for(m0), scales[1] = m_pow2_for(m1), scales[2] = m_pow2_for(m2), scales[3] = m_pow2_for(m3);
    *hdr0 = __float_as_uint(m0), *hdr1 = __float_as_uint(m1), *hdr2 = __float_as_uint(m2), *hdr3 = __float_as_uint(m3);
  }
}


// ---- backward (data gradients) -----------------------------------------------------------------------------------
// g_out -> TMEM; g_h3 = g_out W4, g_pre3 = g_h3 [h3 > 0] in place; ... ; g_x = g_pre1 W1: four contractions per tile with
// the split weights read MN-major (rows = contraction index) from the same global buffers the forward used.  The ReLU
// masks are the hi halves of the activation operands the forward wrote.  dz4 = g_out, dz3, dz2, dz1 are left behind as
// row-major fp16 hi/lo operands for the weight-gradient GEMMs (gemm_tc.cu mode 2); bias gradients are butterfly lane
// sums accumulated in registers over the CTA's tiles.  Scales from bounds: |dz_{l-1}| <= |dz_l|max max_k sum_n |W_l[n,k]|.
struct Mlp4BwdParams {
  const float *gy;           // [M, OUT]
  float *gx;                 // [M, K0] (nullable)
  const __half *mask[3];     // hi halves of h1, h2, h3 [M, 256]
  __half *dz[4][2];          // split (hi, lo) of dz1, dz2, dz3 [M, 256] and dz4 [M, ldo], row-major
  float *g_b[4];             // += (nullable)
  const float *scales;       // [0] dz1 [1] dz2 [2] dz3 [3] dz4
  const unsigned *sc_w[4];
  int M, K0, ld0, OUT, ldo;
  long long n_tiles;
};

__device__ __forceinline__ float m_lane_sum(float (&v)[32], int lane) {      // lane t returns sum over lanes of v[t]
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

__global__ void __launch_bounds__(kMThreads, 1)
mlp4_bwd_kernel(const __grid_constant__ CUtensorMap mn_w1_hi, const __grid_constant__ CUtensorMap mn_w1_lo,
                const __grid_constant__ CUtensorMap mn_w2_hi, const __grid_constant__ CUtensorMap mn_w2_lo,
                const __grid_constant__ CUtensorMap mn_w3_hi, const __grid_constant__ CUtensorMap mn_w3_lo,
                const __grid_constant__ CUtensorMap mn_w4_hi, const __grid_constant__ CUtensorMap mn_w4_lo,
                const Mlp4BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kMRing;
  const uint32_t bFull = bars, bEmpty = bars + 40, bReady = bars + 80, bAcc = bars + 112;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kMRing + 160);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // steps s = 0..3 walk the layers backwards: W4, W3, W2, W1
  // contraction length (rows of W): ldo, 256, 256, 256; output width (cols of W): 256, 256, 256, ld0
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMStages; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bReady + 8 * s, 4 * kMWG);
      mbar_init(bAcc + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(gen + kMRing + 160), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    const uint32_t leader = elect_one();
    uint32_t n = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int s = 0; s < 4; ++s) {
        const int n_kb = s == 0 ? p.ldo / 64 : 4;
        const int n_nb = s == 3 ? p.ld0 / 64 : 4;             // 64-column blocks of the output
        for (int un = 0; un < 2 * n_kb; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t st = n % kMStages, round = n / kMStages;
          if (round > 0) mbar_wait(bEmpty + 8 * st, (round - 1) & 1);
          if (leader) {
            const CUtensorMap *m = s == 0 ? (lo ? &mn_w4_lo : &mn_w4_hi)
                                   : s == 1 ? (lo ? &mn_w3_lo : &mn_w3_hi)
                                   : s == 2 ? (lo ? &mn_w2_lo : &mn_w2_hi) : (lo ? &mn_w1_lo : &mn_w1_hi);
            mbar_expect_tx(bFull + 8 * st, (uint32_t)n_nb * 8192u);
            for (int nb = 0; nb < n_nb; ++nb) tma_load_2d(base + st * kMUnit + nb * 8192, m, bFull + 8 * st, nb * 64, kb * 64);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    uint32_t n = 0, it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      for (int s = 0; s < 4; ++s) {
        mbar_wait(bReady + 8 * s, par);
        tc_fence_after();
        const uint32_t ra = tmem_base + ((s & 1) ? kMR1 : kMR0), rd = tmem_base + ((s & 1) ? kMR0 : kMR1);
        const int n_kb = s == 0 ? p.ldo / 64 : 4;
        const uint32_t idesc = instr_desc_f16_bmn(128, s == 3 ? p.ld0 : 256);
        for (int un = 0; un < 2 * n_kb; ++un, ++n) {
          const int kb = un >> 1, lo = un & 1;
          const uint32_t st = n % kMStages, round = n / kMStages;
          mbar_wait(bFull + 8 * st, round & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t b0 = base + st * kMUnit;
            for (int a_lo = 0; a_lo < (lo ? 1 : 2); ++a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(rd, ra + m_a_col(kb * 4 + k, a_lo), smem_desc_sw128_mn(b0 + k * 2048, 8192, 1024), idesc,
                            (un | a_lo | k) ? 1u : 0u);
            }
            umma_commit(bEmpty + 8 * st);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bAcc + 8 * s);
        __syncwarp();
      }
    }
  } else {
    const int e = warp - 2, q = warp & 3, g = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // step s multiplies by W_{4-s}: accumulator = dz operand (scale of step s) x split weight
    float inv[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) inv[s] = 1.f / (p.scales[3 - s] * scale_from_absmax(p.sc_w[3 - s][0]));
    const bool vec = (p.OUT & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gy) & 15) == 0;
    float acc_b4 = 0.f, acc_b[3][kMCh] = {};
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const long long row = tile * 128 + q * 32 + lane;
      const bool ok = row < p.M;
      const float *gr = p.gy + (size_t)(ok ? row : 0) * p.OUT;
      // ---- dz4 = g_out -> TMEM region 0
      if (g * 32 < p.ldo) {
        const int c = g;
        float v[32];
        if (vec) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int a = 32 * c + 4 * t;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && a < p.OUT) f = __ldg(reinterpret_cast<const float4 *>(gr + a));
            v[4 * t] = f.x, v[4 * t + 1] = f.y, v[4 * t + 2] = f.z, v[4 * t + 3] = f.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (ok && 32 * c + j < p.OUT) ? __ldg(gr + 32 * c + j) : 0.f;
        }
        float sc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sc[j] = v[j] * p.scales[3];
        uint32_t hi[16], lo[16];
        m_split32(sc, hi, lo);
        tmem_st16(tmem_base + lane_off + kMR0 + 32 * c, hi);
        tmem_st16(tmem_base + lane_off + kMR0 + 32 * c + 16, lo);
        if (ok) {
          m_store64(p.dz[3][0] + (size_t)row * p.ldo + 32 * c, hi);
          m_store64(p.dz[3][1] + (size_t)row * p.ldo + 32 * c, lo);
        }
        acc_b4 += m_lane_sum(v, lane);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bReady + 0);
      // ---- dz3, dz2, dz1: previous product masked by the ReLU of the layer, in place
      for (int s = 0; s < 3; ++s) {
        const int layer = 2 - s;                                 // hidden layer index 2, 1, 0 (h3, h2, h1)
        mbar_wait(bAcc + 8 * s, par);
        tc_fence_after();
        const uint32_t reg = (s & 1) ? kMR0 : kMR1;
        const float iv = inv[s], sn = p.scales[layer];
        for (int cc = 0; cc < kMCh; ++cc) {
          const int c = kMCh * g + cc;
          uint32_t d[32];
          tmem_ld32(tmem_base + lane_off + reg + 32 * c, d);
          uint4 mk[4];
          if (ok) {
            const uint4 *mp = reinterpret_cast<const uint4 *>(p.mask[layer] + (size_t)row * kHidden + 32 * c);
#pragma unroll
            for (int t = 0; t < 4; ++t) mk[t] = __ldg(mp + t);
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) mk[t] = make_uint4(0u, 0u, 0u, 0u);
          }
          tmem_ld_wait();
          const uint32_t *mw = reinterpret_cast<const uint32_t *>(mk);
          float v[32], sc[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            // positive fp16 <=> sign bit clear and magnitude bits non-zero
            const uint32_t hbits = (mw[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
            const bool on = hbits != 0u && hbits < 0x8000u;
            v[t] = on ? __uint_as_float(d[t]) * iv : 0.f;
            sc[t] = v[t] * sn;
          }
          uint32_t hi[16], lo[16];
          m_split32(sc, hi, lo);
          tmem_st16(tmem_base + lane_off + reg + 32 * c, hi);
          tmem_st16(tmem_base + lane_off + reg + 32 * c + 16, lo);
          if (ok) {
            m_store64(p.dz[layer][0] + (size_t)row * kHidden + 32 * c, hi);
            m_store64(p.dz[layer][1] + (size_t)row * kHidden + 32 * c, lo);
          }
          acc_b[layer][cc] += m_lane_sum(v, lane);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bReady + 8 * (s + 1));
      }
      // ---- g_x = dz1 W1 (accumulator in region 0, ld0 columns)
      mbar_wait(bAcc + 24, par);
      tc_fence_after();
      if (g * 32 < p.ld0) {
        const int c = g;
        uint32_t d[32];
        tmem_ld32(tmem_base + lane_off + kMR0 + 32 * c, d);
        tmem_ld_wait();
        if (ok && p.gx) {
          float *xr = p.gx + (size_t)row * p.K0;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (32 * c + t < p.K0) xr[32 * c + t] = __uint_as_float(d[t]) * inv[3];
        }
      }
      tc_fence_before();
    }
    if (p.g_b[3] && g * 32 < p.ldo && 32 * g + lane < p.OUT) atomicAdd(p.g_b[3] + 32 * g + lane, acc_b4);
#pragma unroll
    for (int layer = 0; layer < 3; ++layer)
#pragma unroll
      for (int cc = 0; cc < kMCh; ++cc)
        if (p.g_b[layer]) atomicAdd(p.g_b[layer] + 32 * (kMCh * g + cc) + lane, acc_b[layer][cc]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// one block of 256 threads: gradient bounds -> scales and headers.  Thread k owns column k of W3, W2 (and W4).
__global__ void mlp4_bwd_scales_kernel(const unsigned *absmax_g, const float *w2, const float *w3, const float *w4, int OUT,
                                       float *scales, unsigned *hdr1, unsigned *hdr2, unsigned *hdr3, unsigned *hdr4) {
  __shared__ float red[256];
  const int k = threadIdx.x;
  auto block_max = [&](float v) {
    __syncthreads();
    red[k] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if (k < o) red[k] = fmaxf(red[k], red[k + o]);
      __syncthreads();
    }
    return red[0];
  };
  float c4 = 0.f, c3 = 0.f, c2 = 0.f;
  for (int n = 0; n < OUT; ++n) c4 += fabsf(w4[(size_t)n * kHidden + k]);
  for (int n = 0; n < kHidden; ++n) {
    c3 += fabsf(w3[(size_t)n * kHidden + k]);
    c2 += fabsf(w2[(size_t)n * kHidden + k]);
  }
  const float c4m = block_max(c4), c3m = block_max(c3), c2m = block_max(c2);
  if (k == 0) {
    const float m4 = __uint_as_float(absmax_g[0]);
    const float m3 = m4 * c4m, m2 = m3 * c3m, m1 = m2 * c2m;
    scales[0] = m_pow2_for(m1), scales[1] = m_pow2_for(m2), scales[2] = m_pow2_for(m3), scales[3] = m_pow2_for(m4);
    *hdr1 = __float_as_uint(m1), *hdr2 = __float_as_uint(m2), *hdr3 = __float_as_uint(m3), *hdr4 = __float_as_uint(m4);
  }
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" size_t mimrl_split_bytes(int rows, int cols);
extern "C" int mimrl_split_f32(const float *src, const float *mask, int rows, int cols, void *out, float *colsum,
                               void *stream);

extern "C" int mimrl_mlp4_supported(int d_in, int hidden, int d_out) {
  return d_in >= 1 && d_in <= 128 && hidden == kHidden && d_out >= 1 && d_out <= 128;
}

// y = W4 relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3) + b4 for x [M, d_in]; hidden width 256.
// Outputs besides y: op_x [M, d_in], op_h1, op_h2, op_h3 [M, 256] and ws_w1..4 (the split weights), all in the
// mimrl_split_f32 format (caller-allocated with mimrl_split_bytes): the operands of the per-layer backward
// (mimrl_gemm_split modes 1 and 2); the hi half of op_h* is also the ReLU mask (mimrl_split_f32_hmask).
extern "C" int mimrl_mlp4_fwd(const float *x, int M, int d_in, const float *w1, const float *b1, const float *w2,
                              const float *b2, const float *w3, const float *b3, const float *w4, const float *b4,
                              int d_out, float *y, void *op_x, void *op_h1, void *op_h2, void *op_h3, void *ws_w1,
                              void *ws_w2, void *ws_w3, void *ws_w4, void *scratch256, void *stream) {
  MIMRL_REQUIRE(mimrl_mlp4_supported(d_in, kHidden, d_out), "mlp4_fwd: sizes %d -> 256 -> %d not supported", d_in, d_out);
  MIMRL_REQUIRE(M > 0 && x && w1 && w2 && w3 && w4 && y && op_x && op_h1 && op_h2 && op_h3 && ws_w1 && ws_w2 && ws_w3 &&
                    ws_w4 && scratch256,
                "mlp4_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned *absmax = reinterpret_cast<unsigned *>(scratch256);
  float *scales = reinterpret_cast<float *>((unsigned char *)scratch256 + 64);
  cudaMemsetAsync(absmax, 0, 16, st);
  const size_t nx = (size_t)M * d_in;
  int blocks = (int)((nx + 4095) / 4096);
  blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
  mlp4_absmax_kernel<<<blocks, 256, 0, st>>>(x, nx, absmax);
  if (check_launch("mlp4 absmax")) return 1;
  mlp4_scales_kernel<<<1, 256, 0, st>>>(absmax, w1, b1, d_in, w2, b2, w3, b3, scales, (unsigned *)op_x, (unsigned *)op_h1,
                                        (unsigned *)op_h2, (unsigned *)op_h3);
  if (check_launch("mlp4 scales")) return 1;
  if (int rc = mimrl_split_f32(w1, nullptr, kHidden, d_in, ws_w1, nullptr, stream)) return rc;
  if (int rc = mimrl_split_f32(w2, nullptr, kHidden, kHidden, ws_w2, nullptr, stream)) return rc;
  if (int rc = mimrl_split_f32(w3, nullptr, kHidden, kHidden, ws_w3, nullptr, stream)) return rc;
  if (int rc = mimrl_split_f32(w4, nullptr, d_out, kHidden, ws_w4, nullptr, stream)) return rc;
  const int ld0 = (d_in + 63) & ~63;
  auto maps = [&](void *s, int rows, int cols, int box_rows, CUtensorMap *hi, CUtensorMap *lo) {
    const int ld = (cols + 63) & ~63;
    unsigned char *b = (unsigned char *)s;
    const size_t off_lo = 256 + align256((size_t)rows * ld * 2);
    if (make_map(hi, b + 256, cols, rows, ld, box_rows)) return 1;
    return make_map(lo, b + off_lo, cols, rows, ld, box_rows);
  };
  CUtensorMap m[8];
  if (maps(ws_w1, kHidden, d_in, 256, &m[0], &m[1]) || maps(ws_w2, kHidden, kHidden, 256, &m[2], &m[3]) ||
      maps(ws_w3, kHidden, kHidden, 256, &m[4], &m[5]) || maps(ws_w4, d_out, kHidden, 128, &m[6], &m[7]))
    return 1;
  Mlp4Params p;
  p.x = x, p.y = y, p.M = M, p.K0 = d_in, p.ld0 = ld0, p.OUT = d_out;
  p.b[0] = b1, p.b[1] = b2, p.b[2] = b3, p.b[3] = b4;
  p.scales = scales;
  void *ops[4] = {op_x, op_h1, op_h2, op_h3};
  void *wsp[4] = {ws_w1, ws_w2, ws_w3, ws_w4};
  for (int t = 0; t < 4; ++t) {
    const int ld = t == 0 ? ld0 : kHidden;
    p.op[t][0] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256);
    p.op[t][1] = reinterpret_cast<__half *>((unsigned char *)ops[t] + 256 + align256((size_t)M * ld * 2));
    p.sc_w[t] = reinterpret_cast<const unsigned *>(wsp[t]);
  }
  p.n_tiles = ((long long)M + 127) / 128;
  cudaFuncSetAttribute(mlp4_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMSmem);
  const int grid = (int)(p.n_tiles < 148 ? p.n_tiles : 148);
  mlp4_fwd_kernel<<<grid, kMThreads, kMSmem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], p);
  return check_launch("mlp4_fwd");
}

// Data-gradient pass of mimrl_mlp4_fwd.  gy [M, d_out]; ws_w1..4 and op_h1..3 as left by the forward.  Writes gx [M, d_in]
// (nullable) and the row-major operands dz1, dz2, dz3 [M, 256], dz4 [M, d_out] (mimrl_split_f32 format, caller-allocated):
// gW_l = dz_l^T input_l through mimrl_gemm_split(mode 2).  Accumulates (+=) the bias gradients g_b1..4 (nullable).
extern "C" int mimrl_mlp4_bwd(const float *gy, int M, int d_in, int d_out, const float *w2, const float *w3, const float *w4,
                              const void *ws_w1, const void *ws_w2, const void *ws_w3, const void *ws_w4, const void *op_h1,
                              const void *op_h2, const void *op_h3, float *gx, void *dz1, void *dz2, void *dz3, void *dz4,
                              float *g_b1, float *g_b2, float *g_b3, float *g_b4, void *scratch256, void *stream) {
  MIMRL_REQUIRE(mimrl_mlp4_supported(d_in, kHidden, d_out), "mlp4_bwd: sizes %d -> 256 -> %d not supported", d_in, d_out);
  MIMRL_REQUIRE(M > 0 && gy && w2 && w3 && w4 && ws_w1 && ws_w2 && ws_w3 && ws_w4 && op_h1 && op_h2 && op_h3 && dz1 && dz2 &&
                    dz3 && dz4 && scratch256,
                "mlp4_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned *absmax = reinterpret_cast<unsigned *>(scratch256);
  float *scales = reinterpret_cast<float *>((unsigned char *)scratch256 + 64);
  cudaMemsetAsync(absmax, 0, 16, st);
  const size_t ng = (size_t)M * d_out;
  int blocks = (int)((ng + 4095) / 4096);
  blocks = blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
  mlp4_absmax_kernel<<<blocks, 256, 0, st>>>(gy, ng, absmax);
  if (check_launch("mlp4 absmax (gy)")) return 1;
  mlp4_bwd_scales_kernel<<<1, 256, 0, st>>>(absmax, w2, w3, w4, d_out, scales, (unsigned *)dz1, (unsigned *)dz2, (unsigned *)dz3,
                                            (unsigned *)dz4);
  if (check_launch("mlp4 bwd scales")) return 1;
  const int ld0 = (d_in + 63) & ~63, ldo = (d_out + 63) & ~63;
  auto maps = [&](const void *s, int rows, int cols, CUtensorMap *hi, CUtensorMap *lo) {
    const int ld = (cols + 63) & ~63;
    const unsigned char *b = (const unsigned char *)s;
    const size_t off_lo = 256 + align256((size_t)rows * ld * 2);
    if (make_map(hi, b + 256, cols, rows, ld, 64)) return 1;
    return make_map(lo, b + off_lo, cols, rows, ld, 64);
  };
  CUtensorMap m[8];
  if (maps(ws_w1, kHidden, d_in, &m[0], &m[1]) || maps(ws_w2, kHidden, kHidden, &m[2], &m[3]) ||
      maps(ws_w3, kHidden, kHidden, &m[4], &m[5]) || maps(ws_w4, d_out, kHidden, &m[6], &m[7]))
    return 1;
  Mlp4BwdParams p;
  p.gy = gy, p.gx = gx, p.M = M, p.K0 = d_in, p.ld0 = ld0, p.OUT = d_out, p.ldo = ldo;
  const void *hs[3] = {op_h1, op_h2, op_h3};
  for (int t = 0; t < 3; ++t) p.mask[t] = reinterpret_cast<const __half *>((const unsigned char *)hs[t] + 256);
  void *dzs[4] = {dz1, dz2, dz3, dz4};
  const void *wsp[4] = {ws_w1, ws_w2, ws_w3, ws_w4};
  for (int t = 0; t < 4; ++t) {
    const int ld = t == 3 ? ldo : kHidden;
    p.dz[t][0] = reinterpret_cast<__half *>((unsigned char *)dzs[t] + 256);
    p.dz[t][1] = reinterpret_cast<__half *>((unsigned char *)dzs[t] + 256 + align256((size_t)M * ld * 2));
    p.sc_w[t] = reinterpret_cast<const unsigned *>(wsp[t]);
  }
  p.g_b[0] = g_b1, p.g_b[1] = g_b2, p.g_b[2] = g_b3, p.g_b[3] = g_b4;
  p.scales = scales;
  p.n_tiles = ((long long)M + 127) / 128;
  cudaFuncSetAttribute(mlp4_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMSmem);
  const int grid = (int)(p.n_tiles < 148 ? p.n_tiles : 148);
  mlp4_bwd_kernel<<<grid, kMThreads, kMSmem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], p);
  return check_launch("mlp4_bwd");
}
