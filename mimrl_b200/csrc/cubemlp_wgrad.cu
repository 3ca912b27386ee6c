// The three weight-gradient contractions of a CubeMLP axis mix (MLPProcess.py:64-122 backward) in ONE kernel:
//   gW1 [H, A] = gpre^T x,   gW2 [Q, H] = gz^T h,   gWres [Q, A] = gz^T x      (sums over the R fibres)
// Three separate split-K GEMMs read six operands where four exist.  Here a k-block stage holds the stacked gradient
// operand G = [gpre | gz] (MMA rows) and the stacked activation operand P = [x | h] (MMA columns), both as TMA boxes
// from the operands' own buffers placed next to each other in shared memory, so one accumulator G^T P carries all three
// gradients (plus the unused gpre^T h block) and every operand byte is read once.  When the two gradient operands do not
// fit 128 MMA rows together (the 128-wide channel mix) the same k-range is swept twice, once per gradient operand; the
// second sweep finds P in L2.
// Operands: fp16 hi / lo planes in the blocked-K layout the backward kernels write (tiles of 64 fibres, each
// [features][64] contiguous; 256-byte header with the absmax the scale derives from), three products hi.hi + hi.lo +
// lo.hi per k-step, split-K partial sums added in place with red.global.add.
// CTA: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (thread = accumulator row).
#include "tc_common.cuh"

namespace mimrl {
namespace {

constexpr int kWgThreads = 192;
constexpr uint32_t kWgRing = 216 * 1024;
constexpr uint32_t kWgSmem = kWgRing + 512 + 1024;
constexpr int kChunk = 4;          // k-blocks per turn when two gradient tiles share a sweep

struct WgParams {
  int n_mtiles;                  // 1: G = [gpre | gz] stacked in one 128-row tile; 2: tile 0 = gpre, tile 1 = gz
  int rows_gp, rows_gz;          // box heights of gpre / gz (multiples of 8; 128 each when n_mtiles == 2)
  int cols_x, cols_h;            // box heights of x / h (multiples of 16): MMA N = cols_x + cols_h <= 256
  int A, H, Q;                   // true feature counts of x, h (= gpre), gz
  int n_kb, kb_per_split, splits, n_stages, chunk;
  uint32_t stage_bytes, a_plane, b_plane;
  const unsigned *hdr_x, *hdr_h, *hdr_gz, *hdr_gp;
  float *gw1, *gw2, *gwr;        // [H, A], [Q, H], [Q, A] (gwr NULL without res_projection)
};

__device__ __forceinline__ void wg_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
cube_wgrad_kernel(const __grid_constant__ CUtensorMap m_x_hi, const __grid_constant__ CUtensorMap m_x_lo,
                  const __grid_constant__ CUtensorMap m_h_hi, const __grid_constant__ CUtensorMap m_h_lo,
                  const __grid_constant__ CUtensorMap m_gz_hi, const __grid_constant__ CUtensorMap m_gz_lo,
                  const __grid_constant__ CUtensorMap m_gp_hi, const __grid_constant__ CUtensorMap m_gp_lo, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - raw);
  const uint32_t bars = base + kWgRing;
  const uint32_t bFull = bars, bEmpty = bars + 128, bAccFull = bars + 256, bAccEmpty = bars + 272;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kWgRing + 384);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int n_st = p.n_stages;
  const int n_mma = p.cols_x + p.cols_h;

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
    tmem_alloc(smem_u32(gen + kWgRing + 384), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto range_of = [&](int split, int &kb0, int &T) {
    kb0 = split * p.kb_per_split;
    const int kb1 = min(p.n_kb, kb0 + p.kb_per_split);
    T = kb1 > kb0 ? kb1 - kb0 : 0;
  };

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&m_x_hi), prefetch_tmap(&m_h_hi), prefetch_tmap(&m_gz_hi), prefetch_tmap(&m_gp_hi);
      uint32_t n = 0;
      for (int split = blockIdx.x; split < p.splits; split += gridDim.x) {
        int kb0, T;
        range_of(split, kb0, T);
        const int CH = p.n_mtiles == 2 ? p.chunk : (T > 0 ? T : 1);
        for (int c0 = 0; c0 < T; c0 += CH)
        for (int mt = 0; mt < p.n_mtiles; ++mt)
          for (int i = c0; i < T && i < c0 + CH; ++i, ++n) {
            const int stage = n % n_st, kb = kb0 + i;
            mbar_wait(bEmpty + 8 * stage, ((n / n_st) & 1) ^ 1);
            const uint32_t fb = bFull + 8 * stage;
            // (two tiles: the gpre tile needs x only -- gpre^T h is not a gradient)
            const bool with_h = p.n_mtiles == 1 || mt == 1;
            mbar_expect_tx(fb, 2u * p.a_plane + (with_h ? 2u * p.b_plane : 2u * (uint32_t)p.cols_x * 128u));
            const uint32_t dst = base + stage * p.stage_bytes;
            if (p.n_mtiles == 1) {
              tma_load_3d(dst, &m_gp_hi, fb, 0, 0, kb);
              tma_load_3d(dst + p.rows_gp * 128u, &m_gz_hi, fb, 0, 0, kb);
              tma_load_3d(dst + p.a_plane, &m_gp_lo, fb, 0, 0, kb);
              tma_load_3d(dst + p.a_plane + p.rows_gp * 128u, &m_gz_lo, fb, 0, 0, kb);
            } else {
              tma_load_3d(dst, mt ? &m_gz_hi : &m_gp_hi, fb, 0, 0, kb);
              tma_load_3d(dst + p.a_plane, mt ? &m_gz_lo : &m_gp_lo, fb, 0, 0, kb);
            }
            const uint32_t db = dst + 2u * p.a_plane;
            tma_load_3d(db, &m_x_hi, fb, 0, 0, kb);
            tma_load_3d(db + p.b_plane, &m_x_lo, fb, 0, 0, kb);
            if (with_h) {
              tma_load_3d(db + p.cols_x * 128u, &m_h_hi, fb, 0, 0, kb);
              tma_load_3d(db + p.b_plane + p.cols_x * 128u, &m_h_lo, fb, 0, 0, kb);
            }
          }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc_full = instr_desc_f16(128, n_mma), idesc_x = instr_desc_f16(128, p.cols_x);
    uint32_t n = 0, j = 0;
    for (int split = blockIdx.x; split < p.splits; split += gridDim.x) {
      int kb0, T;
      range_of(split, kb0, T);
      // one gradient tile: accumulators alternate between splits; two tiles: accumulator = tile, and the sweep alternates
      // between the tiles every kChunk k-blocks so that the second tile finds the activation operand in L2
      const int CH = p.n_mtiles == 2 ? p.chunk : (T > 0 ? T : 1);
      for (int mt = 0; mt < p.n_mtiles; ++mt) {
        const uint32_t acc = p.n_mtiles == 2 ? (uint32_t)mt : (j & 1);
        mbar_wait(bAccEmpty + 8 * acc, ((p.n_mtiles == 2 ? j : (j >> 1)) & 1) ^ 1);
      }
      tc_fence_after();
      for (int c0 = 0; c0 < T; c0 += CH)
      for (int mt = 0; mt < p.n_mtiles; ++mt) {
        const uint32_t acc = p.n_mtiles == 2 ? (uint32_t)mt : (j & 1), tacc = tmem_base + acc * 256;
        const uint32_t idesc = (p.n_mtiles == 2 && mt == 0) ? idesc_x : idesc_full;
        for (int i = c0; i < T && i < c0 + CH; ++i, ++n) {
          const int stage = n % n_st;
          mbar_wait(bFull + 8 * stage, (n / n_st) & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t s0 = base + stage * p.stage_bytes;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
              const uint32_t a_base = s0 + (prod == 2 ? p.a_plane : 0u);                            // G hi, hi, lo
              const uint32_t b_base = s0 + 2u * p.a_plane + (prod == 1 ? p.b_plane : 0u);           // P hi, lo, hi
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tacc, smem_desc_sw128(a_base + k * 32), smem_desc_sw128(b_base + k * 32), idesc, (i | prod | k) ? 1u : 0u);
            }
            umma_commit(bEmpty + 8 * stage);
          }
          __syncwarp();
        }
      }
      if (leader)
        for (int mt = 0; mt < p.n_mtiles; ++mt) umma_commit(bAccFull + 8 * (p.n_mtiles == 2 ? (uint32_t)mt : (j & 1)));
      __syncwarp();
      ++j;
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const float s_x = scale_from_absmax(p.hdr_x[0]), s_h = scale_from_absmax(p.hdr_h[0]);
    const float s_gz = scale_from_absmax(p.hdr_gz[0]), s_gp = scale_from_absmax(p.hdr_gp[0]);
    uint32_t j = 0;
    for (int split = blockIdx.x; split < p.splits; split += gridDim.x) {
      int kb0, T;
      range_of(split, kb0, T);
      for (int mt = 0; mt < p.n_mtiles; ++mt) {
        const uint32_t acc = p.n_mtiles == 2 ? (uint32_t)mt : (j & 1);
        // which gradient operand and which of its features this accumulator row is
        bool is_gz;
        int f;
        if (p.n_mtiles == 1) is_gz = r >= p.rows_gp, f = is_gz ? r - p.rows_gp : r;
        else is_gz = mt == 1, f = r;
        const bool row_ok = T > 0 && f < (is_gz ? p.Q : p.H) && (p.n_mtiles == 2 || r < p.rows_gp + p.rows_gz);
        const float s_row = is_gz ? s_gz : s_gp;
        mbar_wait(bAccFull + 8 * acc, (p.n_mtiles == 2 ? j : (j >> 1)) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < n_mma; c0 += 16) {
          const bool in_x = c0 < p.cols_x;
          const int c = in_x ? c0 : c0 - p.cols_x;                   // first feature of this chunk in x / h
          const int n_valid = (in_x ? p.A : p.H) - c;
          // gpre^T h is not a gradient; chunks past the operand's features are padding.  A stacked tile has rows of both
          // gradient operands in one warp: the (warp-collective) TMEM load is decided for the warp, the use per lane.
          const bool want = n_valid > 0 && row_ok && !(!in_x && !is_gz) && !(in_x && is_gz && !p.gwr);
          if (!__any_sync(0xffffffffu, want)) continue;
          uint32_t v[16];
          wg_tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + c0, v);
          tmem_ld_wait();
          __syncwarp();
          if (!want) continue;
          const int ld = in_x ? p.A : p.H;
          float *dst = (in_x ? (is_gz ? p.gwr : p.gw1) : p.gw2) + (size_t)f * ld + c;
          const float inv = 1.f / (s_row * (in_x ? s_x : s_h));
          if (n_valid >= 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + jj), "f"(__uint_as_float(v[jj]) * inv),
                           "f"(__uint_as_float(v[jj + 1]) * inv), "f"(__uint_as_float(v[jj + 2]) * inv),
                           "f"(__uint_as_float(v[jj + 3]) * inv)
                           : "memory");
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (jj < n_valid) atomicAdd(dst + jj, __uint_as_float(v[jj]) * inv);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bAccEmpty + 8 * acc);
      }
      ++j;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// op_*: the operand buffers of mimrl_cubemlp_mix_bwd_tc (256-byte header, hi plane, lo plane; `feat` feature rows over R
// fibres in blocked-K order).  Returns 0 and sets *handled when the shapes fit (otherwise the caller runs the three
// separate contractions).
int cube_wgrad_fused(const void *op_x, const void *op_h, const void *op_gz, const void *op_gpre, int A, int H, int Q, long long R,
                     float *gw1, float *gw2, float *gwr, cudaStream_t st, int *handled) {
  *handled = 0;
  static const bool off = getenv("MIMRL_CUBE_WGRAD_FUSED_OFF") != nullptr;
  if (off || R % 64 != 0 || A > 128 || H > 128 || Q > 128) return 0;
  WgParams p{};
  const int r_gp = (H + 7) & ~7, r_gz = (Q + 7) & ~7;
  p.n_mtiles = r_gp + r_gz <= 128 ? 1 : 2;
  p.rows_gp = p.n_mtiles == 1 ? r_gp : 128, p.rows_gz = p.n_mtiles == 1 ? r_gz : 128;
  p.cols_x = (A + 15) & ~15, p.cols_h = (H + 15) & ~15;
  p.A = A, p.H = H, p.Q = Q;
  p.a_plane = (uint32_t)(p.n_mtiles == 1 ? p.rows_gp + p.rows_gz : 128) * 128u;
  p.b_plane = (uint32_t)(p.cols_x + p.cols_h) * 128u;
  p.stage_bytes = 2u * p.a_plane + 2u * p.b_plane;
  // (the MMA reads 128 rows = 16 KB from each G plane; behind a shorter plane that runs into the P planes of the same
  // stage -- a_plane + 16 KB <= stage_bytes for every shape accepted here -- and only feeds accumulator rows nobody reads)
  if (p.a_plane + 16384u > p.stage_bytes) return 0;
  p.n_stages = (int)(kWgRing / p.stage_bytes);
  if (p.n_stages > 16) p.n_stages = 16;
  if (p.n_stages < 2) return 0;
  static const int chunk = getenv("MIMRL_WGRAD_CHUNK") ? atoi(getenv("MIMRL_WGRAD_CHUNK")) : kChunk;
  p.chunk = chunk > 0 ? chunk : kChunk;
  p.n_kb = (int)(R / 64);
  const int want = (p.n_kb + 31) / 32;            // <= 32 k-blocks per accumulator (the tensor core adds with truncation)
  int splits = want <= 148 ? 148 : 148 * ((want + 147) / 148);
  if (splits > p.n_kb) splits = p.n_kb;
  p.kb_per_split = (p.n_kb + splits - 1) / splits;
  p.splits = (p.n_kb + p.kb_per_split - 1) / p.kb_per_split;
  auto planes = [&](const void *op, int feat, const unsigned char *&hi, const unsigned char *&lo) {
    const unsigned char *b = static_cast<const unsigned char *>(op);
    hi = b + 256;
    lo = b + 256 + align256((size_t)feat * (size_t)R * 2);
  };
  const unsigned char *xh, *xl, *hh, *hl, *zh, *zl, *ph, *pl;
  planes(op_x, A, xh, xl), planes(op_h, H, hh, hl), planes(op_gz, Q, zh, zl), planes(op_gpre, H, ph, pl);
  CUtensorMap mx[2], mh[2], mz[2], mp[2];
  const uint64_t kt = (uint64_t)p.n_kb;
  if (make_map_blocked(&mx[0], xh, A, kt, p.cols_x) || make_map_blocked(&mx[1], xl, A, kt, p.cols_x) ||
      make_map_blocked(&mh[0], hh, H, kt, p.cols_h) || make_map_blocked(&mh[1], hl, H, kt, p.cols_h) ||
      make_map_blocked(&mz[0], zh, Q, kt, p.rows_gz) || make_map_blocked(&mz[1], zl, Q, kt, p.rows_gz) ||
      make_map_blocked(&mp[0], ph, H, kt, p.rows_gp) || make_map_blocked(&mp[1], pl, H, kt, p.rows_gp))
    return 1;
  p.hdr_x = static_cast<const unsigned *>(op_x), p.hdr_h = static_cast<const unsigned *>(op_h);
  p.hdr_gz = static_cast<const unsigned *>(op_gz), p.hdr_gp = static_cast<const unsigned *>(op_gpre);
  p.gw1 = gw1, p.gw2 = gw2, p.gwr = gwr;
  cudaFuncSetAttribute(cube_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
  cube_wgrad_kernel<<<p.splits < 148 ? p.splits : 148, kWgThreads, kWgSmem, st>>>(mx[0], mx[1], mh[0], mh[1], mz[0], mz[1], mp[0],
                                                                                mp[1], p);
  *handled = 1;
  return check_launch("cube_wgrad");
}

}  // namespace mimrl
