// Feature heads either side of the CubeMLP fusion encoder (reference Model.py:466-475 and 489-507):
//   stack : the unmasked temporal means T_F, A_F, V_F = t.mean(1), a.mean(1), v.mean(1), zero-padding of each
//           modality to time_len and torch.stack([t, a, v], dim=2) -> x [bs, time_len, 3, D], in ONE pass over t, a, v;
//   reduce: features_compose_k / features_compose_t = mean | sum over the modality and time axes of the encoder
//           output [bs, L', K', D] -> F_F [bs, D], in one pass.
// HBM-bound row kernels: a warp moves one 512-byte row (D = 128) per instruction with 16-byte accesses.
#include "common.cuh"

namespace mimrl {
namespace {

constexpr int kRowGroups = 8;       // row groups per block (one warp each)

// grid = (bs, 3); block = 32 x kRowGroups.  src_m [bs, len_m, D]; x [bs, time_len, 3, D]; mean_m [bs, D]
__global__ void __launch_bounds__(32 * kRowGroups)
feature_stack_fwd_kernel(const float *__restrict__ t, const float *__restrict__ a, const float *__restrict__ v, int len_t,
                         int len_a, int len_v, int time_len, int D, float *__restrict__ x, float *__restrict__ mean_t,
                         float *__restrict__ mean_a, float *__restrict__ mean_v) {
  extern __shared__ float4 part[];                      // [kRowGroups][D/4]
  const int b = blockIdx.x, m = blockIdx.y;
  const float *src = m == 0 ? t : (m == 1 ? a : v);
  const int len = m == 0 ? len_t : (m == 1 ? len_a : len_v);
  float *mean = m == 0 ? mean_t : (m == 1 ? mean_a : mean_v);
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  const float4 *s4 = reinterpret_cast<const float4 *>(src) + (size_t)b * len * D4;
  float4 *x4 = reinterpret_cast<float4 *>(x) + ((size_t)b * time_len * 3 + m) * D4;
  for (int d4 = lane; d4 < D4; d4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = grp; l < time_len; l += kRowGroups) {
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l < len) {
        val = __ldg(s4 + (size_t)l * D4 + d4);
        acc.x += val.x, acc.y += val.y, acc.z += val.z, acc.w += val.w;
      }
      x4[(size_t)l * 3 * D4 + d4] = val;
    }
    part[grp * D4 + d4] = acc;
  }
  __syncthreads();
  if (grp == 0) {
    const float inv = 1.f / (float)len;
    for (int d4 = lane; d4 < D4; d4 += 32) {
      float4 s = part[d4];
#pragma unroll
      for (int g = 1; g < kRowGroups; ++g) {
        const float4 p = part[g * D4 + d4];
        s.x += p.x, s.y += p.y, s.z += p.z, s.w += p.w;
      }
      reinterpret_cast<float4 *>(mean)[(size_t)b * D4 + d4] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    }
  }
}

// g_src_m[b, l, :] = g_x[b, l, m, :] + g_mean_m[b, :] / len_m      (g_x / g_mean may be NULL)
__global__ void __launch_bounds__(32 * kRowGroups)
feature_stack_bwd_kernel(const float *__restrict__ g_x, const float *__restrict__ g_mean_t, const float *__restrict__ g_mean_a,
                         const float *__restrict__ g_mean_v, int len_t, int len_a, int len_v, int time_len, int D,
                         float *__restrict__ g_t, float *__restrict__ g_a, float *__restrict__ g_v) {
  const int b = blockIdx.x, m = blockIdx.y;
  float *dst = m == 0 ? g_t : (m == 1 ? g_a : g_v);
  if (dst == nullptr) return;
  const float *gm = m == 0 ? g_mean_t : (m == 1 ? g_mean_a : g_mean_v);
  const int len = m == 0 ? len_t : (m == 1 ? len_a : len_v);
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  const float inv = 1.f / (float)len;
  float4 *d4p = reinterpret_cast<float4 *>(dst) + (size_t)b * len * D4;
  const float4 *x4 = g_x ? reinterpret_cast<const float4 *>(g_x) + ((size_t)b * time_len * 3 + m) * D4 : nullptr;
  for (int d4 = lane; d4 < D4; d4 += 32) {
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gm) {
      base = __ldg(reinterpret_cast<const float4 *>(gm) + (size_t)b * D4 + d4);
      base.x *= inv, base.y *= inv, base.z *= inv, base.w *= inv;
    }
    for (int l = grp; l < len; l += kRowGroups) {
      float4 o = base;
      if (x4) {
        const float4 gx = __ldg(x4 + (size_t)l * 3 * D4 + d4);
        o.x += gx.x, o.y += gx.y, o.z += gx.z, o.w += gx.w;
      }
      d4p[(size_t)l * D4 + d4] = o;
    }
  }
}

// out[b, :] = scale * sum over the `rows` rows of x[b] ([rows, D])
__global__ void __launch_bounds__(32 * kRowGroups)
feature_reduce_fwd_kernel(const float *__restrict__ x, int rows, int D, float scale, float *__restrict__ out) {
  extern __shared__ float4 part[];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  const float4 *x4 = reinterpret_cast<const float4 *>(x) + (size_t)b * rows * D4;
  for (int d4 = lane; d4 < D4; d4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = grp; r < rows; r += kRowGroups) {
      const float4 val = __ldg(x4 + (size_t)r * D4 + d4);
      acc.x += val.x, acc.y += val.y, acc.z += val.z, acc.w += val.w;
    }
    part[grp * D4 + d4] = acc;
  }
  __syncthreads();
  if (grp == 0) {
    for (int d4 = lane; d4 < D4; d4 += 32) {
      float4 s = part[d4];
#pragma unroll
      for (int g = 1; g < kRowGroups; ++g) {
        const float4 p = part[g * D4 + d4];
        s.x += p.x, s.y += p.y, s.z += p.z, s.w += p.w;
      }
      reinterpret_cast<float4 *>(out)[(size_t)b * D4 + d4] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
    }
  }
}

__global__ void __launch_bounds__(32 * kRowGroups)
feature_reduce_bwd_kernel(const float *__restrict__ g_out, int rows, int D, float scale, float *__restrict__ g_x) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  float4 *x4 = reinterpret_cast<float4 *>(g_x) + (size_t)b * rows * D4;
  for (int d4 = lane; d4 < D4; d4 += 32) {
    float4 g = __ldg(reinterpret_cast<const float4 *>(g_out) + (size_t)b * D4 + d4);
    g.x *= scale, g.y *= scale, g.z *= scale, g.w *= scale;
    for (int r = grp; r < rows; r += kRowGroups) x4[(size_t)r * D4 + d4] = g;
  }
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" int mimrl_feature_stack_fwd(const float *t, const float *a, const float *v, int bs, int len_t, int len_a,
                                       int len_v, int time_len, int d, float *x, float *mean_t, float *mean_a,
                                       float *mean_v, void *stream) {
  MIMRL_REQUIRE(bs > 0 && d > 0 && d % 4 == 0, "feature_stack: need bs > 0 and d %% 4 == 0 (got bs=%d d=%d)", bs, d);
  MIMRL_REQUIRE(len_t >= 1 && len_a >= 1 && len_v >= 1 && len_t <= time_len && len_a <= time_len && len_v <= time_len,
                "feature_stack: sequence lengths (%d, %d, %d) must lie in [1, time_len=%d]", len_t, len_a, len_v, time_len);
  const size_t smem = (size_t)kRowGroups * d * sizeof(float);
  feature_stack_fwd_kernel<<<dim3(bs, 3), 32 * kRowGroups, smem, (cudaStream_t)stream>>>(t, a, v, len_t, len_a, len_v,
                                                                                          time_len, d, x, mean_t, mean_a,
                                                                                          mean_v);
  return check_launch("feature_stack_fwd_kernel");
}

extern "C" int mimrl_feature_stack_bwd(const float *g_x, const float *g_mean_t, const float *g_mean_a,
                                       const float *g_mean_v, int bs, int len_t, int len_a, int len_v, int time_len, int d,
                                       float *g_t, float *g_a, float *g_v, void *stream) {
  MIMRL_REQUIRE(bs > 0 && d > 0 && d % 4 == 0, "feature_stack: need bs > 0 and d %% 4 == 0 (got bs=%d d=%d)", bs, d);
  feature_stack_bwd_kernel<<<dim3(bs, 3), 32 * kRowGroups, 0, (cudaStream_t)stream>>>(g_x, g_mean_t, g_mean_a, g_mean_v,
                                                                                       len_t, len_a, len_v, time_len, d, g_t,
                                                                                       g_a, g_v);
  return check_launch("feature_stack_bwd_kernel");
}

extern "C" int mimrl_feature_reduce_fwd(const float *x, int bs, int rows, int d, float scale, float *out, void *stream) {
  MIMRL_REQUIRE(bs > 0 && rows > 0 && d > 0 && d % 4 == 0, "feature_reduce: need d %% 4 == 0 (got bs=%d rows=%d d=%d)", bs,
                rows, d);
  const size_t smem = (size_t)kRowGroups * d * sizeof(float);
  feature_reduce_fwd_kernel<<<bs, 32 * kRowGroups, smem, (cudaStream_t)stream>>>(x, rows, d, scale, out);
  return check_launch("feature_reduce_fwd_kernel");
}

extern "C" int mimrl_feature_reduce_bwd(const float *g_out, int bs, int rows, int d, float scale, float *g_x,
                                        void *stream) {
  MIMRL_REQUIRE(bs > 0 && rows > 0 && d > 0 && d % 4 == 0, "feature_reduce: need d %% 4 == 0 (got bs=%d rows=%d d=%d)", bs,
                rows, d);
  feature_reduce_bwd_kernel<<<bs, 32 * kRowGroups, 0, (cudaStream_t)stream>>>(g_out, rows, d, scale, g_x);
  return check_launch("feature_reduce_bwd_kernel");
}
