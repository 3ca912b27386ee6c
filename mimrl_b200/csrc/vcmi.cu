// Conditional-MI classifier head (reference Model.py:69-70, 198, 203-219):
// clamp(-10,10) -> Hardtanh(1e-4,1-1e-4) | Sigmoid -> BCE against the
// joint/product one-hot targets, plus estimate_cmi's log-odds sums, in one
// pass over the [2n,2] logits; reductions in fp64.
#include "common.cuh"

namespace mimrl {
namespace {

__device__ __forceinline__ float head_act(float l, int act, float &dout) {
  const bool in_clamp = l >= -10.f && l <= 10.f;     // torch.clamp passes gradient on the closed interval
  const float c = fminf(fmaxf(l, -10.f), 10.f);
  float out;
  if (act == 0) {
    const float lo = 1e-4f, hi = 1.f - 1e-4f;
    out = fminf(fmaxf(c, lo), hi);
    dout = (c > lo && c < hi && in_clamp) ? 1.f : 0.f;   // Hardtanh passes gradient on the open interval
  } else {
    out = 1.f / (1.f + expf(-c));
    dout = in_clamp ? out * (1.f - out) : 0.f;
  }
  return out;
}

__global__ void __launch_bounds__(1024) vcmi_head_fwd_kernel(const float *__restrict__ logits, int n, int act,
                                                             float *__restrict__ result) {
  __shared__ double sh[3][32];
  double bce = 0, sj = 0, sp = 0;
  for (int r = threadIdx.x; r < 2 * n; r += blockDim.x) {
    const bool joint = r < n;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float d;
      const float out = head_act(logits[2 * r + c], act, d);
      const float t = (joint ? c == 0 : c == 1) ? 1.f : 0.f;
      bce -= (double)(t * fmaxf(logf(out), -100.f) + (1.f - t) * fmaxf(logf(1.f - out), -100.f));
      if (c == 0) {
        const float odds = logf(out / (1.f - out + 1e-6f));
        if (joint) sj += odds; else sp += odds;
      }
    }
  }
  double v[3] = {bce, sj, sp};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    if (lane == 0) sh[q][w] = v[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[3] = {0, 0, 0};
    for (int q = 0; q < 3; ++q)
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a[q] += sh[q][i];
    const double n2 = 2.0 * n;
    result[0] = (float)(1.0 + a[1] / n2 - a[2] / n2);   // Model.py:219 normalises by batch.shape[0] = 2n (N3)
    result[1] = (float)(a[0] / (2.0 * n2));
  }
}

__global__ void vcmi_head_bwd_kernel(const float *__restrict__ logits, int n, int act,
                                     const float *__restrict__ grad, float *__restrict__ gl) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * n) return;
  const int r = idx >> 1, c = idx & 1;
  const bool joint = r < n;
  float d;
  const float out = head_act(logits[idx], act, d);
  const float t = (joint ? c == 0 : c == 1) ? 1.f : 0.f;
  const float g_cmi = grad[0], g_loss = grad[1];
  float g = g_loss * (-(t / out) + (1.f - t) / (1.f - out)) / (4.f * n);
  if (c == 0) {
    const float om = 1.f - out + 1e-6f;
    g += g_cmi * (joint ? 1.f : -1.f) * (1.f / out + 1.f / om) / (2.f * n);
  }
  gl[idx] = g * d;
}

}  // namespace
}  // namespace mimrl

using namespace mimrl;

extern "C" int mimrl_vcmi_head_fwd(const float *logits, int n, int act, float *result, void *stream) {
  MIMRL_REQUIRE(n > 0, "vcmi_head_fwd: n=%d", n);
  MIMRL_REQUIRE(act == 0 || act == 1, "vcmi_head_fwd: unknown activation %d", act);
  vcmi_head_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, n, act, result);
  return check_launch("vcmi_head_fwd");
}

extern "C" int mimrl_vcmi_head_bwd(const float *logits, int n, int act, const float *grad, float *grad_logits,
                                   void *stream) {
  MIMRL_REQUIRE(n > 0, "vcmi_head_bwd: n=%d", n);
  MIMRL_REQUIRE(act == 0 || act == 1, "vcmi_head_bwd: unknown activation %d", act);
  vcmi_head_bwd_kernel<<<ceil_div(4 * n, 256), 256, 0, (cudaStream_t)stream>>>(logits, n, act, grad, grad_logits);
  return check_launch("vcmi_head_bwd");
}
