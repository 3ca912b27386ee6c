"""Linear (+ReLU) layers of the critic / baseline / classifier MLP stacks on the
tensor cores: ``mimrl_gemm_f32x3`` (csrc/gemm_tc.cu) for the forward product and
both backward products, with bias + ReLU and the ReLU mask fused in.

``mlp_apply(seq, x)`` evaluates an ``nn.Sequential`` of ``nn.Linear`` / ``nn.ReLU``
(the module layout of VMI.py:13-22, kept for state_dict compatibility) through
these kernels.  Layers outside the kernel's envelope (fewer than 512 rows, or an
output / input width below 32 such as the 1- and 2-wide heads) are plain library
GEMMs (``F.linear``)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

MIN_ROWS, MIN_WIDTH = 512, 32


def _gemm(mode, A, mask, B, M, N, K, bias=None, relu=False):
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    nbytes = L.lib.mimrl_gemm_workspace_bytes(mode, M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    L.check(L.lib.mimrl_gemm_f32x3(mode, L.ptr(A), L.ptr(mask), L.ptr(B), M, N, K, L.ptr(bias), int(relu), L.ptr(C),
                                   L.ptr(ws), ws.numel(), L.stream()))
    return C


def _split(t, mask=None, colsum=None):
    """fp16 hi/lo split of a row-major fp32 matrix (optionally masked), reusable by several products."""
    rows, cols = t.shape
    buf = torch.empty(L.lib.mimrl_split_bytes(rows, cols), dtype=torch.uint8, device=t.device)
    L.check(L.lib.mimrl_split_f32(L.ptr(t), L.ptr(mask), rows, cols, L.ptr(buf), L.ptr(colsum), L.stream()))
    return buf


def _gemm_split(mode, a_buf, b_buf, M, N, K, bias=None, relu=False):
    C = torch.empty(M, N, dtype=torch.float32, device=a_buf.device)
    ws = torch.empty(L.lib.mimrl_gemm_split_workspace_bytes(mode, M, N, K), dtype=torch.uint8, device=a_buf.device)
    L.check(L.lib.mimrl_gemm_split(mode, L.ptr(a_buf), L.ptr(b_buf), M, N, K, L.ptr(bias), int(relu), L.ptr(C),
                                   L.ptr(ws), ws.numel(), L.stream()))
    return C


class _LinearTC(torch.autograd.Function):
    """x [M,K], w [N,K]: the split of x serves the forward and the weight gradient, the split of w the forward and
    the input gradient, the split of the masked output gradient both backward products and the bias gradient."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        x, w = L.f32(x), L.f32(w)
        b = L.f32(b) if b is not None else None
        M, K = x.shape
        N = w.shape[0]
        xs, wsp = _split(x), _split(w)
        y = _gemm_split(0, xs, wsp, M, N, K, b, relu)
        ctx.save_for_backward(xs, wsp, y if relu else x.new_empty(0))
        ctx.cfg = (relu, b is not None, M, N, K)
        return y

    @staticmethod
    def backward(ctx, gy):
        xs, wsp, y = ctx.saved_tensors
        relu, has_b, M, N, K = ctx.cfg
        gy = L.f32(gy)
        gb = torch.zeros(N, dtype=torch.float32, device=gy.device) if has_b and ctx.needs_input_grad[2] else None
        dzs = _split(gy, y if relu else None, gb)
        gx = _gemm_split(1, dzs, wsp, M, K, N) if ctx.needs_input_grad[0] else None       # dz [M,N] . W [N,K]
        gw = _gemm_split(2, dzs, xs, N, K, M) if ctx.needs_input_grad[1] else None        # dz^T [N,M] . x [M,K]
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    """y = relu?(x W^T + b) for 2-D x; tensor-core path when the shape qualifies."""
    if (x.dim() == 2 and x.is_cuda and x.shape[0] >= MIN_ROWS and weight.shape[0] >= MIN_WIDTH
            and weight.shape[1] >= MIN_WIDTH):
        return _LinearTC.apply(x, weight, bias, relu)
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


def mlp_apply(seq: nn.Sequential, x):
    """Evaluate a Linear/activation stack; Linear+ReLU pairs run as one fused call."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x = linear(x, m.weight, m.bias, relu=fuse)
            i += 2 if fuse else 1
        else:
            x = m(x)
            i += 1
    return x
