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


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        x, w = L.f32(x), L.f32(w)
        b = L.f32(b) if b is not None else None
        M, K = x.shape
        N = w.shape[0]
        y = _gemm(0, x, None, w, M, N, K, b, relu)
        ctx.save_for_backward(x, w, y if relu else x.new_empty(0))
        ctx.cfg = (relu, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        relu, has_b = ctx.cfg
        gy = L.f32(gy)
        mask = y if relu else None
        M, K = x.shape
        N = w.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _gemm(1, gy, mask, w, M, K, N)            # dz [M,N] . W [N,K]
        if ctx.needs_input_grad[1]:
            gw = _gemm(2, gy, mask, x, N, K, M)            # dz^T [N,M] . x [M,K]
        if has_b and ctx.needs_input_grad[2]:
            gb = (gy * (y > 0) if relu else gy).sum(dim=0)
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
