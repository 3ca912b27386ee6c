"""Linear (+ReLU) layers of the critic / baseline / classifier MLP stacks on the
tensor cores: ``mimrl_gemm_f32x3`` (csrc/gemm_tc.cu) for the forward product and
both backward products, with bias + ReLU and the ReLU mask fused in.

``mlp_apply(seq, x)`` evaluates an ``nn.Sequential`` of ``nn.Linear`` / ``nn.ReLU``
(the module layout of VMI.py:13-22, kept for state_dict compatibility) through
these kernels.  Layers outside the tensor-core envelope (fewer than 512 rows -- the
reference trains at batch 128 -- or an output / input width below 32 such as the
1- and 2-wide heads) run on the CUDA-core kernels of csrc/linear_small.cu
(``mimrl_linear_small``, exact fp32).  No library GEMM is left on the path."""
from __future__ import annotations

import torch
import torch.nn as nn

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


def _small(mode, A, mask, B, M, N, K, bias=None, relu=False, colsum=None):
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    ws = torch.empty(L.lib.mimrl_linear_small_workspace_bytes(mode, M, N, K), dtype=torch.uint8, device=A.device)
    L.check(L.lib.mimrl_linear_small(mode, L.ptr(A), L.ptr(mask), L.ptr(B), M, N, K, L.ptr(bias), int(relu), L.ptr(C),
                                     L.ptr(colsum), L.ptr(ws), ws.numel(), L.stream()))
    return C


class _LinearSmall(torch.autograd.Function):
    """x [M,K], w [N,K] on the CUDA-core kernels (csrc/linear_small.cu): forward with bias + ReLU fused, backward with
    the ReLU mask applied on load and the bias gradient accumulated by the weight-gradient kernel."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        x, w = L.f32(x), L.f32(w)
        b = L.f32(b) if b is not None else None
        M, K = x.shape
        N = w.shape[0]
        y = _small(0, x, None, w, M, N, K, b, relu)
        ctx.save_for_backward(x, w, y if relu else x.new_empty(0))
        ctx.cfg = (relu, b is not None, M, N, K)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        relu, has_b, M, N, K = ctx.cfg
        gy = L.f32(gy)
        mask = y if relu else None
        gx = _small(1, gy, mask, w, M, K, N) if ctx.needs_input_grad[0] else None         # dz [M,N] . W [N,K]
        gb = torch.zeros(N, dtype=torch.float32, device=gy.device) if has_b and ctx.needs_input_grad[2] else None
        gw = None
        if ctx.needs_input_grad[1] or gb is not None:
            gw = _small(2, gy, mask, x, N, K, M, colsum=gb)                                # dz^T [N,M] . x [M,K]
        return gx, gw, gb, None


USE_FUSED_MLP = True      # mlps(dim<=128, 256, out<=128, layers=2, relu) in one forward kernel (csrc/mlp_tc.cu)


class _MLP4TC(torch.autograd.Function):
    """Linear+ReLU, Linear+ReLU, Linear+ReLU, Linear with hidden width 256 (the critic MLP of VMI.py:13-22).

    Forward: mimrl_mlp4_fwd, activations in TMEM between the layers.  Backward: the per-layer tensor-core products of
    ``_LinearTC`` on the operands the forward left behind (the hi half of an activation operand is its ReLU mask)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4):
        x = L.f32(x)
        ws = [L.f32(w) for w in (w1, w2, w3, w4)]
        bs = [L.f32(b) if b is not None else None for b in (b1, b2, b3, b4)]
        M, d_in = x.shape
        d_out = ws[3].shape[0]
        dev = x.device
        buf = lambda r, c: torch.empty(L.lib.mimrl_split_bytes(r, c), dtype=torch.uint8, device=dev)
        ops = [buf(M, d_in), buf(M, 256), buf(M, 256), buf(M, 256)]
        wsp = [buf(256, d_in), buf(256, 256), buf(256, 256), buf(d_out, 256)]
        scratch = torch.empty(256, dtype=torch.uint8, device=dev)
        y = torch.empty(M, d_out, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_mlp4_fwd(L.ptr(x), M, d_in, L.ptr(ws[0]), L.ptr(bs[0]), L.ptr(ws[1]), L.ptr(bs[1]), L.ptr(ws[2]),
                                     L.ptr(bs[2]), L.ptr(ws[3]), L.ptr(bs[3]), d_out, L.ptr(y), L.ptr(ops[0]), L.ptr(ops[1]),
                                     L.ptr(ops[2]), L.ptr(ops[3]), L.ptr(wsp[0]), L.ptr(wsp[1]), L.ptr(wsp[2]),
                                     L.ptr(wsp[3]), L.ptr(scratch), L.stream()))
        ctx.save_for_backward(*ops, *wsp, ws[1], ws[2], ws[3])
        ctx.cfg = (M, d_in, d_out, [b is not None for b in bs])
        return y

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        ops, wsp, wts = saved[:4], saved[4:8], saved[8:]
        M, d_in, d_out, has_b = ctx.cfg
        dims = [(256, d_in), (256, 256), (256, 256), (d_out, 256)]          # (N_l, K_l) of layer l
        g = L.f32(gy)
        dev = g.device
        buf = lambda r, c: torch.empty(L.lib.mimrl_split_bytes(r, c), dtype=torch.uint8, device=dev)
        dzs = [buf(M, 256), buf(M, 256), buf(M, 256), buf(M, d_out)]
        gbs = [torch.zeros(n, dtype=torch.float32, device=dev) if hb else None for (n, _), hb in zip(dims, has_b)]
        gx = torch.empty(M, d_in, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        scratch = torch.empty(256, dtype=torch.uint8, device=dev)
        # one kernel: the four data-gradient products, ReLU masks from the activation operands, bias gradients
        L.check(L.lib.mimrl_mlp4_bwd(L.ptr(g), M, d_in, d_out, L.ptr(wts[0]), L.ptr(wts[1]), L.ptr(wts[2]), L.ptr(wsp[0]),
                                     L.ptr(wsp[1]), L.ptr(wsp[2]), L.ptr(wsp[3]), L.ptr(ops[1]), L.ptr(ops[2]), L.ptr(ops[3]),
                                     L.ptr(gx), L.ptr(dzs[0]), L.ptr(dzs[1]), L.ptr(dzs[2]), L.ptr(dzs[3]), L.ptr(gbs[0]),
                                     L.ptr(gbs[1]), L.ptr(gbs[2]), L.ptr(gbs[3]), L.ptr(scratch), L.stream()))
        gws = [_gemm_split(2, dzs[l], ops[l], dims[l][0], dims[l][1], M) if ctx.needs_input_grad[1 + 2 * l] else None
               for l in range(4)]                                            # dz^T [N,M] . input [M,K]
        return (gx, gws[0], gbs[0], gws[1], gbs[1], gws[2], gbs[2], gws[3], gbs[3])


class _MLP4Small(torch.autograd.Function):
    """The same four-layer stack for SMALL batches (< 512 rows: the reference trains at 128) on the CUDA cores, three
    launches per forward + backward instead of sixteen: mimrl_mlp4_small_fwd (all layers, a row group's activations in
    shared memory) and mimrl_mlp4_small_bwd (data gradients of all layers, then one grouped launch for the four weight and
    bias gradients).  Inputs up to 384 wide (the CMI classifier), outputs 1 ... 256 wide (baseline head, classifier head)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4):
        x = L.f32(x)
        ws = [L.f32(w) for w in (w1, w2, w3, w4)]
        bs = [L.f32(b) if b is not None else None for b in (b1, b2, b3, b4)]
        M, d_in = x.shape
        d_out = ws[3].shape[0]
        dev = x.device
        hs = torch.empty(3, M, 256, dtype=torch.float32, device=dev)
        y = torch.empty(M, d_out, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_mlp4_small_fwd(L.ptr(x), M, d_in, L.ptr(ws[0]), L.ptr(bs[0]), L.ptr(ws[1]), L.ptr(bs[1]),
                                           L.ptr(ws[2]), L.ptr(bs[2]), L.ptr(ws[3]), L.ptr(bs[3]), d_out, L.ptr(hs[0]),
                                           L.ptr(hs[1]), L.ptr(hs[2]), L.ptr(y), L.stream()))
        ctx.save_for_backward(x, hs, *ws)
        ctx.cfg = (M, d_in, d_out, [b is not None for b in bs])
        return y

    @staticmethod
    def backward(ctx, gy):
        x, hs, w1, w2, w3, w4 = ctx.saved_tensors
        M, d_in, d_out, has_b = ctx.cfg
        gy = L.f32(gy)
        dev = gy.device
        need = ctx.needs_input_grad
        dz = torch.empty(3, M, 256, dtype=torch.float32, device=dev)
        dz4 = torch.empty(M, d_out, dtype=torch.float32, device=dev)
        gx = torch.empty(M, d_in, dtype=torch.float32, device=dev) if need[0] else None
        dims = [(256, d_in), (256, 256), (256, 256), (d_out, 256)]
        gws = [torch.empty(n, k, dtype=torch.float32, device=dev) if (need[1 + 2 * l] or (has_b[l] and need[2 + 2 * l])) else None
               for l, (n, k) in enumerate(dims)]
        gbs = [torch.empty(n, dtype=torch.float32, device=dev) if (has_b[l] and need[2 + 2 * l]) else None
               for l, (n, _) in enumerate(dims)]
        L.check(L.lib.mimrl_mlp4_small_bwd(L.ptr(gy), L.ptr(x), M, d_in, d_out, L.ptr(w1), L.ptr(w2), L.ptr(w3), L.ptr(w4),
                                           L.ptr(hs[0]), L.ptr(hs[1]), L.ptr(hs[2]), L.ptr(dz[0]), L.ptr(dz[1]), L.ptr(dz[2]),
                                           L.ptr(dz4), L.ptr(gx), L.ptr(gws[0]), L.ptr(gbs[0]), L.ptr(gws[1]), L.ptr(gbs[1]),
                                           L.ptr(gws[2]), L.ptr(gbs[2]), L.ptr(gws[3]), L.ptr(gbs[3]), L.stream()))
        return (gx, gws[0], gbs[0], gws[1], gbs[1], gws[2], gbs[2], gws[3], gbs[3])


def _is_mlp4_small(mods, x):
    if not (USE_FUSED_MLP and len(mods) == 7 and x.dim() == 2 and x.is_cuda and x.shape[0] < MIN_ROWS):
        return False
    if not all(isinstance(mods[i], nn.Linear) for i in (0, 2, 4, 6)) or not all(isinstance(mods[i], nn.ReLU) for i in (1, 3, 5)):
        return False
    l1, l2, l3, l4 = mods[0], mods[2], mods[4], mods[6]
    return (l1.out_features == 256 and l2.in_features == 256 and l2.out_features == 256 and l3.in_features == 256
            and l3.out_features == 256 and l4.in_features == 256 and l1.in_features == x.shape[1]
            and bool(L.lib.mimrl_mlp4_small_supported(l1.in_features, 256, l4.out_features)))


def _is_mlp4(mods, x):
    if not (USE_FUSED_MLP and len(mods) == 7 and x.dim() == 2 and x.is_cuda and x.shape[0] >= MIN_ROWS):
        return False
    if not all(isinstance(mods[i], nn.Linear) for i in (0, 2, 4, 6)) or not all(isinstance(mods[i], nn.ReLU) for i in (1, 3, 5)):
        return False
    l1, l2, l3, l4 = mods[0], mods[2], mods[4], mods[6]
    return (l1.out_features == 256 and l2.in_features == 256 and l2.out_features == 256 and l3.in_features == 256
            and l3.out_features == 256 and l4.in_features == 256 and l4.out_features >= MIN_WIDTH
            and l1.in_features >= MIN_WIDTH and l1.in_features == x.shape[1]
            and bool(L.lib.mimrl_mlp4_supported(l1.in_features, 256, l4.out_features)))


def linear(x, weight, bias=None, relu=False):
    """y = relu?(x W^T + b) for 2-D x; tensor-core path when the shape qualifies."""
    if (x.dim() == 2 and x.is_cuda and x.shape[0] >= MIN_ROWS and weight.shape[0] >= MIN_WIDTH
            and weight.shape[1] >= MIN_WIDTH):
        return _LinearTC.apply(x, weight, bias, relu)
    if x.dim() != 2:
        return linear(x.reshape(-1, x.shape[-1]), weight, bias, relu).reshape(*x.shape[:-1], weight.shape[0])
    return _LinearSmall.apply(x, weight, bias, relu)


def mlp_apply(seq: nn.Sequential, x):
    """Evaluate a Linear/activation stack; Linear+ReLU pairs run as one fused call."""
    mods = list(seq)
    if _is_mlp4_small(mods, x):
        l1, l2, l3, l4 = mods[0], mods[2], mods[4], mods[6]
        return _MLP4Small.apply(x, l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, l4.weight, l4.bias)
    if _is_mlp4(mods, x):
        l1, l2, l3, l4 = mods[0], mods[2], mods[4], mods[6]
        return _MLP4TC.apply(x, l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, l4.weight, l4.bias)
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
