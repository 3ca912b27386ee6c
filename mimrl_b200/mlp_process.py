"""CubeMLP fusion encoder: drop-in for the reference's ``MLPProcess.py``
(``MLP``, ``MLPsBlock``, ``MLPEncoder`` with the same constructor arguments,
``forward(x, mask=None)`` and parameter paths ``layers_stack.N.mlp_{l,k,d}.fc{1,2}``,
``res_projection_{l,k,d}``, ``ln_{l,k,d}``).

Each of the three axis mixes of a block (MLPProcess.py:64-122) runs as ONE
fused kernel over ``x [bs, L, K, D]`` in place of the reference's permute ->
Linear -> act -> Linear -> permute -> residual(-projection) -> LayerNorm chain;
see csrc/cubemlp.cu.  The fused path covers dropout p = 0 (every reference
launch command) or eval mode, activations gelu / relu / tanh and mask=None
(the reference itself only warns about masks and never passes one,
Model.py:481).  Other settings run the reference's op order with stock torch
ops on the GPU and say so once.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

_ACT_IDS = {"gelu": 0, "relu": 1, "tanh": 2}
USE_TC = True        # tensor-core forward for the large mixes (set False to force the CUDA-core kernels)
_ACT_FNS = {  # Utils.py:85-98 get_activation_function
    "elu": F.elu, "gelu": F.gelu, "hardshrink": F.hardshrink, "hardtanh": F.hardtanh, "leakyrelu": F.leaky_relu,
    "prelu": F.prelu, "relu": F.relu, "rrelu": F.rrelu, "tanh": torch.tanh,
}


def get_activation_function(activation):
    return _ACT_FNS[activation]


class _AxisMix(torch.autograd.Function):
    """y = mix along `axis` of x [bs, L, K, D]; parameters in nn.Linear layout."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, wres, ln_w, ln_b, axis, ln_first, act_id, prev_w=None, prev_b=None, ws_prepared=None):
        x = L.f32(x)
        shape = list(x.shape)
        A = shape[axis]
        outer = 1
        for s in shape[:axis]:
            outer *= s
        inner = 1
        for s in shape[axis + 1:]:
            inner *= s
        H, A2 = w1.shape[0], w2.shape[0]
        prm = [None if t is None else L.f32(t.detach()) for t in (w1, b1, w2, b2, wres, ln_w, ln_b)]
        oshape = shape[:axis] + [A2] + shape[axis + 1:]
        y = torch.empty(oshape, dtype=torch.float32, device=x.device)
        small = bool(L.lib.mimrl_cubemlp_small_supported(A, H, A2))
        # LayerNorm statistics for the backward; the tiny-axis backward (modality mix) recomputes them instead
        saved = None if small else torch.empty(2 * outer * inner, dtype=torch.float32, device=x.device)
        ws = None
        if (USE_TC and outer * inner >= 1024
                and L.lib.mimrl_cubemlp_tc_supported(A, H, A2, int(ln_first), act_id)):
            # ws_prepared: this mix's slice of the encoder-wide workspace, already filled by mimrl_cubemlp_prep_many
            ws = ws_prepared if ws_prepared is not None else torch.empty(
                L.lib.mimrl_cubemlp_tc_workspace_bytes(A, H, A2), dtype=torch.uint8, device=x.device)
            L.check(L.lib.mimrl_cubemlp_mix_fwd_tc(L.ptr(x), outer, A, inner, L.ptr(prm[0]), L.ptr(prm[1]), H,
                                                   L.ptr(prm[2]), L.ptr(prm[3]), A2, L.ptr(prm[4]), L.ptr(prm[5]),
                                                   L.ptr(prm[6]), act_id, L.ptr(y), L.ptr(saved), L.ptr(ws), ws.numel(),
                                                   L.ptr(L.f32(prev_w.detach())) if prev_w is not None else None,
                                                   L.ptr(L.f32(prev_b.detach())) if prev_b is not None else None,
                                                   prev_w.numel() if prev_w is not None else 0,
                                                   int(ws_prepared is not None), L.stream()))
        else:
            L.check(L.lib.mimrl_cubemlp_mix_fwd(L.ptr(x), outer, A, inner, L.ptr(prm[0]), L.ptr(prm[1]), H,
                                                L.ptr(prm[2]), L.ptr(prm[3]), A2, L.ptr(prm[4]), L.ptr(prm[5]),
                                                L.ptr(prm[6]), int(ln_first), act_id, L.ptr(y), L.ptr(saved),
                                                L.stream()))
        ctx.save_for_backward(x, saved if saved is not None else x.new_empty(0), *[p if p is not None else x.new_empty(0) for p in prm])
        ctx.cfg = (outer, A, inner, H, A2, int(ln_first), act_id, [p is not None for p in prm])
        ctx.use_tc = ws is not None
        ctx.ws = ws          # split weights, max|x| and max rstd of this call: the backward reuses them
        return y

    @staticmethod
    def backward(ctx, gy):
        x, saved, *prm = ctx.saved_tensors
        outer, A, inner, H, A2, ln_first, act_id, present = ctx.cfg
        prm = [p if ok else None for p, ok in zip(prm, present)]
        w1, b1, w2, b2, wres, ln_w, ln_b = prm
        gy = L.f32(gy)
        dev = x.device
        gx = torch.empty_like(x)
        if ctx.use_tc:
            return _AxisMix._backward_tc(x, gy, saved, prm, ctx.cfg, gx, ctx.ws)
        if L.lib.mimrl_cubemlp_small_supported(A, H, A2):
            # the modality mix: data, weight, bias and LayerNorm gradients from one register-resident kernel
            sizes = [w1.numel(), w2.numel(), H, A2, wres.numel() if wres is not None else 0, A2, A2]
            parts = torch.split(torch.zeros(sum(sizes), device=dev), sizes)       # one fill launch for all of them
            gw1, gw2 = parts[0].view_as(w1), parts[1].view_as(w2)
            gb1 = parts[2] if b1 is not None else None
            gb2 = parts[3] if b2 is not None else None
            gwres = parts[4].view_as(wres) if wres is not None else None
            gln = (parts[5], parts[6])
            L.check(L.lib.mimrl_cubemlp_small_bwd(L.ptr(x), L.ptr(gy), outer, A, inner, L.ptr(w1), L.ptr(b1), H, L.ptr(w2),
                                                  L.ptr(b2), A2, L.ptr(wres), L.ptr(ln_w), L.ptr(ln_b), ln_first, act_id,
                                                  L.ptr(gx), L.ptr(gw1), L.ptr(gb1), L.ptr(gw2), L.ptr(gb2), L.ptr(gwres),
                                                  L.ptr(gln[0]), L.ptr(gln[1]), L.stream()))
            return gx, gw1, gb1, gw2, gb2, gwres, gln[0], gln[1], None, None, None, None, None, None
        s_gz = torch.empty(outer, A2, inner, device=dev)
        s_h = torch.empty(outer, H, inner, device=dev)
        s_gpre = torch.empty(outer, H, inner, device=dev)
        s_u = torch.empty(outer, A, inner, device=dev) if ln_first else None
        gln = torch.zeros(2, ln_w.numel(), device=dev)
        L.check(L.lib.mimrl_cubemlp_mix_bwd(L.ptr(x), L.ptr(gy), outer, A, inner, L.ptr(w1), L.ptr(b1), H, L.ptr(w2),
                                            L.ptr(b2), A2, L.ptr(wres), L.ptr(ln_w), L.ptr(ln_b), ln_first, act_id,
                                            L.ptr(saved), L.ptr(gx), L.ptr(s_gz), L.ptr(s_h), L.ptr(s_gpre),
                                            L.ptr(s_u), L.ptr(gln[0]), L.ptr(gln[1]), L.stream()))
        x3 = x.view(outer, A, inner)
        u3 = s_u if ln_first else x3
        # weight gradients: contractions over the fibres (outer, inner) on the CUDA-core product kernel
        # (csrc/linear_small.cu, mode 2: C[M,N] = A[R,M]^T . B[R,N]); the bias gradients ride along as column sums
        from .linear import _small
        rows = lambda t: t.permute(0, 2, 1).reshape(-1, t.shape[1]).contiguous()          # [outer, F, inner] -> [R, F]
        r_gz, r_h, r_gpre, r_u = rows(s_gz), rows(s_h), rows(s_gpre), rows(u3)
        R = r_gz.shape[0]
        gb2 = torch.zeros(A2, device=dev) if b2 is not None else None
        gb1 = torch.zeros(H, device=dev) if b1 is not None else None
        gw2 = _small(2, r_gz, None, r_h, A2, H, R, colsum=gb2)
        gw1 = _small(2, r_gpre, None, r_u, H, A, R, colsum=gb1)
        gwres = _small(2, r_gz, None, r_u if not ln_first else rows(x3), A2, A, R) if wres is not None else None
        return gx, gw1, gb1, gw2, gb2, gwres, gln[0], gln[1], None, None, None, None, None, None


def _backward_tc(x, gy, saved, prm, cfg, gx, ws_fwd=None):
    """Tensor-core backward: one fused data-gradient kernel, then three split-K GEMMs over the feature-major
    fp16 hi/lo operands it leaves behind (csrc/cubemlp_tc.cu)."""
    outer, A, inner, H, A2, ln_first, act_id, present = cfg
    w1, b1, w2, b2, wres, ln_w, ln_b = prm
    dev = x.device
    st = L.stream()
    R = L.lib.mimrl_cubemlp_tc_fibre_rows(outer, inner)
    ops = [torch.empty(L.lib.mimrl_cubemlp_tc_op_bytes(n, R), dtype=torch.uint8, device=dev) for n in (A, H, A2, H)]
    ws = ws_fwd if ws_fwd is not None else torch.empty(L.lib.mimrl_cubemlp_tc_workspace_bytes(A, H, A2), dtype=torch.uint8,
                                                       device=dev)
    # every accumulated output of this mix lives in ONE zero-filled buffer (one fill launch instead of seven)
    sizes = [H, A2, A2, A2, H * A, A2 * H, A2 * A if wres is not None else 0]
    zbuf = torch.zeros(sum(sizes), device=dev)
    parts = torch.split(zbuf, sizes)
    gb1 = parts[0] if b1 is not None else None
    gb2 = parts[1] if b2 is not None else None
    gln = (parts[2], parts[3])
    gw1, gw2 = parts[4].view(H, A), parts[5].view(A2, H)
    gwres = parts[6].view(A2, A) if wres is not None else None
    # one call: the fused data-gradient kernel, then the three weight-gradient contractions over the operands it leaves
    L.check(L.lib.mimrl_cubemlp_mix_bwd_tc(L.ptr(x), L.ptr(gy), outer, A, inner, L.ptr(w1), L.ptr(b1), H, L.ptr(w2), L.ptr(b2),
                                           A2, L.ptr(wres), L.ptr(ln_w), L.ptr(ln_b), act_id, L.ptr(saved), L.ptr(gx),
                                           L.ptr(gb1), L.ptr(gb2), L.ptr(gln[0]), L.ptr(gln[1]), L.ptr(gw1), L.ptr(gw2),
                                           L.ptr(gwres), L.ptr(ops[0]), L.ptr(ops[1]), L.ptr(ops[2]), L.ptr(ops[3]),
                                           L.ptr(ws), ws.numel(), int(ws_fwd is not None), st))
    return gx, gw1, gb1, gw2, gb2, gwres, gln[0], gln[1], None, None, None, None, None, None


_AxisMix._backward_tc = staticmethod(_backward_tc)


class MLP(nn.Module):
    """MLPProcess.py:9-21: fc2(act(fc1(x))) on the last axis."""

    def __init__(self, activate, d_in, d_hidden, d_out, bias):
        super().__init__()
        self.fc1 = nn.Linear(d_in, d_hidden, bias=bias)
        self.fc2 = nn.Linear(d_hidden, d_out, bias=bias)
        self.activate = activate
        self.activation = get_activation_function(activate)

    def forward(self, x, mask=None):
        return self.fc2(self.activation(self.fc1(x)))


class MLPsBlock(nn.Module):
    """MLPProcess.py:25-122."""

    def __init__(self, activate, d_ins, d_hiddens, d_outs, dropouts, bias, ln_first=False, res_project=False):
        super().__init__()
        self.mlp_l = MLP(activate, d_ins[0], d_hiddens[0], d_outs[0], bias)
        self.mlp_k = MLP(activate, d_ins[1], d_hiddens[1], d_outs[1], bias)
        self.mlp_d = MLP(activate, d_ins[2], d_hiddens[2], d_outs[2], bias)
        self.dropout_l = nn.Dropout(p=dropouts[0])
        self.dropout_k = nn.Dropout(p=dropouts[1])
        self.dropout_d = nn.Dropout(p=dropouts[2])
        dims = d_ins if ln_first else d_outs
        self.ln_l = nn.LayerNorm(dims[0], eps=1e-6)
        self.ln_k = nn.LayerNorm(dims[1], eps=1e-6)
        self.ln_d = nn.LayerNorm(dims[2], eps=1e-6)
        self.ln_fist = ln_first          # attribute name as in the reference (MLPProcess.py:43)
        self.res_project = res_project
        self.activate = activate
        if not res_project:
            for a in range(3):
                assert d_ins[a] == d_outs[a], \
                    "Error from MLPsBlock: If using projection for residual, d_in should be equal to d_out."
        else:
            self.res_projection_l = nn.Linear(d_ins[0], d_outs[0], bias=False)
            self.res_projection_k = nn.Linear(d_ins[1], d_outs[1], bias=False)
            self.res_projection_d = nn.Linear(d_ins[2], d_outs[2], bias=False)
        self._warned = False

    def _fusable(self, mask):
        drop = self.training and any(d.p > 0 for d in (self.dropout_l, self.dropout_k, self.dropout_d))
        return mask is None and not drop and self.activate in _ACT_IDS

    def _mix(self, x, axis, ax, prev_ln=None, ws=None):
        """prev_ln: the LayerNorm whose unmodified output x is (ln_last order); the kernels then bound |x| from its
        parameters instead of reading x once more."""
        mlp, ln = getattr(self, "mlp_" + ax), getattr(self, "ln_" + ax)
        wres = getattr(self, "res_projection_" + ax).weight if self.res_project else None
        pw = pb = None
        if prev_ln is not None and not self.ln_fist and prev_ln.weight is not None:
            pw, pb = prev_ln.weight, prev_ln.bias
        return _AxisMix.apply(x, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias, wres, ln.weight, ln.bias,
                              axis, self.ln_fist, _ACT_IDS[self.activate], pw, pb, ws)

    def forward(self, x, mask=None, _prev_ln=None, _ws=None):
        """_prev_ln (internal, set by MLPEncoder): x is the unmodified output of that LayerNorm.  _ws (internal): the
        prepared workspaces of this block's sequence and channel mix."""
        if mask is not None:
            print("Warning from MLPsBlock: If using mask, d_in should be equal to d_out.")   # MLPProcess.py:56-57
        if self._fusable(mask):
            x = self._mix(x, 1, "l", _prev_ln, _ws[0] if _ws else None)
            x = self._mix(x, 2, "k")
            return self._mix(x, 3, "d", self.ln_k, _ws[1] if _ws else None)
        if not self._warned:
            warnings.warn("MLPsBlock: dropout>0 in training / mask / this activation are outside the fused CubeMLP "
                          "kernels; running the reference op order with torch ops")
            self._warned = True
        return self._forward_torch(x, mask)

    def _forward_torch(self, x, mask):
        """Op order of MLPProcess.py:64-122 for the settings the fused kernels do not cover."""
        plan = (("l", (0, 2, 3, 1), (0, 3, 1, 2)), ("k", (0, 1, 3, 2), (0, 1, 3, 2)), ("d", None, None))
        for ax, fwd, back in plan:
            mlp, ln, drop = getattr(self, "mlp_" + ax), getattr(self, "ln_" + ax), getattr(self, "dropout_" + ax)
            xp = x.permute(*fwd) if fwd else x
            res = getattr(self, "res_projection_" + ax)(xp) if self.res_project else xp
            h = mlp(ln(xp) if self.ln_fist else xp)
            if back:
                h, res = h.permute(*back), res.permute(*back)
            if ax == "l" and mask is not None:
                h = h.masked_fill(mask.unsqueeze(-1).unsqueeze(-1).bool(), 0.0)
            x = drop(h) + res
            if not self.ln_fist:
                x = ln(x.permute(*fwd)).permute(*back) if fwd else ln(x)
        return x


class MLPEncoder(nn.Module):
    """MLPProcess.py:126-137."""

    def __init__(self, activate, d_in, d_hiddens, d_outs, dropouts, bias, ln_first=False,
                 res_project=[False, False, True]):
        super().__init__()
        assert len(d_hiddens) == len(d_outs) == len(res_project)
        self.layers_stack = nn.ModuleList([
            MLPsBlock(activate=activate, d_ins=d_in if i == 0 else d_outs[i - 1], d_hiddens=d_hiddens[i],
                      d_outs=d_outs[i], dropouts=dropouts, bias=bias, ln_first=ln_first, res_project=res_project[i])
            for i in range(len(d_hiddens))])

    def forward(self, x, mask=None):
        ws = self._prepare(x, mask)
        prev = None
        for i, enc_layer in enumerate(self.layers_stack):
            fused = enc_layer._fusable(mask) and not enc_layer.ln_fist
            x = enc_layer(x, mask, _prev_ln=prev, _ws=ws[i] if ws else None) if fused else enc_layer(x, mask)
            prev = enc_layer.ln_d if fused else None          # the block's output is ln_d's output (MLPProcess.py:120)
        return x

    def _prepare(self, x, mask):
        """Weight split and operand scales of EVERY tensor-core mix of the encoder in one launch (every mix but the first
        depends on weights and on the LayerNorm bound of its input only): one zero-fill and one kernel instead of a memset
        and a preparation launch per mix.  Returns [(ws_l, ws_d)] per block, or None when the plan does not apply."""
        if not (USE_TC and x.is_cuda and x.dim() == 4 and len(self.layers_stack) * 2 <= 8):
            return None
        bs, Ld, Kd, Dd = x.shape
        plan = []          # (A, H, A2, n_cols, mlp, wres, ln, prev_ln, inner, act)
        prev = None
        for blk in self.layers_stack:
            if not blk._fusable(mask) or blk.ln_fist:
                return None
            for ax in "lkd":
                mlp, ln = getattr(blk, "mlp_" + ax), getattr(blk, "ln_" + ax)
                A, H, A2 = mlp.fc1.in_features, mlp.fc1.out_features, mlp.fc2.out_features
                n_cols = bs * Ld * Kd * Dd // A
                if ax != "k":
                    if n_cols < 1024 or not L.lib.mimrl_cubemlp_tc_supported(A, H, A2, 0, _ACT_IDS[blk.activate]):
                        return None
                    wres = getattr(blk, "res_projection_" + ax).weight if blk.res_project else None
                    plan.append((A, H, A2, n_cols, mlp, wres, ln, prev if ax == "l" else blk.ln_k,
                                 Kd * Dd if ax == "l" else 1, _ACT_IDS[blk.activate]))
                if ax == "l":
                    Ld = A2
                elif ax == "k":
                    Kd = A2
                else:
                    Dd = A2
            prev = blk.ln_d
        import ctypes
        n = len(plan)
        sizes = [-(-L.lib.mimrl_cubemlp_tc_workspace_bytes(A, H, A2) // 256) * 256 for A, H, A2, *_ in plan]
        ws_all = torch.zeros(sum(sizes), dtype=torch.uint8, device=x.device)
        slices = list(torch.split(ws_all, sizes))
        f = lambda t: None if t is None else L.ptr(L.f32(t.detach()))
        arr_p = lambda vals: (ctypes.c_void_p * n)(*vals)
        arr_i = lambda vals: (ctypes.c_int * n)(*vals)
        x32 = L.f32(x)
        L.check(L.lib.mimrl_cubemlp_prep_many(
            n, L.ptr(x32), (ctypes.c_longlong * n)(*[p_[3] for p_ in plan]), arr_i([p_[0] for p_ in plan]),
            arr_i([p_[1] for p_ in plan]), arr_i([p_[2] for p_ in plan]), arr_i([p_[8] for p_ in plan]),
            arr_i([p_[9] for p_ in plan]),
            arr_p([f(p_[4].fc1.weight) for p_ in plan]), arr_p([f(p_[4].fc1.bias) for p_ in plan]),
            arr_p([f(p_[4].fc2.weight) for p_ in plan]), arr_p([f(p_[5]) for p_ in plan]),
            arr_p([f(p_[6].weight) for p_ in plan]),
            arr_p([f(p_[7].weight) if p_[7] is not None else None for p_ in plan]),
            arr_p([f(p_[7].bias) if p_[7] is not None else None for p_ in plan]),
            arr_i([p_[7].weight.numel() if p_[7] is not None else 0 for p_ in plan]),
            arr_p([L.ptr(s_) for s_ in slices]), L.stream()))
        return [(slices[2 * i], slices[2 * i + 1]) for i in range(len(self.layers_stack))]
