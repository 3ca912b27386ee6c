"""Variational MI estimators: drop-in for the reference's ``VMI.py``.

Same names, constructor arguments, parameter paths and error behaviour as
VMI.py:13-250, so ``Model.py:10`` (``from VMI import CriticModel, BaselineModel,
dv_lower_bound, ...``) can import from here unchanged.  The arithmetic runs in
the sm_100a kernels of ``libmimrl_b200.so``:

* ``separable_bound`` is the fused fast path used by ``VMIEstimator``
  (mimrl_b200/model.py): score sweep + bound reductions + both gradient sweeps
  without the B x B matrix ever reaching HBM, sharded by row blocks over ranks.
* the free ``*_lower_bound(scores)`` functions keep the reference's
  materialised-matrix signatures (VMI.py:136-250) and run the row-statistics
  and gradient kernels over the given matrix.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import rowblock as RB
from .linear import linear, mlp_apply

_ACTIVATIONS = {  # Utils.py:70-82 get_activation
    "elu": nn.ELU, "gelu": nn.GELU, "hardshrink": nn.Hardshrink, "hardtanh": nn.Hardtanh,
    "leakyrelu": nn.LeakyReLU, "prelu": nn.PReLU, "relu": nn.ReLU, "rrelu": nn.RReLU, "tanh": nn.Tanh,
}


def get_activation(activation):
    return _ACTIVATIONS[activation]


def mlps(dim, hidden_dim, output_dim, layers, activation):
    """VMI.py:13-22: Linear+act, `layers` x (Linear+act), Linear."""
    act = get_activation(activation)
    seq = [nn.Linear(dim, hidden_dim), act()]
    for _ in range(layers):
        seq += [nn.Linear(hidden_dim, hidden_dim), act()]
    seq += [nn.Linear(hidden_dim, output_dim)]
    return nn.Sequential(*seq)


def _zero_biases(stack):
    """VMI.py:47-51 init_mlp_params: weights keep torch's default init, biases are zeroed."""
    for layer in stack:
        if isinstance(layer, nn.Linear):
            nn.init.constant_(layer.bias, 0)


# --------------------------------------------------------------------------
# kernels behind autograd
# --------------------------------------------------------------------------


def _family(bound_id):
    fam, inc, fl = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.check(L.lib.mimrl_bound_weight_family(bound_id, ctypes.byref(fam), ctypes.byref(inc), ctypes.byref(fl)))
    return fam.value, inc.value, fl.value


def _grad_pair(g_mi, g_loss, like):
    z = like.new_zeros(())
    return torch.stack([g_mi if g_mi is not None else z, g_loss if g_loss is not None else z]).float().contiguous()


class _SeparableBound(torch.autograd.Function):
    """(x_emb [n,E] = g(x), y_emb [n,E] = h(y), log_baseline [n,1] | None) -> (mi, mi_loss).

    Rows of the score matrix index y (VMI.py:57: ``y_ @ x_.T``).  Under a
    sharded RowBlock each rank passes its own rows and receives gradients for
    its own rows; embeddings and per-row statistics are all-gathered inside."""

    @staticmethod
    def forward(ctx, x_emb, y_emb, log_baseline, bound_id, impl, rb):
        x_emb, y_emb = L.f32(x_emb), L.f32(y_emb)
        n_own, embed = y_emb.shape
        if rb is None:
            rb = RB.single(n_own)
        fam, inc, flags = _family(bound_id)
        all_x = RB.all_gather_rows(x_emb, rb)
        n_all = all_x.shape[0]
        st = L.stream()
        ws_bytes = L.lib.mimrl_sep_workspace_bytes(n_own, n_all, embed)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=y_emb.device)
        stats = torch.empty(4, n_own, dtype=torch.float32, device=y_emb.device)  # max, sum, softplus, diag
        L.check(L.lib.mimrl_sep_row_stats(L.ptr(y_emb), L.ptr(all_x), n_own, n_all, embed, rb.offset, flags, impl,
                                          L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(stats[3]),
                                          L.ptr(ws), ws.numel(), st))
        base = None
        if log_baseline is not None:
            base = RB.all_gather_rows(L.f32(log_baseline).reshape(-1), rb)
        all_stats = RB.all_gather_rows(stats.t().contiguous(), rb).t().contiguous() if rb.sharded else stats
        result = torch.zeros(16, dtype=torch.float32, device=y_emb.device)
        L.check(L.lib.mimrl_bound_finalize(bound_id, L.ptr(all_stats[0]), L.ptr(all_stats[1]), L.ptr(all_stats[2]),
                                           L.ptr(all_stats[3]), L.ptr(base), n_all, L.ptr(result), st))
        ctx.save_for_backward(x_emb, y_emb, all_x, all_stats, result, base if base is not None else result.new_empty(0))
        ctx.cfg = (bound_id, impl, rb, fam, inc, log_baseline is not None, ws)
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(ctx, g_mi, g_loss):
        x_emb, y_emb, all_x, all_stats, result, base = ctx.saved_tensors
        bound_id, impl, rb, fam, inc, has_base, ws = ctx.cfg
        base = base if has_base else None
        n_own, embed = y_emb.shape
        n_all = all_x.shape[0]
        dev = y_emb.device
        st = L.stream()
        grad = _grad_pair(g_mi, g_loss, result)
        coef = torch.empty(1, dtype=torch.float32, device=dev)
        vec = torch.empty(3, n_all, dtype=torch.float32, device=dev)            # shift, dcoef, dbaseline
        L.check(L.lib.mimrl_bound_backward_coef(bound_id, L.ptr(result), L.ptr(grad), L.ptr(all_stats[0]),
                                                L.ptr(all_stats[1]), L.ptr(all_stats[3]), L.ptr(base), n_all,
                                                L.ptr(coef), L.ptr(vec[0]), L.ptr(vec[1]),
                                                L.ptr(vec[2]) if has_base else None, st))
        own = slice(rb.offset, rb.offset + n_own)
        shift_own, dcoef_own = vec[0, own].contiguous(), vec[1, own].contiguous()
        # d/d h(y): rows are owned, x is swept, shift indexed by the owned row
        swept = all_x
        if bound_id == L.BOUND_IDS["infonce"]:
            # InfoNCE rows: sum_j P_ij x_j - x_i with sum_j P_ij = 1 cancels the common mean of the x embeddings.
            # Sweep the CENTRED embeddings instead (S_ij = y_i.(x_j - mu) + y_i.mu, the row constant moves into the
            # shift): algebraically identical, but the cancellation no longer happens in floating point.
            mu = all_x.mean(dim=0)
            swept = (all_x - mu).contiguous()
            shift_own = (shift_own - (y_emb * mu).sum(dim=1)).contiguous()
        dy = torch.empty_like(y_emb)
        L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(y_emb), L.ptr(swept), n_own, n_all, embed, rb.offset, fam, inc,
                                             L.ptr(shift_own), 0, L.ptr(coef), L.ptr(dcoef_own), impl, L.ptr(dy),
                                             L.ptr(ws), ws.numel(), st))
        # d/d g(x): columns are owned, y is swept, shift indexed by the swept row
        all_y = RB.all_gather_rows(y_emb, rb)
        dx = torch.empty_like(x_emb)
        L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(x_emb), L.ptr(all_y), n_own, n_all, embed, rb.offset, fam, inc,
                                             L.ptr(vec[0]), 1, L.ptr(coef), L.ptr(dcoef_own), impl, L.ptr(dx),
                                             L.ptr(ws), ws.numel(), st))
        dbase = vec[2, own].reshape(n_own, 1).clone() if has_base else None
        return dx, dy, dbase, None, None, None


FUSED_FORWARD = True      # exp-family bounds: statistics and the owned-row gradient sum from ONE sweep (see below)
_FUSED_BOUNDS = ("dv", "mine", "tuba", "nwj", "infonce")


class _SeparableBoundFused(torch.autograd.Function):
    """Same contract as ``_SeparableBound`` for the exp-family bounds when a gradient is wanted.

    The backward of these bounds weights row i's swept embeddings by exp(S_ij - shift_i); up to a per-row factor
    that is exp(S_ij - ref_i) for ANY reference point ref_i.  So the forward is ONE sweep (mimrl_sep_online_forward):
    the kernel keeps a running reference per row (online softmax, accumulators rescaled lazily) and returns the
    reference, the exact row statistic sum_{j != i} exp(S_ij - ref_i) and O_i = sum_j exp(S_ij - ref_i) x_j.  The
    backward rescales O_i and only has to sweep for the swept side: 2 + 2 tensor-core units instead of 1 + 2 + 2."""

    @staticmethod
    def forward(ctx, x_emb, y_emb, log_baseline, bound_id, impl, rb):
        x_emb, y_emb = L.f32(x_emb), L.f32(y_emb)
        n_own, embed = y_emb.shape
        if rb is None:
            rb = RB.single(n_own)
        fam, inc, flags = _family(bound_id)
        all_x = RB.all_gather_rows(x_emb, rb)
        n_all = all_x.shape[0]
        dev = y_emb.device
        st = L.stream()
        ws = torch.empty(max(L.lib.mimrl_sep_workspace_bytes(n_own, n_all, embed), 16), dtype=torch.uint8, device=dev)
        # InfoNCE: sweep the CENTRED embeddings (see _SeparableBound.backward); S_ij = y_i.(x_j - mu) + y_i.mu
        centred = bound_id == L.BOUND_IDS["infonce"]
        if centred:
            mu = all_x.mean(dim=0)
            swept = (all_x - mu).contiguous()
            ymu = (y_emb * mu).sum(dim=1)
        else:
            swept, ymu = all_x, None
        # ONE sweep, no reference point supplied: the kernel keeps a running reference per row (online softmax with
        # lazily rescaled accumulators) and returns it together with the sums referred to it
        wsum = torch.empty(n_own, embed, dtype=torch.float32, device=dev)
        stats = torch.zeros(4, n_own, dtype=torch.float32, device=dev)      # max, sum, softplus, diag (true S units)
        ref = torch.empty(n_own, dtype=torch.float32, device=dev)
        diag = torch.empty(n_own, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_sep_online_forward(L.ptr(y_emb), L.ptr(swept), n_own, n_all, embed, rb.offset, inc, L.ptr(ref),
                                               L.ptr(wsum), L.ptr(stats[1]), L.ptr(diag), L.ptr(ws), ws.numel(), st))
        pre = (None, None, None, diag)
        stats[0] = ref if ymu is None else ref + ymu
        stats[3] = pre[3] if ymu is None else pre[3] + ymu
        base = None
        if log_baseline is not None:
            base = RB.all_gather_rows(L.f32(log_baseline).reshape(-1), rb)
        all_stats = RB.all_gather_rows(stats.t().contiguous(), rb).t().contiguous() if rb.sharded else stats
        result = torch.zeros(16, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_bound_finalize(bound_id, L.ptr(all_stats[0]), L.ptr(all_stats[1]), L.ptr(all_stats[2]),
                                           L.ptr(all_stats[3]), L.ptr(base), n_all, L.ptr(result), st))
        own_rows = swept[rb.offset: rb.offset + n_own]
        ctx.save_for_backward(x_emb, y_emb, all_stats, result, base if base is not None else result.new_empty(0), wsum,
                              own_rows)
        ctx.cfg = (bound_id, impl, rb, fam, inc, log_baseline is not None, ws)
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(ctx, g_mi, g_loss):
        x_emb, y_emb, all_stats, result, base, wsum, own_rows = ctx.saved_tensors
        bound_id, impl, rb, fam, inc, has_base, ws = ctx.cfg
        base = base if has_base else None
        n_own, embed = y_emb.shape
        n_all = all_stats.shape[1]
        dev = y_emb.device
        st = L.stream()
        grad = _grad_pair(g_mi, g_loss, result)
        coef = torch.empty(1, dtype=torch.float32, device=dev)
        vec = torch.empty(3, n_all, dtype=torch.float32, device=dev)            # shift, dcoef, dbaseline
        L.check(L.lib.mimrl_bound_backward_coef(bound_id, L.ptr(result), L.ptr(grad), L.ptr(all_stats[0]),
                                                L.ptr(all_stats[1]), L.ptr(all_stats[3]), L.ptr(base), n_all,
                                                L.ptr(coef), L.ptr(vec[0]), L.ptr(vec[1]),
                                                L.ptr(vec[2]) if has_base else None, st))
        own = slice(rb.offset, rb.offset + n_own)
        # d/d h(y): the forward's weighted sum, moved from its reference point to the final shift
        dy = None
        if ctx.needs_input_grad[1]:
            scale = coef * torch.exp(all_stats[0, own] - vec[0, own])
            dy = scale[:, None] * wsum + vec[1, own][:, None] * own_rows
        # d/d g(x): columns are owned, y is swept, shift indexed by the swept row.  Under sharding every rank must take
        # part in the gather, so only the sweep itself is skipped when x needs no gradient.
        all_y = RB.all_gather_rows(y_emb, rb)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x_emb)
            dcoef_own = vec[1, own].contiguous()
            L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(x_emb), L.ptr(all_y), n_own, n_all, embed, rb.offset, fam, inc,
                                                 L.ptr(vec[0]), 1, L.ptr(coef), L.ptr(dcoef_own), impl, L.ptr(dx),
                                                 L.ptr(ws), ws.numel(), st))
        dbase = vec[2, own].reshape(n_own, 1).clone() if has_base else None
        return dx, dy, dbase, None, None, None


def separable_bound(x_emb, y_emb, bound_type, log_baseline=None, rowblock=None, impl=L.IMPL_AUTO):
    """Fused ``bound(h(y) @ g(x).T)`` for every bound with a single-sweep form
    (all of VMI.py:136-198 plus the MINE branch of Model.py:121-124).
    Returns ``(mi, mi_loss)`` exactly as ``VMIEstimator.forward`` does."""
    if bound_type not in L.BOUND_IDS:
        raise NotImplementedError
    if (FUSED_FORWARD and bound_type in _FUSED_BOUNDS and torch.is_grad_enabled()
            and (x_emb.requires_grad or y_emb.requires_grad)):
        n_own = y_emb.shape[0]
        n_all = rowblock.n_all if rowblock is not None else n_own
        if L.lib.mimrl_sep_selected_impl(n_own, n_all, y_emb.shape[1], impl) == L.IMPL_TCGEN05:
            return _SeparableBoundFused.apply(x_emb, y_emb, log_baseline, L.BOUND_IDS[bound_type], impl, rowblock)
    return _SeparableBound.apply(x_emb, y_emb, log_baseline, L.BOUND_IDS[bound_type], impl, rowblock)


class _SeparableInterp(torch.autograd.Function):
    """Interpolated bound (VMI.py:201-250, alpha = sigmoid(alpha_logit)) over the separable critic, fused: the score
    matrix only exists as TMEM tiles.  With L_i = logsumexp_j S_ij, p_ij = exp(S_ij - L_i), a_i the learnt log-baseline,
    c_i = logit(alpha) + L_i - log(n-1) - a_i, C_i = exp(c_i), sigma_i = C_i / (1 + C_i) the reference's quantities are

        I_ij  = log(alpha * loo-mean_i,j + (1-alpha) e^{a_i}) = log(1-alpha) + a_i + log(1 + C_i) + log(1 - sigma_i p_ij)
        joint = sum_{i != j} (S_jj - I_ij) / (n(n-1)),   marg = sum_{i != j} exp(S_ij - I_jj) / (n(n-1)),   mi = 1 + joint - marg

    Forward, three statistics sweeps: rows of S (L_i, diagonal), columns of S (the operands swapped: sum_i e^{S_ij}), and
    -- once sigma_i and L_i are known -- the row sums T_i = sum_{j != i} log(1 - sigma_i p_ij), Q_i = sum_{j != i}
    p_ij / (1 - sigma_i p_ij) (mimrl_sep_interp_stats).  Backward, per side two weighted-sum sweeps: the row-parameterised
    part lambda_i p_ij + sigma_i p_ij / (1 - sigma_i p_ij) (MIMRL_WEIGHT_INTERP) and the column-shifted part
    -exp(S_ij - I_jj) (MIMRL_WEIGHT_EXP).  Per-row vectors are handled in float64 torch ops (n numbers each).
    Shardable by row blocks like every other bound: per-row vectors are all-gathered."""

    @staticmethod
    def forward(ctx, x_emb, y_emb, log_baseline, alpha_logit, impl, rb):
        x_emb, y_emb = L.f32(x_emb), L.f32(y_emb)
        n_own, embed = y_emb.shape
        if rb is None:
            rb = RB.single(n_own)
        dev = y_emb.device
        st = L.stream()
        all_x = RB.all_gather_rows(x_emb, rb)
        all_y = RB.all_gather_rows(y_emb, rb)
        n = all_x.shape[0]
        ws = torch.empty(max(L.lib.mimrl_sep_workspace_bytes(n_own, n, embed), 16), dtype=torch.uint8, device=dev)
        own = slice(rb.offset, rb.offset + n_own)

        def stats(own_rows, swept):
            out = torch.empty(4, n_own, dtype=torch.float32, device=dev)
            L.check(L.lib.mimrl_sep_row_stats(L.ptr(own_rows), L.ptr(swept), n_own, n, embed, rb.offset, 0, impl,
                                              L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]), L.ptr(ws),
                                              ws.numel(), st))
            return out.double()
        r = stats(y_emb, all_x)                                   # rows i of S = y_i . x_j
        c = stats(x_emb, all_y)                                   # columns j of S: own = x_j, swept = y_i
        d = r[3]
        lse_off = r[0] + torch.log(r[1])                          # off-diagonal part; -inf when n == 1
        lse = torch.logaddexp(lse_off, d)
        a = L.f32(log_baseline).reshape(-1).double()
        log_alpha = -math.log1p(math.exp(-float(alpha_logit)))
        log_1m_alpha = -math.log1p(math.exp(float(alpha_logit)))
        ci = (log_alpha - log_1m_alpha) + lse - math.log(n - 1.0) - a
        sig = torch.sigmoid(ci)
        row_par = torch.stack([lse, sig]).float().contiguous()
        qt = torch.empty(2, n_own, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_sep_interp_stats(L.ptr(y_emb), L.ptr(all_x), n_own, n, embed, rb.offset, L.ptr(row_par[0]),
                                             L.ptr(row_par[1]), L.ptr(qt[0]), L.ptr(qt[1]), L.ptr(ws), ws.numel(), st))
        Q, T = qt[0].double(), qt[1].double()
        p_dd = torch.exp(d - lse)
        # (VMI.py:219-223 substitutes d = 1 where an entry carries its whole row to fp32 precision, lse - S == 0; the exact
        # limit log(1 - sigma_i) is evaluated here instead -- DESIGN.md, quirk N9)
        x_dd = (sig * p_dd).clamp(max=1.0 - 1e-12)
        base_i = log_1m_alpha + a + F.softplus(ci)                # log(1-alpha) + a_i + log(1 + C_i)
        I_dd = base_i + torch.log1p(-x_dd)
        col_log = c[0] + torch.log(c[1])                          # log sum_{i != j} e^{S_ij} for the own columns j
        Mj = torch.exp(col_log - I_dd)
        nn_ = n * (n - 1.0)
        part = torch.stack([((n - 1.0) * d - (n - 1.0) * base_i - T - Mj).sum()])
        if rb.sharded:
            torch.distributed.all_reduce(part, group=rb.group)
        mi = (1.0 + part[0] / nn_).float()
        ctx.save_for_backward(x_emb, y_emb, all_x, all_y, lse, sig, Q, d, I_dd, Mj, c[0], p_dd)
        ctx.cfg = (impl, rb, n, ws)
        return mi, -mi

    @staticmethod
    def backward(ctx, g_mi, g_loss):
        x_emb, y_emb, all_x, all_y, lse, sig, Q, d, I_dd, Mj, col_max, p_dd = ctx.saved_tensors
        impl, rb, n, ws = ctx.cfg
        n_own, embed = y_emb.shape
        dev = y_emb.device
        st = L.stream()
        g = (g_mi if g_mi is not None else 0.0) - (g_loss if g_loss is not None else 0.0)
        g = torch.as_tensor(g, dtype=torch.float64, device=dev)
        nn_ = n * (n - 1.0)
        r_dd = p_dd / (1.0 - (sig * p_dd).clamp(max=1.0 - 1e-12))        # r_ii = p_ii / (1 - sigma_i p_ii)
        lam = -(n - 1.0) * sig - sig * sig * Q + Mj * sig * (1.0 + sig * r_dd)           # coefficient of dL_i
        diag_c = (n - 1.0) + lam * p_dd - Mj * sig * r_dd                                # coefficient of dS_ii
        da = -(n - 1.0) * (1.0 - sig) - sig * (1.0 - sig) * Q + Mj * (1.0 - sig) * (1.0 + sig * r_dd)
        # all-gather the per-row vectors both sweeps need
        def allrows(v):
            return RB.all_gather_rows(v.float().contiguous(), rb).double() if rb.sharded else v
        lam_all, sig_all, lse_all, I_all = allrows(lam), allrows(sig), allrows(lse), allrows(I_dd)
        colmax_all = allrows(col_max)
        # scale of the row-parameterised weights: |lambda_i p + sigma_i p / (1 - sigma_i p)| <= |lambda_i| + sigma_i / (1 - sigma_i)
        Lam = (lam_all.abs() + sig_all / (1.0 - sig_all).clamp(min=1e-12)).max().clamp(min=1e-30)
        par_all = torch.stack([lse_all, lam_all / Lam, sig_all / Lam, sig_all]).float().contiguous()     # [4][n]
        par_own = par_all[:, rb.offset: rb.offset + n_own].contiguous()
        # column-shifted exp weights, referred to their global maximum so that they stay <= 1
        wmax = (colmax_all - I_all).max()
        shift_all = (I_all + wmax).float().contiguous()
        shift_own = shift_all[rb.offset: rb.offset + n_own].contiguous()
        coef_i = (g * Lam / nn_).float().reshape(1)
        coef_e = (-g * torch.exp(wmax) / nn_).float().reshape(1)
        dcoef = (g * diag_c / nn_).float().contiguous()
        zero = torch.zeros(n_own, dtype=torch.float32, device=dev)

        def wsum(own_rows, swept, fam, shift, by_swept, coef, dc):
            out = torch.empty(n_own, embed, dtype=torch.float32, device=dev)
            L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(own_rows), L.ptr(swept), n_own, n, embed, rb.offset, fam, 0, L.ptr(shift),
                                                 by_swept, L.ptr(coef), L.ptr(dc), impl, L.ptr(out), L.ptr(ws), ws.numel(), st))
            return out
        dy = dx = None
        if ctx.needs_input_grad[1]:
            dy = wsum(y_emb, all_x, L.WEIGHT_INTERP, par_own, 0, coef_i, dcoef) + \
                wsum(y_emb, all_x, L.WEIGHT_EXP, shift_all, 1, coef_e, zero)
        if ctx.needs_input_grad[0]:
            dx = wsum(x_emb, all_y, L.WEIGHT_INTERP, par_all, 1, coef_i, dcoef) + \
                wsum(x_emb, all_y, L.WEIGHT_EXP, shift_own, 0, coef_e, zero)
        dbase = (g * da / nn_).float().reshape(n_own, 1) if ctx.needs_input_grad[2] else None
        return dx, dy, dbase, None, None, None


def separable_interp_bound(x_emb, y_emb, log_baseline, alpha_logit, rowblock=None, impl=L.IMPL_AUTO):
    """Fused ``interp_lower_bound(h(y) @ g(x).T, log_baseline, alpha_logit)`` -> (mi, mi_loss)."""
    return _SeparableInterp.apply(x_emb, y_emb, log_baseline, alpha_logit, impl, rowblock)


class _ScoresBound(torch.autograd.Function):
    """Bound over a materialised score matrix (API parity with VMI.py:136-198).

    Single rank: scores is [n, n].  Under a sharded RowBlock each rank passes its own rows [n_own, n_all] of the
    global matrix; per-row statistics are all-gathered so every rank obtains the global-batch value and the
    gradient of its own rows."""

    @staticmethod
    def forward(ctx, scores, log_baseline, bound_id, rb=None):
        scores = L.f32(scores)
        if scores.dim() != 2:
            raise ValueError("scores must be a [batch, batch] matrix")
        n_own, n_all = scores.shape
        if rb is None:
            rb = RB.single(n_own)
        if n_all != rb.n_all or n_own != rb.n_own:
            raise ValueError("scores must be a square [batch, batch] matrix (or this rank's rows of it)")
        fam, inc, flags = _family(bound_id)
        st = L.stream()
        stats = torch.empty(4, n_own, dtype=torch.float32, device=scores.device)
        L.check(L.lib.mimrl_scores_row_stats(L.ptr(scores), n_own, n_all, rb.offset, flags, L.ptr(stats[0]),
                                             L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(stats[3]), st))
        base = None
        if log_baseline is not None:
            base = RB.all_gather_rows(L.f32(log_baseline).reshape(-1), rb)
        all_stats = RB.all_gather_rows(stats.t().contiguous(), rb).t().contiguous() if rb.sharded else stats
        result = torch.zeros(16, dtype=torch.float32, device=scores.device)
        L.check(L.lib.mimrl_bound_finalize(bound_id, L.ptr(all_stats[0]), L.ptr(all_stats[1]), L.ptr(all_stats[2]),
                                           L.ptr(all_stats[3]), L.ptr(base), n_all, L.ptr(result), st))
        ctx.save_for_backward(scores, all_stats, result, base if base is not None else result.new_empty(0))
        ctx.cfg = (bound_id, fam, inc, base is not None, rb)
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(ctx, g_mi, g_loss):
        scores, stats, result, base = ctx.saved_tensors
        bound_id, fam, inc, has_base, rb = ctx.cfg
        base = base if has_base else None
        n_own, n_all = scores.shape
        st = L.stream()
        grad = _grad_pair(g_mi, g_loss, result)
        coef = torch.empty(1, dtype=torch.float32, device=scores.device)
        vec = torch.empty(3, n_all, dtype=torch.float32, device=scores.device)
        L.check(L.lib.mimrl_bound_backward_coef(bound_id, L.ptr(result), L.ptr(grad), L.ptr(stats[0]), L.ptr(stats[1]),
                                                L.ptr(stats[3]), L.ptr(base), n_all, L.ptr(coef), L.ptr(vec[0]),
                                                L.ptr(vec[1]), L.ptr(vec[2]) if has_base else None, st))
        own = slice(rb.offset, rb.offset + n_own)
        shift_own, dcoef_own = vec[0, own].contiguous(), vec[1, own].contiguous()
        g = torch.empty_like(scores)
        L.check(L.lib.mimrl_scores_grad(L.ptr(scores), n_own, n_all, rb.offset, fam, inc, L.ptr(shift_own), L.ptr(coef),
                                        L.ptr(dcoef_own), L.ptr(g), st))
        return g, (vec[2, own].reshape(n_own, 1).clone() if has_base else None), None, None


def _scores_bound(scores, bound_type, log_baseline=None, rowblock=None):
    return _ScoresBound.apply(scores, log_baseline, L.BOUND_IDS[bound_type], rowblock)


class _GatherRows(torch.autograd.Function):
    """All-gather of row blocks that is differentiable: the backward sums every rank's gradient of the gathered
    matrix and keeps the rows this rank owns (the swept operand of a row-block sharded score matrix)."""

    @staticmethod
    def forward(ctx, local, rb):
        ctx.rb = rb
        return RB.all_gather_rows(local, rb)

    @staticmethod
    def backward(ctx, g):
        rb = ctx.rb
        g = g.contiguous()
        if rb.sharded:
            torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.SUM, group=rb.group)
        return RB.own_slice(g, rb).contiguous(), None


def gather_rows(local, rowblock):
    if rowblock is None or not rowblock.sharded:
        return local
    return _GatherRows.apply(local, rowblock)


# --------------------------------------------------------------------------
# reference-named free functions (VMI.py:136-250)
# --------------------------------------------------------------------------


def dv_lower_bound(scores):
    return _scores_bound(scores, "dv")[0]


def mine_lower_bound_test(scores):
    """VMI.py:142-145: (dv bound, diag(scores), exp of the off-diagonal entries)."""
    mi = _scores_bound(scores, "dv")[0]
    n = scores.size(0)
    et = torch.exp(scores).masked_fill(torch.eye(n, dtype=torch.bool, device=scores.device), 0.0)
    return mi, scores.diag(), et


def tuba_lower_bound(scores, log_baseline=None):
    if log_baseline is not None and not torch.is_tensor(log_baseline):
        scores = scores - log_baseline          # VMI.py:151 accepts a python scalar too
        log_baseline = None
    return _scores_bound(scores, "tuba", log_baseline)[0]


def nwj_lower_bound(scores):
    return _scores_bound(scores, "nwj")[0]


def infonce_lower_bound(scores):
    return _scores_bound(scores, "infonce")[0]


def js_fgan_lower_bound(scores):
    return _scores_bound(scores, "js_fgan")[0]


def js_lower_bound(scores):
    return _scores_bound(scores, "js")[0]


def smile_lower_bound(scores, clip=None):
    return _scores_bound(scores, "smile")[0]     # VMI.py:186 pins clip to 1 whatever is passed


def log_interpolate(log_a, log_b, alpha_logit: float):
    """VMI.py:201-210: log(alpha*a + (1-alpha)*b), alpha = sigmoid(alpha_logit)."""
    alpha_logit = float(alpha_logit)
    log_alpha = -math.log1p(math.exp(-alpha_logit))
    log_1m_alpha = -math.log1p(math.exp(alpha_logit))
    return torch.logaddexp(log_alpha + log_a, log_1m_alpha + log_b)


def compute_log_loomean(scores):
    """VMI.py:213-226: log of the leave-one-out mean of exp(scores) along dim 1."""
    lse = torch.logsumexp(scores, dim=1, keepdim=True)
    d = lse - scores
    safe = torch.where(d == 0, torch.ones_like(d), d)
    return scores + safe + torch.log(-torch.expm1(-safe)) - math.log(scores.size(1) - 1.0)


def interp_lower_bound(scores, baseline, alpha_logit):
    """VMI.py:229-250 (two dependent sweeps; runs on the materialised matrix)."""
    n = scores.size(0)
    n_off = n * (n - 1.0)
    ib = log_interpolate(compute_log_loomean(scores), baseline.reshape(n, 1).expand(n, n), alpha_logit)
    off = ~torch.eye(n, dtype=torch.bool, device=scores.device)
    marg = torch.exp(torch.logsumexp((scores - ib.diag()[None, :])[off], dim=0) - math.log(n_off))
    joint = ((scores.diag()[None, :] - ib) * off).sum() / n_off
    return 1 + joint - marg


# --------------------------------------------------------------------------
# concat critic, all pairs (VMI.py:58-65) on the tensor cores: csrc/concat_tc.cu
# --------------------------------------------------------------------------

CONCAT_GRAD_PAIRS = 1 << 21      # pairs per backward pass (4 KB of weight-gradient operands each)


class _ConcatPairMLP(torch.autograd.Function):
    """scores[i, j] = w4 . relu(W3 relu(W2 relu(u_i + v_j) + b2) + b3) + b4 for every pair.

    Forward: mimrl_concat_scores (activations stay in TMEM).  Backward: row chunks of mimrl_concat_grad
    (recompute + both data-gradient contractions) followed by the two split-K weight-gradient GEMMs."""

    @staticmethod
    def forward(ctx, u, v, w2, b2, w3, b3, w4, b4):
        u, w2, w3 = L.f32(u), L.f32(w2), L.f32(w3)
        vt = L.f32(v).t().contiguous()
        w4 = L.f32(w4).reshape(-1)
        n_own, n_all, hid = u.shape[0], vt.shape[1], u.shape[1]
        scores = torch.empty(n_own, n_all, device=u.device, dtype=torch.float32)
        wsb = L.lib.mimrl_concat_workspace_bytes(hid)
        ws = torch.empty(wsb, dtype=torch.uint8, device=u.device)
        L.check(L.lib.mimrl_concat_scores(L.ptr(u), L.ptr(vt), n_own, n_all, n_all, hid, L.ptr(w2), L.ptr(b2), L.ptr(w3),
                                          L.ptr(b3), L.ptr(w4), L.ptr(b4), L.ptr(scores), L.ptr(ws), wsb, L.stream()))
        ctx.save_for_backward(u, vt, w2, b2, w3, b3, w4)
        return scores

    @staticmethod
    def backward(ctx, g):
        u, vt, w2, b2, w3, b3, w4 = ctx.saved_tensors
        g = L.f32(g)
        dev = u.device
        n_own, n_all, hid = u.shape[0], vt.shape[1], u.shape[1]
        g_u = torch.zeros_like(u)
        g_vt = torch.zeros_like(vt)
        g_b2, g_b3, g_w4 = (torch.zeros(hid, device=dev) for _ in range(3))
        g_w2, g_w3 = torch.zeros_like(w2), torch.zeros_like(w3)
        rows = max(4, (CONCAT_GRAD_PAIRS // max(n_all, 1)) // 4 * 4)
        rows = min(rows, n_own)
        pair_rows = L.lib.mimrl_concat_pair_rows(rows, n_all)
        opb = L.lib.mimrl_split_bytes(hid, pair_rows)
        ops = [torch.empty(opb, dtype=torch.uint8, device=dev) for _ in range(4)]
        wsb = L.lib.mimrl_concat_workspace_bytes(hid)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        gwb = L.lib.mimrl_gemm_split_workspace_bytes(0, hid, hid, pair_rows)
        gws = torch.empty(gwb, dtype=torch.uint8, device=dev)
        tmp = torch.empty_like(w2)
        st = L.stream()
        for r0 in range(0, n_own, rows):
            r = min(rows, n_own - r0)
            pr = L.lib.mimrl_concat_pair_rows(r, n_all)
            L.check(L.lib.mimrl_concat_grad(L.ptr(u[r0:r0 + r]), L.ptr(vt), r, n_all, n_all, hid, L.ptr(w2), L.ptr(b2), L.ptr(w3),
                                            L.ptr(b3), L.ptr(w4), L.ptr(g[r0:r0 + r]), L.ptr(g_u[r0:r0 + r]), L.ptr(g_vt),
                                            L.ptr(g_b2), L.ptr(g_b3), L.ptr(g_w4), L.ptr(ops[0]), L.ptr(ops[1]), L.ptr(ops[2]),
                                            L.ptr(ops[3]), L.ptr(ws), wsb, st))
            for a, b, acc in ((ops[2], ops[0], g_w2), (ops[3], ops[1], g_w3)):
                L.check(L.lib.mimrl_gemm_split_blocked(L.ptr(a), L.ptr(b), hid, hid, pr, L.ptr(tmp), L.ptr(gws), gwb, st))
                acc += tmp
        return (g_u, g_vt.t(), g_w2, g_b2 if b2 is not None else None, g_w3, g_b3 if b3 is not None else None,
                g_w4.reshape(1, -1), g.sum().reshape(1))


class _ConcatBoundFused(torch.autograd.Function):
    """(u [n_own,256], v_all [n_all,256], diag [n_own], W2, b2, W3, b3, w4, b4, log_baseline | None) -> (mi, mi_loss) for
    the concat critic (VMI.py:58-65) under any bound of VMI.py:136-198, with neither the score matrix nor dL/dscores in
    memory: the forward kernel reduces the off-diagonal row statistics, the backward kernel forms each pair's gradient
    weight from its recomputed score.  ``diag`` are the scores of the n_own diagonal pairs (i, own_offset + i), computed
    by the caller with the same MLP as an ordinary row batch; their gradient is returned like any other input's.

    The diagonal weights of a bound are ~n times larger than the off-diagonal ones; keeping them out of the all-pairs
    sweep lets the sweep's fp16 hi/lo operands be scaled to the off-diagonal magnitude."""

    @staticmethod
    def forward(ctx, u, v_all, diag, w2, b2, w3, b3, w4, b4, log_baseline, bound_id, rb):
        u, w2, w3 = L.f32(u), L.f32(w2), L.f32(w3)
        vt = L.f32(v_all).t().contiguous()
        w4 = L.f32(w4).reshape(-1)
        n_own, n_all, hid = u.shape[0], vt.shape[1], u.shape[1]
        if rb is None:
            rb = RB.single(n_own)
        dev = u.device
        fam, inc, flags = _family(bound_id)
        st = L.stream()
        wsb = L.lib.mimrl_concat_stats_workspace_bytes(hid, n_own, n_all)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        stats = torch.zeros(4, n_own, dtype=torch.float32, device=dev)          # max, sum, softplus, diag
        L.check(L.lib.mimrl_concat_row_stats(L.ptr(u), L.ptr(vt), n_own, n_all, n_all, hid, rb.offset, flags, L.ptr(w2),
                                             L.ptr(b2), L.ptr(w3), L.ptr(b3), L.ptr(w4), L.ptr(b4), L.ptr(stats[0]),
                                             L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(ws), wsb, st))
        stats[3] = L.f32(diag).reshape(-1)
        base = None
        if log_baseline is not None:
            base = RB.all_gather_rows(L.f32(log_baseline).reshape(-1), rb)
        all_stats = RB.all_gather_rows(stats.t().contiguous(), rb).t().contiguous() if rb.sharded else stats
        result = torch.zeros(16, dtype=torch.float32, device=dev)
        L.check(L.lib.mimrl_bound_finalize(bound_id, L.ptr(all_stats[0]), L.ptr(all_stats[1]), L.ptr(all_stats[2]),
                                           L.ptr(all_stats[3]), L.ptr(base), n_all, L.ptr(result), st))
        ctx.save_for_backward(u, vt, w2, b2, w3, b3, w4, b4, all_stats, result, base if base is not None else result.new_empty(0))
        ctx.cfg = (bound_id, fam, inc, base is not None, rb)
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(ctx, g_mi, g_loss):
        u, vt, w2, b2, w3, b3, w4, b4, stats, result, base = ctx.saved_tensors
        bound_id, fam, inc, has_base, rb = ctx.cfg
        base = base if has_base else None
        dev = u.device
        n_own, n_all, hid = u.shape[0], vt.shape[1], u.shape[1]
        st = L.stream()
        grad = _grad_pair(g_mi, g_loss, result)
        coef = torch.empty(1, dtype=torch.float32, device=dev)
        vec = torch.empty(3, n_all, dtype=torch.float32, device=dev)            # shift, dcoef, dbaseline
        L.check(L.lib.mimrl_bound_backward_coef(bound_id, L.ptr(result), L.ptr(grad), L.ptr(stats[0]), L.ptr(stats[1]),
                                                L.ptr(stats[3]), L.ptr(base), n_all, L.ptr(coef), L.ptr(vec[0]),
                                                L.ptr(vec[1]), L.ptr(vec[2]) if has_base else None, st))
        own = slice(rb.offset, rb.offset + n_own)
        shift_own, dcoef_own = vec[0, own].contiguous(), vec[1, own].contiguous()
        # the diagonal pairs: dcoef, plus the pair weight itself where the bound's sum includes the diagonal (InfoNCE)
        g_diag = dcoef_own
        if inc:
            d_own = stats[3, own]
            g_diag = g_diag + coef * (torch.exp(d_own - shift_own) if fam == L.WEIGHT_EXP else torch.sigmoid(d_own))
        g_u = torch.zeros_like(u)
        g_vt = torch.zeros_like(vt)
        g_b2, g_b3, g_w4 = (torch.zeros(hid, device=dev) for _ in range(3))
        g_b4 = torch.zeros(1, device=dev)
        g_w2, g_w3 = torch.zeros_like(w2), torch.zeros_like(w3)
        rows = max(4, (CONCAT_GRAD_PAIRS // max(n_all, 1)) // 4 * 4)
        rows = min(rows, n_own)
        opb = L.lib.mimrl_split_bytes(hid, L.lib.mimrl_concat_pair_rows(rows, n_all))
        ops = [torch.empty(opb, dtype=torch.uint8, device=dev) for _ in range(4)]
        wsb = L.lib.mimrl_concat_workspace_bytes(hid)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        gwb = L.lib.mimrl_gemm_split_workspace_bytes(0, hid, hid, L.lib.mimrl_concat_pair_rows(rows, n_all))
        gws = torch.empty(gwb, dtype=torch.uint8, device=dev)
        tmp = torch.empty_like(w2)
        for r0 in range(0, n_own, rows):
            r = min(rows, n_own - r0)
            pr = L.lib.mimrl_concat_pair_rows(r, n_all)
            L.check(L.lib.mimrl_concat_grad_fused(L.ptr(u[r0:r0 + r]), L.ptr(vt), r, n_all, n_all, hid, rb.offset + r0,
                                                  L.ptr(w2), L.ptr(b2), L.ptr(w3), L.ptr(b3), L.ptr(w4), L.ptr(b4), fam,
                                                  L.ptr(coef), L.ptr(shift_own[r0:r0 + r]), L.ptr(g_u[r0:r0 + r]), L.ptr(g_vt),
                                                  L.ptr(g_b2), L.ptr(g_b3), L.ptr(g_w4), L.ptr(g_b4), L.ptr(ops[0]),
                                                  L.ptr(ops[1]), L.ptr(ops[2]), L.ptr(ops[3]), L.ptr(ws), wsb, st))
            for a, b, acc in ((ops[2], ops[0], g_w2), (ops[3], ops[1], g_w3)):
                L.check(L.lib.mimrl_gemm_split_blocked(L.ptr(a), L.ptr(b), hid, hid, pr, L.ptr(tmp), L.ptr(gws), gwb, st))
                acc += tmp
        dbase = vec[2, own].reshape(n_own, 1).clone() if has_base else None
        return (g_u, g_vt.t(), g_diag, g_w2, g_b2 if b2 is not None else None, g_w3, g_b3 if b3 is not None else None,
                g_w4.reshape(1, -1), g_b4 if b4 is not None else None, dbase, None, None)


def concat_bound(critic, x_own, y_own, bound_type, log_baseline=None, rowblock=None):
    """Fused ``bound(CriticModel('concat')(x, y))`` (VMI.py:58-65 + VMI.py:136-198) for the reference's default concat
    critic (ReLU, hidden 256, two hidden layers): returns ``(mi, mi_loss)`` exactly as ``VMIEstimator.forward`` does.
    Rows of the score matrix index x (VMI.py:65), the rank's own rows; columns index every rank's y."""
    if bound_type not in L.BOUND_IDS:
        raise NotImplementedError
    f = critic.MLP_f
    first = f[0]
    dx = x_own.shape[1]
    u = linear(x_own, first.weight[:, :dx], first.bias)                  # layer 1 factorised: W1 [x; y] = W1x x + W1y y
    v_own = linear(y_own, first.weight[:, dx:])
    v_all = gather_rows(v_own, rowblock)
    # the diagonal pairs (i, own_offset + i) are an ordinary row batch of the same MLP
    diag = mlp_apply(f[1:], u + v_own).reshape(-1)
    return _ConcatBoundFused.apply(u, v_all, diag, f[2].weight, f[2].bias, f[4].weight, f[4].bias, f[6].weight, f[6].bias,
                                   log_baseline, L.BOUND_IDS[bound_type], rowblock)


# --------------------------------------------------------------------------
# modules (VMI.py:25-110)
# --------------------------------------------------------------------------


class CriticModel(nn.Module):
    """VMI.py:25-69.  ``forward`` returns the materialised [B,B] matrix like
    the reference; ``embed`` exposes (g(x), h(y)) for the fused path."""

    def __init__(self, critic_type, dim_x, dim_y, hidden_dim=256, embed_dim=128, layers=2, activation='relu'):
        super().__init__()
        self.critic_type = critic_type
        if critic_type == 'separate':
            self.MLP_g = mlps(dim_x, hidden_dim, embed_dim, layers, activation)
            self.MLP_h = mlps(dim_y, hidden_dim, embed_dim, layers, activation)
            _zero_biases(self.MLP_g)
            _zero_biases(self.MLP_h)
        elif critic_type == 'concat':
            self.MLP_f = mlps(dim_x + dim_y, hidden_dim, 1, layers, activation)
            _zero_biases(self.MLP_f)
        else:
            raise NotImplementedError
        self.pair_chunk = 1 << 20          # pairs scored per step of the concat critic

    def embed(self, x, y):
        return mlp_apply(self.MLP_g, x), mlp_apply(self.MLP_h, y)

    def _concat_rows(self, x_rows, y):
        """scores[i, :] = f([x_i, y_j]) for a block of rows, first layer factorised
        (W1 [x;y] = W1x x + W1y y) so the B^2 x (dx+dy) pair matrix is never built."""
        first = self.MLP_f[0]
        dx = x_rows.shape[1]
        u = linear(x_rows, first.weight[:, :dx], first.bias)
        v = linear(y, first.weight[:, dx:])
        if self._fused_pairs():
            f = self.MLP_f
            return _ConcatPairMLP.apply(u, v, f[2].weight, f[2].bias, f[4].weight, f[4].bias, f[6].weight, f[6].bias)
        h = (u[:, None, :] + v[None, :, :]).reshape(-1, u.shape[1])
        h = mlp_apply(self.MLP_f[1:], h)          # activation of layer 1, then the hidden layers on the tensor cores
        return h.reshape(x_rows.shape[0], y.shape[0])

    def _fused_pairs(self):
        """The tensor-core all-pairs kernels cover the reference default: ReLU, hidden 256, two hidden layers."""
        f = self.MLP_f
        return (f[0].weight.is_cuda and len(f) == 7 and all(isinstance(f[i], nn.ReLU) for i in (1, 3, 5))
                and L.lib.mimrl_concat_tc_supported(f[2].weight.shape[0], 2) and f[2].weight.shape == (256, 256)
                and f[4].weight.shape == (256, 256) and f[6].weight.shape == (1, 256))

    def forward(self, x, y):
        if self.critic_type == 'separate':
            x_, y_ = self.embed(x, y)
            return linear(y_, x_)                      # scores = y_ @ x_.T (VMI.py:57) on the repo's product kernels
        if self.critic_type == 'concat':
            if self._fused_pairs():
                return self._concat_rows(x, y)
            from torch.utils.checkpoint import checkpoint
            n = x.shape[0]
            rows = max(1, self.pair_chunk // max(n, 1))
            if rows >= n:
                return self._concat_rows(x, y)
            blocks = [checkpoint(self._concat_rows, x[r:r + rows], y, use_reentrant=False)
                      for r in range(0, n, rows)]
            return torch.cat(blocks, dim=0)
        raise NotImplementedError


class BaselineModel(nn.Module):
    """VMI.py:72-110 (``gaussain`` spelled as in the reference)."""

    def __init__(self, baseline_type, dim_y, hidden_dim=256, layers=2, activation='relu', mu=0, rho=1):
        super().__init__()
        self.baseline_type = baseline_type
        if baseline_type == 'unnormalized':
            self.MLP = mlps(dim_y, hidden_dim, 1, layers, activation)
            _zero_biases(self.MLP)
        elif baseline_type == 'constant':
            pass
        elif baseline_type == 'gaussain':
            self.gaussain_dist = torch.distributions.Normal(mu, rho)
        else:
            raise NotImplementedError

    def forward(self, y):
        n = y.shape[0]
        if self.baseline_type == 'unnormalized':
            return mlp_apply(self.MLP, y).reshape(n, 1)
        if self.baseline_type == 'constant':
            return torch.zeros(n, 1, device=y.device, dtype=y.dtype)
        if self.baseline_type == 'gaussain':
            return torch.sum(self.gaussain_dist.log_prob(y), -1).reshape(n, 1)
        raise NotImplementedError
