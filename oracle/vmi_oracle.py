"""CPU oracle for the variational-MI path: a numpy restatement of the
reference's critics, baselines and lower bounds, with hand-derived gradients.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never by mimrl_b200/.

Pinned against the reference itself: tests/test_oracle_vmi.py compares every
function here with tests/golden/{bounds,vmi}.npz, which oracle/gen_golden.py
produced by running /root/reference/VMI.py and Model.py unmodified.

Conventions
-----------
* ``S`` is the B x B score matrix exactly as the reference critic returns it:
  separable critic ``S[i,j] = h(y_i) . g(x_j)`` (VMI.py:54-57); concat critic
  ``S[i,j] = f([x_i, y_j])`` (VMI.py:58-65, note the final ``.t()``).
* ``a`` is the log-baseline column [B,1] (VMI.py:99-110); it is subtracted
  row-wise, ``S - a`` (VMI.py:151).
* Every ``bound_*`` returns ``(mi, dmi_dS, dmi_da)``; ``dmi_da`` is None when
  the bound takes no baseline.  Closed forms: SURVEY.md Appendix A.
* dtype follows the inputs (float64 for checking, float32 for CPU timing).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# small numerics helpers
# ----------------------------------------------------------------------------


def _softplus(z):
    return np.logaddexp(0.0, z)


def _sigmoid(z):
    return 0.5 * (1.0 + np.tanh(0.5 * z))


def _offdiag_mask(B):
    return ~np.eye(B, dtype=bool)


def _lse_offdiag(S):
    """log sum_{i != j} exp(S_ij)  (VMI.py:121-126 without the log n_off)."""
    B = S.shape[0]
    off = _offdiag_mask(B)
    if B < 2:
        return S.dtype.type(-np.inf)
    m = S[off].max()
    return m + np.log(np.exp(S[off] - m).sum())


# ----------------------------------------------------------------------------
# bounds: value and gradient w.r.t. the score matrix
# ----------------------------------------------------------------------------


def bound_infonce(S):
    """VMI.py:162-166: log B + mean_i(S_ii - logsumexp_j S_ij)."""
    B = S.shape[0]
    m = S.max(axis=1, keepdims=True)
    e = np.exp(S - m)
    z = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(z))[:, 0]
    mi = np.log(S.dtype.type(B)) + np.mean(np.diag(S) - lse)
    G = -(e / z) / B
    G[np.arange(B), np.arange(B)] += 1.0 / B
    return mi, G, None


def bound_dv(S):
    """VMI.py:136-139."""
    B = S.shape[0]
    n_off = B * (B - 1.0)
    L = _lse_offdiag(S)
    mi = np.mean(np.diag(S)) - (L - np.log(n_off))
    G = -np.exp(S - L)
    G[np.arange(B), np.arange(B)] = 1.0 / B
    return mi, G, None


def bound_tuba(S, a=None):
    """VMI.py:148-154: 1 + mean diag(S-a) - mean_offdiag exp(S-a)."""
    B = S.shape[0]
    n_off = B * (B - 1.0)
    Sp = S if a is None else S - a.reshape(B, 1)
    L = _lse_offdiag(Sp)
    mi = 1.0 + np.mean(np.diag(Sp)) - np.exp(L - np.log(n_off))
    G = -np.exp(Sp) / n_off
    G[np.arange(B), np.arange(B)] = 1.0 / B
    da = None if a is None else -G.sum(axis=1, keepdims=True)
    return mi, G, da


def bound_nwj(S):
    """VMI.py:157-159: TUBA on S - 1."""
    mi, G, _ = bound_tuba(S - 1.0)
    return mi, G, None


def bound_js_fgan(S):
    """VMI.py:169-174."""
    B = S.shape[0]
    n_off = B * (B - 1.0)
    d = np.diag(S)
    mi = np.mean(-_softplus(-d)) - (_softplus(S).sum() - _softplus(d).sum()) / n_off
    G = -_sigmoid(S) / n_off
    G[np.arange(B), np.arange(B)] = _sigmoid(-d) / B
    return mi, G, None


def bound_js(S):
    """VMI.py:177-182: value of NWJ, gradient of JS-fGAN."""
    nwj, _, _ = bound_nwj(S)
    _, G, _ = bound_js_fgan(S)
    return nwj, G, None


def bound_smile(S):
    """VMI.py:185-198: DV with the partition term on clamp(S,-1,1); gradient of
    JS-fGAN.  The ``clip`` argument is overwritten with 1 at VMI.py:186."""
    B = S.shape[0]
    n_off = B * (B - 1.0)
    z = _lse_offdiag(np.clip(S, -1.0, 1.0)) - np.log(n_off)
    dv = np.mean(np.diag(S)) - z
    _, G, _ = bound_js_fgan(S)
    return dv, G, None


def bound_mine(S, ma_et=1.0, ma_rate=0.01):
    """VMI.py:142-145 + Model.py:121-124.  Returns the DV value as ``mi`` and
    ALSO the training loss and its gradient, because for this bound alone
    ``mi_loss`` is not ``-mi``: positive sign, moving-average
    denominator treated as a constant, plain (unstabilised) exp, mean over B^2
    entries with zeros on the diagonal."""
    B = S.shape[0]
    mi, G_dv, _ = bound_dv(S)
    off = _offdiag_mask(B)
    et = np.where(off, np.exp(S), 0.0)
    mean_et = et.sum() / (B * B)
    ma = (1.0 - ma_rate) * ma_et + ma_rate * mean_et
    loss = np.mean(np.diag(S)) - mean_et / ma
    G_loss = -et / (B * B * ma)
    G_loss[np.arange(B), np.arange(B)] = 1.0 / B
    return mi, G_dv, None, loss, G_loss


def bound_interpolate(S, a, alpha_logit=0.01):
    """VMI.py:201-250.  ``a`` [B,1] is the learnt log-baseline.

    loo_ij = log sum_{k != j} exp S_ik - log(B-1)            (VMI.py:213-226)
    ib_ij  = logaddexp(log alpha + loo_ij, log(1-alpha) + a_i) (VMI.py:201-210,241)
    marg   = sum_{i != j} exp(S_ij - ib_jj) / n_off           (VMI.py:244-245)
    joint  = sum_{i != j} (S_jj - ib_ij) / n_off              (VMI.py:248-249)
    """
    B = S.shape[0]
    n_off = B * (B - 1.0)
    off = _offdiag_mask(B)
    a = a.reshape(B)
    log_alpha = -_softplus(-alpha_logit)
    log_1m_alpha = -_softplus(alpha_logit)
    m = S.max(axis=1, keepdims=True)
    e = np.exp(S - m)                       # scaled exp, row-wise
    E = e.sum(axis=1, keepdims=True)
    D = E - e                               # D_ij = sum_{k != j} e_ik (scaled)
    loo = m + np.log(D) - np.log(B - 1.0)
    t1 = log_alpha + loo
    t2 = (log_1m_alpha + a)[:, None] + np.zeros_like(S)
    ib = np.logaddexp(t1, t2)
    omega = _sigmoid(t1 - t2)               # d ib / d loo
    ibd = np.diag(ib)
    marg_mat = np.where(off, np.exp(S - ibd[None, :]), 0.0)
    marg = marg_mat.sum() / n_off
    joint = ((np.diag(S)[None, :] - ib) * off).sum() / n_off
    mi = 1.0 + joint - marg

    # gradient
    w = np.where(off, omega / D, 0.0)       # omega_ij / D_ij, j != i
    T = w.sum(axis=1, keepdims=True)
    GJ = -(e / n_off) * (T - w)             # w_ik already 0 at k == i
    GJ[np.arange(B), np.arange(B)] += 1.0 / B
    C = marg_mat.sum(axis=0) / n_off        # C_j
    od = np.diag(omega)
    Dd = np.diag(D)
    GM = marg_mat / n_off - np.where(off, (C * od / Dd)[:, None] * e, 0.0)
    G = GJ - GM
    da = (-((1.0 - omega) * off).sum(axis=1) / n_off + C * (1.0 - od)).reshape(B, 1)
    return mi, G, da


BOUND_TYPES = ("dv", "mine", "tuba", "nwj", "infonce", "js_fgan", "js", "smile", "interpolate")


def bound(bound_type, S, a=None, alpha_logit=0.01):
    """Dispatch mirroring Model.py:121-146.  Returns
    ``(mi, mi_loss, dloss_dS, dloss_da)``."""
    if bound_type == "mine":
        mi, _, _, loss, G_loss = bound_mine(S)
        return mi, loss, G_loss, None
    if bound_type == "dv":
        mi, G, da = bound_dv(S)
    elif bound_type == "tuba":
        mi, G, da = bound_tuba(S, a)
    elif bound_type == "nwj":
        mi, G, da = bound_nwj(S)
    elif bound_type == "infonce":
        mi, G, da = bound_infonce(S)
    elif bound_type == "js":
        mi, G, da = bound_js(S)
    elif bound_type == "js_fgan":
        mi, G, da = bound_js_fgan(S)
    elif bound_type == "smile":
        mi, G, da = bound_smile(S)
    elif bound_type == "interpolate":
        mi, G, da = bound_interpolate(S, a, alpha_logit)
    else:
        raise NotImplementedError(bound_type)
    return mi, -mi, -G, (None if da is None else -da)


# ----------------------------------------------------------------------------
# relu MLP stacks (VMI.py:13-22) with manual backward
# ----------------------------------------------------------------------------


def mlp_forward(stack, x):
    """stack = [(W [out,in], b [out]), ...]; ReLU between layers, none after
    the last.  Returns (out, cache)."""
    acts = [x]
    h = x
    n = len(stack)
    for i, (w, b) in enumerate(stack):
        h = h @ w.T.astype(h.dtype)
        if b is not None:
            h = h + b.astype(h.dtype)
        if i < n - 1:
            h = np.maximum(h, 0)
        acts.append(h)
    return h, acts


def mlp_backward(stack, acts, gout):
    """Returns (gx, [(gW, gb), ...])."""
    n = len(stack)
    grads = [None] * n
    g = gout
    for i in range(n - 1, -1, -1):
        if i < n - 1:
            g = g * (acts[i + 1] > 0)
        w, b = stack[i]
        gw = g.T @ acts[i]
        gb = g.sum(axis=0) if b is not None else None
        grads[i] = (gw, gb)
        g = g @ w.astype(g.dtype)
    return g, grads


def cast_stack(stack, dtype):
    return [(w.astype(dtype), None if b is None else b.astype(dtype)) for w, b in stack]


# ----------------------------------------------------------------------------
# baselines (VMI.py:72-110)
# ----------------------------------------------------------------------------


def baseline_forward(baseline_type, y, stack=None, mu=0.0, rho=1.0):
    """Returns (a [B,1], cache).  'gaussain' [sic]: sum_d log N(y_d; mu, rho)."""
    B = y.shape[0]
    if baseline_type == "unnormalized":
        out, acts = mlp_forward(stack, y)
        return out.reshape(B, 1), acts
    if baseline_type == "constant":
        return np.zeros((B, 1), dtype=y.dtype), None
    if baseline_type == "gaussain":
        lp = -((y - mu) ** 2) / (2.0 * rho * rho) - np.log(rho) - 0.5 * np.log(2.0 * np.pi)
        return lp.sum(axis=-1).reshape(B, 1).astype(y.dtype), None
    raise NotImplementedError(baseline_type)


# ----------------------------------------------------------------------------
# critics (VMI.py:25-69)
# ----------------------------------------------------------------------------


def separable_scores(g_stack, h_stack, x, y):
    xe, cx = mlp_forward(g_stack, x)
    ye, cy = mlp_forward(h_stack, y)
    return ye @ xe.T, (xe, ye, cx, cy)


def concat_scores(f_stack, x, y):
    """S[i,j] = f([x_i, y_j]).  Materialises all B^2 rows like the reference
    (small B only)."""
    B = x.shape[0]
    pairs = np.concatenate([np.repeat(x, B, axis=0), np.tile(y, (B, 1))], axis=1)  # row i*B+j = [x_i, y_j]
    out, acts = mlp_forward(f_stack, pairs)
    return out.reshape(B, B), acts


# ----------------------------------------------------------------------------
# VMIEstimator (Model.py:108-148): forward + backward of mi_loss
# ----------------------------------------------------------------------------


def vmi_estimator(params, critic_type, baseline_type, bound_type, x, y, dtype=np.float64,
                  mu=0.0, rho=1.0, alpha_logit=0.01, want_grads=True):
    """params: dict from oracle.params.vmi_params.  Returns a dict with
    ``mi``, ``loss`` and (if want_grads) gradients of ``loss`` w.r.t. x, y and
    every parameter, keyed like the reference state_dict."""
    x = x.astype(dtype)
    y = y.astype(dtype)
    B = x.shape[0]
    p = {k: cast_stack(v, dtype) for k, v in params.items()}
    if critic_type == "separate":
        S, (xe, ye, cx, cy) = separable_scores(p["g"], p["h"], x, y)
    elif critic_type == "concat":
        S, cf = concat_scores(p["f"], x, y)
    else:
        raise NotImplementedError(critic_type)
    a = acache = None
    if bound_type in ("tuba", "interpolate"):
        a, acache = baseline_forward(baseline_type, y, p.get("a"), mu, rho)
    mi, loss, G, ga = bound(bound_type, S, a, alpha_logit)
    res = dict(mi=mi, loss=loss, scores=S)
    if not want_grads:
        return res
    gx = np.zeros_like(x)
    gy = np.zeros_like(y)
    pg = {}
    if critic_type == "separate":
        g_ye = G @ xe
        g_xe = G.T @ ye
        dx, gg = mlp_backward(p["g"], cx, g_xe)
        dy, gh = mlp_backward(p["h"], cy, g_ye)
        gx += dx
        gy += dy
        for i, (gw, gb) in enumerate(gg):
            pg[f"critic_model.MLP_g.{2 * i}.weight"] = gw
            pg[f"critic_model.MLP_g.{2 * i}.bias"] = gb
        for i, (gw, gb) in enumerate(gh):
            pg[f"critic_model.MLP_h.{2 * i}.weight"] = gw
            pg[f"critic_model.MLP_h.{2 * i}.bias"] = gb
    else:
        gpairs, gf = mlp_backward(p["f"], cf, G.reshape(B * B, 1))
        d = x.shape[1]
        gx += gpairs[:, :d].reshape(B, B, d).sum(axis=1)
        gy += gpairs[:, d:].reshape(B, B, d).sum(axis=0)
        for i, (gw, gb) in enumerate(gf):
            pg[f"critic_model.MLP_f.{2 * i}.weight"] = gw
            pg[f"critic_model.MLP_f.{2 * i}.bias"] = gb
    if ga is not None and baseline_type == "unnormalized":
        dy, gs = mlp_backward(p["a"], acache, ga.reshape(B, 1))
        gy += dy
        for i, (gw, gb) in enumerate(gs):
            pg[f"baseline_model.MLP.{2 * i}.weight"] = gw
            pg[f"baseline_model.MLP.{2 * i}.bias"] = gb
    elif ga is not None and baseline_type == "gaussain":
        gy += ga.reshape(B, 1) * (-(y - mu) / (rho * rho))
    res.update(gx=gx, gy=gy, pg=pg)
    return res


# ----------------------------------------------------------------------------
# large-B separable InfoNCE without materialising S: row-block streaming.
# Used for the cpu_baseline timing and for parity at sizes where B x B does not
# fit; arithmetic identical to bound_infonce + the separable critic above.
# ----------------------------------------------------------------------------


def separable_infonce_streamed(params, x, y, dtype=np.float32, block=2048, want_grads=True):
    x = x.astype(dtype)
    y = y.astype(dtype)
    B = x.shape[0]
    g = cast_stack(params["g"], dtype)
    h = cast_stack(params["h"], dtype)
    xe, cx = mlp_forward(g, x)
    ye, cy = mlp_forward(h, y)
    acc = 0.0
    g_ye = np.empty_like(ye)
    g_xe = np.zeros_like(xe)
    for r0 in range(0, B, block):
        r1 = min(B, r0 + block)
        Sb = ye[r0:r1] @ xe.T
        m = Sb.max(axis=1, keepdims=True)
        e = np.exp(Sb - m)
        z = e.sum(axis=1, keepdims=True)
        d = Sb[np.arange(r1 - r0), np.arange(r0, r1)]
        acc += float((d - (m[:, 0] + np.log(z[:, 0]))).sum(dtype=np.float64))
        if want_grads:
            Gl = e / z                       # d loss / dS = (softmax - I)/B
            Gl[np.arange(r1 - r0), np.arange(r0, r1)] -= 1.0
            Gl /= B
            g_ye[r0:r1] = Gl @ xe
            g_xe += Gl.T @ ye[r0:r1]
    mi = float(np.log(dtype(B))) + acc / B
    res = dict(mi=mi, loss=-mi)
    if want_grads:
        gx, gg = mlp_backward(g, cx, g_xe)
        gy, gh = mlp_backward(h, cy, g_ye)
        res.update(gx=gx, gy=gy, g_xe=g_xe, g_ye=g_ye, pg_g=gg, pg_h=gh)
    return res
