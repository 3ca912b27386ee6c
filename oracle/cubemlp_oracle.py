"""CPU oracle for the CubeMLP fusion encoder (reference MLPProcess.py:9-137).

TEST INFRASTRUCTURE ONLY.  Permute-free numpy restatement (SURVEY.md
Appendix D) with hand-derived backward, pinned by tests/test_oracle_cubemlp.py
against tests/golden/cubemlp.npz (the reference run unmodified).

x is [bs, L, K, D].  One block applies three axis-mixes in the order L, K, D
(axes 1, 2, 3).  For axis A with the default ``ln_first=False``
(MLPProcess.py:94-122):

    y = LayerNorm_{A'}( W2 . act(W1 . x + b1) + b2  +  (Wres . x | x) ),  eps = 1e-6

and with ``ln_first=True`` (MLPProcess.py:64-92):

    y = W2 . act(W1 . LayerNorm_A(x) + b1) + b2  +  (Wres . x | x)

Dropout is the identity here (p = 0 in every reference launch command, and
the oracle is only used in eval-equivalent mode).
"""
from __future__ import annotations

import math

import numpy as np

_erf = np.vectorize(math.erf, otypes=[np.float64])


def _act(name, z):
    """Returns (act(z), act'(z)).  'gelu' is the exact erf form (Utils.py:88 -> F.gelu)."""
    if name == "gelu":
        cdf = 0.5 * (1.0 + _erf(z / math.sqrt(2.0)))
        pdf = np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
        return z * cdf, cdf + z * pdf
    if name == "relu":
        return np.maximum(z, 0), (z > 0).astype(z.dtype)
    if name == "tanh":
        t = np.tanh(z)
        return t, 1.0 - t * t
    raise NotImplementedError(name)


def _ln_fwd(z, w, b, eps=1e-6):
    mu = z.mean(axis=-1, keepdims=True)
    var = ((z - mu) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    zh = (z - mu) * rstd
    return zh * w + b, (zh, rstd)


def _ln_bwd(g, w, cache):
    zh, rstd = cache
    gw = (g * zh).reshape(-1, zh.shape[-1]).sum(axis=0)
    gb = g.reshape(-1, zh.shape[-1]).sum(axis=0)
    gz = g * w
    gzc = gz - gz.mean(axis=-1, keepdims=True) - zh * (gz * zh).mean(axis=-1, keepdims=True)
    return gzc * rstd, gw, gb


def _mix_fwd(x, prm, ax, act, ln_first, res_project):
    """One axis-mix on the LAST axis of x (caller has moved the axis there)."""
    w1, b1 = prm[f"mlp_{ax}.fc1.weight"], prm.get(f"mlp_{ax}.fc1.bias")
    w2, b2 = prm[f"mlp_{ax}.fc2.weight"], prm.get(f"mlp_{ax}.fc2.bias")
    lw, lb = prm[f"ln_{ax}.weight"], prm[f"ln_{ax}.bias"]
    cache = {}
    r = x @ prm[f"res_projection_{ax}.weight"].T if res_project else x
    u = x
    if ln_first:
        u, cache["ln"] = _ln_fwd(x, lw, lb)
    z = u @ w1.T + (0 if b1 is None else b1)
    h, cache["dact"] = _act(act, z)
    o = h @ w2.T + (0 if b2 is None else b2)
    s = o + r
    if not ln_first:
        s, cache["ln"] = _ln_fwd(s, lw, lb)
    cache.update(x=x, u=u, h=h)
    return s, cache


def _mix_bwd(g, prm, ax, ln_first, res_project, cache, pg):
    w1, w2 = prm[f"mlp_{ax}.fc1.weight"], prm[f"mlp_{ax}.fc2.weight"]
    lw = prm[f"ln_{ax}.weight"]
    flat = lambda a: a.reshape(-1, a.shape[-1])  # noqa: E731
    if not ln_first:
        g, gw, gb = _ln_bwd(g, lw, cache["ln"])
        pg[f"ln_{ax}.weight"], pg[f"ln_{ax}.bias"] = gw, gb
    gx = np.zeros_like(cache["x"])
    if res_project:
        wr = prm[f"res_projection_{ax}.weight"]
        pg[f"res_projection_{ax}.weight"] = flat(g).T @ flat(cache["x"])
        gx += g @ wr
    else:
        gx += g
    pg[f"mlp_{ax}.fc2.weight"] = flat(g).T @ flat(cache["h"])
    if f"mlp_{ax}.fc2.bias" in prm:
        pg[f"mlp_{ax}.fc2.bias"] = flat(g).sum(axis=0)
    gz = (g @ w2) * cache["dact"]
    pg[f"mlp_{ax}.fc1.weight"] = flat(gz).T @ flat(cache["u"])
    if f"mlp_{ax}.fc1.bias" in prm:
        pg[f"mlp_{ax}.fc1.bias"] = flat(gz).sum(axis=0)
    gu = gz @ w1
    if ln_first:
        gu, gw, gb = _ln_bwd(gu, lw, cache["ln"])
        pg[f"ln_{ax}.weight"], pg[f"ln_{ax}.bias"] = gw, gb
    return gx + gu


_AXES = (("l", 1), ("k", 2), ("d", 3))


def encoder_forward(blocks, x, act, ln_first, res_project, dtype=np.float64):
    """blocks: list of per-block param dicts (oracle.params.cubemlp_params).
    Returns (y, caches)."""
    x = x.astype(dtype)
    caches = []
    for bi, blk in enumerate(blocks):
        prm = {k: v.astype(dtype) for k, v in blk.items()}
        for ax, dim in _AXES:
            xm = np.moveaxis(x, dim, -1)
            ym, c = _mix_fwd(xm, prm, ax, act, ln_first, res_project[bi])
            x = np.moveaxis(ym, -1, dim)
            caches.append((bi, ax, dim, prm, c))
    return x, caches


def encoder_backward(caches, gy, ln_first, res_project):
    """Returns (gx, {state_dict-style name: grad})."""
    g = gy
    pgs = {}
    for bi, ax, dim, prm, c in reversed(caches):
        pg = {}
        gm = _mix_bwd(np.moveaxis(g, dim, -1), prm, ax, ln_first, res_project[bi], c, pg)
        g = np.moveaxis(gm, -1, dim)
        for k, v in pg.items():
            pgs[f"layers_stack.{bi}.{k}"] = v
    return g, pgs
