"""CPU oracle for the classifier-based conditional-MI estimator
(reference Model.py:47-72 ``MLP_For_CMI`` and Model.py:150-225 ``VCMIEstimator``).

TEST INFRASTRUCTURE ONLY.  numpy restatement with hand-derived gradients,
pinned by tests/test_oracle_vcmi.py against tests/golden/vcmi.npz (the
reference run unmodified).
"""
from __future__ import annotations

import numpy as np

from .vmi_oracle import cast_stack, mlp_backward, mlp_forward


def _widen(f, embed):
    """Model.py:161-166: narrow inputs are tiled to embed_dim columns (N5)."""
    return np.tile(f, (1, embed // f.shape[1])) if f.shape[1] != embed else f


def _head(logits, act):
    """Model.py:69-70: clamp(-10,10) then Hardtanh(1e-4,1-1e-4) or Sigmoid.
    Returns (out, d out / d logits)."""
    inside = (logits > -10.0) & (logits < 10.0)
    c = np.clip(logits, -10.0, 10.0)
    if act == "hardtanh":
        lo, hi = 1e-4, 1.0 - 1e-4
        out = np.clip(c, lo, hi)
        d = ((c > lo) & (c < hi)).astype(logits.dtype)
    elif act == "sigmoid":
        out = 1.0 / (1.0 + np.exp(-c))
        d = out * (1.0 - out)
    else:
        raise NotImplementedError(act)
    return out, d * inside


def vcmi_estimator(stack, act, embed, fx, fy, fz, kx, ky, kz, dtype=np.float64, w_cmi=0.0, w_loss=1.0):
    """Returns dict(cmi, loss, grads of  w_cmi*cmi + w_loss*loss).

    cmi  = 1 + (sum log-odds(gamma_joint) - sum log-odds(gamma_prod)) / (2n)     (Model.py:203-219, N3)
    loss = mean BCE(out, one-hot)                                              (Model.py:176-198)
    with n = number of product rows; the joint batch is truncated to n (N4).
    The reference runs the classifier twice on the same batch (Model.py:189,206);
    both passes produce identical values, so gradients simply add.
    """
    st = cast_stack(stack, dtype)
    ins = [a.astype(dtype) for a in (fx, fy, fz, kx, ky, kz)]
    wide = [_widen(a, embed) for a in ins[:3]]
    joint = np.concatenate(wide, axis=1)
    prod = np.concatenate(ins[3:], axis=1)
    n = prod.shape[0]
    joint = joint[:n]
    batch = np.concatenate([joint, prod], axis=0)
    logits, acts = mlp_forward(st, batch)
    out, dout = _head(logits, act)
    tgt = np.zeros_like(out)
    tgt[:n, 0] = 1.0
    tgt[n:, 1] = 1.0
    logo = np.maximum(np.log(out), -100.0)
    log1m = np.maximum(np.log(1.0 - out), -100.0)
    loss = -(tgt * logo + (1.0 - tgt) * log1m).mean()
    gam = out[:, 0]
    n2 = 2 * n
    half = int(n2 / 2)
    odds = np.log(gam / (1.0 - gam + 1e-6))
    cmi = 1.0 + odds[:half].sum() / n2 - odds[half:2 * half].sum() / n2

    g_out = w_loss * (-(tgt / out) + (1.0 - tgt) / (1.0 - out)) / out.size
    dodds = 1.0 / gam + 1.0 / (1.0 - gam + 1e-6)
    sign = np.where(np.arange(n2) < half, 1.0, -1.0)
    g_out[:, 0] += w_cmi * sign * dodds / n2
    g_logits = g_out * dout
    g_batch, pg = mlp_backward(st, acts, g_logits)
    e = embed
    gj, gp = g_batch[:n], g_batch[n:]
    grads = {}
    for i, nm in enumerate(("fx", "fy", "fz")):
        g = np.zeros((ins[i].shape[0], e), dtype=dtype)
        g[:n] = gj[:, i * e:(i + 1) * e]
        w = ins[i].shape[1]
        grads[nm] = g.reshape(g.shape[0], e // w, w).sum(axis=1)
    for i, nm in enumerate(("kx", "ky", "kz")):
        grads[nm] = gp[:, i * e:(i + 1) * e]
    pgd = {}
    for i, (gw, gb) in enumerate(pg):
        pgd[f"classifier.mlp.{2 * i}.weight"] = gw
        pgd[f"classifier.mlp.{2 * i}.bias"] = gb
    return dict(cmi=cmi, loss=loss, grads=grads, pg=pgd)
