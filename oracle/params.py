"""Deterministic synthetic parameters and inputs shared by the golden generator,
the oracle tests and the GPU parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under mimrl_b200/ may
import this package.

Every array is drawn from ``np.random.default_rng(seed)`` in a fixed order so
that a golden file only has to store the seed, not the weights.  Shapes follow
the reference's module layout:

* relu MLP stack ``Linear(d,h) [+ Linear(h,h)]*layers + Linear(h,out)``
  -> VMI.py:13-22 (``mlps``) and Model.py:52-57 (``MLP_For_CMI``)
* CubeMLP block parameters -> MLPProcess.py:26-52
"""
from __future__ import annotations

import numpy as np


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def linear(rng, d_in, d_out, bias=True, bias_scale=0.05):
    """One nn.Linear worth of parameters: W [d_out, d_in], b [d_out]."""
    w = _uniform(rng, (d_out, d_in), 1.0 / np.sqrt(d_in))
    b = _uniform(rng, (d_out,), bias_scale) if bias else None
    return w, b


def mlp_stack(rng, d_in, hidden, d_out, layers):
    """Parameters of VMI.py:13-22 ``mlps``: list of (W, b), len = layers + 2."""
    dims = [d_in] + [hidden] * (layers + 1) + [d_out]
    return [linear(rng, dims[i], dims[i + 1]) for i in range(len(dims) - 1)]


def stack_to_state(prefix, stack):
    """(W,b) list -> torch state_dict keys of an nn.Sequential with ReLUs
    interleaved (Linear at even positions 0,2,4,...)."""
    out = {}
    for i, (w, b) in enumerate(stack):
        out[f"{prefix}{2 * i}.weight"] = w
        out[f"{prefix}{2 * i}.bias"] = b
    return out


def vmi_params(seed, critic_type, baseline_type, d, hidden, embed, layers):
    """All parameters of one Model.py:108-113 ``VMIEstimator``."""
    rng = np.random.default_rng(seed)
    p = {}
    if critic_type == "separate":
        p["g"] = mlp_stack(rng, d, hidden, embed, layers)
        p["h"] = mlp_stack(rng, d, hidden, embed, layers)
    else:
        p["f"] = mlp_stack(rng, 2 * d, hidden, 1, layers)
    if baseline_type == "unnormalized":
        p["a"] = mlp_stack(rng, d, hidden, 1, layers)
    return p


def vmi_state_dict(p):
    sd = {}
    if "g" in p:
        sd.update(stack_to_state("critic_model.MLP_g.", p["g"]))
        sd.update(stack_to_state("critic_model.MLP_h.", p["h"]))
    if "f" in p:
        sd.update(stack_to_state("critic_model.MLP_f.", p["f"]))
    if "a" in p:
        sd.update(stack_to_state("baseline_model.MLP.", p["a"]))
    return sd


def vcmi_params(seed, embed, hidden):
    """Model.py:52-57: Linear(3e,h) Linear(h,h) Linear(h,h) Linear(h,2)."""
    rng = np.random.default_rng(seed)
    dims = [3 * embed, hidden, hidden, hidden, 2]
    return [linear(rng, dims[i], dims[i + 1]) for i in range(4)]


def vcmi_state_dict(stack):
    return stack_to_state("classifier.mlp.", stack)


def cubemlp_params(seed, d_in, d_hiddens, d_outs, bias, ln_first, res_project):
    """Per-block dict of MLPProcess.py:26-52 parameters, axis order l,k,d."""
    rng = np.random.default_rng(seed)
    blocks = []
    for i in range(len(d_hiddens)):
        ins = d_in if i == 0 else d_outs[i - 1]
        blk = {}
        for a, ax in enumerate("lkd"):
            w1, b1 = linear(rng, ins[a], d_hiddens[i][a], bias)
            w2, b2 = linear(rng, d_hiddens[i][a], d_outs[i][a], bias)
            blk[f"mlp_{ax}.fc1.weight"] = w1
            blk[f"mlp_{ax}.fc2.weight"] = w2
            if bias:
                blk[f"mlp_{ax}.fc1.bias"] = b1
                blk[f"mlp_{ax}.fc2.bias"] = b2
            n_ln = ins[a] if ln_first else d_outs[i][a]
            blk[f"ln_{ax}.weight"] = (1.0 + _uniform(rng, (n_ln,), 0.2)).astype(np.float32)
            blk[f"ln_{ax}.bias"] = _uniform(rng, (n_ln,), 0.1)
            if res_project[i]:
                blk[f"res_projection_{ax}.weight"] = linear(rng, ins[a], d_outs[i][a], False)[0]
        blocks.append(blk)
    return blocks


def cubemlp_state_dict(blocks):
    sd = {}
    for i, blk in enumerate(blocks):
        for k, v in blk.items():
            sd[f"layers_stack.{i}.{k}"] = v
    return sd


def features(seed, n, d, scale=1.0, corr=None):
    """Synthetic feature rows ~N(0,1)*scale.  With ``corr`` given, returns a
    pair (x, y) with y = corr*x + sqrt(1-corr^2)*eps (dependent pairs)."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, d)) * scale).astype(np.float32)
    if corr is None:
        return x
    eps = (rng.standard_normal((n, d)) * scale).astype(np.float32)
    y = (corr * x + np.sqrt(1.0 - corr * corr) * eps).astype(np.float32)
    return x, y
