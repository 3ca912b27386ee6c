"""Two-stage MI/CMI training step: the per-batch body of the reference's
``Solver.train`` (Solver.py:194-248) with ``compute_custumized_loss``
(Customization.py:91-115), hosted around the B200 estimator modules.

Stage 1 (Solver.py:204-216): features are computed by the model, only the MI /
CMI estimators are updated (``optimizer_vmi``) with
``loss = sum_i coef1[i] * mi_losses[i]`` over the 11 stage-1 terms.
Stage 2 (Solver.py:220-242): the main model is updated (``optimizer_main``) with
``loss = task_loss + sum_i coef2[i] * mi_losses[i]`` over the 8 stage-2 terms,
and the batch features are appended to next epoch's k-NN pool.

The encoders (BERT / GRU / CubeMLP) are the caller's business: this driver
takes a callable ``features(batch) -> (prediction, F_F, T_F, A_F, V_F)`` so it
works both with a full reference-style ``Model.forward`` and with the
synthetic feature heads used by the benchmark.  ``mis`` stay on the device (the
reference ``.cpu().item()``s them every batch, Solver.py:229, a sync per step).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn


@dataclass
class FeaturePool:
    """Previous epoch's stage-2 features (Solver.py:219-244), device-resident."""
    C: Optional[torch.Tensor] = None
    F: Optional[torch.Tensor] = None
    T: Optional[torch.Tensor] = None
    A: Optional[torch.Tensor] = None
    V: Optional[torch.Tensor] = None
    _next: dict = field(default_factory=lambda: {k: [] for k in "CFTAV"})
    _fit: dict = field(default_factory=dict)

    def __len__(self):
        return 0 if self.C is None else int(self.C.shape[0])

    def keys(self, name):
        """Pool ``name`` as a fitted k-NN key pool (model.KnnPool): fitted on first use, reused by every sampler call until
        the pool tensor is replaced (``roll``, assignment) or modified in place through torch (version counter).  The
        reference refits inside every prod_knn_sample call (Model.py:82-85)."""
        from .model import KnnPool
        t = getattr(self, name)
        hit = self._fit.get(name)
        if hit is None or hit[0] is not t or hit[1] != t._version:
            hit = self._fit[name] = (t, t._version, KnnPool(t))
        return hit[2]

    def append(self, labels, F_F, T_F, A_F, V_F):
        for k, v in zip("CFTAV", (labels.reshape(-1, 1), F_F, T_F, A_F, V_F)):
            self._next[k].append(v.detach())

    def roll(self):
        """End of epoch: the collected features become the pool (Solver.py:244)."""
        if self._next["C"]:
            for k in "CFTAV":
                setattr(self, k, torch.cat(self._next[k], 0))
                self._next[k] = []
            self._fit = {}


class TwoStageStep:
    def __init__(self, heads: nn.Module, features: Callable, task_loss: Callable, optimizer_main, optimizer_vmi,
                 coef1: Sequence[float] = (0.1,) * 11, coef2: Sequence[float] = (0.1,) * 8, gradient_clip: float = 1.0,
                 clip_params: Optional[List[nn.Parameter]] = None):
        self.heads, self.features, self.task_loss = heads, features, task_loss
        self.opt_main, self.opt_vmi = optimizer_main, optimizer_vmi
        self.coef1, self.coef2, self.clip = list(coef1), list(coef2), gradient_clip
        self.clip_params = clip_params

    def _clip(self):
        if self.clip > 0 and self.clip_params:
            torch.nn.utils.clip_grad_value_([p for p in self.clip_params if p.requires_grad], self.clip)

    def stage1(self, batch, labels, pool: FeaturePool):
        """One optimizer_vmi step (Solver.py:204-214).  Returns (loss, mis) on the device."""
        pred, F_F, T_F, A_F, V_F = self.features(batch)
        if len(pool) == 0:                                   # Customization.py:97-98
            return torch.zeros((), device=pred.device), []
        mis, losses = self.heads.compute_vmi_loss_stage1(pred.reshape(-1, 1), labels.reshape(-1, 1), F_F, T_F, A_F, V_F,
                                                         pool.C, pool.F, pool.keys("T"), pool.keys("A"),
                                                         pool.keys("V"))
        loss = sum(l * c for l, c in zip(losses, self.coef1))
        self.opt_vmi.zero_grad(set_to_none=True)
        loss.backward()
        self._clip()
        self.opt_vmi.step()
        return loss.detach(), [m.detach() for m in mis]

    def stage2(self, batch, labels, pool: FeaturePool):
        """One optimizer_main step (Solver.py:220-236); appends the features to next epoch's pool."""
        pred, F_F, T_F, A_F, V_F = self.features(batch)
        pool.append(labels, F_F, T_F, A_F, V_F)
        loss = self.task_loss(pred.reshape(-1), labels.reshape(-1))
        mis = []
        if len(pool) > 0:                                    # Customization.py:105-106
            mis, losses = self.heads.compute_vmi_loss_stage2(pred.reshape(-1, 1), labels.reshape(-1, 1), F_F, T_F, A_F,
                                                             V_F, pool.C, pool.F, pool.keys("T"), pool.keys("A"),
                                                             pool.keys("V"))
            loss = loss + sum(l * c for l, c in zip(losses, self.coef2))
        self.opt_main.zero_grad(set_to_none=True)
        loss.backward()
        self._clip()
        self.opt_main.step()
        return loss.detach(), [m.detach() for m in mis]


class GraphedTwoStageStep:
    """``TwoStageStep`` with each stage captured into one CUDA graph (graphs.py): the launch-bound small-batch regime.

    The pool tensors are static for an epoch (Solver.py:244 swaps them at the epoch boundary): call ``recapture`` after
    ``pool.roll()``.  Optimisers must be built with ``capturable=True``.  The k-NN query ids are drawn on the host from
    numpy's global RNG before every replay, in the reference's call order."""

    def __init__(self, step: TwoStageStep, batch: torch.Tensor, labels: torch.Tensor, pool: FeaturePool,
                 parallel_branches: bool = True):
        """parallel_branches: capture the eleven independent estimators of a stage as parallel branches of the graph
        (model.run_branches): same values, gradients and RNG consumption, ~2x shorter replays at small batch."""
        self.step, self.pool, self.parallel_branches = step, pool, parallel_branches
        self.recapture(batch, labels)

    def recapture(self, batch, labels):
        from .graphs import GraphedCallable, HostIdSource
        step, pool = self.step, self.pool
        frozen = FeaturePool(C=pool.C, F=pool.F, T=pool.T, A=pool.A, V=pool.V)     # appends go to a scratch list
        self._feats = None

        def s1(b, l):
            loss, mis = step.stage1(b, l, frozen)
            return (loss,) + tuple(mis)

        def s2(b, l):
            frozen._next = {k: [] for k in "CFTAV"}
            loss, mis = step.stage2(b, l, frozen)
            feats = tuple(frozen._next[k][-1] for k in "CFTAV")
            return (loss,) + tuple(mis) + feats

        # warm-up and capture run real optimiser steps and consume numpy's RNG: put everything back afterwards
        # (in place -- the graphs hold the addresses of the parameters and of the optimiser state)
        import numpy as np
        opts = (step.opt_main, step.opt_vmi)
        params = [p for o in opts for gr in o.param_groups for p in gr["params"]]
        p_saved = [p.detach().clone() for p in params]
        o_saved = [{id(p): {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in o.state.items()}
                   for o in opts]
        rng = np.random.get_state()
        # module buffers (e.g. BatchNorm running statistics inside `features`) and torch's CUDA RNG (dropout) advance
        # during warm-up and capture as well: snapshot them so that a recapture leaves no trace
        mods = [step.heads] + [m for m in (getattr(step.features, "__self__", None), step.features)
                               if isinstance(m, nn.Module)]
        b_saved = [(b, b.detach().clone()) for m in mods for b in m.buffers()]
        cuda_rng = torch.cuda.get_rng_state()
        was = getattr(step.heads, "parallel_branches", False)
        if hasattr(step.heads, "parallel_branches") or hasattr(type(step.heads), "parallel_branches"):
            step.heads.parallel_branches = bool(self.parallel_branches) or was
        try:
            self.g1 = GraphedCallable(s1, [batch, labels], id_source=HostIdSource())
            self.g2 = GraphedCallable(s2, [batch, labels], id_source=HostIdSource())
        finally:
            if hasattr(step.heads, "parallel_branches"):
                step.heads.parallel_branches = was
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, v in zip(params, p_saved):
                p.copy_(v)
            for o, saved in zip(opts, o_saved):
                for p, st in o.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(saved[id(p)][k]) if id(p) in saved and k in saved[id(p)] else v.zero_()
            for b, v in b_saved:
                b.copy_(v)
        torch.cuda.set_rng_state(cuda_rng)
        np.random.set_state(rng)

    def stage1(self, batch, labels):
        out = self.g1(batch, labels)
        return out[0], list(out[1:])

    def stage2(self, batch, labels):
        out = self.g2(batch, labels)
        feats = out[-5:]
        self.pool.append(feats[0].clone(), *(f.clone() for f in feats[1:]))      # next epoch's pool (Solver.py:237-241)
        return out[0], list(out[1:-5])
