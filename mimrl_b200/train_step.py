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

    def __len__(self):
        return 0 if self.C is None else int(self.C.shape[0])

    def append(self, labels, F_F, T_F, A_F, V_F):
        for k, v in zip("CFTAV", (labels.reshape(-1, 1), F_F, T_F, A_F, V_F)):
            self._next[k].append(v.detach())

    def roll(self):
        """End of epoch: the collected features become the pool (Solver.py:244)."""
        if self._next["C"]:
            for k in "CFTAV":
                setattr(self, k, torch.cat(self._next[k], 0))
                self._next[k] = []


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
                                                         pool.C, pool.F, pool.T, pool.A, pool.V)
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
                                                             V_F, pool.C, pool.F, pool.T, pool.A, pool.V)
            loss = loss + sum(l * c for l, c in zip(losses, self.coef2))
        self.opt_main.zero_grad(set_to_none=True)
        loss.backward()
        self._clip()
        self.opt_main.step()
        return loss.detach(), [m.detach() for m in mis]
