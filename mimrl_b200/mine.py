"""Standalone MI estimator: drop-in for ``EMA``, ``train_MINE``, ``compute_MI``,
``sample_correlated_gaussian`` and ``rho_to_mi`` of the reference's VMI.py:253-396.

Same signatures and training procedure (Adamax, the weight averaging applied
after every step, the per-epoch MI history and the three read-outs); the
per-batch estimate runs on the fused sm_100a sweeps of ``vmi.separable_bound``
for the separable critic (the batch x batch score matrix is never written) and
on the fused all-pairs kernels + materialised-score bounds for the concat
critic.  ``sample_correlated_gaussian`` / ``rho_to_mi`` give the analytic
known answer ``-0.5 d log(1 - rho^2)`` the tests train against.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from .vmi import BaselineModel, CriticModel, _scores_bound, interp_lower_bound, separable_bound

_BOUNDS = ("dv", "tuba", "nwj", "infonce", "js", "js_fgan", "smile", "interpolate", "mine")


class EMA:
    """VMI.py:253-284.  ``shadow`` <- (1 - decay) * param + decay * shadow on ``update``; ``apply_shadow`` swaps the
    averaged tensors in (keeping the live ones in ``backup``) and ``restore`` swaps them back."""

    def __init__(self, model, decay):
        self.model, self.decay = model, decay
        self.shadow, self.backup = {}, {}

    def _trainable(self):
        return ((n, p) for n, p in self.model.named_parameters() if p.requires_grad)

    def register(self):
        for name, p in self._trainable():
            self.shadow[name] = p.data.clone()

    def update(self):
        for name, p in self._trainable():
            assert name in self.shadow
            self.shadow[name] = ((1.0 - self.decay) * p.data + self.decay * self.shadow[name]).clone()

    def apply_shadow(self):
        for name, p in self._trainable():
            assert name in self.shadow
            self.backup[name] = p.data
            p.data = self.shadow[name]

    def restore(self):
        for name, p in self._trainable():
            assert name in self.backup
            p.data = self.backup[name]
        self.backup = {}


def _batch_estimate(critic_model, baseline_model, bound_type, x, y, alpha_logit):
    """(mi, mi_loss-ready pieces) for one batch.  Separable critic: fused sweeps; concat critic: all-pairs kernels."""
    if bound_type not in _BOUNDS:
        raise NotImplementedError
    needs_base = bound_type in ("tuba", "interpolate")
    base = baseline_model(y) if needs_base else None
    if critic_model.critic_type == "separate" and bound_type != "interpolate":
        x_, y_ = critic_model.embed(x, y)
        return (lambda b: separable_bound(x_, y_, b, base if b == "tuba" else None)[0]), (x_ * y_).sum(dim=1)
    scores = critic_model(x, y)
    if bound_type == "interpolate":
        return (lambda b: interp_lower_bound(scores, base, alpha_logit)), scores.diag()
    return (lambda b: _scores_bound(scores, b, base if b == "tuba" else None)[0]), scores.diag()


def train_MINE(critic_model, baseline_model, bound_type, xy_loader, epochs, lr=5e-4, alpha_logit=0.0, log=False, ma_et=1,
               ma_rate=0.01, weight_decay=0.999):
    """VMI.py:287-347.  ``weight_decay`` is the EMA decay, as in the reference."""
    if baseline_model.baseline_type == 'unnormalized':
        optimizer = torch.optim.Adamax(list(critic_model.parameters()) + list(baseline_model.parameters()), lr=lr)
        emas = [EMA(critic_model, weight_decay), EMA(baseline_model, weight_decay)]
    else:
        optimizer = torch.optim.Adamax(critic_model.parameters(), lr=lr)
        emas = [EMA(critic_model, weight_decay)]
    for ema in emas:
        ema.register()
    if bound_type == 'interpolated':
        assert baseline_model.baseline_type != 'constant', "If using Interpolate bound, baseline should not be none!"
    dev = next(critic_model.parameters()).device
    history_mi = []
    for epoch in range(epochs):
        mi_epoch = 0.0
        for features in xy_loader:
            x, y = features[0].to(dev), features[1].to(dev)
            bound, diag = _batch_estimate(critic_model, baseline_model, bound_type, x, y, alpha_logit)
            if bound_type == 'mine':
                # VMI.py:304-307: mi = dv bound; loss = -(mean(t) - mean(et) / ma_et) with the moving average ma_et carried
                # over the steps and treated as a constant.  mean(et) = sum_{i != j} exp(S_ij) / n^2 comes from the fused
                # NWJ sweep: nwj = mean(diag) - e^-1 sum_{i != j} exp(S_ij) / (n (n - 1)).
                n = diag.shape[0]
                with torch.no_grad():
                    mi = bound("dv")
                mean_t = diag.mean()
                mean_et = (mean_t - bound("nwj")) * (math.e * (n - 1) / n)
                ma_et = ((1 - ma_rate) * ma_et + ma_rate * mean_et).detach()   # only ever used as a constant (VMI.py:307)
                mi_loss = -(mean_t - mean_et / ma_et)
            else:
                mi = bound(bound_type)
                mi_loss = -mi
            optimizer.zero_grad()
            mi_loss.backward()
            optimizer.step()
            for ema in emas:
                ema.update()
                ema.apply_shadow()
            mi_epoch += mi.detach().cpu().numpy()
        mi_epoch = mi_epoch / len(xy_loader)
        if log and epoch % 50 == 0:
            print('Epoch', epoch, ':', np.round(mi_epoch, 3))
        history_mi.append(mi_epoch)
    return np.asarray(history_mi)


def compute_MI(critic_type, baseline_type, bound_type, features_x, features_y, dim_x, dim_y, hidden_dim=256,
               embed_dim=128, layers=2, activation='relu', mu=0, rho=1, epochs=100, batch_size=128, lr=5e-4,
               alpha_logit=0.0, log=False, ma_et=1, ma_rate=0.01, weight_decay=0.999, estimation='mean'):
    """VMI.py:350-378: train a fresh critic (+ baseline) on (features_x, features_y), read the estimate off the history."""
    dev = features_x.device if features_x.is_cuda else torch.device("cuda")
    critic_model = CriticModel(critic_type, dim_x, dim_y, hidden_dim=hidden_dim, embed_dim=embed_dim, layers=layers,
                               activation=activation).to(dev)
    baseline_model = BaselineModel(baseline_type, dim_y, hidden_dim=hidden_dim, layers=layers, activation=activation,
                                   mu=mu, rho=rho).to(dev)
    xy_loader = DataLoader(TensorDataset(features_x.clone().detach(), features_y.clone().detach()), batch_size=batch_size)
    history_mi = train_MINE(critic_model, baseline_model, bound_type, xy_loader, epochs, lr, alpha_logit, log, ma_et,
                            ma_rate, weight_decay=weight_decay)
    del critic_model, baseline_model, xy_loader
    if estimation == 'max':
        mi_score = np.max(history_mi)
    elif estimation == 'mean':
        mi_score = np.mean(history_mi[-50:-1])
    elif estimation == 'smooth':
        from scipy.signal import savgol_filter
        history_mi = savgol_filter(history_mi, 51, 3)
        mi_score = np.mean(history_mi[-50:-1])
    else:
        raise NotImplementedError
    return mi_score, history_mi


def sample_correlated_gaussian(rho=0.5, dim=20, num_samples=1000):
    """VMI.py:389-393: (x, y) with per-coordinate correlation rho."""
    x, eps = torch.split(torch.normal(0, 1, size=(num_samples, 2 * dim)), dim, dim=1)
    y = rho * x + torch.sqrt(torch.tensor(1. - rho ** 2, dtype=torch.float32)) * eps
    return x, y


def rho_to_mi(dim, rho):
    """VMI.py:395-396."""
    return -0.5 * np.log(1 - rho ** 2) * dim
