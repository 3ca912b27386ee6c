"""Standalone estimator (VMI.py:253-396): compute_MI / train_MINE / EMA on the fused kernels follow the reference's own
training trajectory on the correlated-Gaussian known answer (same seeds: same data, same initial weights)."""
import numpy as np
import pytest
import torch

from conftest import cfg_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["mine_00_infonce", "mine_01_mine", "mine_02_tuba", "mine_03_smile"])
def test_compute_mi_tracks_reference_history(golden, case):
    from mimrl_b200 import mine as M
    rec = golden("mine")[case]
    c = cfg_of(rec)
    seed = int(rec["seed"])
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    x, y = M.sample_correlated_gaussian(rho=float(c["rho"]), dim=int(c["dim"]), num_samples=int(c["n"]))
    torch.manual_seed(seed + 1)
    score, hist = M.compute_MI(str(c["critic"]), str(c["baseline"]), str(c["bound"]), x.cuda(), y.cuda(), int(c["dim"]),
                               int(c["dim"]), hidden_dim=64, embed_dim=32, epochs=int(c["epochs"]), batch_size=int(c["bs"]),
                               lr=5e-4, estimation="mean")
    want = rec["history"]
    assert hist.shape == want.shape
    # the same optimisation path up to fp32 rounding: 480 Adamax steps apart, the histories agree to a few per cent
    assert np.abs(hist - want).max() <= 0.03 * np.abs(want).max(), (hist[-5:], want[-5:])
    assert abs(score - float(rec["score"])) <= 0.03 * abs(float(rec["score"]))
    true_mi = float(rec["true_mi"])
    assert M.rho_to_mi(int(c["dim"]), float(c["rho"])) == pytest.approx(true_mi)
    assert 0.15 * true_mi < hist[-1] < true_mi          # a lower bound, still climbing towards -0.5 d log(1 - rho^2)
    assert hist[-1] > hist[len(hist) // 2] > hist[0]


def test_ema_matches_reference_semantics():
    from mimrl_b200.mine import EMA
    lin = torch.nn.Linear(3, 2).cuda()
    ema = EMA(lin, 0.9)
    ema.register()
    w0 = lin.weight.data.clone()
    with torch.no_grad():
        lin.weight.add_(1.0)
    ema.update()
    assert torch.allclose(ema.shadow["weight"], 0.1 * (w0 + 1.0) + 0.9 * w0)
    ema.apply_shadow()
    assert torch.equal(lin.weight.data, ema.shadow["weight"])
    ema.restore()
    assert torch.allclose(lin.weight.data, w0 + 1.0) and ema.backup == {}
