"""Full Model.forward (Model.py:388-519) on the B200 kernels against goldens produced by the unmodified reference:
stock encoders (tiny BERT, GRU / LSTM / Conv1d) + the feature-head kernels (csrc/features.cu) + CubeMLP + composition +
classifier; outputs [output, F_F, T_F, A_F, V_F], input gradients and every non-estimator parameter gradient."""
import numpy as np
import pytest
import torch

from conftest import cfg_of, rel_err

pytestmark = pytest.mark.gpu


def _opts(c):
    from types import SimpleNamespace
    return SimpleNamespace(
        d_common=32, encoders=c["encoders"], features_compose_t=c["compose_t"], features_compose_k=c["compose_k"],
        num_class=1, activate="gelu", time_len=20, d_hiddens=[[10, 3, 32], [5, 3, 32]], d_outs=[[10, 3, 32], [5, 3, 32]],
        dropout_mlp=[0.0, 0.0, 0.0], dropout=[0.0, 0.0, 0.0, 0.0], bias=True, ln_first=False, res_project=[True, True],
        critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
        cmi_last_acticate="sigmoid")


def _inputs(c, seed, vocab):
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, vocab, size=(c["bs"], c["lt"])).astype(np.int64)
    a = rng.standard_normal((c["bs"], c["la"], 5)).astype(np.float32)
    v = rng.standard_normal((c["bs"], c["lv"], 7)).astype(np.float32)
    if c["encoders"] != "conv":
        for b in range(c["bs"]):
            a[b, c["la"] - (b % 4):] = 0
            v[b, c["lv"] - (b % 3):] = 0
    return ids, a, v


@pytest.mark.parametrize("case", ["model_00_gru_mean_mean", "model_01_lstm_sum_cat", "model_02_conv_cat_sum"])
def test_model_forward_backward_matches_reference(golden, tiny_bert, case):
    from mimrl_b200.full_model import Model
    from oracle import params as P
    rec = golden("model")[case]
    c = cfg_of(rec)
    c = {k: (str(v) if isinstance(v, (np.str_, str)) else int(v)) for k, v in c.items()}
    seed = int(rec["seed"])
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = Model(_opts(c), tiny_bert["hidden_size"], 5, 7)
    sd = {k[4:]: torch.tensor(v) for k, v in rec.items() if k.startswith("sd__")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(("vmi" in k or "vcmi" in k) for k in missing)
    model = model.cuda().train()
    ids, a, v = _inputs(c, seed, tiny_bert["vocab_size"])
    at = torch.tensor(a, device="cuda", requires_grad=True)
    vt = torch.tensor(v, device="cuda", requires_grad=True)
    mask = torch.ones(ids.shape, dtype=torch.long, device="cuda")
    outs = model(torch.tensor(ids, device="cuda"), torch.zeros_like(mask), mask, at, vt, return_features=True)
    names = ("output", "F_F", "T_F", "A_F", "V_F")
    for n, o in zip(names, outs):
        assert tuple(o.shape) == rec["out_" + n].shape
        assert rel_err(o.detach().cpu().numpy(), rec["out_" + n]) <= 1e-4, n
    w = [torch.tensor(P.features(seed + 10 + i, int(o.shape[0]), int(np.prod(o.shape[1:]))).reshape(tuple(o.shape)),
                      device="cuda") for i, o in enumerate(outs)]
    sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
    assert rel_err(at.grad.cpu().numpy(), rec["ga"]) <= 2e-4
    assert rel_err(vt.grad.cpu().numpy(), rec["gv"]) <= 2e-4
    checked = 0
    for name, p in model.named_parameters():
        key = "pg__" + name
        if key in rec:
            assert p.grad is not None, name
            want = rec[key]
            floor = 1e-6 * max(1.0, float(np.abs(want).max()))       # gradients that are exactly zero in the reference
            assert np.abs(p.grad.cpu().numpy() - want).max() <= 2e-4 * np.abs(want).max() + floor, name
            checked += 1
    assert checked > 20


def test_feature_heads_against_torch():
    """csrc/features.cu against the reference's op chain (mean / F.pad / stack; mean over k then t) on random sizes."""
    import torch.nn.functional as F
    from mimrl_b200.full_model import compose_features, feature_stack
    g = torch.Generator(device="cuda").manual_seed(0)
    for bs, lt, la, lv, T, D in ((3, 7, 20, 1, 20, 128), (130, 50, 33, 50, 50, 64), (2, 4, 4, 4, 9, 4)):
        srcs = [torch.randn(bs, n, D, device="cuda", generator=g, requires_grad=True) for n in (lt, la, lv)]
        refs = [s.detach().clone().requires_grad_(True) for s in srcs]
        x, tf, af, vf = feature_stack(*srcs, T)
        xr = torch.stack([F.pad(r, (0, 0, 0, T - r.shape[1], 0, 0), "constant", 0) for r in refs], dim=2)
        assert torch.equal(x, xr)
        for got, r in zip((tf, af, vf), refs):
            assert torch.allclose(got, r.mean(1), rtol=1e-5, atol=1e-6)
        wx = torch.randn(x.shape, device="cuda", generator=g)
        wm = torch.randn(3, bs, D, device="cuda", generator=g)
        ((x * wx).sum() + (tf * wm[0]).sum() + (af * wm[1]).sum() + (vf * wm[2]).sum()).backward()
        ((xr * wx).sum() + sum((r.mean(1) * wm[i]).sum() for i, r in enumerate(refs))).backward()
        for s, r in zip(srcs, refs):
            assert torch.allclose(s.grad, r.grad, rtol=1e-5, atol=1e-6)
    for ck, ct in (("mean", "mean"), ("sum", "mean"), ("cat", "sum"), ("mean", "cat"), ("cat", "cat")):
        x = torch.randn(5, 10, 3, 128, device="cuda", generator=g, requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        got = compose_features(x, ck, ct)
        f = {"mean": xr.mean(dim=2), "sum": xr.sum(dim=2), "cat": torch.cat(torch.split(xr, 1, dim=2), dim=-1).squeeze(2)}[ck]
        f = {"mean": f.mean(dim=1), "sum": f.sum(dim=1), "cat": torch.cat(torch.split(f, 1, dim=1), dim=-1).squeeze(1)}[ct]
        assert got.shape == f.shape and torch.allclose(got, f, rtol=1e-5, atol=1e-5), (ck, ct)
        w = torch.randn(f.shape, device="cuda", generator=g)
        (got * w).sum().backward()
        (f * w).sum().backward()
        assert torch.allclose(x.grad, xr.grad, rtol=1e-5, atol=1e-6), (ck, ct)
