"""The drop-in claim of INTEGRATION.md, executed: the reference's OWN `Model.py` (oracle/_ref, unmodified on disk) with the
names of INTEGRATION.md section 2 rebound to mimrl_b200 -- exactly what the maintainer's import patch does -- must reproduce
the goldens that the untouched reference produced:

* `Model.Model.compute_vmi_loss_stage1/2` (Model.py:305-386, the reference's code) driving our `VMIEstimator`,
  `VCMIEstimator` and `prod_knn_sample` -> the `stage` goldens (values, feature gradients, numpy RNG consumption);
* `Model.Model(opt, ...)` built by the reference's constructor out of our `MLPEncoder` / estimators, its own `forward`
  (Model.py:388-519) -> the `model` goldens (outputs and input gradients).
"""
import types

import numpy as np
import pytest
import torch

from conftest import cfg_of, load_golden, rel_err
from oracle import params as P

pytestmark = pytest.mark.gpu
TOL = 1e-4
STAGE = load_golden("stage")
MODEL = load_golden("model")
SWAPPED = ("MLPEncoder", "MLP_For_CMI", "prod_knn_sample", "VMIEstimator", "VCMIEstimator", "CriticModel", "BaselineModel",
           "dv_lower_bound", "mine_lower_bound_test", "tuba_lower_bound", "nwj_lower_bound", "infonce_lower_bound",
           "js_fgan_lower_bound", "js_lower_bound", "smile_lower_bound", "interp_lower_bound")


@pytest.fixture()
def patched_reference():
    """The reference's Model module with the INTEGRATION.md names rebound (and restored afterwards)."""
    import __graft_entry__ as g
    g.build()
    from oracle import ref_shim as R
    if R.locate() is None:
        pytest.skip("oracle/_ref not present (make -C oracle _ref needs /root/reference)")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = R.import_reference(cpu=False, random_bert=False)
    import mimrl_b200.mlp_process as OP
    import mimrl_b200.model as OM
    import mimrl_b200.vmi as OV
    saved = {n: getattr(ref.Model, n) for n in SWAPPED if hasattr(ref.Model, n)}
    from_model = ("MLP_For_CMI", "prod_knn_sample", "VMIEstimator", "VCMIEstimator")
    for n in SWAPPED:
        src = OP if n == "MLPEncoder" else (OM if n in from_model else OV)
        setattr(ref.Model, n, getattr(src, n))
    yield ref
    for n in SWAPPED:
        if n in saved:
            setattr(ref.Model, n, saved[n])
        else:
            delattr(ref.Model, n)


def T(a, **kw):
    return torch.tensor(a, device="cuda", **kw)


@pytest.mark.parametrize("case", sorted(STAGE))
def test_reference_stage_functions_on_swapped_estimators(patched_reference, case):
    Model = patched_reference.Model
    rec = STAGE[case]
    c = cfg_of(rec)
    seed = int(rec["seed"])
    d, hidden = c["d"], c["hidden"]
    ns = types.SimpleNamespace(d_common=d, k_neighbor=c["k"], radius=1.0)
    for i, n in enumerate(["f_t", "f_a", "f_v", "t_a", "t_v"]):
        est = Model.VMIEstimator(c["critic"], c["baseline"], c["bound"], d, hidden, d, 2, "relu", 0, 1)
        assert type(est).__module__.startswith("mimrl_b200")
        sd = P.vmi_state_dict(P.vmi_params(seed + 10 + i, c["critic"], c["baseline"], d, hidden, d, 2))
        est.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
        setattr(ns, "vmi_estimator_" + n, est.cuda())
    for i, n in enumerate(["ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"]):
        est = Model.VCMIEstimator(d, hidden, 2, "relu", c["k"], 1.0, c["act"])
        stack = P.vcmi_params(seed + 30 + i, d, hidden)
        if c["act"] == "hardtanh":
            stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
        est.load_state_dict({k: torch.tensor(v) for k, v in P.vcmi_state_dict(stack).items()})
        setattr(ns, "vcmi_estimator_" + n, est.cuda())
    feats = {n: P.features(seed + 50 + i, c["bs"], d) for i, n in enumerate(["F", "T", "A", "V"])}
    labels = P.features(seed + 60, c["bs"], 1)[:, 0]
    pools = {n: T(P.features(seed + 70 + i, c["N"], d)) for i, n in enumerate(["F", "T", "A", "V"])}
    pool_c = T(P.features(seed + 80, c["N"], 1))
    for stage in (1, 2):
        fn = Model.Model.compute_vmi_loss_stage1 if stage == 1 else Model.Model.compute_vmi_loss_stage2   # the reference's code
        ft = {n: T(v, requires_grad=True) for n, v in feats.items()}
        np.random.seed(seed + stage)
        mis, losses = fn(ns, None, T(labels), ft["F"], ft["T"], ft["A"], ft["V"], pool_c, pools["F"], pools["T"],
                         pools["A"], pools["V"])
        got_mis = np.array([float(m) for m in mis])
        got_losses = np.array([float(m) for m in losses])
        assert np.allclose(got_mis, rec[f"s{stage}_mis"], rtol=TOL, atol=TOL), (got_mis, rec[f"s{stage}_mis"])
        assert np.allclose(got_losses, rec[f"s{stage}_losses"], rtol=TOL, atol=TOL)
        total = sum(l * (0.1 * (i + 1)) for i, l in enumerate(losses))
        gs = torch.autograd.grad(total, [ft[n] for n in "FTAV"])
        for n, g in zip("FTAV", gs):
            want = rec[f"s{stage}_g{n}"]
            assert np.abs(g.cpu().numpy() - want).max() <= 2 * TOL * np.abs(want).max() + 1e-7, (stage, n)


def test_reference_model_class_on_swapped_modules(patched_reference, tiny_bert):
    """Model.Model.__init__ / forward are the reference's; MLPEncoder and the eleven estimators it constructs are ours."""
    Model = patched_reference.Model
    case = "model_00_gru_mean_mean"
    rec = MODEL[case]
    c = {k: (str(v) if isinstance(v, (np.str_, str)) else int(v)) for k, v in cfg_of(rec).items()}
    seed = int(rec["seed"])
    opt = types.SimpleNamespace(
        d_common=32, encoders=c["encoders"], features_compose_t=c["compose_t"], features_compose_k=c["compose_k"],
        num_class=1, activate="gelu", time_len=20, d_hiddens=[[10, 3, 32], [5, 3, 32]], d_outs=[[10, 3, 32], [5, 3, 32]],
        dropout_mlp=[0.0, 0.0, 0.0], dropout=[0.0, 0.0, 0.0, 0.0], bias=True, ln_first=False, res_project=[True, True],
        critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
        cmi_last_acticate="sigmoid")
    model = Model.Model(opt, tiny_bert["hidden_size"], 5, 7)
    assert type(model.mlp_encoder).__module__.startswith("mimrl_b200")
    assert type(model.vmi_estimator_f_t).__module__.startswith("mimrl_b200")
    sd = {k[4:]: torch.tensor(v) for k, v in rec.items() if k.startswith("sd__")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(("vmi" in k or "vcmi" in k) for k in missing)
    model = model.cuda().train()
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, tiny_bert["vocab_size"], size=(c["bs"], c["lt"])).astype(np.int64)
    a = rng.standard_normal((c["bs"], c["la"], 5)).astype(np.float32)
    v = rng.standard_normal((c["bs"], c["lv"], 7)).astype(np.float32)
    for b in range(c["bs"]):
        a[b, c["la"] - (b % 4):] = 0
        v[b, c["lv"] - (b % 3):] = 0
    at, vt = T(a, requires_grad=True), T(v, requires_grad=True)
    mask = torch.ones(ids.shape, dtype=torch.long, device="cuda")
    outs = model(T(ids), torch.zeros_like(mask), mask, at, vt, return_features=True)
    for n, o in zip(("output", "F_F", "T_F", "A_F", "V_F"), outs):
        assert rel_err(o.detach().cpu().numpy(), rec["out_" + n]) <= TOL, n
    w = [T(P.features(seed + 10 + i, int(o.shape[0]), int(np.prod(o.shape[1:]))).reshape(tuple(o.shape))) for i, o in enumerate(outs)]
    sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
    assert rel_err(at.grad.cpu().numpy(), rec["ga"]) <= 2e-4
    assert rel_err(vt.grad.cpu().numpy(), rec["gv"]) <= 2e-4
