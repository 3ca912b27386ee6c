#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, read-only) on CPU at fixed seeds.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no
/root/reference):

    python oracle/gen_golden.py

Shims: matplotlib is absent -> empty stub modules;
no CUDA here -> Tensor.cuda / Module.cuda are identity.  Nothing else of the
reference is altered; every value stored below is produced by the reference's
own VMI.py / Model.py / MLPProcess.py code.

Inputs and weights are NOT stored when they can be regenerated from
oracle/params.py with the stored seed; outputs and gradients are stored.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MIMRL_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import params as P  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import VMI, Model, MLPProcess  # noqa: E401
    return VMI, Model, MLPProcess


def t(a):
    return torch.tensor(np.asarray(a))


def sd_t(sd):
    return {k: t(v) for k, v in sd.items()}


BOUNDS = ["dv", "mine", "tuba", "nwj", "infonce", "js_fgan", "js", "smile", "interpolate"]


def baseline_for(bound):
    return "unnormalized" if bound in ("tuba", "interpolate") else "constant"


def gen_vmi(Model, out):
    """VMIEstimator (Model.py:108-148) over every critic x bound, small dims."""
    cases = []
    for critic in ("separate", "concat"):
        for bound in BOUNDS:
            cases.append(dict(critic=critic, bound=bound, baseline=baseline_for(bound),
                              B=24, d=16, hidden=32, embed=16, layers=2, scale=1.0))
    # baseline variants and a ragged / larger-score case
    cases.append(dict(critic="separate", bound="tuba", baseline="gaussain", B=19, d=16, hidden=32, embed=16, layers=2, scale=0.3))
    cases.append(dict(critic="separate", bound="tuba", baseline="constant", B=19, d=16, hidden=32, embed=16, layers=2, scale=1.0))
    cases.append(dict(critic="separate", bound="infonce", baseline="constant", B=37, d=16, hidden=32, embed=16, layers=1, scale=3.0))
    cases.append(dict(critic="separate", bound="smile", baseline="constant", B=37, d=16, hidden=32, embed=16, layers=2, scale=4.0))
    cases.append(dict(critic="separate", bound="infonce", baseline="constant", B=2, d=16, hidden=32, embed=16, layers=2, scale=1.0))
    # Model.py:285 sizes: d_common=embed=128, hidden=256, layers=2 (config 1 shape, bs=128)
    for bound in ("infonce", "nwj", "js", "dv", "smile"):
        cases.append(dict(critic="separate", bound=bound, baseline="constant", B=128, d=128, hidden=256, embed=128, layers=2, scale=1.0))
    cases.append(dict(critic="separate", bound="infonce", baseline="constant", B=200, d=128, hidden=256, embed=128, layers=2, scale=2.0))
    cases.append(dict(critic="concat", bound="nwj", baseline="constant", B=48, d=128, hidden=256, embed=128, layers=2, scale=1.0))
    cases.append(dict(critic="concat", bound="js", baseline="constant", B=48, d=128, hidden=256, embed=128, layers=2, scale=1.0))

    for ci, c in enumerate(cases):
        seed = 1000 + ci
        prm = P.vmi_params(seed, c["critic"], c["baseline"], c["d"], c["hidden"], c["embed"], c["layers"])
        x, y = P.features(seed + 7, c["B"], c["d"], scale=c["scale"], corr=0.6)
        est = Model.VMIEstimator(c["critic"], c["baseline"], c["bound"], c["d"], c["hidden"], c["embed"],
                                 c["layers"], "relu", 0, 1)
        missing = est.load_state_dict(sd_t(P.vmi_state_dict(prm)), strict=True)
        xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
        mi, loss = est(xt, yt)
        loss.backward()
        rec = dict(mi=mi.detach().numpy(), loss=loss.detach().numpy(),
                   gx=xt.grad.numpy(), gy=yt.grad.numpy(), seed=seed)
        big = c["hidden"] >= 256
        for name, p_ in est.named_parameters():
            g = p_.grad
            if g is None:
                continue
            if big:   # keep fixtures small: moments + a strided sample of each grad
                gn = g.numpy().ravel()
                rec["pgs__" + name] = np.array([gn.sum(dtype=np.float64), np.abs(gn).sum(dtype=np.float64),
                                                np.sqrt((gn.astype(np.float64) ** 2).sum())])
                rec["pgx__" + name] = gn[:: max(1, gn.size // 64)][:64].copy()
            else:
                rec["pg__" + name] = g.numpy()
        for k, v in c.items():
            rec["cfg_" + k] = np.array(v)
        out[f"vmi_{ci:02d}_{c['critic']}_{c['bound']}_{c['baseline']}_B{c['B']}_d{c['d']}"] = rec


def gen_bounds(VMI, out):
    """Free bound functions (VMI.py:136-250) on a raw score matrix."""
    rng = np.random.default_rng(77)
    for B in (5, 33):
        S = (rng.standard_normal((B, B)) * 1.5).astype(np.float32)
        a = (rng.standard_normal((B, 1)) * 0.5).astype(np.float32)
        rec = dict(S=S, a=a)
        fns = dict(dv=lambda s: VMI.dv_lower_bound(s), tuba=lambda s: VMI.tuba_lower_bound(s, t(a)),
                   tuba_nobase=lambda s: VMI.tuba_lower_bound(s),
                   nwj=VMI.nwj_lower_bound, infonce=VMI.infonce_lower_bound, js_fgan=VMI.js_fgan_lower_bound,
                   js=VMI.js_lower_bound, smile=VMI.smile_lower_bound,
                   interpolate=lambda s: VMI.interp_lower_bound(s, t(a), 0.01))
        for name, fn in fns.items():
            st = t(S).requires_grad_(True)
            v = fn(st)
            v.backward()
            rec["val_" + name] = v.detach().numpy()
            rec["grad_" + name] = st.grad.numpy()
        out[f"bounds_B{B}"] = rec


class _Recorder:
    """Wrap sklearn's NearestNeighbors.kneighbors to record what the reference
    asked for and got (Model.py:82-86), without changing the result."""

    def __init__(self):
        import sklearn.neighbors as skn
        self.skn = skn
        self.calls = []
        self._orig = skn.NearestNeighbors.kneighbors

    def __enter__(self):
        rec = self

        def wrapped(self_nn, X=None, n_neighbors=None, return_distance=True):
            r = rec._orig(self_nn, X, n_neighbors, return_distance)
            rec.calls.append(dict(idx=np.asarray(r).copy(), method=self_nn._fit_method))
            return r
        self.skn.NearestNeighbors.kneighbors = wrapped
        return self

    def __exit__(self, *a):
        self.skn.NearestNeighbors.kneighbors = self._orig


def gen_knn(Model, out):
    """prod_knn_sample (Model.py:75-106)."""
    cases = [
        dict(N=1284, bs=128, k=2, wx=128, wy=1, wz=128, dup=0),     # config 1 shape, brute route
        dict(N=1284, bs=128, k=2, wx=128, wy=128, wz=1, dup=0),     # Z = labels -> kd_tree route (F4)
        dict(N=700, bs=100, k=16, wx=128, wy=1, wz=128, dup=0),     # bs % k != 0 (N4)
        dict(N=300, bs=64, k=4, wx=16, wy=16, wz=16, dup=0),        # narrow equal widths (brute: 16 > 15)
        dict(N=300, bs=32, k=3, wx=8, wy=1, wz=8, dup=0),           # width 8 -> kd_tree, multi-dim
        dict(N=400, bs=64, k=4, wx=128, wy=1, wz=128, dup=60),      # duplicated key rows (ties, H3)
        dict(N=40, bs=16, k=2, wx=128, wy=1, wz=128, dup=0),        # tiny pool
    ]
    for ci, c in enumerate(cases):
        seed = 2000 + ci
        X = P.features(seed, c["N"], c["wx"])
        Y = P.features(seed + 1, c["N"], c["wy"])
        Z = P.features(seed + 2, c["N"], c["wz"])
        if c["dup"]:
            Z[c["N"] - c["dup"]:] = Z[: c["dup"]]       # exact duplicate keys
        np.random.seed(seed)
        with _Recorder() as r:
            bx, by, bz = Model.prod_knn_sample(t(X), t(Y), t(Z), c["bs"], c["k"], 1.0)
        state_after = np.random.get_state()
        np.random.seed(seed)
        ids = np.random.permutation(c["N"])[: c["bs"] // c["k"]]
        assert np.array_equal(np.random.get_state()[1], state_after[1]) and np.random.get_state()[2] == state_after[2]
        rec = dict(seed=seed, ids=ids, nbr=r.calls[0]["idx"], method=np.array(r.calls[0]["method"]),
                   bx=bx.detach().numpy(), by=by.detach().numpy(), bz=bz.detach().numpy(),
                   leaf=np.array([bx.is_leaf and bx.requires_grad, by.is_leaf and by.requires_grad,
                                  bz.is_leaf and bz.requires_grad]))
        for k_, v in c.items():
            rec["cfg_" + k_] = np.array(v)
        out[f"knn_{ci:02d}_N{c['N']}_bs{c['bs']}_k{c['k']}_wz{c['wz']}"] = rec


def gen_vcmi(Model, out):
    """VCMIEstimator.forward + estimate_cmi (Model.py:150-225)."""
    cases = [
        dict(act="hardtanh", bs=20, nprod=20, embed=16, hidden=32, wy=16, scale=1.0),
        dict(act="sigmoid", bs=20, nprod=20, embed=16, hidden=32, wy=16, scale=1.0),
        dict(act="hardtanh", bs=21, nprod=16, embed=16, hidden=32, wy=1, scale=1.0),    # N4 + N5
        dict(act="sigmoid", bs=21, nprod=16, embed=16, hidden=32, wy=1, scale=6.0),     # clamp active
        dict(act="hardtanh", bs=128, nprod=128, embed=128, hidden=256, wy=1, scale=1.0),
    ]
    for ci, c in enumerate(cases):
        seed = 3000 + ci
        stack = P.vcmi_params(seed, c["embed"], c["hidden"])
        if c["act"] == "hardtanh":
            # centre the logits in (1e-4, 1-1e-4) so the hardtanh is not saturated everywhere
            stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
        fx = P.features(seed + 1, c["bs"], c["embed"], c["scale"])
        fy = P.features(seed + 2, c["bs"], c["wy"], c["scale"])
        fz = P.features(seed + 3, c["bs"], c["embed"], c["scale"])
        kx = P.features(seed + 4, c["nprod"], c["embed"], c["scale"])
        ky = P.features(seed + 5, c["nprod"], c["embed"], c["scale"])
        kz = P.features(seed + 6, c["nprod"], c["embed"], c["scale"])
        est = Model.VCMIEstimator(c["embed"], c["hidden"], 2, "relu", 2, 1.0, c["act"])
        est.load_state_dict(sd_t(P.vcmi_state_dict(stack)), strict=True)
        ins = [t(a).requires_grad_(True) for a in (fx, fy, fz, kx, ky, kz)]
        cmi, loss = est(*ins)
        # stage-1 uses loss, stage-2 uses cmi: store grads of both, separately
        names = [n for n, _ in est.named_parameters()]
        g_loss = torch.autograd.grad(loss, ins + list(est.parameters()), retain_graph=True, allow_unused=True)
        g_cmi = torch.autograd.grad(cmi, ins + list(est.parameters()), allow_unused=True)
        rec = dict(seed=seed, cmi=cmi.detach().numpy(), loss=loss.detach().numpy())
        big = c["hidden"] >= 256
        for tag, gs in (("gl", g_loss), ("gc", g_cmi)):
            for i, nm in enumerate(["fx", "fy", "fz", "kx", "ky", "kz"]):
                rec[f"{tag}_{nm}"] = gs[i].numpy() if gs[i] is not None else np.zeros(0, np.float32)
            for i, nm in enumerate(names):
                g = gs[6 + i].numpy()
                if big:
                    gn = g.ravel()
                    rec[f"{tag}s__{nm}"] = np.array([gn.sum(dtype=np.float64), np.abs(gn).sum(dtype=np.float64),
                                                     np.sqrt((gn.astype(np.float64) ** 2).sum())])
                else:
                    rec[f"{tag}p__{nm}"] = g
        for k_, v in c.items():
            rec["cfg_" + k_] = np.array(v)
        out[f"vcmi_{ci:02d}_{c['act']}_bs{c['bs']}_np{c['nprod']}"] = rec


def gen_cubemlp(MLPProcess, out):
    """MLPEncoder (MLPProcess.py:126-137), both LN placements."""
    cases = [
        dict(act="gelu", bs=3, d_in=[10, 3, 16], d_hiddens=[[7, 4, 16], [4, 3, 8]], d_outs=[[5, 3, 16], [4, 3, 8]],
             bias=True, ln_first=False, res=[True, True]),
        dict(act="gelu", bs=3, d_in=[10, 3, 16], d_hiddens=[[7, 4, 16], [4, 3, 8]], d_outs=[[5, 3, 16], [4, 3, 8]],
             bias=True, ln_first=True, res=[True, True]),
        dict(act="relu", bs=2, d_in=[6, 3, 8], d_hiddens=[[12, 5, 16]], d_outs=[[6, 3, 8]],
             bias=False, ln_first=False, res=[False]),
        # README.md:17 configuration: --d_hiddens 50-3-128=10-3-128, time_len 100
        dict(act="gelu", bs=2, d_in=[100, 3, 128], d_hiddens=[[50, 3, 128], [10, 3, 128]],
             d_outs=[[50, 3, 128], [10, 3, 128]], bias=True, ln_first=False, res=[True, True]),
    ]
    for ci, c in enumerate(cases):
        seed = 4000 + ci
        blocks = P.cubemlp_params(seed, c["d_in"], c["d_hiddens"], c["d_outs"], c["bias"], c["ln_first"], c["res"])
        enc = MLPProcess.MLPEncoder(activate=c["act"], d_in=c["d_in"], d_hiddens=c["d_hiddens"], d_outs=c["d_outs"],
                                    dropouts=[0.0, 0.0, 0.0], bias=c["bias"], ln_first=c["ln_first"],
                                    res_project=c["res"])
        enc.load_state_dict(sd_t(P.cubemlp_state_dict(blocks)), strict=True)
        x = P.features(seed + 1, c["bs"] * c["d_in"][0] * c["d_in"][1], c["d_in"][2]).reshape(
            c["bs"], c["d_in"][0], c["d_in"][1], c["d_in"][2])
        xt = t(x).requires_grad_(True)
        y = enc(xt, mask=None)
        w = P.features(seed + 2, int(np.prod(y.shape[:-1])), y.shape[-1]).reshape(tuple(y.shape))
        (y * t(w)).sum().backward()
        rec = dict(seed=seed, y=y.detach().numpy(), gx=xt.grad.numpy())
        big = c["d_in"][2] >= 128
        for name, p_ in enc.named_parameters():
            g = p_.grad.numpy()
            if big:
                gn = g.ravel()
                rec["pgs__" + name] = np.array([gn.sum(dtype=np.float64), np.abs(gn).sum(dtype=np.float64),
                                                np.sqrt((gn.astype(np.float64) ** 2).sum())])
            else:
                rec["pg__" + name] = g
        if big:
            rec["y"] = rec["y"][:, :, :, ::8].copy()
            rec["gx"] = rec["gx"][:, ::5, :, ::8].copy()
        rec["cfg"] = np.array(repr(c))
        out[f"cubemlp_{ci:02d}_{c['act']}_{'lnfirst' if c['ln_first'] else 'lnlast'}_d{c['d_in'][2]}"] = rec


def gen_stage(Model, out):
    """compute_vmi_loss_stage1/2 (Model.py:305-386) driven on a stand-in
    ``self`` that owns real reference estimators; avoids building BERT."""
    for ci, c in enumerate([
        dict(critic="separate", bound="infonce", baseline="constant", k=2, act="hardtanh", bs=32, N=200, d=16, hidden=32),
        dict(critic="separate", bound="nwj", baseline="constant", k=4, act="sigmoid", bs=30, N=150, d=16, hidden=32),
    ]):
        seed = 5000 + ci
        d, hidden = c["d"], c["hidden"]
        ns = types.SimpleNamespace(d_common=d, k_neighbor=c["k"], radius=1.0)
        vmi_names = ["f_t", "f_a", "f_v", "t_a", "t_v"]
        vcmi_names = ["ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"]
        for i, n in enumerate(vmi_names):
            est = Model.VMIEstimator(c["critic"], c["baseline"], c["bound"], d, hidden, d, 2, "relu", 0, 1)
            est.load_state_dict(sd_t(P.vmi_state_dict(P.vmi_params(seed + 10 + i, c["critic"], c["baseline"], d, hidden, d, 2))))
            setattr(ns, "vmi_estimator_" + n, est)
        for i, n in enumerate(vcmi_names):
            est = Model.VCMIEstimator(d, hidden, 2, "relu", c["k"], 1.0, c["act"])
            stack = P.vcmi_params(seed + 30 + i, d, hidden)
            if c["act"] == "hardtanh":
                stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
            est.load_state_dict(sd_t(P.vcmi_state_dict(stack)))
            setattr(ns, "vcmi_estimator_" + n, est)
        feats = {n: P.features(seed + 50 + i, c["bs"], d) for i, n in enumerate(["F", "T", "A", "V"])}
        labels = P.features(seed + 60, c["bs"], 1)[:, 0]
        pools = {n: P.features(seed + 70 + i, c["N"], d) for i, n in enumerate(["F", "T", "A", "V"])}
        pool_c = P.features(seed + 80, c["N"], 1)
        rec = dict(seed=seed)
        for stage in (1, 2):
            fn = Model.Model.compute_vmi_loss_stage1 if stage == 1 else Model.Model.compute_vmi_loss_stage2
            ft = {n: t(v).requires_grad_(True) for n, v in feats.items()}
            np.random.seed(seed + stage)
            mis, losses = fn(ns, None, t(labels), ft["F"], ft["T"], ft["A"], ft["V"], t(pool_c),
                             t(pools["F"]), t(pools["T"]), t(pools["A"]), t(pools["V"]))
            rec[f"s{stage}_mis"] = np.array([float(m) for m in mis], dtype=np.float32)
            rec[f"s{stage}_losses"] = np.array([float(m) for m in losses], dtype=np.float32)
            total = sum(l * (0.1 * (i + 1)) for i, l in enumerate(losses))
            gs = torch.autograd.grad(total, [ft[n] for n in "FTAV"])
            for n, g in zip("FTAV", gs):
                rec[f"s{stage}_g{n}"] = g.numpy()
        for k_, v in c.items():
            rec["cfg_" + k_] = np.array(v)
        out[f"stage_{ci:02d}_{c['bound']}_k{c['k']}"] = rec


TINY_BERT = dict(vocab_size=120, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                 max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def tiny_bert_patch():
    """Model.py:243-244 load bert-base-uncased from the local cache; the goldens use a tiny random-init BERT of the
    same class instead (d_t = 32), so the fixture stays small.  The BERT encoder itself is out of scope."""
    import transformers
    transformers.BertConfig.from_pretrained = staticmethod(
        lambda *a, **k: transformers.BertConfig(output_hidden_states=True, **TINY_BERT))
    transformers.BertModel.from_pretrained = staticmethod(lambda *a, config=None, **k: transformers.BertModel(config))


MODEL_CASES = [
    dict(encoders="gru", compose_t="mean", compose_k="mean", bs=6, lt=12, la=15, lv=9),
    dict(encoders="lstm", compose_t="sum", compose_k="cat", bs=5, lt=20, la=11, lv=20),
    dict(encoders="conv", compose_t="cat", compose_k="sum", bs=4, lt=7, la=7, lv=7),
]


def model_opts(c):
    return types.SimpleNamespace(
        d_common=32, encoders=c["encoders"], features_compose_t=c["compose_t"], features_compose_k=c["compose_k"],
        num_class=1, activate="gelu", time_len=20, d_hiddens=[[10, 3, 32], [5, 3, 32]], d_outs=[[10, 3, 32], [5, 3, 32]],
        dropout_mlp=[0.0, 0.0, 0.0], dropout=[0.0, 0.0, 0.0, 0.0], bias=True, ln_first=False, res_project=[True, True],
        critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
        cmi_last_acticate="sigmoid")


def model_inputs(c, seed):
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, TINY_BERT["vocab_size"], size=(c["bs"], c["lt"])).astype(np.int64)
    a = rng.standard_normal((c["bs"], c["la"], 5)).astype(np.float32)
    v = rng.standard_normal((c["bs"], c["lv"], 7)).astype(np.float32)
    if c["encoders"] != "conv":                      # ragged sequences: trailing all-zero frames (Model.py:425-432)
        for b in range(c["bs"]):
            a[b, c["la"] - (b % 4):] = 0
            v[b, c["lv"] - (b % 3):] = 0
    return ids, a, v


def gen_model(Model, out):
    """Model.forward (Model.py:388-519): encoders (tiny BERT / GRU|LSTM|Conv, stock torch) + feature heads + CubeMLP +
    composition + classifier.  Stores every non-estimator weight, the five outputs and input/parameter gradients."""
    tiny_bert_patch()
    for ci, c in enumerate(MODEL_CASES):
        seed = 6000 + ci
        torch.manual_seed(seed)
        model = Model.Model(model_opts(c), TINY_BERT["hidden_size"], 5, 7)
        model.train()
        ids, a, v = model_inputs(c, seed)
        at, vt = t(a).requires_grad_(True), t(v).requires_grad_(True)
        mask = torch.ones(ids.shape, dtype=torch.long)
        outs = model(t(ids), torch.zeros_like(mask), mask, at, vt, return_features=True)
        w = [t(P.features(seed + 10 + i, int(o.shape[0]), int(np.prod(o.shape[1:]))).reshape(tuple(o.shape)))
             for i, o in enumerate(outs)]
        sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
        rec = dict(seed=seed, ga=at.grad.numpy(), gv=vt.grad.numpy())
        for name, o in zip(("output", "F_F", "T_F", "A_F", "V_F"), outs):
            rec["out_" + name] = o.detach().numpy()
        for name, p_ in model.state_dict().items():
            if "vmi" not in name and "vcmi" not in name:
                rec["sd__" + name] = p_.numpy()
        for name, p_ in model.named_parameters():
            if p_.grad is not None and "vmi" not in name and "vcmi" not in name:
                rec["pg__" + name] = p_.grad.numpy()
        for k_, v_ in c.items():
            rec["cfg_" + k_] = np.array(v_)
        out[f"model_{ci:02d}_{c['encoders']}_{c['compose_t']}_{c['compose_k']}"] = rec


def gen_mine(VMI, out):
    """compute_MI / train_MINE / EMA (VMI.py:253-378) on the correlated-Gaussian known answer (VMI.py:389-396)."""
    for ci, c in enumerate([
        dict(critic="separate", baseline="constant", bound="infonce", dim=20, rho=0.8, n=1024, epochs=60, bs=128),
        dict(critic="separate", baseline="constant", bound="mine", dim=20, rho=0.8, n=1024, epochs=60, bs=128),
        dict(critic="separate", baseline="unnormalized", bound="tuba", dim=20, rho=0.8, n=1024, epochs=60, bs=128),
        dict(critic="separate", baseline="constant", bound="smile", dim=20, rho=0.8, n=1024, epochs=60, bs=128),
    ]):
        seed = 7000 + ci
        torch.manual_seed(seed)
        x, y = VMI.sample_correlated_gaussian(rho=c["rho"], dim=c["dim"], num_samples=c["n"])
        torch.manual_seed(seed + 1)
        score, hist = VMI.compute_MI(c["critic"], c["baseline"], c["bound"], x, y, c["dim"], c["dim"], hidden_dim=64,
                                     embed_dim=32, epochs=c["epochs"], batch_size=c["bs"], lr=5e-4, estimation="mean")
        rec = dict(seed=seed, score=np.array(score), history=np.asarray(hist, dtype=np.float64),
                   true_mi=np.array(VMI.rho_to_mi(c["dim"], c["rho"])))
        for k_, v_ in c.items():
            rec["cfg_" + k_] = np.array(v_)
        out[f"mine_{ci:02d}_{c['bound']}"] = rec


def main():
    VMI, Model, MLPProcess = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)
    groups = dict(vmi=lambda o: gen_vmi(Model, o), bounds=lambda o: gen_bounds(VMI, o),
                  knn=lambda o: gen_knn(Model, o), vcmi=lambda o: gen_vcmi(Model, o),
                  cubemlp=lambda o: gen_cubemlp(MLPProcess, o), stage=lambda o: gen_stage(Model, o),
                  mine=lambda o: gen_mine(VMI, o), model=lambda o: gen_model(Model, o))
    only = sys.argv[1:]
    for gname, fn in groups.items():
        if only and gname not in only:
            continue
        recs = {}
        fn(recs)
        flat = {}
        for case, rec in recs.items():
            for k, v in rec.items():
                flat[f"{case}::{k}"] = np.asarray(v)
        path = os.path.join(OUT, f"{gname}.npz")
        np.savez_compressed(path, **flat)
        print(f"{gname}: {len(recs)} cases -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
