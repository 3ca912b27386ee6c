"""Import and drive the UNMODIFIED reference (kiva12138/MIMRL) on CPU.

TEST / BASELINE INFRASTRUCTURE ONLY: used by ``bench.py --impl reference``,
``bench.py``'s ``cpu_baseline`` leg and ``oracle/gen_golden.py``.  Nothing under
``mimrl_b200/`` imports it.

The reference is four plain Python files.  ``make -C oracle _ref`` copies them,
byte for byte, from ``/root/reference`` into the git-ignored ``oracle/_ref/``
(so they travel to the GPU box like a built ``.so`` and stay out of history);
this module puts that directory on ``sys.path`` and applies the import shims of
SURVEY.md section 8(c):

* ``matplotlib`` is not installed            -> empty stub modules (VMI.py:4 only plots in show_history_mi)
* no CUDA on the CPU arm                     -> ``Tensor.cuda`` / ``Module.cuda`` are identity (Model.py:106,179,187)
* ``bert-base-uncased`` weights not cached   -> ``from_pretrained`` builds a random-init bert-base (Model.py:243-244)

No reference arithmetic is altered.
"""
from __future__ import annotations

import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILES = ("VMI.py", "Model.py", "MLPProcess.py", "Utils.py", "Customization.py")
_CANDIDATES = (os.path.join(HERE, "_ref"), os.environ.get("MIMRL_REFERENCE", "/root/reference"))


def locate():
    """Directory holding the reference sources, or None (-> the caller falls back to the numpy port)."""
    for d in _CANDIDATES:
        if d and all(os.path.exists(os.path.join(d, f)) for f in REF_FILES[:4]):
            return d
    return None


def import_reference(path=None, cpu=True, random_bert=True):
    """Returns the reference modules as a namespace (VMI, Model, MLPProcess, Customization | None)."""
    import torch
    path = path or locate()
    if path is None:
        raise ImportError("reference sources not found: run `make -C oracle _ref` where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if random_bert:
        import transformers

        def _cfg(*a, **k):
            return transformers.BertConfig(output_hidden_states=bool(k.get("output_hidden_states", False)))

        def _model(*a, config=None, **k):
            return transformers.BertModel(config if config is not None else transformers.BertConfig())
        transformers.BertConfig.from_pretrained = staticmethod(_cfg)
        transformers.BertModel.from_pretrained = staticmethod(_model)
    if path not in sys.path:
        sys.path.insert(0, path)
    import MLPProcess, Model, VMI  # noqa: E401
    try:
        import Customization
    except Exception:                                   # needs BertTokenizer only at call time; optional here
        Customization = None
    return types.SimpleNamespace(VMI=VMI, Model=Model, MLPProcess=MLPProcess, Customization=Customization, path=path)


def set_threads(n=None):
    """Use every host core even under torchrun (which exports OMP_NUM_THREADS=1 to its workers)."""
    import torch
    n = n or os.cpu_count() or 1
    torch.set_num_threads(n)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return torch.get_num_threads()


# --------------------------------------------------------------------------
# BASELINE configs[1] on the reference: VMIEstimator separate / constant / infonce, forward + backward
# --------------------------------------------------------------------------


def vmi_step_fn(ref, B, d=128, hidden=256, embed=128, layers=2, seed=0):
    """Model.VMIEstimator.forward + mi_loss.backward() (Model.py:108-148) on torch CPU, fp32."""
    import torch
    torch.manual_seed(seed)
    est = ref.Model.VMIEstimator("separate", "constant", "infonce", d, hidden, embed, layers, "relu", 0, 1)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, d, generator=g)
    y = 0.6 * x + 0.8 * torch.randn(B, d, generator=g)
    x.requires_grad_(True)
    y.requires_grad_(True)
    params = list(est.parameters())

    def step():
        x.grad = y.grad = None
        for p in params:
            p.grad = None
        mi, loss = est(x, y)
        loss.backward()
        return float(mi)
    return step


def time_fn(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


# --------------------------------------------------------------------------
# BASELINE configs[0]: one stage-1 + one stage-2 step of the full reference Model on CPU
# --------------------------------------------------------------------------


def cfg1_opts():
    """README.md:16-26 (the one documented launch command), as the namespace Model.__init__ reads."""
    return types.SimpleNamespace(
        d_common=128, encoders="gru", features_compose_t="mean", features_compose_k="mean", num_class=1,
        activate="gelu", time_len=100, d_hiddens=[[50, 3, 128], [10, 3, 128]], d_outs=[[50, 3, 128], [10, 3, 128]],
        dropout_mlp=[0.0, 0.0, 0.0], dropout=[0.1, 0.1, 0.1, 0.1], bias=True, ln_first=False, res_project=[True, True],
        critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
        cmi_last_acticate="sigmoid", loss_mi_coefficient1=[1.0] * 11, loss_mi_coefficient2=[0.01] * 8,
        gradient_clip=1.5, learning_rate=4e-3, bert_lr_rate=0.01, mi_lr_rate=1.0, weight_decay=0.0)


def cfg1_batch(bs=128, time_len=100, d_a=5, d_v=20, n_pool=1284, seed=0):
    """MOSI-shaped synthetic batch and pools (SURVEY 8(d) cfg 1; dims Config.py:75)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 20000, (bs, time_len), generator=g)
    mask = torch.ones(bs, time_len, dtype=torch.long)
    types_ = torch.zeros(bs, time_len, dtype=torch.long)
    a = torch.randn(bs, time_len, d_a, generator=g)
    v = torch.randn(bs, time_len, d_v, generator=g)
    labels = torch.randn(bs, generator=g).clamp(-3, 3)
    pools = [torch.randn(n_pool, 1, generator=g)] + [torch.randn(n_pool, 128, generator=g) for _ in range(4)]
    return (ids, types_, mask, a, v), labels, pools


def cfg1_step_fn(ref, bs=128, n_pool=1284, seed=0):
    """One stage-1 step then one stage-2 step, the loop bodies of Solver.train (Solver.py:204-216, 220-236) with
    compute_loss (Solver.py:317-342, MAE) and compute_custumized_loss (Customization.py:91-115) restated: Solver
    itself cannot be imported (it needs the unshipped DataLoaderLocal)."""
    import numpy as np
    import torch
    opt = cfg1_opts()
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = ref.Model.Model(opt, 768, 5, 20)
    model.train()
    inputs, labels, pools = cfg1_batch(bs, opt.time_len, 5, 20, n_pool, seed)
    bert, vmi, main = [], [], []
    for name, p in model.named_parameters():                      # Solver.py:124-133
        (bert if "bert" in name else vmi if ("vmi" in name or "vcmi" in name) else main).append(p)
    opt_main = torch.optim.Adam([{"params": bert, "lr": opt.learning_rate * opt.bert_lr_rate},
                                 {"params": main, "lr": opt.learning_rate}], lr=opt.learning_rate)
    opt_vmi = torch.optim.Adam([{"params": vmi, "lr": opt.learning_rate * opt.mi_lr_rate}], lr=opt.learning_rate)
    every = [p for p in model.parameters() if p.requires_grad]

    def stage(which):
        out = model(*inputs, return_features=True)
        task = torch.nn.functional.l1_loss(out[0].reshape(-1), labels.reshape(-1))
        fn = model.compute_vmi_loss_stage1 if which == 1 else model.compute_vmi_loss_stage2
        mis, mi_losses = fn(out[0].reshape(-1, 1), labels.reshape(-1, 1), out[1], out[2], out[3], out[4], *pools)
        coef = opt.loss_mi_coefficient1 if which == 1 else opt.loss_mi_coefficient2
        loss = 0.0 if which == 1 else task
        for c, l in zip(coef, mi_losses):
            loss = loss + l * c
        o = opt_vmi if which == 1 else opt_main
        o.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_value_(every, opt.gradient_clip)
        o.step()
        return float(loss)

    def step():
        return stage(1), stage(2)
    return step
