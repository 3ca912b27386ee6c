"""The reference ``Model`` (Model.py:227-519) hosted on the B200 kernels.

Only the hot path is rebuilt: the feature heads either side of the fusion
encoder (temporal means + pad + stack, Model.py:466-475; modality/time
reduction, Model.py:489-507: ``csrc/features.cu``), the CubeMLP encoder
(``mlp_process.MLPEncoder``) and the eleven MI / CMI estimators with the two
stage functions (``model.MIStageMixin``).  The BERT text encoder, the GRU /
LSTM / Conv1d encoders, LayerNorm/dropout and the task classifier are the
reference's stock torch modules under the reference's attribute names, so a
reference ``state_dict`` loads with ``strict=True`` and
``Solver.get_optimizer``'s substring grouping (Solver.py:124-133) applies.

``forward`` keeps the reference signature and return value
``[output, F_F, T_F, A_F, V_F]`` (``[output]`` without ``return_features``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import _lib as L
from .mlp_process import MLPEncoder
from .model import MIStageMixin, VCMIEstimator, VMIEstimator


def get_output_dim(features_compose_t, features_compose_k, d_out, t_out, k_out):
    """Model.py:12-27."""
    if features_compose_t not in ('mean', 'sum', 'cat') or features_compose_k not in ('mean', 'sum', 'cat'):
        raise NotImplementedError
    dim = d_out * (t_out if features_compose_t == 'cat' else 1)
    return dim * (k_out if features_compose_k == 'cat' else 1)


class _FeatureStack(torch.autograd.Function):
    """(t, a, v) -> (x [bs, time_len, 3, D], T_F, A_F, V_F): Model.py:466-475 in one pass."""

    @staticmethod
    def forward(ctx, t, a, v, time_len):
        t, a, v = L.f32(t), L.f32(a), L.f32(v)
        bs, d = t.shape[0], t.shape[2]
        lens = (t.shape[1], a.shape[1], v.shape[1])
        x = torch.empty(bs, time_len, 3, d, dtype=torch.float32, device=t.device)
        means = torch.empty(3, bs, d, dtype=torch.float32, device=t.device)
        L.check(L.lib.mimrl_feature_stack_fwd(L.ptr(t), L.ptr(a), L.ptr(v), bs, *lens, time_len, d, L.ptr(x),
                                              L.ptr(means[0]), L.ptr(means[1]), L.ptr(means[2]), L.stream()))
        ctx.cfg = (bs, lens, time_len, d)
        return x, means[0], means[1], means[2]

    @staticmethod
    def backward(ctx, g_x, g_t, g_a, g_v):
        bs, lens, time_len, d = ctx.cfg
        dev = next(g for g in (g_x, g_t, g_a, g_v) if g is not None).device
        gs = [L.f32(g) if g is not None else None for g in (g_x, g_t, g_a, g_v)]
        outs = [torch.empty(bs, n, d, dtype=torch.float32, device=dev) if need else None
                for n, need in zip(lens, ctx.needs_input_grad[:3])]
        L.check(L.lib.mimrl_feature_stack_bwd(L.ptr(gs[0]), L.ptr(gs[1]), L.ptr(gs[2]), L.ptr(gs[3]), bs, *lens, time_len,
                                              d, L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]), L.stream()))
        return outs[0], outs[1], outs[2], None


class _FeatureReduce(torch.autograd.Function):
    """x [bs, rows, D] -> scale * sum over rows (the mean/sum compositions of Model.py:489-504)."""

    @staticmethod
    def forward(ctx, x, scale):
        x = L.f32(x)
        bs, rows, d = x.shape
        out = torch.empty(bs, d, dtype=torch.float32, device=x.device)
        L.check(L.lib.mimrl_feature_reduce_fwd(L.ptr(x), bs, rows, d, float(scale), L.ptr(out), L.stream()))
        ctx.cfg = (bs, rows, d, float(scale))
        return out

    @staticmethod
    def backward(ctx, g):
        bs, rows, d, scale = ctx.cfg
        g = L.f32(g)
        gx = torch.empty(bs, rows, d, dtype=torch.float32, device=g.device)
        L.check(L.lib.mimrl_feature_reduce_bwd(L.ptr(g), bs, rows, d, scale, L.ptr(gx), L.stream()))
        return gx, None


def feature_stack(t, a, v, time_len):
    return _FeatureStack.apply(t, a, v, time_len)


def compose_features(x, compose_k, compose_t):
    """Model.py:489-504 on the encoder output x [bs, L', K', D]."""
    bs, lo, ko, d = x.shape
    if compose_k in ('mean', 'sum') and compose_t in ('mean', 'sum'):
        scale = (1.0 / ko if compose_k == 'mean' else 1.0) * (1.0 / lo if compose_t == 'mean' else 1.0)
        return _FeatureReduce.apply(x.reshape(bs, lo * ko, d), scale)
    # a 'cat' composition is a pure re-layout; the remaining reduction (if any) is one row kernel
    if compose_k == 'cat':
        fused = x.reshape(bs, lo, ko * d)
    else:
        fused = _FeatureReduce.apply(x.reshape(bs * lo, ko, d), 1.0 / ko if compose_k == 'mean' else 1.0).reshape(bs, lo, d)
    if compose_t == 'cat':
        return fused.reshape(bs, -1)
    return _FeatureReduce.apply(fused, 1.0 / lo if compose_t == 'mean' else 1.0)


def _length_mask(seq):
    """Utils.get_mask_from_sequence(seq, dim=-1) inverted: 1 where the frame is not all-zero (Model.py:425-426)."""
    return (seq.abs().sum(dim=-1) != 0).int()


class Model(nn.Module, MIStageMixin):
    def __init__(self, opt, d_t, d_a, d_v):
        super().__init__()
        from transformers import BertConfig, BertModel
        d_common = opt.d_common
        self.time_len = opt.time_len
        self.opt = opt
        self.d_t, self.d_a, self.d_v, self.d_common = d_t, d_a, d_v, d_common
        self.encoders = opt.encoders
        assert self.encoders in ['lstm', 'gru', 'conv']
        self.features_compose_t, self.features_compose_k = opt.features_compose_t, opt.features_compose_k
        assert self.features_compose_t in ['mean', 'cat', 'sum']
        assert self.features_compose_k in ['mean', 'cat', 'sum']

        # out of scope (stock modules, reference names): text / audio / video encoders
        bertconfig = BertConfig.from_pretrained('bert-base-uncased', output_hidden_states=True, local_files_only=True)
        self.bertmodel = BertModel.from_pretrained('bert-base-uncased', config=bertconfig, local_files_only=True)
        if self.encoders == 'conv':
            self.conv_a = nn.Conv1d(d_a, d_common, kernel_size=3, stride=1, padding=1)
            self.conv_v = nn.Conv1d(d_v, d_common, kernel_size=3, stride=1, padding=1)
        elif self.encoders == 'lstm':
            self.rnn_v = nn.LSTM(d_v, d_common, 1, bidirectional=True, batch_first=True)
            self.rnn_a = nn.LSTM(d_a, d_common, 1, bidirectional=True, batch_first=True)
        else:
            self.rnn_v = nn.GRU(d_v, d_common, 2, bidirectional=True, batch_first=True)
            self.rnn_a = nn.GRU(d_a, d_common, 2, bidirectional=True, batch_first=True)
        self.ln_a, self.ln_v = nn.LayerNorm(d_common, eps=1e-6), nn.LayerNorm(d_common, eps=1e-6)
        self.dropout_t, self.dropout_a, self.dropout_v = (nn.Dropout(opt.dropout[i]) for i in range(3))
        self.W_t = nn.Linear(d_t, d_common, bias=False)

        # hot path: CubeMLP fusion encoder
        self.mlp_encoder = MLPEncoder(activate=opt.activate, d_in=[opt.time_len, 3, d_common], d_hiddens=opt.d_hiddens,
                                      d_outs=opt.d_outs, dropouts=opt.dropout_mlp, bias=opt.bias, ln_first=opt.ln_first,
                                      res_project=opt.res_project)
        classify_dim = get_output_dim(self.features_compose_t, self.features_compose_k, opt.d_outs[-1][2],
                                      opt.d_outs[-1][0], opt.d_outs[-1][1])
        if classify_dim <= 128:
            self.classifier = nn.Sequential(nn.Linear(classify_dim, opt.num_class))
        else:
            self.classifier = nn.Sequential(nn.Linear(classify_dim, 128), nn.ReLU(), nn.Dropout(opt.dropout[3]),
                                            nn.Linear(128, opt.num_class))

        # hot path: the MI / CMI estimators (Model.py:283-303; sizes hard-coded there)
        hidden_dim, embed_dim, layers, activation, mu, rho = 256, 128, 2, 'relu', 0, 1
        self.critic_type, self.baseline_type, self.bound_type = opt.critic_type, opt.baseline_type, opt.bound_type
        self.k_neighbor, self.radius, self.last_acticate = opt.k_neighbor, opt.radius, opt.cmi_last_acticate
        for n in ("f_t", "f_a", "f_v", "t_a", "t_v"):
            setattr(self, "vmi_estimator_" + n,
                    VMIEstimator(opt.critic_type, opt.baseline_type, opt.bound_type, d_common, hidden_dim, embed_dim, layers,
                                 activation, mu, rho))
        for n in ("ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"):
            setattr(self, "vcmi_estimator_" + n,
                    VCMIEstimator(embed_dim, hidden_dim, layers, activation, self.k_neighbor, self.radius,
                                  self.last_acticate))

    def set_rowblock(self, rb):
        """Shard the global batch of every MI estimator by row blocks (rowblock.RowBlock) over the ranks."""
        for n in ("f_t", "f_a", "f_v", "t_a", "t_v"):
            getattr(self, "vmi_estimator_" + n).rowblock = rb

    def encode(self, bert_sentences, bert_sentence_types, bert_sentence_att_mask, a, v):
        """Model.py:391-461: stock encoders -> t, a, v [bs, L_m, d_common]."""
        t = self.bertmodel(input_ids=bert_sentences, attention_mask=bert_sentence_att_mask,
                           token_type_ids=bert_sentence_types)[0]
        t = self.W_t(t)
        l_a, l_v = a.shape[1], v.shape[1]
        if self.encoders == 'conv':
            a = self.conv_a(a.transpose(1, 2)).transpose(1, 2)
            v = self.conv_v(v.transpose(1, 2)).transpose(1, 2)
        else:
            lengths_a, lengths_v = _length_mask(a).sum(dim=1).cpu(), _length_mask(v).sum(dim=1).cpu()
            lengths_a[lengths_a == 0] = 1
            lengths_v[lengths_v == 0] = 1
            self.rnn_a.flatten_parameters()
            self.rnn_v.flatten_parameters()
            pa, _ = self.rnn_a(pack_padded_sequence(a, lengths_a, batch_first=True, enforce_sorted=False))
            pv, _ = self.rnn_v(pack_padded_sequence(v, lengths_v, batch_first=True, enforce_sorted=False))
            a, _ = pad_packed_sequence(pa, batch_first=True, total_length=l_a)
            v, _ = pad_packed_sequence(pv, batch_first=True, total_length=l_v)
            a = a[..., :self.d_common] + a[..., self.d_common:]                  # forward + backward halves
            v = v[..., :self.d_common] + v[..., self.d_common:]
        a, v = F.relu(self.ln_a(a)), F.relu(self.ln_v(v))
        return self.dropout_t(t), self.dropout_a(a), self.dropout_v(v)

    def fuse(self, t, a, v):
        """Model.py:466-517 on the encoder outputs: feature heads + CubeMLP + composition + classifier."""
        x, T_F, A_F, V_F = feature_stack(t, a, v, self.time_len)
        x = self.mlp_encoder(x, mask=None)
        fused = compose_features(x, self.features_compose_k, self.features_compose_t)
        F_F = fused                                                               # features.unsqueeze(1).mean(1)
        return self.classifier(fused), F_F, T_F, A_F, V_F

    def forward(self, bert_sentences, bert_sentence_types, bert_sentence_att_mask, a, v, return_features=False,
                debug=False):
        t, a, v = self.encode(bert_sentences, bert_sentence_types, bert_sentence_att_mask, a, v)
        output, F_F, T_F, A_F, V_F = self.fuse(t, a, v)
        return [output, F_F, T_F, A_F, V_F] if return_features else [output]


# --------------------------------------------------------------------------
# optimisers and checkpoints (Solver.py:119-151, 60-65, 526-531)
# --------------------------------------------------------------------------


def build_optimizers(model, opt):
    """Solver.get_optimizer's parameter grouping: names containing 'bert' / 'vmi' / 'vcmi' / the rest, Adam or SGD."""
    bert, vmi, main = [], [], []
    for name, p in model.named_parameters():
        if p.requires_grad:
            (bert if 'bert' in name else vmi if ('vmi' in name or 'vcmi' in name) else main).append(p)
    lr = float(opt.learning_rate)
    bert_rate = getattr(opt, "bert_lr_rate", -1)
    main_groups = [{'params': bert, 'lr': lr if bert_rate <= 0 else lr * bert_rate}, {'params': main, 'lr': lr}]
    vmi_groups = [{'params': vmi, 'lr': lr * getattr(opt, "mi_lr_rate", 1.0)}]
    wd = getattr(opt, "weight_decay", 0.0)
    optm = getattr(opt, "optm", "Adam")
    if optm == "Adam":
        return torch.optim.Adam(main_groups, lr=lr, weight_decay=wd), torch.optim.Adam(vmi_groups, lr=lr, weight_decay=wd)
    if optm == "SGD":
        return (torch.optim.SGD(main_groups, lr=lr, weight_decay=wd, momentum=0.9),
                torch.optim.SGD(vmi_groups, lr=lr, weight_decay=wd, momentum=0.9))
    raise NotImplementedError


def checkpoint_state(epoch, model, optimizer_main, optimizer_vmi):
    """The dict the reference keeps for its best-valid / best-test states (Solver.py:60-65)."""
    module = model.module if hasattr(model, "module") else model
    return {"epoch": epoch, "model": module.state_dict(), "optim_main": optimizer_main.state_dict(),
            "optim_vmi": optimizer_vmi.state_dict()}


def save_checkpoint(path, epoch, model, optimizer_main, optimizer_vmi):
    torch.save(checkpoint_state(epoch, model, optimizer_main, optimizer_vmi), path)      # Solver.py:530-531


def load_checkpoint(path_or_state, model, optimizer_main=None, optimizer_vmi=None, map_location=None, strict=True):
    """Load a checkpoint written by the reference (or by save_checkpoint).  Keys saved through ``nn.DataParallel``
    carry a ``module.`` prefix (Solver.py:33-35 wraps the model before ``state_dict()`` is taken); it is stripped."""
    state = path_or_state
    if not isinstance(state, dict):
        state = torch.load(path_or_state, map_location=map_location, weights_only=False)
    sd = state["model"]
    if sd and all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    module = model.module if hasattr(model, "module") else model
    module.load_state_dict(sd, strict=strict)
    if optimizer_main is not None:
        optimizer_main.load_state_dict(state["optim_main"])
    if optimizer_vmi is not None:
        optimizer_vmi.load_state_dict(state["optim_vmi"])
    return state["epoch"]


# --------------------------------------------------------------------------
# BASELINE configs[4]: one stage-1 + one stage-2 step of the full model, MOSEI-shaped synthetic data
# --------------------------------------------------------------------------


def mosei_opts(**over):
    from types import SimpleNamespace
    o = dict(d_common=128, encoders="gru", features_compose_t="mean", features_compose_k="mean", num_class=1,
             activate="gelu", time_len=100, d_hiddens=[[50, 3, 128], [10, 3, 128]], d_outs=[[50, 3, 128], [10, 3, 128]],
             dropout_mlp=[0.0, 0.0, 0.0], dropout=[0.1, 0.1, 0.1, 0.1], bias=True, ln_first=False, res_project=[True, True],
             critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
             cmi_last_acticate="sigmoid", loss_mi_coefficient1=[1.0] * 11, loss_mi_coefficient2=[0.01] * 8,
             gradient_clip=1.5, learning_rate=4e-3, bert_lr_rate=0.01, mi_lr_rate=1.0, weight_decay=0.0, optm="Adam")
    o.update(over)
    return SimpleNamespace(**o)


def random_init_bert():
    """No network, no cached weights: ``from_pretrained`` builds a random-init bert-base (same architecture and
    FLOPs; the same shim the CPU reference arm uses).  Returns a restore() callable."""
    import transformers
    classes = (transformers.BertConfig, transformers.BertModel)
    saved = [cls.__dict__.get("from_pretrained") for cls in classes]

    def _cfg(*a, **k):
        return transformers.BertConfig(output_hidden_states=bool(k.get("output_hidden_states", False)))

    def _model(*a, config=None, **k):
        return transformers.BertModel(config if config is not None else transformers.BertConfig())
    transformers.BertConfig.from_pretrained = staticmethod(_cfg)
    transformers.BertModel.from_pretrained = staticmethod(_model)

    def restore():
        for cls, old in zip(classes, saved):                  # inherited classmethod: remove the override again
            if old is None:
                delattr(cls, "from_pretrained")
            else:
                setattr(cls, "from_pretrained", old)
    return restore


class TrainStep:
    """One stage-1 step then one stage-2 step: the loop bodies of Solver.train (Solver.py:204-216 and 220-236) with
    compute_loss / compute_custumized_loss (Solver.py:317-342, Customization.py:91-115) for the MAE task loss.
    Data parallel over ranks: every rank encodes its own bs rows; the MI estimators see the GLOBAL batch through the
    row-block all-gathers (the reference's DataParallel semantics); the CMI classifiers, CubeMLP, heads and
    encoders are plain data parallel.  Parameter gradients are summed over ranks with one flat all-reduce."""

    def __init__(self, model, opt, rb=None):
        from . import rowblock as RB
        self.model, self.opt = model, opt
        self.rb = rb if rb is not None else None
        self.opt_main, self.opt_vmi = build_optimizers(model, opt)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.RB = RB
        if rb is not None and rb.sharded:
            model.set_rowblock(rb)

    def _stage(self, which, inputs, labels, pools):
        opt, model = self.opt, self.model
        out = model(*inputs, return_features=True)
        task = F.l1_loss(out[0].reshape(-1), labels.reshape(-1))
        fn = model.compute_vmi_loss_stage1 if which == 1 else model.compute_vmi_loss_stage2
        mis, mi_losses = fn(out[0].reshape(-1, 1), labels.reshape(-1, 1), out[1], out[2], out[3], out[4], *pools)
        coef = opt.loss_mi_coefficient1 if which == 1 else opt.loss_mi_coefficient2
        loss = task * 0.0 if which == 1 else task
        for c, l in zip(coef, mi_losses):
            loss = loss + l * c
        o = self.opt_vmi if which == 1 else self.opt_main
        o.zero_grad(set_to_none=True)
        loss.backward()
        if self.rb is not None and self.rb.sharded:
            self.RB.all_reduce_param_grads(self.params, self.rb)
        if opt.gradient_clip > 0:
            torch.nn.utils.clip_grad_value_(self.params, opt.gradient_clip)
        o.step()
        return loss.detach(), [m.detach() for m in mis]

    def __call__(self, inputs, labels, pools):
        l1, _ = self._stage(1, inputs, labels, pools)
        l2, mis = self._stage(2, inputs, labels, pools)
        return l1, l2, mis


def synthetic_mosei_batch(bs, dev, seed=0, time_len=100, d_a=74, d_v=35, n_pool=16326):
    """MOSEI-shaped synthetic inputs (Config.py:76: dims [768, 74, 35]) and feature pools (SURVEY 8(d) cfg 5)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 20000, (bs, time_len), generator=g)
    mask = torch.ones(bs, time_len, dtype=torch.long)
    types_ = torch.zeros(bs, time_len, dtype=torch.long)
    a = torch.randn(bs, time_len, d_a, generator=g)
    v = torch.randn(bs, time_len, d_v, generator=g)
    labels = torch.randn(bs, generator=g).clamp(-3, 3)
    gp = torch.Generator().manual_seed(12345)                       # pools identical on every rank
    pools = [torch.randn(n_pool, 1, generator=gp)] + [torch.randn(n_pool, 128, generator=gp) for _ in range(4)]
    to = lambda t_: t_.to(dev)
    return tuple(to(t_) for t_ in (ids, types_, mask, a, v)), to(labels), [to(p) for p in pools]


def bench_config5(world, rank, dev, timed, bs=1024):
    """Full MIMRL training step (random-init bert-base + GRU encoders in stock torch, CubeMLP + all MI/CMI losses on the
    B200 kernels), bs per GPU, global-batch MI: steps/s at this N, with the hot-path share of the step."""
    import numpy as np
    from . import rowblock as RB
    restore = random_init_bert()
    try:
        torch.manual_seed(0)
        opt = mosei_opts()
        model = Model(opt, 768, 74, 35).to(dev)
    finally:
        restore()
    model.train()
    rb = RB.RowBlock(rank, world, tuple([bs] * world), None) if world > 1 else None
    step = TrainStep(model, opt, rb)
    inputs, labels, pools = synthetic_mosei_batch(bs, dev, seed=100 + rank)
    np.random.seed(0)
    ms = timed(lambda: step(inputs, labels, pools), warm=1, reps=2)
    # share of the step outside BERT/GRU: the same two stages on precomputed encoder outputs
    with torch.no_grad():
        t, a, v = model.encode(*inputs)
    t, a, v = (z.detach().requires_grad_(True) for z in (t, a, v))

    def hot():
        for which in (1, 2):
            out = model.fuse(t, a, v)
            fn = model.compute_vmi_loss_stage1 if which == 1 else model.compute_vmi_loss_stage2
            _, mi_losses = fn(out[0].reshape(-1, 1), labels.reshape(-1, 1), out[1], out[2], out[3], out[4], *pools)
            loss = F.l1_loss(out[0].reshape(-1), labels.reshape(-1))
            for l in mi_losses:
                loss = loss + l
            model.zero_grad(set_to_none=True)
            loss.backward()
    ms_hot = timed(hot, warm=1, reps=2)
    # the same hot path as ONE CUDA graph with the estimators of a stage as parallel branches (single GPU: the sharded
    # estimators all-gather through NCCL, which this capture leaves alone)
    ms_graph = None
    if world == 1:
        try:
            from .graphs import GraphedCallable, HostIdSource
            model.parallel_branches = True
            graphed = GraphedCallable(lambda: (hot(), t.grad)[1], [], warmup=2, id_source=HostIdSource())
            ms_graph = timed(lambda: graphed(), warm=1, reps=3)
        except Exception as e:       # reported, never fatal for the bench line
            ms_graph = f"{type(e).__name__}: {e}"[:200]
        finally:
            model.parallel_branches = False
    return {"bs_per_gpu": bs, "global_batch": bs * world, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "samples_per_s": bs * world * 1e3 / ms, "hot_path_ms": ms_hot, "hot_path_share": ms_hot / ms,
            "hot_path_cuda_graph_ms": ms_graph,
            "what": "stage-1 + stage-2 step (Solver.py:204-236): random-init bert-base + 2-layer bidirectional GRUs "
                    "(stock torch, fp32, TF32 off) + feature heads + CubeMLP 50-3-128=10-3-128 + 5 VMI (separate/infonce, "
                    "global batch) + 6 k-NN samplers (pool 16326) + 6 VCMI, Adam on both optimisers, gradient all-reduce; "
                    "hot_path_ms = the same two stages forward+backward on precomputed encoder outputs (no BERT/GRU, no optimiser)"}
