"""Checkpoint compatibility with the reference (Solver.py:60-65, 526-531): a `{epoch, model, optim_main, optim_vmi}` dict
produced by the unmodified reference modules loads into mimrl_b200.full_model.Model and its optimisers with
strict=True, and a checkpoint saved here loads back into the reference.  CPU only (no kernel runs)."""
import os
from types import SimpleNamespace

import pytest
import torch


def _opts():
    return SimpleNamespace(
        d_common=32, encoders="gru", features_compose_t="mean", features_compose_k="mean", num_class=1, activate="gelu",
        time_len=20, d_hiddens=[[10, 3, 32], [5, 3, 32]], d_outs=[[10, 3, 32], [5, 3, 32]], dropout_mlp=[0.0, 0.0, 0.0],
        dropout=[0.0, 0.0, 0.0, 0.0], bias=True, ln_first=False, res_project=[True, True], critic_type="separate",
        baseline_type="unnormalized", bound_type="tuba", k_neighbor=2, radius=1.0, cmi_last_acticate="sigmoid",
        learning_rate=4e-3, bert_lr_rate=0.01, mi_lr_rate=1.0, weight_decay=0.0, optm="Adam")


def _reference():
    from oracle import ref_shim as R
    if R.locate() is None:
        pytest.skip("reference sources not available (make -C oracle _ref)")
    return R.import_reference(random_bert=False)


def _ref_optimizers(model, opt):
    """Solver.get_optimizer (Solver.py:119-151), Adam branch, restated (Solver cannot be imported)."""
    bert, vmi, main = [], [], []
    for name, p in model.named_parameters():
        if p.requires_grad:
            (bert if 'bert' in name else vmi if ('vmi' in name or 'vcmi' in name) else main).append(p)
    lr = float(opt.learning_rate)
    return (torch.optim.Adam([{'params': bert, 'lr': lr * opt.bert_lr_rate}, {'params': main, 'lr': lr}], lr=lr),
            torch.optim.Adam([{'params': vmi, 'lr': lr * opt.mi_lr_rate}], lr=lr))


def test_reference_checkpoint_round_trip(tiny_bert, tmp_path):
    ref = _reference()
    from mimrl_b200.full_model import Model, build_optimizers, load_checkpoint, save_checkpoint
    opt = _opts()
    torch.manual_seed(0)
    ref_model = ref.Model.Model(opt, tiny_bert["hidden_size"], 5, 7)
    om, ov = _ref_optimizers(ref_model, opt)
    for p in ref_model.parameters():                       # one step of each optimiser so that Adam state exists
        p.grad = torch.randn_like(p) * 1e-3
    om.step()
    ov.step()
    path = os.path.join(tmp_path, "best_valid.pt")
    # the reference wraps the model in DataParallel before taking state_dict() (Solver.py:33-35,60-65)
    wrapped = {"module." + k: v for k, v in ref_model.state_dict().items()}
    torch.save({"epoch": 7, "model": wrapped, "optim_main": om.state_dict(), "optim_vmi": ov.state_dict()}, path)

    torch.manual_seed(1)
    ours = Model(opt, tiny_bert["hidden_size"], 5, 7)
    assert list(ours.state_dict()) == list(ref_model.state_dict())          # same names, same order
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref_model.state_dict().values()]
    m2, v2 = build_optimizers(ours, opt)
    assert load_checkpoint(path, ours, m2, v2) == 7
    for (k, a), b in zip(ours.state_dict().items(), ref_model.state_dict().values()):
        assert torch.equal(a, b), k
    for mine, theirs in ((m2, om), (v2, ov)):
        assert [len(g["params"]) for g in mine.param_groups] == [len(g["params"]) for g in theirs.param_groups]
        assert [g["lr"] for g in mine.param_groups] == [g["lr"] for g in theirs.param_groups]
        for k, st in theirs.state_dict()["state"].items():
            assert torch.equal(mine.state_dict()["state"][k]["exp_avg"], st["exp_avg"])

    back = os.path.join(tmp_path, "ours.pt")
    save_checkpoint(back, 8, ours, m2, v2)
    state = torch.load(back, weights_only=False)
    assert sorted(state) == ["epoch", "model", "optim_main", "optim_vmi"] and state["epoch"] == 8
    ref_model.load_state_dict(state["model"], strict=True)
    om.load_state_dict(state["optim_main"])
    ov.load_state_dict(state["optim_vmi"])
