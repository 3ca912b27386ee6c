"""The two-stage step driver (Solver.py:194-248 semantics) runs, updates the right parameter groups, keeps the
reference's RNG consumption, and builds next epoch's pool."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_stage_step_updates_the_right_groups():
    import __graft_entry__ as g
    g.build()
    from mimrl_b200.model import MIHeads
    from mimrl_b200.train_step import FeaturePool, TwoStageStep
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
                          cmi_last_acticate="hardtanh", d_common=32, mi_hidden_dim=64, mi_embed_dim=32)
    heads = MIHeads(opt).to(dev)
    enc = torch.nn.Linear(32, 4 * 32).to(dev)
    cls = torch.nn.Linear(32, 1).to(dev)

    def features(batch):
        f = enc(batch).view(-1, 4, 32)
        return cls(f[:, 0]), f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), f[:, 3].contiguous()
    main = list(enc.parameters()) + list(cls.parameters())
    step = TwoStageStep(heads, features, torch.nn.L1Loss(), torch.optim.SGD(main, 1e-2), torch.optim.SGD(heads.parameters(), 1e-2),
                        clip_params=main + list(heads.parameters()))
    pool = FeaturePool()
    batch, labels = torch.randn(64, 32, device=dev), torch.randn(64, device=dev)
    # epoch 0: empty pool -> stage 1 is a no-op, stage 2 is the task loss only (Customization.py:97-98,105-106)
    loss1, mis1 = step.stage1(batch, labels, pool)
    assert float(loss1) == 0.0 and mis1 == []
    step.stage2(batch, labels, pool)
    pool.roll()
    assert len(pool) == 64 and pool.C.shape == (64, 1) and pool.T.shape == (64, 32)
    pool.C, pool.F, pool.T, pool.A, pool.V = (torch.cat([t] * 4) + 0.01 * torch.randn(256, t.shape[1], device=dev)
                                              for t in (pool.C, pool.F, pool.T, pool.A, pool.V))
    h0 = [p.detach().clone() for p in heads.parameters()]
    m0 = [p.detach().clone() for p in main]
    np.random.seed(3)
    loss, mis = step.stage1(batch, labels, pool)
    state = np.random.get_state()[1].copy()
    np.random.seed(3)
    for _ in range(6):
        np.random.choice(range(256), size=32, replace=False)      # six sampler draws per stage call
    assert np.array_equal(state, np.random.get_state()[1])
    assert len(mis) == 11 and torch.isfinite(loss)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(h0, heads.parameters()))      # estimators moved
    assert all(torch.equal(a, b.detach()) for a, b in zip(m0, main))                        # main model did not
    h1 = [p.detach().clone() for p in heads.parameters()]
    loss, mis = step.stage2(batch, labels, pool)
    assert len(mis) == 8 and torch.isfinite(loss)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(m0, main))                    # main model moved
    assert all(torch.equal(a, b.detach()) for a, b in zip(h1, heads.parameters()))          # estimators did not


def test_graphed_two_stage_step_matches_eager():
    """One CUDA graph per stage (graphs.py) reproduces the eager step: same numpy RNG stream for the k-NN query ids,
    same losses, same parameters after three steps."""
    import copy
    from types import SimpleNamespace
    from mimrl_b200.model import MIHeads
    from mimrl_b200.train_step import FeaturePool, GraphedTwoStageStep, TwoStageStep
    dev = torch.device("cuda:0")
    opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2,
                          radius=1.0, cmi_last_acticate="hardtanh", d_common=128)
    torch.manual_seed(0)
    heads0 = MIHeads(opt).to(dev)
    enc0 = torch.nn.Linear(128, 4 * 128).to(dev)
    cls0 = torch.nn.Linear(128, 1).to(dev)
    g = torch.Generator(device="cuda").manual_seed(0)
    N, bs = 700, 96
    pool_t = [torch.randn(N, 1, device=dev, generator=g).clamp(-3, 3)] + [torch.randn(N, 128, device=dev, generator=g)
                                                                           for _ in range(4)]
    batches = [(torch.randn(bs, 128, device=dev, generator=g), torch.randn(bs, device=dev, generator=g).clamp(-3, 3))
               for _ in range(3)]

    def make():
        heads, enc, cls = copy.deepcopy(heads0), copy.deepcopy(enc0), copy.deepcopy(cls0)

        def features(batch):
            f = enc(batch).view(-1, 4, 128)
            return cls(f[:, 0]), f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), f[:, 3].contiguous()
        main = list(enc.parameters()) + list(cls.parameters())
        step = TwoStageStep(heads, features, torch.nn.L1Loss(), torch.optim.Adam(main, 1e-3, capturable=True),
                            torch.optim.Adam(heads.parameters(), 1e-3, capturable=True),
                            clip_params=main + list(heads.parameters()))
        pool = FeaturePool(C=pool_t[0], F=pool_t[1], T=pool_t[2], A=pool_t[3], V=pool_t[4])
        names = [n for n, _ in heads.named_parameters()] + ["main"] * len(main)
        return step, pool, list(zip(names, list(heads.parameters()) + main))

    step_e, pool_e, params_e = make()
    np.random.seed(3)
    eager = []
    for b, l in batches:
        l1, _ = step_e.stage1(b, l, pool_e)
        l2, mis = step_e.stage2(b, l, pool_e)
        eager.append((float(l1), float(l2), [float(m) for m in mis]))
    step_g, pool_g, params_g = make()
    graphed = GraphedTwoStageStep(step_g, batches[0][0], batches[0][1], pool_g)
    np.random.seed(3)
    got = []
    for b, l in batches:
        l1, _ = graphed.stage1(b, l)
        l2, mis = graphed.stage2(b, l)
        got.append((float(l1), float(l2), [float(m) for m in mis]))
    for (a1, a2, am), (b1, b2, bm) in zip(eager, got):
        assert abs(a1 - b1) <= 1e-4 * max(1.0, abs(a1)) and abs(a2 - b2) <= 1e-4 * max(1.0, abs(a2))
        assert np.allclose(am, bm, rtol=1e-4, atol=1e-5)
    for (name, pe), (_, pg) in zip(params_e, params_g):
        # InfoNCE is invariant to a per-row shift of the scores, so the exact gradient of the x-side output bias
        # (S_ij + y_i . b) is zero: Adam normalises pure rounding noise there, and the graph (three-sweep forward) and
        # eager (fused forward) paths round differently
        if name.endswith("critic_model.MLP_g.6.bias"):
            continue
        assert torch.allclose(pe, pg, rtol=1e-4, atol=1e-6), name
    assert len(pool_g._next["C"]) == 3 and pool_g._next["F"][0].shape == (bs, 128)


def test_stage_branches_on_side_streams_match_the_sequential_run():
    """MIStageMixin.parallel_branches: the eleven estimators of a stage enqueued on side streams give the values,
    feature gradients, estimator gradients and numpy RNG state of the sequential run."""
    import __graft_entry__ as g
    g.build()
    from mimrl_b200.model import MIHeads
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2, radius=1.0,
                          cmi_last_acticate="hardtanh", d_common=128)
    heads = MIHeads(opt).to(dev)
    gen = torch.Generator(device="cuda").manual_seed(1)
    bs, N = 128, 1284
    feats = [torch.randn(bs, 128, device=dev, generator=gen) for _ in range(4)]
    labels = torch.randn(bs, device=dev, generator=gen)
    pools = [torch.randn(N, 1, device=dev, generator=gen)] + [torch.randn(N, 128, device=dev, generator=gen) for _ in range(4)]

    def run(parallel, stage):
        heads.parallel_branches = parallel
        for p in heads.parameters():
            p.grad = None
        fs = [f.clone().requires_grad_(True) for f in feats]
        np.random.seed(11)
        fn = heads.compute_vmi_loss_stage1 if stage == 1 else heads.compute_vmi_loss_stage2
        mis, losses = fn(None, labels, *fs, *pools)
        torch.stack([l.reshape(()) for l in losses]).sum().backward()
        torch.cuda.synchronize()
        return (torch.stack([m.detach().reshape(()) for m in mis]), [f.grad.clone() for f in fs],
                [p.grad.clone() if p.grad is not None else None for p in heads.parameters()], np.random.get_state()[1].copy())

    for stage in (1, 2):
        m0, g0, p0, r0 = run(False, stage)
        for _ in range(3):                       # a race would not show every time
            m1, g1, p1, r1 = run(True, stage)
            assert np.array_equal(r0, r1)
            assert torch.allclose(m0, m1, rtol=1e-6, atol=1e-7)
            for a, b in zip(g0, g1):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)
            for a, b in zip(p0, p1):
                assert (a is None) == (b is None)
                if a is not None:
                    assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)
    heads.parallel_branches = False
