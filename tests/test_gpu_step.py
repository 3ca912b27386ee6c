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
