"""Host-side query draw of the k-NN sampler (csrc/host_rng.cu, mimrl_legacy_permutation_head) against numpy itself:
Model.py:81 draws np.random.choice(range(N), size=m, replace=False) from the GLOBAL legacy generator, i.e.
permutation(N)[:m]; the drop-in must return the same ids AND leave the generator in the same state."""
import numpy as np
import pytest


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import __graft_entry__ as g
    g.build()


@pytest.mark.parametrize("N,m,seed", [(1, 1, 0), (2, 1, 1), (5, 5, 2), (63, 3, 8), (64, 64, 9), (65, 1, 10), (129, 0, 11),
                                      (1284, 64, 3), (65536, 100, 4), (65537, 1000, 5), (1 << 20, 4096, 6),
                                      (1000003, 7, 7)])
def test_draw_and_state_match_numpy(N, m, seed):
    from mimrl_b200.model import legacy_permutation_head
    np.random.seed(seed)
    np.random.rand(seed % 5)                           # start somewhere inside a block of the generator
    want = np.random.permutation(N)[:m]
    after = np.random.randint(0, 1 << 30, 700)         # crosses a block boundary
    np.random.seed(seed)
    np.random.rand(seed % 5)
    got = legacy_permutation_head(N, m, _min_n=0)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    assert np.array_equal(np.random.randint(0, 1 << 30, 700), after)


def test_consecutive_draws_and_cached_gaussian():
    """Six sampler draws in a row (one stage of a step) and a pending cached normal variate survive the state round trip."""
    from mimrl_b200.model import legacy_permutation_head
    np.random.seed(5)
    np.random.randn(3)                                 # leaves has_gauss = 1
    want = [np.random.permutation(50000)[:64] for _ in range(6)] + [np.random.randn(2)]
    np.random.seed(5)
    np.random.randn(3)
    got = [legacy_permutation_head(50000, 64) for _ in range(6)] + [np.random.randn(2)]
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_other_bit_generators_fall_back_to_numpy():
    from mimrl_b200.model import legacy_permutation_head
    np.random.seed(1)
    a = legacy_permutation_head(100, 10)              # small pools: numpy's own call
    np.random.seed(1)
    assert np.array_equal(a, np.random.permutation(100)[:10])


def test_feature_pool_refits_keys_when_the_pool_changes(monkeypatch):
    """train_step.FeaturePool.keys: one fit per pool tensor, refitted after roll(), after assignment of a new tensor and
    after an in-place torch write (version counter) -- host logic, KnnPool replaced by a counter."""
    import torch
    import mimrl_b200.model as M
    from mimrl_b200.train_step import FeaturePool
    made = []

    class FakePool:
        def __init__(self, Z):
            made.append(Z)
    monkeypatch.setattr(M, "KnnPool", FakePool)
    pool = FeaturePool()
    pool.T = torch.zeros(8, 4)
    a = pool.keys("T")
    assert pool.keys("T") is a and len(made) == 1                 # cached
    pool.T.add_(1.0)                                               # in-place write through torch
    b = pool.keys("T")
    assert b is not a and len(made) == 2
    pool.T = torch.ones(8, 4)                                      # replaced
    assert pool.keys("T") is not b and len(made) == 3
    pool.append(torch.zeros(3), torch.zeros(3, 4), torch.zeros(3, 4), torch.zeros(3, 4), torch.zeros(3, 4))
    pool.roll()                                                    # epoch boundary
    c = pool.keys("T")
    assert len(made) == 4 and made[-1] is pool.T and pool.keys("T") is c
