"""Pin the k-NN sampler oracle to the reference (prod_knn_sample on sklearn 1.9.0)."""
import numpy as np
import pytest

from conftest import cfg_of, load_golden
from oracle import knn_oracle as K
from oracle import params as P

KNN = load_golden("knn")


def _inputs(rec):
    c = cfg_of(rec)
    seed = int(rec["seed"])
    X = P.features(seed, c["N"], c["wx"])
    Y = P.features(seed + 1, c["N"], c["wy"])
    Z = P.features(seed + 2, c["N"], c["wz"])
    if c["dup"]:
        Z[c["N"] - c["dup"]:] = Z[: c["dup"]]
    return c, seed, X, Y, Z


@pytest.mark.parametrize("case", sorted(KNN))
def test_sampler_matches_reference(case):
    rec = KNN[case]
    c, seed, X, Y, Z = _inputs(rec)
    np.random.seed(seed)
    bx, by, bz, ids, nbr = K.prod_knn_sample(X, Y, Z, c["bs"], c["k"], 1.0)
    assert np.array_equal(ids, rec["ids"])
    assert str(rec["method"]) == K.sklearn_route(c["wz"], c["k"], c["N"] - len(ids))
    if c["dup"] == 0:
        assert np.array_equal(nbr, rec["nbr"])              # bit-exact indices, order included
        assert np.array_equal(bx, rec["bx"])
    else:
        # exact ties: sklearn guarantees the SET (lowest indices win), not the order inside a tie
        assert np.array_equal(np.sort(nbr, axis=1), np.sort(rec["nbr"], axis=1))
    assert np.array_equal(by, rec["by"]) and np.array_equal(bz, rec["bz"])
    # rec["leaf"]: widened outputs went through .repeat() in the reference (Model.py:99-104)
    # and are therefore non-leaf; un-widened ones are leaves.  All require grad.
    widths = np.array([c["wx"], c["wy"], c["wz"]])
    assert np.array_equal(rec["leaf"], widths == widths.max())


def test_rng_state_advances_like_reference():
    np.random.seed(123)
    K.draw_ids(1000, 10)
    s1 = np.random.get_state()
    np.random.seed(123)
    np.random.choice(range(1000), size=10, replace=False)
    s2 = np.random.get_state()
    assert np.array_equal(s1[1], s2[1]) and s1[2] == s2[2]


def test_too_many_neighbours_raises():
    X = P.features(1, 10, 16)
    with pytest.raises(ValueError):
        K.prod_knn_sample(X, X, X, 100, 2)                  # m=50 > N
    with pytest.raises(ValueError):
        K.prod_knn_sample(X, X, X, 8, 8)                    # k=8 > N-m=9? no: m=1, fit on 9 rows, k=8 ok
        K.prod_knn_sample(X, X, X, 9, 9)                    # m=1, 9 rows left, k=9 ok
        K.prod_knn_sample(X, X, X, 10, 10)                  # m=1, 9 rows left, k=10 -> error


@pytest.mark.parametrize("N,width,m,k,dup", [(120000, 128, 48, 16, 0), (262144, 128, 24, 2, 0), (200000, 1, 64, 4, 0),
                                             (50000, 8, 64, 3, 0), (60000, 128, 32, 4, 3000)])
def test_oracle_matches_sklearn_at_scale(N, width, m, k, dup):
    """The oracle against scikit-learn itself (the un-vendored dependency behind Model.py:82-86) at pool sizes far above
    the goldens' N <= 1284: both routes (brute for width > 15, kd_tree below), neighbour lists bit-exact, order included
    (sets inside exact ties)."""
    sk = pytest.importorskip("sklearn.neighbors")
    Z = np.random.default_rng(N + width).standard_normal((N, width), dtype=np.float32)
    if dup:
        Z[N - dup:] = Z[:dup]
    ids = np.random.RandomState(1).permutation(N)[:m]
    if dup:
        ids[:8] = np.arange(8)
        ids = np.unique(ids)
    keep = np.ones(N, bool)
    keep[ids] = False
    Z2 = Z[keep]
    nn_ = sk.NearestNeighbors(n_neighbors=k, radius=1.0, metric="euclidean").fit(Z2)      # Model.py:82-85
    want = nn_.kneighbors(Z[ids], return_distance=False)                                  # Model.py:86 (compacted ids)
    route = K.sklearn_route(width, k, N - len(ids))
    assert nn_._fit_method == route
    got, _ = K.knn(Z, Z[ids], k, (~keep).astype(np.uint8), route)
    got_comp = got - np.searchsorted(np.sort(ids), got)
    if dup:
        assert np.array_equal(np.sort(got_comp, 1), np.sort(want, 1))
    else:
        assert np.array_equal(got_comp, want)
