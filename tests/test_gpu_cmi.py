"""GPU parity of the conditional-MI path: k-NN sampler (bit-exact indices
against scikit-learn via the golden vectors and the float64 oracle), the
classifier estimator, and the two stage functions."""
import numpy as np
import pytest
import torch

from conftest import cfg_of, load_golden, rel_err
from oracle import knn_oracle as K
from oracle import params as P
from oracle import vcmi_oracle as VO

pytestmark = pytest.mark.gpu
TOL = 1e-4
KNN = load_golden("knn")
VCMI = load_golden("vcmi")
STAGE = load_golden("stage")


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import __graft_entry__ as g
    g.build()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def dev():
    return torch.device("cuda:0")


def T(a, **kw):
    return torch.tensor(a, device=dev(), **kw)


@pytest.mark.parametrize("case", sorted(KNN))
def test_sampler_matches_reference_golden(case):
    from mimrl_b200.model import knn_search, prod_knn_sample
    rec = KNN[case]
    c = cfg_of(rec)
    seed = int(rec["seed"])
    X = P.features(seed, c["N"], c["wx"])
    Y = P.features(seed + 1, c["N"], c["wy"])
    Z = P.features(seed + 2, c["N"], c["wz"])
    if c["dup"]:
        Z[c["N"] - c["dup"]:] = Z[: c["dup"]]
    np.random.seed(seed)
    bx, by, bz = prod_knn_sample(T(X), T(Y), T(Z), c["bs"], c["k"], 1.0)
    state = np.random.get_state()
    np.random.seed(seed)
    np.random.choice(range(c["N"]), size=c["bs"] // c["k"], replace=False)
    assert np.array_equal(state[1], np.random.get_state()[1])       # same RNG consumption as the reference
    _, nbr = knn_search(T(Z), T(rec["ids"]), c["k"])
    nbr = nbr.cpu().numpy()
    if c["dup"] == 0:
        assert np.array_equal(nbr, rec["nbr"])                       # bit-exact, order included
        assert np.array_equal(bx.detach().cpu().numpy(), rec["bx"])
    else:
        assert np.array_equal(np.sort(nbr, 1), np.sort(rec["nbr"], 1))
    assert np.array_equal(by.detach().cpu().numpy(), rec["by"])
    assert np.array_equal(bz.detach().cpu().numpy(), rec["bz"])
    assert bx.requires_grad and by.requires_grad and bz.requires_grad and bx.is_cuda


@pytest.mark.parametrize("N,width,m,k", [(20000, 128, 300, 2), (20000, 128, 70, 16), (50000, 1, 257, 4),
                                         (3000, 8, 64, 3), (5000, 16, 100, 32), (130, 128, 64, 2),
                                         (400000, 1, 300, 2), (300000, 128, 200, 4), (100000, 40, 129, 8)])
def test_knn_indices_vs_float64_oracle(N, width, m, k):
    from mimrl_b200.model import knn_search, sklearn_route
    Z = P.features(N + width, N, width)
    rng = np.random.default_rng(N)
    ids = rng.permutation(N)[:m]
    exc = np.zeros(N, np.uint8)
    exc[ids] = 1
    want, wdist = K.knn(Z, Z[ids], k, exc, sklearn_route(width, k, N - m))
    got, comp, dist = knn_search(T(Z), T(ids), k, return_distance=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(comp.cpu().numpy(), want - np.searchsorted(np.sort(ids), want))
    assert np.allclose(dist.cpu().numpy(), wdist, rtol=1e-12, atol=1e-12)


_POOL_1M = {}


def _pool_1m(width):
    """BASELINE configs[3] pool: 1,048,576 rows ~N(0,1) (float32), generated once per width."""
    if width not in _POOL_1M:
        _POOL_1M[width] = np.random.default_rng(2 + width).standard_normal((1 << 20, width), dtype=np.float32)
    return _POOL_1M[width]


@pytest.mark.parametrize("width,bs,k,dup", [(128, 8192, 2, 0), (128, 8192, 4, 0), (128, 8192, 8, 0), (128, 8192, 16, 0),
                                            (1, 8192, 2, 0), (1, 8192, 16, 0), (128, 8192, 4, 4096)])
def test_knn_config4_scale_vs_float64_oracle(width, bs, k, dup):
    """BASELINE configs[3]: N = 1,048,576 keys, batch 8192 -> m = bs // k queries (4096 ... 512), k in {2,4,8,16}; the
    width-1 label pool (scikit-learn's kd_tree route) and a pool with exact duplicate rows (tie rule).  The GPU search
    runs all m queries; the float64 oracle checks a 96-query subset (every neighbour list bit-exact, order included,
    distances to 1e-12) -- the queries are independent, so a subset checks the same code path at full pool size."""
    from mimrl_b200.model import knn_search, sklearn_route
    N = 1 << 20
    Z = _pool_1m(width)
    if dup:
        Z = Z.copy()
        Z[N - dup:] = Z[:dup]                       # rows N-dup.. are exact copies of rows 0..dup-1
    m = bs // k
    ids = np.random.RandomState(0).permutation(N)[:m]
    if dup:
        ids[:32] = np.arange(32)                    # 32 queries whose twin (row N-dup+i) stays in the pool at distance 0
        ids = np.unique(ids)
        m = len(ids)
    exc = np.zeros(N, np.uint8)
    exc[ids] = 1
    route = sklearn_route(width, k, N - m)
    assert route == ("brute" if width > 15 else "kd_tree")
    got, comp, dist = knn_search(T(Z), T(ids), k, return_distance=True)
    got, comp, dist = got.cpu().numpy(), comp.cpu().numpy(), dist.cpu().numpy()
    sub = np.r_[0:32, np.linspace(32, m - 1, 64).astype(int)]
    want, wdist = K.knn(Z, Z[ids[sub]], k, exc, route)
    if dup:
        # inside an exact tie scikit-learn guarantees the set (lowest indices), not the order: compare sets and distances
        assert np.array_equal(np.sort(got[sub], 1), np.sort(want, 1))
        assert np.array_equal(got[:32, 0] % (N - dup), ids[:32])          # the twin comes first, at distance 0
        assert np.all(dist[:32, 0] < 1e-5)          # GEMM-form float64 distance of an exact copy: 0 up to rounding
    else:
        assert np.array_equal(got[sub], want)
    assert np.allclose(dist[sub], wdist, rtol=1e-12, atol=1e-12)
    assert np.array_equal(comp, got - np.searchsorted(np.sort(ids), got))
    assert not np.isin(got, ids).any() and np.all(np.diff(dist, axis=1) >= 0)


@pytest.mark.parametrize("kind,N,m,k", [("labels", 200000, 512, 2), ("labels", 200000, 300, 16), ("signed_zero", 5000, 64, 8),
                                       ("colliding", 60000, 100, 12), ("few_keys", 60, 30, 7), ("few_keys", 50, 36, 6), ("constant", 30000, 200, 5)])
def test_knn_width1_sorted_route_ties(kind, N, m, k):
    """Width-1 pools (the label pools, knn_1d.cu: sort + two-sided walk) where ties decide the answer: label-like
    pools with a handful of distinct values (runs of thousands of equal distances, symmetric ties q - a / q + a),
    +0.0 / -0.0 mixed, distinct values whose float64 distances collide (large q, tiny distinct keys), fewer valid keys
    than 2k, one constant value.  Indices bit-exact and in the oracle's (distance, row) order, distances bit-equal;
    then the same queries through mimrl_knn_search_rows on a key shard with an index offset."""
    import mimrl_b200._lib as L
    from mimrl_b200.model import knn_search, sklearn_route
    rng = np.random.default_rng(N + k)
    if kind == "labels":
        Z = rng.integers(-3, 4, N).astype(np.float32) + rng.integers(0, 3, N).astype(np.float32) * 0.2
    elif kind == "signed_zero":
        Z = rng.choice(np.array([0.0, -0.0, 1.0, -1.0, 2.5], np.float32), N)
    elif kind == "colliding":
        Z = (rng.integers(1, 2000, N) * 1e-9).astype(np.float32)
        Z[rng.permutation(N)[: N // 50]] = 3.0          # queries drawn below include rows at 3.0: 3 - 1e-9 rounds in float64
    elif kind == "constant":
        Z = np.full(N, 0.7, np.float32)
    else:
        Z = rng.integers(0, 4, N).astype(np.float32)
    Z = Z.reshape(N, 1)
    ids = rng.permutation(N)[:m]
    if kind == "colliding":
        ids[:20] = np.flatnonzero(Z[:, 0] == 3.0)[:20]
        ids = np.unique(ids)
        m = len(ids)
    exc = np.zeros(N, np.uint8)
    exc[ids] = 1
    route = sklearn_route(1, k, N - m)
    if route != "kd_tree":
        pytest.skip("brute route")
    want, wdist = K.knn(Z, Z[ids], k, exc, route)
    got, comp, dist = knn_search(T(Z), T(ids), k, return_distance=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(dist.cpu().numpy(), wdist)
    assert np.array_equal(comp.cpu().numpy(), want - np.searchsorted(np.sort(ids), want))
    # key shard [lo, hi) with global row numbers, explicit queries, excluded ids as a global list
    lo, hi = N // 3, N - N // 5
    Zs = np.ascontiguousarray(Z[lo:hi])
    e2 = np.zeros(hi - lo, np.uint8)
    inside = ids[(ids >= lo) & (ids < hi)]
    e2[inside - lo] = 1
    kk = min(k, int((e2 == 0).sum()))
    want2, wd2 = K.knn(Zs, Z[ids], kk, e2, "kd_tree")
    ws = torch.empty(max(L.lib.mimrl_knn_workspace_bytes(hi - lo, m, 1, k), 16), dtype=torch.uint8, device=dev())
    nbr = torch.full((m, k), -1, dtype=torch.int64, device=dev())
    d2 = torch.full((m, k), float("inf"), dtype=torch.float64, device=dev())
    keys_t, q_t, exc_t = T(Zs), T(Z[ids]), T(np.sort(ids))          # (kept alive until the kernels have run)
    L.check(L.lib.mimrl_knn_search_rows(L.ptr(keys_t), hi - lo, 1, lo, L.ptr(q_t), m, L.ptr(exc_t), m, k, 0,
                                        L.ptr(nbr), L.ptr(d2), L.ptr(ws), ws.numel(), L.stream()))
    torch.cuda.synchronize()
    assert np.array_equal(nbr.cpu().numpy()[:, :kk], want2 + lo)
    assert np.array_equal(d2.cpu().numpy()[:, :kk], wd2)
    assert (nbr.cpu().numpy()[:, kk:] == -1).all()


def test_knn_fitted_pool_is_bit_identical():
    """KnnPool (mimrl_knn_fit + mimrl_knn_search_fitted): one fit, several searches with different query draws and k --
    indices, compacted indices and float64 distances equal to the unfitted search bit for bit; the sampler accepts the
    fitted pool in place of the tensor; narrow / small pools are wrapped without a fit."""
    from mimrl_b200.model import KnnPool, knn_search, prod_knn_sample
    N = 150000
    Z = T(P.features(5, N, 128))
    X = T(P.features(6, N, 128))
    Y = T(P.features(7, N, 1))
    pool = KnnPool(Z)
    assert pool.fitted is not None
    for seed, m, k in ((0, 300, 2), (1, 129, 16), (2, 1000, 4)):
        ids = T(np.random.default_rng(seed).permutation(N)[:m])
        a = knn_search(Z, ids, k, return_distance=True)
        b = knn_search(pool, ids, k, return_distance=True)
        for u, v in zip(a, b):
            assert torch.equal(u, v)
    np.random.seed(3)
    ra = prod_knn_sample(X, Y, Z, 512, 4, 1.0)
    np.random.seed(3)
    rb = prod_knn_sample(X, Y, pool, 512, 4, 1.0)
    for u, v in zip(ra, rb):
        assert torch.equal(u, v)
    assert KnnPool(Y).fitted is None and KnnPool(Z[:1000]).fitted is None
    np.random.seed(4)
    rc = prod_knn_sample(X, Z, Y, 512, 4, 1.0)
    np.random.seed(4)
    rd = prod_knn_sample(KnnPool(X), pool, KnnPool(Y), 512, 4, 1.0)
    for u, v in zip(rc, rd):
        assert torch.equal(u, v)


def test_knn_shard_with_fewer_than_k_keys():
    """mimrl_knn_search_rows on a key shard that holds fewer than k keys (other shards fill in): legal, unfilled slots come
    back as index -1 (the global n_neighbors <= n_samples_fit check belongs to the caller)."""
    from mimrl_b200 import _lib as L
    Z = T(P.features(3, 5, 128))
    q = T(P.features(4, 7, 128))
    k = 8
    ws = torch.empty(max(L.lib.mimrl_knn_workspace_bytes(5, 7, 128, k), 16), dtype=torch.uint8, device=dev())
    nbr = torch.full((7, k), -1, dtype=torch.int64, device=dev())
    dist = torch.full((7, k), float("inf"), dtype=torch.float64, device=dev())
    exc = torch.empty(0, dtype=torch.int64, device=dev())
    L.check(L.lib.mimrl_knn_search_rows(L.ptr(Z), 5, 128, 100, L.ptr(q), 7, L.ptr(exc), 0, k, 1, L.ptr(nbr), L.ptr(dist),
                                        L.ptr(ws), ws.numel(), L.stream()))
    nbr = nbr.cpu().numpy()
    d2 = ((q.double()[:, None, :] - Z.double()[None]) ** 2).sum(-1)
    want = 100 + torch.argsort(d2, dim=1, stable=True).cpu().numpy()
    assert np.array_equal(nbr[:, :5], want) and np.all(nbr[:, 5:] == -1)


def test_knn_duplicate_keys_lowest_index_wins():
    from mimrl_b200.model import knn_search
    Z = P.features(5, 600, 128)
    Z[300:] = Z[:300]                        # every key has an exact duplicate
    ids = np.arange(0, 40, dtype=np.int64)   # queries 0..39; their twins 300..339 stay in the pool at distance 0
    got, _ = knn_search(T(Z), T(ids), 3)
    got = got.cpu().numpy()
    exc = np.zeros(600, np.uint8)
    exc[ids] = 1
    want, _ = K.knn(Z, Z[ids], 3, exc, "brute")
    assert np.array_equal(got, want)
    assert np.array_equal(got[:, 0], ids + 300)


def test_knn_errors():
    from mimrl_b200.model import prod_knn_sample
    X = T(P.features(1, 10, 16))
    with pytest.raises(ValueError):
        prod_knn_sample(X, X, X, 100, 2, 1.0)        # m > N          (numpy's error in the reference)
    with pytest.raises(ValueError):
        prod_knn_sample(X, X, X, 10, 10, 1.0)        # k > N - m      (sklearn's error in the reference)


def vcmi_inputs(rec):
    c = cfg_of(rec)
    seed = int(rec["seed"])
    stack = P.vcmi_params(seed, c["embed"], c["hidden"])
    if c["act"] == "hardtanh":
        stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
    ins = [P.features(seed + 1, c["bs"], c["embed"], c["scale"]), P.features(seed + 2, c["bs"], c["wy"], c["scale"]),
           P.features(seed + 3, c["bs"], c["embed"], c["scale"]), P.features(seed + 4, c["nprod"], c["embed"], c["scale"]),
           P.features(seed + 5, c["nprod"], c["embed"], c["scale"]), P.features(seed + 6, c["nprod"], c["embed"], c["scale"])]
    return c, stack, ins


@pytest.mark.parametrize("case", sorted(VCMI))
def test_vcmi_matches_reference_golden(case):
    from mimrl_b200.model import VCMIEstimator
    rec = VCMI[case]
    c, stack, ins = vcmi_inputs(rec)
    est = VCMIEstimator(c["embed"], c["hidden"], 2, "relu", 2, 1.0, c["act"]).to(dev())
    est.load_state_dict({k: torch.tensor(v) for k, v in P.vcmi_state_dict(stack).items()}, strict=True)
    for tag, pick in (("gl", 1), ("gc", 0)):
        ts = [T(a, requires_grad=True) for a in ins]
        out = est(*ts)
        assert abs(float(out[0]) - float(rec["cmi"])) <= TOL * max(1.0, abs(float(rec["cmi"])))
        assert abs(float(out[1]) - float(rec["loss"])) <= TOL * max(1.0, abs(float(rec["loss"])))
        params = list(est.parameters())
        gs = torch.autograd.grad(out[pick], ts + params, allow_unused=True)
        for i, nm in enumerate(["fx", "fy", "fz", "kx", "ky", "kz"]):
            ref = rec[f"{tag}_{nm}"]
            assert np.abs(gs[i].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max() + 1e-8, (tag, nm)
        names = [n for n, _ in est.named_parameters()]
        for i, nm in enumerate(names):
            g = gs[6 + i].cpu().numpy()
            if f"{tag}p__{nm}" in rec:
                v = rec[f"{tag}p__{nm}"]
                assert np.abs(g - v).max() <= TOL * np.abs(v).max() + 2e-7, nm
            else:
                v = rec[f"{tag}s__{nm}"]
                gd = g.ravel().astype(np.float64)
                assert np.allclose([np.abs(gd).sum(), np.sqrt((gd ** 2).sum())], v[1:], rtol=2e-4, atol=2e-6), nm


@pytest.mark.parametrize("case", sorted(STAGE))
def test_stage_functions_match_reference_golden(case):
    """compute_vmi_loss_stage1/2 (Model.py:305-386): eleven estimators + six
    sampler draws in the reference's order, values and feature gradients."""
    from types import SimpleNamespace
    from mimrl_b200.model import MIHeads
    rec = STAGE[case]
    c = cfg_of(rec)
    seed = int(rec["seed"])
    d, hidden = c["d"], c["hidden"]
    opt = SimpleNamespace(critic_type=c["critic"], baseline_type=c["baseline"], bound_type=c["bound"], k_neighbor=c["k"],
                          radius=1.0, cmi_last_acticate=c["act"], d_common=d, mi_hidden_dim=hidden, mi_embed_dim=d)
    heads = MIHeads(opt).to(dev())
    for i, n in enumerate(["f_t", "f_a", "f_v", "t_a", "t_v"]):
        sd = P.vmi_state_dict(P.vmi_params(seed + 10 + i, c["critic"], c["baseline"], d, hidden, d, 2))
        getattr(heads, "vmi_estimator_" + n).load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    for i, n in enumerate(["ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"]):
        stack = P.vcmi_params(seed + 30 + i, d, hidden)
        if c["act"] == "hardtanh":
            stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
        getattr(heads, "vcmi_estimator_" + n).load_state_dict(
            {k: torch.tensor(v) for k, v in P.vcmi_state_dict(stack).items()})
    feats = {n: P.features(seed + 50 + i, c["bs"], d) for i, n in enumerate(["F", "T", "A", "V"])}
    labels = P.features(seed + 60, c["bs"], 1)[:, 0]
    pools = {n: T(P.features(seed + 70 + i, c["N"], d)) for i, n in enumerate(["F", "T", "A", "V"])}
    pool_c = T(P.features(seed + 80, c["N"], 1))
    for stage in (1, 2):
        fn = heads.compute_vmi_loss_stage1 if stage == 1 else heads.compute_vmi_loss_stage2
        ft = {n: T(v, requires_grad=True) for n, v in feats.items()}
        np.random.seed(seed + stage)
        mis, losses = fn(None, T(labels), ft["F"], ft["T"], ft["A"], ft["V"], pool_c, pools["F"], pools["T"],
                         pools["A"], pools["V"])
        got_mis = np.array([float(m) for m in mis])
        got_losses = np.array([float(m) for m in losses])
        assert np.allclose(got_mis, rec[f"s{stage}_mis"], rtol=TOL, atol=TOL), (got_mis, rec[f"s{stage}_mis"])
        assert np.allclose(got_losses, rec[f"s{stage}_losses"], rtol=TOL, atol=TOL)
        total = sum(l * (0.1 * (i + 1)) for i, l in enumerate(losses))
        gs = torch.autograd.grad(total, [ft[n] for n in "FTAV"])
        for n, g in zip("FTAV", gs):
            ref = rec[f"s{stage}_g{n}"]
            assert np.abs(g.cpu().numpy() - ref).max() <= 2 * TOL * np.abs(ref).max() + 1e-7, (stage, n)
