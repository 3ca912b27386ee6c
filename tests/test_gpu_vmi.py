"""GPU parity of the variational-MI path (through the C ABI) against
(1) the golden vectors produced by the reference and (2) the numpy oracle on
seeded inputs, including ragged sizes.  Tolerance: 1e-4 relative on bound
values and gradients (BASELINE.json north_star), fp32, TF32 disabled."""
import numpy as np
import pytest
import torch

from conftest import cfg_of, load_golden, rel_err
from oracle import params as P
from oracle import vmi_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
VMI = load_golden("vmi")
BOUNDS = load_golden("bounds")


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import __graft_entry__ as g
    g.build()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def dev():
    return torch.device("cuda:0")


def make_estimator(c, prm, impl=0):
    from mimrl_b200.model import VMIEstimator
    est = VMIEstimator(c["critic"], c["baseline"], c["bound"], c["d"], c["hidden"], c["embed"], c["layers"], "relu", 0, 1)
    est.load_state_dict({k: torch.tensor(v) for k, v in P.vmi_state_dict(prm).items()}, strict=True)
    est.impl = impl
    return est.to(dev())


def run(est, x, y):
    xt = torch.tensor(x, device=dev(), requires_grad=True)
    yt = torch.tensor(y, device=dev(), requires_grad=True)
    mi, loss = est(xt, yt)
    loss.backward()
    pg = {n: p.grad.detach().cpu().numpy() for n, p in est.named_parameters() if p.grad is not None}
    return float(mi), float(loss), xt.grad.cpu().numpy(), yt.grad.cpu().numpy(), pg


def close_scalar(a, b, tol=TOL):
    return abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("case", sorted(VMI))
def test_estimator_matches_reference_golden(case):
    rec = VMI[case]
    c = cfg_of(rec)
    seed = int(rec["seed"])
    prm = P.vmi_params(seed, c["critic"], c["baseline"], c["d"], c["hidden"], c["embed"], c["layers"])
    x, y = P.features(seed + 7, c["B"], c["d"], scale=c["scale"], corr=0.6)
    mi, loss, gx, gy, pg = run(make_estimator(c, prm), x, y)
    assert close_scalar(mi, float(rec["mi"])), (mi, float(rec["mi"]))
    assert close_scalar(loss, float(rec["loss"]))
    assert rel_err(gx, rec["gx"]) < TOL and rel_err(gy, rec["gy"]) < TOL
    for k, v in rec.items():
        if k.startswith("pg__"):
            # analytically-zero sums (e.g. InfoNCE last-layer bias: rows of G sum to 0) are pure fp32 noise
            assert np.abs(pg[k[4:]] - v).max() <= TOL * np.abs(v).max() + 5e-7, k
        elif k.startswith("pgs__"):
            g = pg[k[5:]].ravel().astype(np.float64)
            got = np.array([np.abs(g).sum(), np.sqrt((g ** 2).sum())])
            assert np.allclose(got, v[1:], rtol=2e-4, atol=1e-5), k   # atol: analytically-zero sums are fp32 noise


@pytest.mark.parametrize("bound", ["infonce", "dv", "mine", "tuba", "nwj", "js_fgan", "js", "smile", "interpolate"])
@pytest.mark.parametrize("B,d,impl", [(1000, 128, 0), (129, 128, 0), (257, 48, 1), (64, 128, 1), (3, 16, 1)])
def test_separable_bounds_vs_oracle(bound, B, d, impl):
    """Seeded inputs, ragged batch sizes, both kernel implementations (the interpolated bound is fused on the tcgen05
    path, impl 0, and runs on the materialised matrix with the CUDA-core implementation, impl 1)."""
    baseline = "unnormalized" if bound in ("tuba", "interpolate") else "constant"
    hidden = 64
    prm = P.vmi_params(7 + B, "separate", baseline, d, hidden, d, 2)
    x, y = P.features(8 + B, B, d, scale=1.5, corr=0.7)
    c = dict(critic="separate", baseline=baseline, bound=bound, d=d, hidden=hidden, embed=d, layers=2)
    mi, loss, gx, gy, pg = run(make_estimator(c, prm, impl), x, y)
    r = O.vmi_estimator(prm, "separate", baseline, bound, x, y)
    assert close_scalar(mi, r["mi"]), (mi, r["mi"])
    assert close_scalar(loss, r["loss"]), (loss, r["loss"])
    assert rel_err(gx, r["gx"]) < TOL, rel_err(gx, r["gx"])
    assert rel_err(gy, r["gy"]) < TOL, rel_err(gy, r["gy"])
    for k, v in r["pg"].items():
        if k.endswith("weight"):
            assert rel_err(pg[k], v) < 2 * TOL, k


@pytest.mark.parametrize("B,scale", [(4096, 1.0), (2500, 4.0)])
def test_interpolate_fused_large_batch_vs_oracle(B, scale):
    """VMI.py:201-250 through the fused sweeps (row / column / interp statistics, MIMRL_WEIGHT_INTERP weighted sums) at a
    batch the goldens do not reach, against the float64 oracle; scale = 4 gives peaked rows (p_ij up to ~1)."""
    prm = P.vmi_params(41, "separate", "unnormalized", 128, 64, 128, 2)
    x, y = P.features(42, B, 128, scale=scale, corr=0.7)
    c = dict(critic="separate", baseline="unnormalized", bound="interpolate", d=128, hidden=64, embed=128, layers=2)
    mi, loss, gx, gy, pg = run(make_estimator(c, prm), x, y)
    r = O.vmi_estimator(prm, "separate", "unnormalized", "interpolate", x, y)
    assert close_scalar(mi, r["mi"]), (mi, r["mi"])
    assert rel_err(gx, r["gx"]) < TOL and rel_err(gy, r["gy"]) < TOL, (rel_err(gx, r["gx"]), rel_err(gy, r["gy"]))
    # (parameter gradients are checked at the oracle sizes above: with thousands of rows a few ReLU units sit within fp32
    # rounding of their kink, and one flipped mask moves a weight gradient by more than 1e-4 in ANY fp32 run -- see
    # tests/test_gpu_reference_ab.py::test_relu_kink_rows_reference_fp32_misses_float64_too)
    for k, v in r["pg"].items():
        if k.endswith("weight"):
            assert np.linalg.norm(pg[k] - v) <= 1e-3 * np.linalg.norm(v), k       # one flipped unit of 4096 x 192 ~ 3e-4


@pytest.mark.parametrize("case", sorted(BOUNDS))
def test_free_bound_functions_match_reference(case):
    import mimrl_b200.vmi as V
    rec = BOUNDS[case]
    a = torch.tensor(rec["a"], device=dev())
    fns = dict(dv=V.dv_lower_bound, tuba=lambda s: V.tuba_lower_bound(s, a), tuba_nobase=V.tuba_lower_bound,
               nwj=V.nwj_lower_bound, infonce=V.infonce_lower_bound, js_fgan=V.js_fgan_lower_bound,
               js=V.js_lower_bound, smile=V.smile_lower_bound,
               interpolate=lambda s: V.interp_lower_bound(s, a, 0.01))
    for name, fn in fns.items():
        s = torch.tensor(rec["S"], device=dev(), requires_grad=True)
        v = fn(s)
        v.backward()
        assert close_scalar(float(v), float(rec["val_" + name])), name
        assert rel_err(s.grad.cpu().numpy(), rec["grad_" + name]) < TOL, name


def test_mine_mi_gradient_is_dv():
    """The MINE branch returns (dv value, its own loss); both outputs carry gradients."""
    import mimrl_b200.vmi as V
    x, y = P.features(3, 50, 32, corr=0.5)
    xe = torch.tensor(x, device=dev(), requires_grad=True)
    ye = torch.tensor(y, device=dev(), requires_grad=True)
    mi, _ = V.separable_bound(xe, ye, "mine")
    mi.backward()
    S = y.astype(np.float64) @ x.astype(np.float64).T
    _, G, _ = O.bound_dv(S)
    assert rel_err(ye.grad.cpu().numpy(), G @ x) < TOL
    assert rel_err(xe.grad.cpu().numpy(), G.T @ y) < TOL


@pytest.mark.parametrize("B", [4096, 20000])
def test_large_batch_against_streamed_oracle(B):
    """Sizes where B x B does not fit comfortably on the host in float64: the
    oracle streams row blocks; the fused kernels never build the matrix.

    Rows whose float64 pre-activations sit within 1e-5 of a ReLU kink are left
    out of the input-gradient comparison: there the fp32 reference itself
    (torch, checked on the B200) flips the ReLU mask against float64 and
    differs by ~5e-3, which says nothing about the kernels under test."""
    prm = P.vmi_params(99, "separate", "constant", 128, 256, 128, 2)
    x, y = P.features(100, B, 128, corr=0.6)
    st = O.separable_infonce_streamed(prm, x, y, dtype=np.float64, block=2048)
    c = dict(critic="separate", baseline="constant", bound="infonce", d=128, hidden=256, embed=128, layers=2)
    mi, loss, gx, gy, _ = run(make_estimator(c, prm), x, y)
    assert close_scalar(mi, st["mi"]), (mi, st["mi"])

    def safe_rows(stack, inp):
        h = inp.astype(np.float64)
        ok = np.ones(len(inp), bool)
        for w, b in stack[:-1]:
            z = h @ w.T.astype(np.float64) + b.astype(np.float64)      # pre-activation of a hidden layer
            ok &= (np.abs(z) > 1e-5).all(axis=1)
            h = np.maximum(z, 0)
        return ok
    okx, oky = safe_rows(prm["g"], x), safe_rows(prm["h"], y)
    assert okx.mean() > 0.9 and oky.mean() > 0.9
    assert rel_err(gx[okx], st["gx"][okx]) < TOL, rel_err(gx[okx], st["gx"][okx])
    assert rel_err(gy[oky], st["gy"][oky]) < TOL, rel_err(gy[oky], st["gy"][oky])


def test_full_size_properties():
    """BASELINE config 2 at its largest size (B = 65536): size-independent
    properties instead of an oracle run.  (1) InfoNCE <= log B; (2) duplicating
    nothing but permuting rows of x and y together leaves mi unchanged;
    (3) row-sums of the InfoNCE gradient vanish: sum_j G_ij = 0 implies
    sum_i dL/dy_emb_i . 1 relation  ->  d mi / d (a constant shift of all scores) = 0,
    checked as <grad_y_emb, y_emb> + <grad_x_emb, x_emb> = 2 * sum_ij G_ij S_ij
    against the same quantity from the two implementations (FFMA vs auto)."""
    import mimrl_b200.vmi as V
    from mimrl_b200 import _lib as L
    B, E = 65536, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    xe = torch.randn(B, E, device=dev(), generator=g) * 0.3
    ye = 0.7 * xe + 0.3 * torch.randn(B, E, device=dev(), generator=g) * 0.3
    outs = []
    for impl in (L.IMPL_AUTO,):
        a = xe.clone().requires_grad_(True)
        b = ye.clone().requires_grad_(True)
        mi, loss = V.separable_bound(a, b, "infonce", impl=impl)
        loss.backward()
        outs.append((float(mi), a.grad, b.grad))
    mi0, gx0, gy0 = outs[0]
    assert mi0 <= np.log(B) + 1e-4
    perm = torch.randperm(B, device=dev(), generator=g)
    mi_p, _ = V.separable_bound(xe[perm].contiguous(), ye[perm].contiguous(), "infonce")
    assert abs(float(mi_p) - mi0) <= 1e-4 * max(1.0, abs(mi0))
    # gradient of a permuted problem is the permuted gradient
    a = xe[perm].clone().requires_grad_(True)
    b = ye[perm].clone().requires_grad_(True)
    V.separable_bound(a, b, "infonce")[1].backward()
    assert rel_err(a.grad[:4096].cpu().numpy(), gx0[perm][:4096].cpu().numpy()) < TOL
    # Euler identity for scores homogeneous of degree 1 in each operand
    lhs = float((gx0.double() * xe.double()).sum())
    rhs = float((gy0.double() * ye.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1e-6)
    # a 2048-row slab of the gradient against a direct float64 evaluation on the GPU
    rows = slice(1000, 1000 + 2048)
    S = ye[rows].double() @ xe.double().t()
    Pm = torch.softmax(S, dim=1)
    Gs = Pm / B
    idx = torch.arange(1000, 1000 + 2048, device=dev())
    Gs[torch.arange(2048, device=dev()), idx] -= 1.0 / B
    want = (Gs @ xe.double()).cpu().numpy()            # d loss / d y_emb rows
    assert rel_err(gy0[rows].cpu().numpy(), want) < TOL


@pytest.mark.parametrize("bound", ["infonce", "dv", "mine", "tuba", "nwj"])
@pytest.mark.parametrize("B,scale", [(700, 1.0), (1500, 6.0), (900, 40.0)])
def test_fused_forward_matches_three_sweep_path(bound, B, scale, monkeypatch):
    """The one-sweep forward (mimrl_sep_online_forward: online reference point, statistics and the owned-row gradient
    sum together) against the exact-statistics path it replaces, and both against the float64 oracle.  scale = 6 makes
    the scores large (|S| ~ 100s), scale = 40 gives |S| ~ 1e4."""
    import mimrl_b200.vmi as V
    if bound not in ("infonce", "dv") and scale > 10:
        pytest.skip("mine / tuba / nwj exponentiate the scores unshifted (VMI.py:148-159, Model.py:121-124): at |S| ~ 1e4 "
                    "they overflow in the reference itself; only the log-domain bounds are meaningful here")
    baseline = "unnormalized" if bound == "tuba" else "constant"
    c = dict(critic="separate", baseline=baseline, bound=bound, d=128, hidden=64, embed=128, layers=2)
    prm = P.vmi_params(17, "separate", baseline, 128, 64, 128, 2)
    x, y = P.features(18, B, 128, scale=scale, corr=0.6)
    out = {}
    for fused in (True, False):
        monkeypatch.setattr(V, "FUSED_FORWARD", fused)
        out[fused] = run(make_estimator(c, prm), x, y)
    ref = O.vmi_estimator(prm, "separate", baseline, bound, x, y)
    for fused in (True, False):
        mi, loss, gx, gy, pg = out[fused]
        assert close_scalar(mi, ref["mi"]), (fused, mi, ref["mi"])
        assert rel_err(gx, ref["gx"]) < TOL and rel_err(gy, ref["gy"]) < TOL, fused
    assert rel_err(out[True][2], out[False][2]) < TOL and rel_err(out[True][3], out[False][3]) < TOL


@pytest.mark.parametrize("n_own,n_all,offset,inc", [(300, 300, 0, 1), (200, 517, 130, 0), (129, 700, 571, 1)])
def test_fused_forward_abi_row_block(n_own, n_all, offset, inc):
    """mimrl_sep_row_stats(MIMRL_STAT_MAXONLY) + mimrl_sep_fused_forward on a row block of a larger batch, straight
    through the C ABI, against float64: approximate off-diagonal row maxima, exact off-diagonal weight sums and the
    weighted sum (diagonal included iff include_diag) relative to the given reference point."""
    from mimrl_b200 import _lib as L
    rng = np.random.default_rng(5)
    E = 128
    own = rng.standard_normal((n_own, E)).astype(np.float32)
    swept = rng.standard_normal((n_all, E)).astype(np.float32)
    T = lambda a: torch.tensor(a, device=dev())
    o, a = T(own), T(swept)
    ws_b = L.lib.mimrl_sep_workspace_bytes(n_own, n_all, E)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=dev())
    pre = torch.empty(4, n_own, device=dev())
    L.check(L.lib.mimrl_sep_row_stats(L.ptr(o), L.ptr(a), n_own, n_all, E, offset, L.STAT_MAXONLY, L.IMPL_TCGEN05, L.ptr(pre[0]),
                                      L.ptr(pre[1]), L.ptr(pre[2]), L.ptr(pre[3]), L.ptr(ws), ws_b, L.stream()))
    S = own.astype(np.float64) @ swept.astype(np.float64).T
    rows = np.arange(n_own)
    diag = S[rows, offset + rows]
    off = S.copy()
    off[rows, offset + rows] = -np.inf
    approx = pre[0].cpu().numpy()
    bound = 2.0 ** -11 * np.linalg.norm(own, axis=1) * np.linalg.norm(swept, axis=1).max()
    assert np.all(np.abs(approx - off.max(axis=1)) <= bound + 1e-4)
    assert np.allclose(pre[3].cpu().numpy(), diag, rtol=1e-5, atol=1e-4)
    ref = (off.max(axis=1) + 0.3).astype(np.float32)                      # any reference point near the maximum
    wsum = torch.empty(n_own, E, device=dev())
    rsum = torch.empty(n_own, device=dev())
    L.check(L.lib.mimrl_sep_fused_forward(L.ptr(o), L.ptr(a), n_own, n_all, E, offset, inc, L.ptr(T(ref)), L.ptr(wsum),
                                          L.ptr(rsum), L.ptr(ws), ws_b, L.stream()))
    W = np.exp(S - ref.astype(np.float64)[:, None])
    Woff = W.copy()
    Woff[rows, offset + rows] = 0.0
    assert rel_err(rsum.cpu().numpy(), Woff.sum(axis=1)) < 5e-5          # fp32 ulp of a score ~ 40 is 4e-6
    want = (W if inc else Woff) @ swept.astype(np.float64)
    assert rel_err(wsum.cpu().numpy(), want) < 5e-5


@pytest.mark.parametrize("n_own,n_all,offset,inc,ramp", [(300, 300, 0, 1, 0.0), (200, 517, 130, 0, 0.0), (129, 700, 571, 1, 0.05),
                                                         (256, 4096, 1024, 1, 0.02), (140, 20000, 7000, 0, 0.004),
                                                         (1, 1, 0, 0, 0.0), (2, 3, 1, 1, 0.0)])
def test_online_forward_abi_row_block(n_own, n_all, offset, inc, ramp):
    """mimrl_sep_online_forward straight through the C ABI against float64: the returned reference point lies within
    kOnlineTau (4.5) below the true row maximum, the off-diagonal weight sum and the weighted sum (diagonal included
    iff include_diag) are exact relative to it.  ramp > 0 scales the swept rows up along the batch, so the running
    maximum keeps growing and the accumulators are rescaled many times during one sweep; the single-pair and
    empty-off-diagonal corners (n_all = 1) are included."""
    from mimrl_b200 import _lib as L
    rng = np.random.default_rng(7)
    E = 128
    own = rng.standard_normal((n_own, E)).astype(np.float32)
    swept = rng.standard_normal((n_all, E)).astype(np.float32)
    if ramp:
        swept *= (1.0 + ramp * np.arange(n_all, dtype=np.float32))[:, None] ** 0.5
    T = lambda a: torch.tensor(a, device=dev())
    o, a = T(own), T(swept)
    ws_b = L.lib.mimrl_sep_workspace_bytes(n_own, n_all, E)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=dev())
    ref, rsum, diag = (torch.empty(n_own, device=dev()) for _ in range(3))
    wsum = torch.empty(n_own, E, device=dev())
    L.check(L.lib.mimrl_sep_online_forward(L.ptr(o), L.ptr(a), n_own, n_all, E, offset, inc, L.ptr(ref), L.ptr(wsum),
                                           L.ptr(rsum), L.ptr(diag), L.ptr(ws), ws_b, L.stream()))
    S = own.astype(np.float64) @ swept.astype(np.float64).T
    rows = np.arange(n_own)
    seen = S.copy()
    if not inc:
        seen[rows, offset + rows] = -np.inf
    ref = ref.cpu().numpy().astype(np.float64)
    top = seen.max(axis=1)
    live = np.isfinite(top)
    assert np.all(ref[live] <= top[live] + 1e-3 * np.abs(top[live]).clip(1.0)) and np.all(ref[live] >= top[live] - 4.5 - 1e-3)
    assert np.all(ref[~live] == 0.0)
    assert np.allclose(diag.cpu().numpy(), S[rows, offset + rows], rtol=1e-5, atol=1e-4)
    W = np.exp(S - ref[:, None])
    Woff = W.copy()
    Woff[rows, offset + rows] = 0.0
    want_sum = Woff.sum(axis=1)
    # a weight is exp(S - ref): an ABSOLUTE error of the score is a RELATIVE error of the weight.  The fp32-class score
    # (three fp16 split products, fp32 accumulation) is good to 2^-21 |own_i| |swept_j| (dropped lo.lo term 2^-22 plus the
    # accumulation), the same order as an fp32 dot product of 128 terms; the ramped cases reach |own||swept| ~ 1e3
    tol = 5e-5 + 2.0 ** -21 * float(np.linalg.norm(own, axis=1).max() * np.linalg.norm(swept, axis=1).max())
    assert np.all(np.abs(rsum.cpu().numpy() - want_sum) <= tol * np.maximum(want_sum, 1e-30))
    want = (W if inc else Woff) @ swept.astype(np.float64)
    if np.abs(want).max() > 0:
        assert rel_err(wsum.cpu().numpy(), want) < tol
    else:
        assert np.all(wsum.cpu().numpy() == 0)


def test_fused_forward_one_sided_gradients():
    """Only one of the two embeddings requires a gradient (stage 1 trains the critics on detached features of one
    side in some configurations): the other gradient is None and the computed one is unchanged."""
    import mimrl_b200.vmi as V
    x, y = P.features(21, 640, 128, corr=0.5)
    full = {}
    for gx, gy in ((True, True), (True, False), (False, True)):
        xe = torch.tensor(x, device=dev(), requires_grad=gx)
        ye = torch.tensor(y, device=dev(), requires_grad=gy)
        mi, loss = V.separable_bound(xe, ye, "infonce")
        loss.backward()
        full[(gx, gy)] = (xe.grad, ye.grad)
    assert full[(True, False)][1] is None and full[(False, True)][0] is None
    assert torch.equal(full[(True, False)][0], full[(True, True)][0])
    assert torch.equal(full[(False, True)][1], full[(True, True)][1])
