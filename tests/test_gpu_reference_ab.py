"""A/B against the UNMODIFIED reference running on the same B200 (`oracle/_ref`, the reference's own
`VMI.py` / `Model.py` copied by `make -C oracle _ref`, imported with real `.cuda()`), at sizes the committed goldens do
not reach:

* concat critic, NWJ and JS, B = 2048 (BASELINE configs[2] parity size) against the reference in float64;
* the ReLU-kink exclusion of tests/test_gpu_vmi.py::test_large_batch_against_streamed_oracle demonstrated instead of
  asserted: on the excluded rows the reference in fp32 (TF32 off) misses its own float64 evaluation, the kernels
  are compared on ALL rows against both.

Norm: max|a - b| / max|b| per tensor (tests/conftest.py::rel_err), tolerance 1e-4.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ref():
    import __graft_entry__ as g
    g.build()
    from oracle import ref_shim as R
    if R.locate() is None:
        pytest.skip("oracle/_ref not present (make -C oracle _ref needs /root/reference)")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return R.import_reference(cpu=False, random_bert=False)


def _pair(ref, critic, bound, dtype, seed=0):
    """(reference estimator in `dtype`, ours in fp32) with identical weights, on the GPU."""
    from mimrl_b200.model import VMIEstimator
    torch.manual_seed(seed)
    theirs = ref.Model.VMIEstimator(critic, "constant", bound, 128, 256, 128, 2, "relu", 0, 1)
    with torch.no_grad():                               # reference biases start at zero (VMI.py:47-51): make them live
        for n, p in theirs.named_parameters():
            if n.endswith("bias"):
                p.uniform_(-0.05, 0.05)
    ours = VMIEstimator(critic, "constant", bound, 128, 256, 128, 2, "relu", 0, 1)
    ours.load_state_dict(theirs.state_dict(), strict=True)
    return theirs.to("cuda", dtype), ours.cuda()


def _run(est, x, y, dtype):
    est.zero_grad(set_to_none=True)
    xt = x.to("cuda", dtype).requires_grad_(True)
    yt = y.to("cuda", dtype).requires_grad_(True)
    mi, loss = est(xt, yt)
    loss.backward()
    pg = {n: p.grad.detach().double().cpu().numpy() for n, p in est.named_parameters() if p.grad is not None}
    return float(mi), xt.grad.double().cpu().numpy(), yt.grad.double().cpu().numpy(), pg


def _inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 128, generator=g)
    return x, 0.6 * x + 0.8 * torch.randn(B, 128, generator=g)


@pytest.mark.parametrize("bound", ["nwj", "js"])
def test_concat_B2048_against_reference_float64(ref, bound):
    """VMI.py:58-65 (all-pairs concat MLP) + VMI.py:157-182, B = 2048: the reference materialises [B^2, 256] tensors
    (float64: ~9 GB each), the fused kernels keep them in TMEM.

    The yardstick is the reference's float64 run.  At this size the input gradient is a small difference of two sums over
    2048 pairs each (the diagonal term against the mean of the off-diagonal ones), so fp32 arithmetic ITSELF is only good
    to ~1e-3 here: the reference's own fp32 run (TF32 off) is measured against its float64 run in the same test, and the
    kernels must be within 1e-4 of float64, or closer to float64 than twice the fp32 reference's own miss."""
    B = 2048
    theirs, ours = _pair(ref, "concat", bound, torch.float64, seed=11)
    x, y = _inputs(B, 12)
    mi_r, gx_r, gy_r, pg_r = _run(theirs, x, y, torch.float64)
    torch.cuda.empty_cache()
    mi32, gx32, gy32, pg32 = _run(theirs.to(torch.float32), x, y, torch.float32)
    del theirs
    torch.cuda.empty_cache()
    mi, gx, gy, pg = _run(ours, x, y, torch.float32)
    assert abs(mi - mi_r) <= TOL * max(1.0, abs(mi_r)), (mi, mi_r)
    report = {}
    for name, got, ref32, want in [("x", gx, gx32, gx_r), ("y", gy, gy32, gy_r)] + [(n, pg[n], pg32[n], pg_r[n]) for n in pg_r]:
        floor = 1e-7 * max(1.0, float(np.abs(want).max()))
        e_ours = (np.abs(got - want).max() - floor) / np.abs(want).max()
        e_ref = np.abs(ref32 - want).max() / np.abs(want).max()
        report[name] = (float(e_ours), float(e_ref))
    print("rel. error vs float64 (kernels, reference fp32):", report)
    assert set(pg) == set(pg_r)
    for name, (e_ours, e_ref) in report.items():
        assert e_ours <= max(TOL, 2 * e_ref), (name, report)
    # the kernels are not riding on that allowance: over all tensors they are at least as close to float64 as fp32 torch
    assert max(e[0] for e in report.values()) <= max(TOL, max(e[1] for e in report.values())), report


def test_relu_kink_rows_reference_fp32_misses_float64_too(ref):
    """Separable InfoNCE, B = 4096.  Rows whose float64 hidden pre-activation lies within 1e-5 of zero are the ones
    tests/test_gpu_vmi.py leaves out of its input-gradient check.  Evidence for that exclusion: on exactly those rows
    the reference's own fp32 arithmetic on this GPU (TF32 off) disagrees with its float64 evaluation, and wherever the
    kernels exceed the tolerance the fp32 reference does as well; on every other row both stay within 1e-4."""
    B = 4096
    theirs64, ours = _pair(ref, "separate", "infonce", torch.float64, seed=21)
    x, y = _inputs(B, 22)
    mi64, gx64, gy64, _ = _run(theirs64, x, y, torch.float64)

    def kink_rows(mlp, inp):
        h = inp.to("cuda", torch.float64)
        bad = torch.zeros(len(inp), dtype=torch.bool, device="cuda")
        for layer in list(mlp)[:-1]:
            h = layer(h)
            if isinstance(layer, torch.nn.Linear):
                bad |= (h.abs() < 1e-5).any(dim=1)
        return bad.cpu().numpy()
    with torch.no_grad():
        bad_x = kink_rows(theirs64.critic_model.MLP_g, x)
        bad_y = kink_rows(theirs64.critic_model.MLP_h, y)
    theirs32 = theirs64.to(torch.float32)
    mi32, gx32, gy32, _ = _run(theirs32, x, y, torch.float32)
    mi, gx, gy, _ = _run(ours, x, y, torch.float32)
    assert abs(mi - mi64) <= TOL * max(1.0, abs(mi64))

    def row_err(a, b):                  # per row: max |a - b| over the row / max |b| over the tensor
        return np.abs(a - b).max(axis=1) / np.abs(b).max()
    report, worst_ours, worst_ref, ref_missed = {}, 0.0, 0.0, 0
    for name, bad, g_ours, g32, g64 in (("x", bad_x, gx, gx32, gx64), ("y", bad_y, gy, gy32, gy64)):
        e_ours, e_ref32 = row_err(g_ours, g64), row_err(g32, g64)
        assert 0 < bad.sum() < 0.1 * B, bad.sum()
        # away from a kink: both fp32 implementations agree with float64
        assert e_ours[~bad].max() < TOL, (name, e_ours[~bad].max())
        assert e_ref32[~bad].max() < TOL, (name, e_ref32[~bad].max())
        report[name] = (int(bad.sum()), int((bad & (e_ours >= TOL)).sum()), int((bad & (e_ref32 >= TOL)).sum()),
                        float(e_ours[bad].max()), float(e_ref32[bad].max()))
        worst_ours, worst_ref = max(worst_ours, float(e_ours[bad].max())), max(worst_ref, float(e_ref32[bad].max()))
        ref_missed += int((bad & (e_ref32 >= TOL)).sum())
    print("kink rows (n, missed by kernels, missed by fp32 reference, worst kernels, worst reference):", report)
    # at a kink the mask of a unit is decided by the rounding of its pre-activation, independently in every fp32
    # implementation: the reference's own fp32 arithmetic misses its float64 evaluation on some of these rows (so the
    # exclusion is needed to compare ANY fp32 run with float64), and a miss of the kernels is of the same size -- one
    # flipped unit -- not a loss of precision
    assert ref_missed > 0 and worst_ref >= TOL, report
    assert worst_ours <= 10 * worst_ref, report
