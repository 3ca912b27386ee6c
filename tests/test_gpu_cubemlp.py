"""GPU parity of the fused CubeMLP kernels against the reference golden
vectors (MLPProcess.py run unmodified) and the numpy oracle."""
import ast

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import cubemlp_oracle as C
from oracle import params as P

pytestmark = pytest.mark.gpu
CUBE = load_golden("cubemlp")
TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import __graft_entry__ as g
    g.build()
    torch.backends.cuda.matmul.allow_tf32 = False


def build(c, blocks):
    from mimrl_b200.mlp_process import MLPEncoder
    enc = MLPEncoder(activate=c["act"], d_in=c["d_in"], d_hiddens=c["d_hiddens"], d_outs=c["d_outs"],
                     dropouts=[0.0, 0.0, 0.0], bias=c["bias"], ln_first=c["ln_first"], res_project=c["res"])
    enc.load_state_dict({k: torch.tensor(v) for k, v in P.cubemlp_state_dict(blocks).items()}, strict=True)
    return enc.cuda()


@pytest.mark.parametrize("case", sorted(CUBE))
def test_encoder_matches_reference_golden(case):
    rec = CUBE[case]
    c = ast.literal_eval(str(rec["cfg"]))
    seed = int(rec["seed"])
    blocks = P.cubemlp_params(seed, c["d_in"], c["d_hiddens"], c["d_outs"], c["bias"], c["ln_first"], c["res"])
    x = P.features(seed + 1, c["bs"] * c["d_in"][0] * c["d_in"][1], c["d_in"][2]).reshape(c["bs"], *c["d_in"])
    oshape = (c["bs"], *c["d_outs"][-1])
    w = P.features(seed + 2, int(np.prod(oshape[:-1])), oshape[-1]).reshape(oshape)
    enc = build(c, blocks)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = enc(xt, mask=None)
    (y * torch.tensor(w, device="cuda")).sum().backward()
    big = c["d_in"][2] >= 128
    yn, gx = y.detach().cpu().numpy(), xt.grad.cpu().numpy()
    if big:
        yn, gx = yn[:, :, :, ::8], gx[:, ::5, :, ::8]
    assert rel_err(yn, rec["y"]) < TOL, rel_err(yn, rec["y"])
    assert rel_err(gx, rec["gx"]) < 2 * TOL, rel_err(gx, rec["gx"])
    pg = {n: p.grad.cpu().numpy() for n, p in enc.named_parameters()}
    for k, v in rec.items():
        if k.startswith("pg__"):
            assert np.abs(pg[k[4:]] - v).max() <= 2 * TOL * np.abs(v).max() + 2e-6, k
        elif k.startswith("pgs__"):
            g = pg[k[5:]].ravel().astype(np.float64)
            assert np.allclose([np.abs(g).sum(), np.sqrt((g ** 2).sum())], v[1:], rtol=5e-4, atol=1e-5), k


@pytest.mark.parametrize("bs,act,ln_first,res", [(37, "gelu", False, True), (5, "relu", True, True),
                                                  (9, "tanh", False, False)])
def test_encoder_vs_oracle_ragged(bs, act, ln_first, res):
    """Column counts that are not multiples of the 32-column tile, all three activations."""
    d_in = [11, 3, 20]
    d_h = [[7, 5, 24]] if res else [[7, 5, 24]]
    d_out = [[6, 3, 12]] if res else [[11, 3, 20]]      # (LN over a 2-wide axis is ill-conditioned)
    c = dict(act=act, d_in=d_in, d_hiddens=d_h, d_outs=d_out, bias=True, ln_first=ln_first, res=[res])
    blocks = P.cubemlp_params(77, d_in, d_h, d_out, True, ln_first, [res])
    x = P.features(78, bs * d_in[0] * d_in[1], d_in[2]).reshape(bs, *d_in)
    oshape = (bs, *d_out[-1])
    w = P.features(79, int(np.prod(oshape[:-1])), oshape[-1]).reshape(oshape)
    enc = build(c, blocks)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = enc(xt)
    (y * torch.tensor(w, device="cuda")).sum().backward()
    yo, caches = C.encoder_forward(blocks, x, act, ln_first, [res])
    gxo, pgo = C.encoder_backward(caches, w.astype(np.float64), ln_first, [res])
    assert rel_err(y.detach().cpu().numpy(), yo) < TOL
    assert rel_err(xt.grad.cpu().numpy(), gxo) < 2 * TOL
    for n, p in enc.named_parameters():
        assert np.abs(p.grad.cpu().numpy() - pgo[n]).max() <= 2 * TOL * np.abs(pgo[n]).max() + 2e-6, n


def test_state_dict_names():
    from mimrl_b200.mlp_process import MLPEncoder
    enc = MLPEncoder("gelu", [10, 3, 16], [[5, 3, 16]], [[5, 3, 16]], [0.0] * 3, True, False, [True])
    want = set(P.cubemlp_state_dict(P.cubemlp_params(0, [10, 3, 16], [[5, 3, 16]], [[5, 3, 16]], True, False, [True])))
    assert set(enc.state_dict()) == want


@pytest.mark.parametrize("bs,d_in,d_h,d_out,act,res", [
    (16, [100, 3, 128], [[50, 3, 128]], [[50, 3, 128]], "gelu", True),          # README block 1: L-mix and D-mix on tcgen05
    (16, [50, 3, 128], [[10, 3, 128]], [[10, 3, 128]], "gelu", True),           # README block 2
    (9, [40, 4, 72], [[24, 4, 100]], [[40, 4, 72]], "relu", False),             # identity residual, odd sizes
    (7, [33, 4, 20], [[17, 4, 40]], [[21, 4, 12]], "tanh", True)])
def test_tensor_core_forward_vs_oracle(bs, d_in, d_h, d_out, act, res):
    """Shapes large enough (>= 1024 fibres per mix) to take the tcgen05 forward; backward is the recompute kernel."""
    c = dict(act=act, d_in=d_in, d_hiddens=d_h, d_outs=d_out, bias=True, ln_first=False, res=[res])
    blocks = P.cubemlp_params(91, d_in, d_h, d_out, True, False, [res])
    x = P.features(92, bs * d_in[0] * d_in[1], d_in[2]).reshape(bs, *d_in)
    oshape = (bs, *d_out[-1])
    w = P.features(93, int(np.prod(oshape[:-1])), oshape[-1]).reshape(oshape)
    enc = build(c, blocks)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = enc(xt)
    (y * torch.tensor(w, device="cuda")).sum().backward()
    yo, caches = C.encoder_forward(blocks, x, act, False, [res])
    gxo, pgo = C.encoder_backward(caches, w.astype(np.float64), False, [res])
    assert rel_err(y.detach().cpu().numpy(), yo) < TOL, rel_err(y.detach().cpu().numpy(), yo)
    assert rel_err(xt.grad.cpu().numpy(), gxo) < 2 * TOL
    for n, p in enc.named_parameters():
        assert np.abs(p.grad.cpu().numpy() - pgo[n]).max() <= 2 * TOL * np.abs(pgo[n]).max() + 2e-6, n


@pytest.mark.parametrize("bs", [24, 131])
def test_specialised_kernels_agree_with_the_general_kernel(bs, monkeypatch):
    """README configuration (two blocks) at batches with whole and ragged channel-mix tiles: the compile-time specialised
    sequence- / channel-mix kernels (cubemlp_tc2.cu, cubemlp_tc3.cu) against the general tensor-core kernel
    (cubemlp_tc.cu), outputs and every gradient."""
    d_in, d_h, d_out = [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]]
    c = dict(act="gelu", d_in=d_in, d_hiddens=d_h, d_outs=d_out, bias=True, ln_first=False, res=[True, True])
    blocks = P.cubemlp_params(123, d_in, d_h, d_out, True, False, [True, True])
    enc = build(c, blocks)
    g = torch.Generator(device="cuda").manual_seed(bs)
    x = torch.randn(bs, *d_in, device="cuda", generator=g)
    w = torch.randn(bs, *d_out[-1], device="cuda", generator=g)

    def run():
        for p_ in enc.parameters():
            p_.grad = None
        xt = x.clone().requires_grad_(True)
        y = enc(xt)
        (y * w).sum().backward()
        return y.detach(), xt.grad, {n: p_.grad.clone() for n, p_ in enc.named_parameters()}

    y1, gx1, pg1 = run()
    monkeypatch.setenv("MIMRL_CUBE2_OFF", "1")
    monkeypatch.setenv("MIMRL_CUBE3_OFF", "1")
    y0, gx0, pg0 = run()
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 2e-5
    assert rel_err(gx1.cpu().numpy(), gx0.cpu().numpy()) < 2 * TOL       # each is ~1e-4 from float64 at this depth (scripts/cube_ab.py)
    for n in pg0:
        a, b = pg1[n].cpu().numpy(), pg0[n].cpu().numpy()
        # (parameter gradients are sums over 1e5-1e6 fibres; each family is within 2 TOL of float64 -- the bound of the
        # oracle tests above, measured in scripts/cube_ab.py -- so the two may differ by twice that)
        assert np.abs(a - b).max() <= 4 * TOL * np.abs(b).max() + 2e-6, n
