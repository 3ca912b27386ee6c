"""fp32-class tensor-core GEMM (mimrl_gemm_f32x3) and the Linear layers built on
it, against float64 torch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import __graft_entry__ as g
    g.build()
    torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K", [(1000, 256, 128), (128, 128, 64), (777, 130, 100), (4096, 128, 384), (65, 40, 72)])
def test_modes_against_float64(M, N, K):
    from mimrl_b200.linear import _gemm
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bt = torch.randn(N, K, device="cuda", generator=g) * 0.3
    bias = torch.randn(N, device="cuda", generator=g)
    want = A.double() @ Bt.double().t() + bias.double()
    got = _gemm(0, A, None, Bt, M, N, K, bias, False)
    assert rel(got, want) < 1e-5
    got = _gemm(0, A, None, Bt, M, N, K, bias, True)
    assert rel(got, want.clamp_min(0)) < 1e-5
    Bn = Bt.t().contiguous()                                       # [K, N]
    mask = torch.randn(M, K, device="cuda", generator=g)
    want = (A.double() * (mask > 0)) @ Bn.double()
    assert rel(_gemm(1, A, mask, Bn, M, N, K), want) < 1e-5
    At = torch.randn(K, M, device="cuda", generator=g)             # mode 2: contraction over the rows
    maskt = torch.randn(K, M, device="cuda", generator=g)
    want = (At.double() * (maskt > 0)).t() @ Bn.double()
    assert rel(_gemm(2, At, maskt, Bn, M, N, K), want) < 1e-5


def test_mode2_long_contraction_split_k():
    from mimrl_b200.linear import _gemm
    g = torch.Generator(device="cuda").manual_seed(1)
    K, M, N = 50000, 256, 128
    At = torch.randn(K, M, device="cuda", generator=g)
    Bn = torch.randn(K, N, device="cuda", generator=g)
    want = At.double().t() @ Bn.double()
    assert rel(_gemm(2, At, None, Bn, M, N, K), want) < 2e-5


@pytest.mark.parametrize("rows", [512, 3000])
def test_mlp_stack_matches_torch(rows):
    """The relu MLP stack of VMI.py:13-22 through mlp_apply vs the same modules in float64."""
    from mimrl_b200.linear import mlp_apply
    from mimrl_b200.vmi import mlps
    torch.manual_seed(0)
    seq = mlps(128, 256, 128, 2, "relu").cuda()
    for m in seq:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.uniform_(m.bias, -0.05, 0.05)
    x = torch.randn(rows, 128, device="cuda", requires_grad=True)
    w = torch.randn(rows, 128, device="cuda")
    y = mlp_apply(seq, x)
    (y * w).sum().backward()
    got = [y.detach(), x.grad.clone()] + [p.grad.clone() for p in seq.parameters()]
    seq64 = mlps(128, 256, 128, 2, "relu").cuda().double()
    seq64.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = seq64(x64)
    (y64 * w.double()).sum().backward()
    want = [y64.detach(), x64.grad] + [p.grad for p in seq64.parameters()]
    for a, b in zip(got, want):
        assert rel(a, b) < 1e-5
