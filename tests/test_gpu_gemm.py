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
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-3))


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


@pytest.mark.parametrize("M,d_in,d_out", [(512, 128, 128), (700, 128, 128), (1000, 64, 32), (3001, 100, 96),
                                          (600, 36, 40), (513, 33, 34)])
def test_fused_mlp4_vs_float64(M, d_in, d_out):
    """mimrl_mlp4_fwd (the critic MLP of VMI.py:13-22 in one kernel) + the per-layer backward on its operands against
    float64.  Rows with a pre-activation within 1e-5 of a ReLU kink get a zero upstream gradient (their mask is
    decided by rounding in any fp32 implementation); the forward is checked on every row."""
    import mimrl_b200.linear as LN
    torch.manual_seed(M)
    dev = "cuda"
    mods = [torch.nn.Linear(d_in, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
            torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, d_out)]
    m = torch.nn.Sequential(*mods).to(dev)
    for p in m.parameters():
        if p.dim() == 1:
            torch.nn.init.normal_(p, std=0.1)
    x = torch.randn(M, d_in, device=dev) * 1.5
    assert LN._is_mlp4(list(m), x)
    m64 = torch.nn.Sequential(*[torch.nn.Linear(l.in_features, l.out_features) if isinstance(l, torch.nn.Linear)
                                else torch.nn.ReLU() for l in mods]).double().to(dev)
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    with torch.no_grad():
        h, near = x.double(), torch.zeros(M, dtype=torch.bool, device=dev)
        for i, l in enumerate(m64):
            h = l(h)
            if isinstance(l, torch.nn.Linear) and i < 6:
                near |= (h.abs() < 1e-5).any(dim=1)
    assert float(near.double().mean()) < 0.2
    w = torch.randn(M, d_out, device=dev).abs()
    w[near] = 0
    xt = x.clone().requires_grad_(True)
    y = LN.mlp_apply(m, xt)
    (y * w).sum().backward()
    x64 = x.double().requires_grad_(True)
    y64 = m64(x64)
    (y64 * w.double()).sum().backward()
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
    assert rel(y.detach(), y64.detach()) < 1e-5
    assert rel(xt.grad, x64.grad) < 1e-4
    for (n, p), p64 in zip(m.named_parameters(), m64.parameters()):
        assert rel(p.grad, p64.grad) < 1e-4, n


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (128, 1, 256), (96, 2, 256), (1, 1, 1), (130, 70, 33), (3000, 2, 256),
                                   (65536, 1, 256), (17, 300, 1000)])
def test_small_linear_modes_against_float64(M, N, K):
    """csrc/linear_small.cu (CUDA cores, exact fp32): the three products of a Linear layer for small batches and the
    1- / 2-wide heads, with bias + ReLU, the ReLU mask on load and the fused bias gradient (colsum)."""
    from mimrl_b200.linear import _small
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bt = torch.randn(N, K, device="cuda", generator=g) * 0.3
    bias = torch.randn(N, device="cuda", generator=g)
    want = A.double() @ Bt.double().t() + bias.double()
    assert rel(_small(0, A, None, Bt, M, N, K, bias, False), want) < 2e-6
    assert rel(_small(0, A, None, Bt, M, N, K, bias, True), want.clamp_min(0)) < 2e-6
    Bn = Bt.t().contiguous()                                       # [K, N]
    mask = torch.randn(M, K, device="cuda", generator=g)
    want = (A.double() * (mask > 0)) @ Bn.double()
    assert rel(_small(1, A, mask, Bn, M, N, K), want) < 2e-6
    At = torch.randn(K, M, device="cuda", generator=g)             # mode 2: contraction over the rows (K may be the batch)
    maskt = torch.randn(K, M, device="cuda", generator=g)
    masked = At.double() * (maskt > 0)
    want = masked.t() @ Bn.double()
    colsum = torch.full((M,), 0.5, device="cuda")
    got = _small(2, At, maskt, Bn, M, N, K, colsum=colsum)
    assert rel(got, want) < 5e-6
    assert rel(colsum, masked.sum(0) + 0.5) < 5e-6                # accumulated (+=) onto what was there


@pytest.mark.parametrize("rows,d_in,d_out", [(128, 128, 128), (96, 128, 1), (2048, 384, 2), (7, 128, 128), (256, 384, 2),
                                            (130, 127, 5)])
def test_small_batch_mlp_stack_matches_torch(rows, d_in, d_out):
    """The relu MLP stacks at the reference's batch size (128) and with the narrow heads (baseline: 1, CMI classifier: 2)
    through mlp_apply vs the same modules in float64; no library GEMM is involved."""
    from mimrl_b200.linear import mlp_apply
    from mimrl_b200.vmi import mlps
    torch.manual_seed(1)
    seq = mlps(d_in, 256, d_out, 2, "relu").cuda()
    for m in seq:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.uniform_(m.bias, -0.1, 0.1)
    import copy
    x = torch.randn(rows, d_in, device="cuda", requires_grad=True)
    w = torch.randn(rows, d_out, device="cuda")
    seq64 = copy.deepcopy(seq).double()
    seq64.zero_grad()
    x64 = x.detach().double().requires_grad_(True)
    # rows with a hidden pre-activation within 1e-5 of zero get no upstream gradient: there the ReLU mask is decided by
    # rounding in ANY fp32 implementation
    h, near = x64, torch.zeros(rows, dtype=torch.bool, device="cuda")
    for m in seq64:
        h = m(h)
        if isinstance(m, torch.nn.Linear) and m is not seq64[-1]:
            near |= (h.abs() < 1e-5).any(dim=1)
    assert float(near.double().mean()) < 0.2
    w = torch.where(near[:, None], torch.zeros_like(w), w)
    y64 = h
    (y64 * w.double()).sum().backward()
    want = [x64.grad] + [p.grad for p in seq64.parameters()]
    y = mlp_apply(seq, x)
    (y * w).sum().backward()
    got = [x.grad.clone()] + [p.grad.clone() for p in seq.parameters()]
    assert rel(y.detach(), y64.detach()) < 1e-5
    for a, b in zip(got, want):
        assert rel(a, b) < 1e-5


def _blocked_operand(mat):
    """[rows, K] fp32 -> the blocked-K fp16 hi/lo operand buffer of mimrl_gemm_split_blocked (header, hi, lo)."""
    import math
    from mimrl_b200 import _lib as L
    rows, K = mat.shape
    amax = float(mat.abs().max())
    _, e = math.frexp(amax)
    v = mat.double() * 2.0 ** (14 - e)
    hi = v.to(torch.float16)
    lo = (v - hi.double()).to(torch.float16)
    blk = lambda t: t.view(rows, K // 64, 64).permute(1, 0, 2).contiguous().view(-1).view(torch.uint8)
    buf = torch.zeros(L.lib.mimrl_split_bytes(rows, K), dtype=torch.uint8, device=mat.device)
    buf[:4] = torch.tensor([amax], dtype=torch.float32, device=mat.device).view(torch.uint8)
    plane = (rows * K * 2 + 255) // 256 * 256
    buf[256:256 + rows * K * 2] = blk(hi)
    buf[256 + plane:256 + plane + rows * K * 2] = blk(lo)
    return buf


@pytest.mark.parametrize("M,N,K", [(10, 50, 64 * 96), (50, 100, 64 * 200), (50, 50, 64 * 7), (128, 128, 64 * 480),
                                   (100, 10, 64 * 33), (3, 128, 64)])
def test_blocked_accumulating_gemm(M, N, K):
    """mimrl_gemm_split_blocked_acc (the CubeMLP weight gradients): narrow operand on the MMA columns, transposed
    output when swapped, one piece of K per SM, partial sums added in place -- against float64, on top of a running sum."""
    from mimrl_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.1
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    op_a, op_b = _blocked_operand(A), _blocked_operand(B)          # (keep both alive across the call)
    L.check(L.lib.mimrl_gemm_split_blocked_acc(L.ptr(op_a), L.ptr(op_b), M, N, K, L.ptr(C), L.stream()))
    torch.cuda.synchronize()
    want = C0.double() + A.double() @ B.double().t()
    assert rel(C, want) < 2e-5
