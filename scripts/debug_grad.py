"""Debug helper: embedding-level InfoNCE gradients from the fused path vs float64 torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimrl_b200 import _lib as L
import mimrl_b200.vmi as V

dev = torch.device("cuda:0")
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bound = sys.argv[2] if len(sys.argv) > 2 else "infonce"
for B in [int(a) for a in sys.argv[3:]] or [20000, 4096]:
    g = torch.Generator(device="cuda").manual_seed(B)
    xe = (torch.randn(B, 128, device=dev, generator=g) * 0.3)
    ye = (0.7 * xe + 0.3 * torch.randn(B, 128, device=dev, generator=g))
    a = xe.clone().requires_grad_(True)
    b = ye.clone().requires_grad_(True)
    mi, loss = V.separable_bound(a, b, bound, impl=impl)
    loss.backward()
    # float64 reference, row-block streamed (infonce only)
    xd, yd = xe.double(), ye.double()
    gy = torch.empty_like(yd)
    gx = torch.zeros_like(xd)
    acc = 0.0
    for r0 in range(0, B, 2048):
        S = yd[r0:r0 + 2048] @ xd.t()
        Pm = torch.softmax(S, dim=1)
        idx = torch.arange(r0, min(B, r0 + 2048), device=dev)
        rr = torch.arange(len(idx), device=dev)
        acc += float((S[rr, idx] - torch.logsumexp(S, 1)).sum())
        G = Pm / B
        G[rr, idx] -= 1.0 / B
        gy[r0:r0 + 2048] = G @ xd
        gx += G.t() @ yd[r0:r0 + 2048]
    ref_mi = torch.log(torch.tensor(float(B))).item() + acc / B
    ey = (b.grad.double() - gy).abs()
    ex = (a.grad.double() - gx).abs()
    print(f"B={B} impl={impl} {bound}: mi {float(mi):.7f} ref {ref_mi:.7f} | gy relerr {float(ey.max() / gy.abs().max()):.3e} "
          f"(row {int(ey.max(1).values.argmax())}) gx relerr {float(ex.max() / gx.abs().max()):.3e} (row {int(ex.max(1).values.argmax())})")
    bad = torch.nonzero(ey.max(1).values > 1e-4 * gy.abs().max()).flatten()
    print("   bad gy rows:", bad.numel(), bad[:12].tolist(), bad[-4:].tolist())
