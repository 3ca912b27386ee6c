"""One forward and one backward of the fused concat critic (for ncu captures)."""
import sys, torch
sys.path.insert(0, ".")
from mimrl_b200.vmi import _ConcatPairMLP
torch.manual_seed(0)
H, n_own, n_all = 256, 2048, 4096
g = lambda *s: torch.randn(*s, device="cuda")
u, v = g(n_own, H).requires_grad_(True), g(n_all, H).requires_grad_(True)
w2, w3 = (g(H, H) / 16).requires_grad_(True), (g(H, H) / 16).requires_grad_(True)
b2, b3 = (g(H) * 0.1).requires_grad_(True), (g(H) * 0.1).requires_grad_(True)
w4, b4 = (g(1, H) / 16).requires_grad_(True), g(1).requires_grad_(True)
for _ in range(2):
    s = _ConcatPairMLP.apply(u, v, w2, b2, w3, b3, w4, b4)
    (s * s).mean().backward()
torch.cuda.synchronize()
