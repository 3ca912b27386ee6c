"""Where does the 1e-4 of the concat JS gradient at B = 2048 come from?  (a) scores, (b) dL/dS from the bound,
(c) the pair-MLP backward fed with the float64 dL/dS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mimrl_b200.vmi as V
from mimrl_b200.model import VMIEstimator

torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(11)
est = VMIEstimator("concat", "constant", "js", 128, 256, 128, 2, "relu", 0, 1).cuda()
with torch.no_grad():
    for n, p in est.named_parameters():
        if n.endswith("bias"):
            p.uniform_(-0.05, 0.05)
g = torch.Generator().manual_seed(12)
x = torch.randn(B, 128, generator=g)
y = 0.6 * x + 0.8 * torch.randn(B, 128, generator=g)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())

f = est.critic_model.MLP_f
W = [m.weight.detach().double() for m in (f[0], f[2], f[4], f[6])]
bb = [m.bias.detach().double() for m in (f[0], f[2], f[4], f[6])]
xd, yd = x.cuda().double().requires_grad_(True), y.cuda().double().requires_grad_(True)
u = xd @ W[0][:, :128].t() + bb[0]
v = yd @ W[0][:, 128:].t()
h = torch.relu(u[:, None, :] + v[None, :, :])
h = torch.relu(h @ W[1].t() + bb[1])
h = torch.relu(h @ W[2].t() + bb[2])
S64 = (h @ W[3].t() + bb[3]).reshape(B, B)          # rows x, cols y
S64 = S64.t()                                        # VMI.py:65 returns scores.t(): rows y? keep the module's orientation below
xs, ys = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
S = est.critic_model(xs, ys)
print("scores orientation/equality  rel(S, S64) =", rel(S, S64), " rel(S, S64.t()) =", rel(S, S64.t()))
Sref = S64 if rel(S, S64) < rel(S, S64.t()) else S64.t()
# (b) bound gradient
S_leaf = S.detach().clone().requires_grad_(True)
mi = V.js_lower_bound(S_leaf)
(-mi).backward()
Sd = Sref.detach().clone().requires_grad_(True)
d = Sd.diag()
first = -F.softplus(-d).mean()
second = (F.softplus(Sd).sum() - F.softplus(d).sum()) / (B * (B - 1.0))
(-(first - second)).backward()
print("dL/dS: rel =", rel(S_leaf.grad, Sd.grad), " offdiag-only rel =",
      rel(S_leaf.grad - torch.diag(S_leaf.grad.diag()), Sd.grad - torch.diag(Sd.grad.diag())))
# (c) pair-MLP backward with the float64 gradient
G = Sd.grad.float()
S.backward(G)
Sref.backward(Sd.grad)
print("grad x rel =", rel(xs.grad, xd.grad), " grad y rel =", rel(ys.grad, yd.grad))
# (d) same with the diagonal of G removed / only the diagonal
for name, Gm in (("offdiag", Sd.grad - torch.diag(Sd.grad.diag())), ("diag", torch.diag(Sd.grad.diag()))):
    xs2, ys2 = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    est.critic_model(xs2, ys2).backward(Gm.float())
    xd2, yd2 = x.cuda().double().requires_grad_(True), y.cuda().double().requires_grad_(True)
    u = xd2 @ W[0][:, :128].t() + bb[0]
    v = yd2 @ W[0][:, 128:].t()
    h = torch.relu(u[:, None, :] + v[None, :, :])
    h = torch.relu(h @ W[1].t() + bb[1])
    h = torch.relu(h @ W[2].t() + bb[2])
    S2 = (h @ W[3].t() + bb[3]).reshape(B, B)
    S2 = S2 if rel(S, S64) < rel(S, S64.t()) else S2.t()
    S2.backward(Gm)
    print(name, "grad x rel =", rel(xs2.grad, xd2.grad), " grad y rel =", rel(ys2.grad, yd2.grad))
