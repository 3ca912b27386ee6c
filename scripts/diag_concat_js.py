"""Concat critic, JS / NWJ at batch B: the fused-bound path and the materialising path against a float64 evaluation of the
reference's op chain (VMI.py:58-65, 157-182) on the GPU, per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mimrl_b200.model as M
from mimrl_b200.model import VMIEstimator

torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bound = sys.argv[2] if len(sys.argv) > 2 else "js"
torch.manual_seed(11)
est = VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1).cuda()
with torch.no_grad():
    for n, p in est.named_parameters():
        if n.endswith("bias"):
            p.uniform_(-0.05, 0.05)
g = torch.Generator().manual_seed(12)
x = torch.randn(B, 128, generator=g)
y = 0.6 * x + 0.8 * torch.randn(B, 128, generator=g)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())

f = est.critic_model.MLP_f
names = [n for n, _ in est.named_parameters()]
P64 = [p.detach().double().requires_grad_(True) for p in est.parameters()]
pd = dict(zip(names, P64))
W = [pd[f"critic_model.MLP_f.{i}.weight"] for i in (0, 2, 4, 6)]
bb = [pd[f"critic_model.MLP_f.{i}.bias"] for i in (0, 2, 4, 6)]
xd, yd = x.cuda().double().requires_grad_(True), y.cuda().double().requires_grad_(True)
u = xd @ W[0][:, :128].t() + bb[0]
v = yd @ W[0][:, 128:].t()
h = torch.relu(u[:, None, :] + v[None, :, :])
h = torch.relu(h @ W[1].t() + bb[1])
h = torch.relu(h @ W[2].t() + bb[2])
S = (h @ W[3].t() + bb[3]).reshape(B, B)          # rows x, cols y (VMI.py:65 after the .t())
d = S.diag()
n = float(B)
if bound == "js":
    val = -F.softplus(-d).mean() - (F.softplus(S).sum() - F.softplus(d).sum()) / (n * (n - 1))
else:   # nwj
    Sm = S - 1.0
    off = ~torch.eye(B, dtype=torch.bool, device="cuda")
    val = 1 + (Sm.diag()).mean() - torch.exp(torch.logsumexp(Sm[off], 0) - torch.log(torch.tensor(n * (n - 1), dtype=torch.float64)))
(-val).backward()
for fused in (True, False):
    M.FUSED_CONCAT_BOUND = fused
    est.zero_grad(set_to_none=True)
    xs, ys = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    mi, loss = est(xs, ys)
    loss.backward()
    print(f"B={B} {bound} fused={fused}: gx {rel(xs.grad, xd.grad):.2e} gy {rel(ys.grad, yd.grad):.2e}", end="")
    for nme, p in est.named_parameters():
        if p.grad is not None:
            print(f" | {nme.split('MLP_f.')[-1]} {rel(p.grad, pd[nme].grad):.2e}", end="")
    print()
