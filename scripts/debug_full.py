"""Debug helper: full-estimator input gradients, ours vs torch fp32 (reference-style, materialised) vs torch fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import params as P
from mimrl_b200.model import VMIEstimator
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
prm = P.vmi_params(99, "separate", "constant", 128, 256, 128, 2)
x, y = P.features(100, B, 128, corr=0.6)
est = VMIEstimator("separate", "constant", "infonce", 128, 256, 128, 2, "relu", 0, 1)
est.load_state_dict({k: torch.tensor(v) for k, v in P.vmi_state_dict(prm).items()})
est = est.to(dev)

def ref(dtype):
    e = est.to(dtype)
    xt = torch.tensor(x, device=dev, dtype=dtype, requires_grad=True)
    yt = torch.tensor(y, device=dev, dtype=dtype, requires_grad=True)
    x_, y_ = e.critic_model.embed(xt, yt)
    S = y_ @ x_.t()
    mi = torch.log(torch.tensor(float(B), device=dev, dtype=dtype)) + (S.diag() - torch.logsumexp(S, 1)).mean()
    (-mi).backward()
    out = (float(mi), xt.grad.double().cpu().numpy(), yt.grad.double().cpu().numpy(), x_.detach().double(), y_.detach().double())
    est.to(torch.float32)
    return out

mi64, gx64, gy64, xe, ye = ref(torch.float64)
mi32, gx32, gy32, _, _ = ref(torch.float32)
xt = torch.tensor(x, device=dev, requires_grad=True)
yt = torch.tensor(y, device=dev, requires_grad=True)
mi, loss = est(xt, yt)
loss.backward()
gx, gy = xt.grad.double().cpu().numpy(), yt.grad.double().cpu().numpy()
re = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
print(f"B={B} mi ours {float(mi):.7f} torch32 {mi32:.7f} torch64 {mi64:.7f}")
print(f"  gx: ours-vs-64 {re(gx, gx64):.3e}  torch32-vs-64 {re(gx32, gx64):.3e}")
print(f"  gy: ours-vs-64 {re(gy, gy64):.3e}  torch32-vs-64 {re(gy32, gy64):.3e}")
print("  embedding stats: |mean x_| %.3e  std x_ %.3e |mean y_| %.3e std y_ %.3e" % (
    float(xe.mean(0).norm()), float(xe.std(0).norm()), float(ye.mean(0).norm()), float(ye.std(0).norm())))
d = np.abs(gy - gy64)
i = np.unravel_index(d.argmax(), d.shape)
print("  worst gy entry", i, "ours", gy[i], "ref64", gy64[i], "torch32", gy32[i], " max|gy64|", np.abs(gy64).max())
print("  rows with err > 1e-4*max:", int((d.max(1) > 1e-4 * np.abs(gy64).max()).sum()))
