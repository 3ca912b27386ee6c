"""Concat critic at batch B: ours vs the reference in float64, next to the reference's own fp32 run vs its float64 run
(how well conditioned is the gradient in fp32 at all?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_shim as R
import mimrl_b200.model as M
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ref = R.import_reference(cpu=False, random_bert=False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bound = sys.argv[2] if len(sys.argv) > 2 else "js"
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
torch.manual_seed(seed)
theirs = ref.Model.VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1)
with torch.no_grad():
    for n, p in theirs.named_parameters():
        if n.endswith("bias"):
            p.uniform_(-0.05, 0.05)
ours = M.VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1)
ours.load_state_dict(theirs.state_dict(), strict=True)
ours = ours.cuda()
g = torch.Generator().manual_seed(12)
x = torch.randn(B, 128, generator=g)
y = 0.6 * x + 0.8 * torch.randn(B, 128, generator=g)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())


def run(est, dtype):
    est.zero_grad(set_to_none=True)
    xt, yt = x.to("cuda", dtype).requires_grad_(True), y.to("cuda", dtype).requires_grad_(True)
    mi, loss = est(xt, yt)
    loss.backward()
    return float(mi), xt.grad.double(), yt.grad.double(), {n: p.grad.double() for n, p in est.named_parameters() if p.grad is not None}


r64 = run(theirs.to("cuda", torch.float64), torch.float64)
r32 = run(theirs.to(torch.float32), torch.float32)
for name, fused in (("ours fused", True), ("ours materialised", False)):
    M.FUSED_CONCAT_BOUND = fused
    o = run(ours, torch.float32)
    print(f"B={B} {bound} {name}: mi {o[0]:.6f} vs {r64[0]:.6f} | gx {rel(o[1], r64[1]):.2e} gy {rel(o[2], r64[2]):.2e}",
          " ".join(f"{k.split('MLP_f.')[-1]} {rel(v, r64[3][k]):.1e}" for k, v in o[3].items()))
print(f"B={B} {bound} reference fp32 vs its fp64: mi {r32[0]:.6f} | gx {rel(r32[1], r64[1]):.2e} gy {rel(r32[2], r64[2]):.2e}",
      " ".join(f"{k.split('MLP_f.')[-1]} {rel(v, r64[3][k]):.1e}" for k, v in r32[3].items()))
