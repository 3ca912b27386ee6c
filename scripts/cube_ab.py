"""CubeMLP at the README configuration against the reference's own MLPEncoder in float64 on the GPU: errors of the
specialised kernels (cubemlp_tc2 / _tc3) and of the general kernel (MIMRL_CUBE2_OFF / MIMRL_CUBE3_OFF) -- usage:
cube_ab.py [bs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.backends.cuda.matmul.allow_tf32 = False
import __graft_entry__ as g; g.build()
from oracle import ref_shim as R
from mimrl_b200.mlp_process import MLPEncoder
ref = R.import_reference(cpu=False, random_bert=False)
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 131
cfg = dict(activate="gelu", d_in=[100, 3, 128], d_hiddens=[[50, 3, 128], [10, 3, 128]], d_outs=[[50, 3, 128], [10, 3, 128]],
           dropouts=[0.0] * 3, bias=True, ln_first=False, res_project=[True, True])
torch.manual_seed(3)
enc = MLPEncoder(**cfg).cuda()
for n, p in enc.named_parameters():            # non-trivial LayerNorm parameters and biases
    if "ln_" in n or "bias" in n:
        p.data.add_(0.3 * torch.randn_like(p))
renc = ref.MLPProcess.MLPEncoder(**cfg).cuda().double()
renc.load_state_dict({k: v.double() for k, v in enc.state_dict().items()})
gen = torch.Generator(device="cuda").manual_seed(bs)
x = torch.randn(bs, 100, 3, 128, device="cuda", generator=gen)
w = torch.randn(bs, 10, 3, 128, device="cuda", generator=gen)
def run(m, xx, ww):
    for p in m.parameters(): p.grad = None
    xt = xx.clone().requires_grad_(True)
    y = m(xt); (y * ww).sum().backward()
    return y.detach(), xt.grad, {n: p.grad.clone() for n, p in m.named_parameters()}
y64, gx64, pg64 = run(renc, x.double(), w.double())
rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
for name, env in (("specialised", {}), ("general", {"MIMRL_CUBE2_OFF": "1", "MIMRL_CUBE3_OFF": "1"})):
    for k in ("MIMRL_CUBE2_OFF", "MIMRL_CUBE3_OFF"): os.environ.pop(k, None)
    os.environ.update(env)
    y, gx, pg = run(enc, x, w)
    worst = max((rel(pg[n], pg64[n]), n) for n in pg)
    print(f"{name:12s} bs={bs}: y {rel(y, y64):.2e}  gx {rel(gx, gx64):.2e}  worst param grad {worst[0]:.2e} ({worst[1]})")
renc32 = ref.MLPProcess.MLPEncoder(**cfg).cuda(); renc32.load_state_dict(enc.state_dict())
y, gx, pg = run(renc32, x, w)
worst = max((rel(pg[n], pg64[n]), n) for n in pg)
print(f"{'torch fp32':12s} bs={bs}: y {rel(y, y64):.2e}  gx {rel(gx, gx64):.2e}  worst param grad {worst[0]:.2e} ({worst[1]})")
