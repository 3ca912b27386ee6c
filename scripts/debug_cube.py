import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import params as P, cubemlp_oracle as C
import mimrl_b200.mlp_process as MP
from test_gpu_cubemlp import build
bs, d_in, d_h, d_out, act, res = 9, [40, 4, 72], [[24, 4, 100]], [[40, 4, 72]], "relu", False
c = dict(act=act, d_in=d_in, d_hiddens=d_h, d_outs=d_out, bias=True, ln_first=False, res=[res])
blocks = P.cubemlp_params(91, d_in, d_h, d_out, True, False, [res])
x = P.features(92, bs * d_in[0] * d_in[1], d_in[2]).reshape(bs, *d_in)
oshape = (bs, *d_out[-1])
w = P.features(93, int(np.prod(oshape[:-1])), oshape[-1]).reshape(oshape)
yo, caches = C.encoder_forward(blocks, x, act, False, [res])
gxo, pgo = C.encoder_backward(caches, w.astype(np.float64), False, [res])
for tc in (True, False):
    MP.USE_TC = tc
    enc = build(c, blocks)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = enc(xt)
    (y * torch.tensor(w, device="cuda")).sum().backward()
    print("TC", tc, "y", np.abs(y.detach().cpu().numpy() - yo).max() / np.abs(yo).max(),
          "gx", np.abs(xt.grad.cpu().numpy() - gxo).max() / np.abs(gxo).max())
    for n, p in enc.named_parameters():
        e = np.abs(p.grad.cpu().numpy() - pgo[n])
        print("   ", n, e.max() / np.abs(pgo[n]).max(), int(e.argmax()))
