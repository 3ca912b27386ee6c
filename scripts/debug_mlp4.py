import sys, torch
sys.path.insert(0, ".")
import mimrl_b200.linear as LN
torch.manual_seed(0)
dev = "cuda"
def mk(d_in, d_out):
    m = torch.nn.Sequential(torch.nn.Linear(d_in, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                            torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, d_out)).to(dev)
    for p in m.parameters():
        if p.dim() == 1: torch.nn.init.normal_(p, std=0.1)
    return m
for M, d_in, d_out in ((512, 128, 128), (700, 128, 128), (1000, 64, 32), (5000, 100, 96), (65536, 128, 128)):
    m = mk(d_in, d_out)
    x = torch.randn(M, d_in, device=dev)
    w = torch.randn(M, d_out, device=dev).abs()
    res = {}
    for fused in (True, False):
        LN.USE_FUSED_MLP = fused
        m.zero_grad()
        xt = x.clone().requires_grad_(True)
        y = LN.mlp_apply(m, xt)
        (y * w).sum().backward()
        res[fused] = (y.detach(), xt.grad, [p.grad.clone() for p in m.parameters()])
    m64 = torch.nn.Sequential(*[torch.nn.Linear(l.in_features, l.out_features).double().to(dev) if isinstance(l, torch.nn.Linear) else torch.nn.ReLU() for l in m])
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    xt = x.double().requires_grad_(True)
    y64 = m64(xt); (y64 * w.double()).sum().backward()
    ref = (y64.detach(), xt.grad, [p.grad for p in m64.parameters()])
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    for fused in (True, False):
        r = res[fused]
        print(f"M={M} {d_in}->{d_out} fused={fused}: y {rel(r[0], ref[0]):.1e} gx {rel(r[1], ref[1]):.1e} params " +
              " ".join(f"{rel(a, b):.1e}" for a, b in zip(r[2], ref[2])), flush=True)
    if M >= 5000:
        for fused in (True, False):
            LN.USE_FUSED_MLP = fused
            def fb():
                m.zero_grad(); xt = x.clone().requires_grad_(True); (LN.mlp_apply(m, xt) * w).sum().backward()
            def f():
                with torch.no_grad(): LN.mlp_apply(m, x)
            for fn, name in ((f, "fwd"), (fb, "fwd+bwd")):
                for _ in range(3): fn()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(5): fn()
                e1.record(); torch.cuda.synchronize()
                print(f"   M={M} fused={fused} {name}: {e0.elapsed_time(e1)/5*1e3:.0f} us", flush=True)
