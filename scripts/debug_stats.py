"""Debug helper: per-row statistics from the sweep kernels vs a float64 torch evaluation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimrl_b200 import _lib as L

dev = torch.device("cuda:0")
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for B in [int(a) for a in sys.argv[2:]] or [20000, 19968, 4096 + 32, 1000]:
    g = torch.Generator(device="cuda").manual_seed(B)
    xe = torch.randn(B, 128, device=dev, generator=g) * 0.5
    ye = 0.7 * xe + 0.3 * torch.randn(B, 128, device=dev, generator=g)
    ws = torch.empty(L.lib.mimrl_sep_workspace_bytes(B, B, 128) + 16, dtype=torch.uint8, device=dev)
    stats = torch.zeros(4, B, device=dev)
    L.check(L.lib.mimrl_sep_row_stats(L.ptr(ye), L.ptr(xe), B, B, 128, 0, 0, impl, L.ptr(stats[0]), L.ptr(stats[1]),
                                      L.ptr(stats[2]), L.ptr(stats[3]), L.ptr(ws), ws.numel(), L.stream()))
    torch.cuda.synchronize()
    lse = torch.empty(B, dtype=torch.float64, device=dev)
    for r0 in range(0, B, 4096):
        S = ye[r0:r0 + 4096].double() @ xe.double().t()
        idx = torch.arange(r0, min(B, r0 + 4096), device=dev)
        S[torch.arange(len(idx), device=dev), idx] = -float("inf")
        lse[r0:r0 + 4096] = torch.logsumexp(S, dim=1)
    got = stats[0].double() + torch.log(stats[1].double())
    err = (got - lse).abs()
    bad = torch.nonzero(err > 1e-4).flatten()
    print(f"B={B} impl={impl}: max lse err {float(err.max()):.3e}, bad rows {bad.numel()}: {bad[:16].tolist()} ... {bad[-4:].tolist()}")
    if bad.numel():
        i = int(bad[0])
        print("   first bad row", i, "got", float(got[i]), "want", float(lse[i]), "m,s", float(stats[0, i]), float(stats[1, i]))
