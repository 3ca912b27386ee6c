"""Fused concat-critic kernels against a float64 torch evaluation of the same MLP."""
import sys, torch
sys.path.insert(0, ".")
from mimrl_b200 import _lib as L
torch.manual_seed(0)
dev = "cuda"
def run(n_own, n_all, time_it=False):
    H = 256
    u = torch.randn(n_own, H, device=dev) * 1.3
    v = torch.randn(n_all, H, device=dev) * 0.8
    w2 = torch.randn(H, H, device=dev) / 16; b2 = torch.randn(H, device=dev) * 0.1
    w3 = torch.randn(H, H, device=dev) / 16; b3 = torch.randn(H, device=dev) * 0.1
    w4 = torch.randn(H, device=dev) / 16; b4 = torch.randn(1, device=dev)
    vt = v.t().contiguous()
    out = torch.empty(n_own, n_all, device=dev)
    wsb = L.lib.mimrl_concat_workspace_bytes(H)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    def call():
        L.check(L.lib.mimrl_concat_scores(L.ptr(u), L.ptr(vt), n_own, n_all, n_all, H, L.ptr(w2), L.ptr(b2), L.ptr(w3), L.ptr(b3),
                                          L.ptr(w4), L.ptr(b4), L.ptr(out), L.ptr(ws), wsb, L.stream()))
    call(); torch.cuda.synchronize()
    if n_own * n_all <= 1 << 22:
        d = torch.float64
        h1 = torch.relu(u.to(d)[:, None, :] + v.to(d)[None, :, :])
        h2 = torch.relu(h1 @ w2.to(d).t() + b2.to(d))
        h3 = torch.relu(h2 @ w3.to(d).t() + b3.to(d))
        ref = h3 @ w4.to(d) + b4.to(d)
        err = (out.to(d) - ref).abs().max().item() / ref.abs().max().item()
        print(f"n_own={n_own} n_all={n_all} rel err {err:.3e}", flush=True)
    if time_it:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"n_own={n_own} n_all={n_all} fwd {ms:.3f} ms  {n_own*n_all/ms*1e3:.3e} pairs/s  "
              f"{n_own*n_all*2*2*65536*3/ms*1e-9:.1f} TF/s executed", flush=True)
for n_own, n_all in ((4, 32), (8, 64), (37, 101), (512, 512), (2048, 2048)):
    run(n_own, n_all)
run(4096, 4096, True)

def run_bwd(n_own, n_all, time_it=False, check=True):
    H = 256
    torch.manual_seed(1)
    u = torch.randn(n_own, H, device=dev) * 1.3
    v = torch.randn(n_all, H, device=dev) * 0.8
    w2 = torch.randn(H, H, device=dev) / 16; b2 = torch.randn(H, device=dev) * 0.1
    w3 = torch.randn(H, H, device=dev) / 16; b3 = torch.randn(H, device=dev) * 0.1
    w4 = torch.randn(H, device=dev) / 16
    G = torch.randn(n_own, n_all, device=dev).abs() / (n_own * n_all)     # one sign: no cancellation, so ReLU-kink flips stay O(1/pairs)
    if check:      # drop pairs with a pre-activation on a ReLU kink (their mask is decided by rounding in ANY fp32 code)
        d = torch.float64
        h1 = torch.relu(u.to(d)[:, None, :] + v.to(d)[None, :, :])
        z2 = h1 @ w2.to(d).t() + b2.to(d)
        z3 = torch.relu(z2) @ w3.to(d).t() + b3.to(d)
        near = (z2.abs() < 1e-5).any(-1) | (z3.abs() < 1e-5).any(-1)
        G = torch.where(near, torch.zeros_like(G), G)
        print("   pairs on a kink:", int(near.sum()), "of", near.numel())
        del h1, z2, z3
    vt = v.t().contiguous()
    rows = L.lib.mimrl_concat_pair_rows(n_own, n_all)
    sb = L.lib.mimrl_split_bytes(H, rows)
    ops = [torch.empty(sb, dtype=torch.uint8, device=dev) for _ in range(4)]
    wsb = L.lib.mimrl_concat_workspace_bytes(H)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    gws_b = L.lib.mimrl_gemm_split_workspace_bytes(0, H, H, rows)
    gws = torch.empty(gws_b, dtype=torch.uint8, device=dev)
    def call(parts=(1, 1)):
        g_u = torch.zeros(n_own, H, device=dev); g_vt = torch.zeros(H, n_all, device=dev)
        gb2 = torch.zeros(H, device=dev); gb3 = torch.zeros(H, device=dev); gw4 = torch.zeros(H, device=dev)
        gW2 = torch.empty(H, H, device=dev); gW3 = torch.empty(H, H, device=dev)
        if parts[0]:
            L.check(L.lib.mimrl_concat_grad(L.ptr(u), L.ptr(vt), n_own, n_all, n_all, H, L.ptr(w2), L.ptr(b2), L.ptr(w3), L.ptr(b3),
                                        L.ptr(w4), L.ptr(G), L.ptr(g_u), L.ptr(g_vt), L.ptr(gb2), L.ptr(gb3), L.ptr(gw4),
                                        L.ptr(ops[0]), L.ptr(ops[1]), L.ptr(ops[2]), L.ptr(ops[3]), L.ptr(ws), wsb, L.stream()))
        if not parts[1]:
            return
        L.check(L.lib.mimrl_gemm_split_blocked(L.ptr(ops[2]), L.ptr(ops[0]), H, H, rows, L.ptr(gW2), L.ptr(gws), gws_b, L.stream()))
        L.check(L.lib.mimrl_gemm_split_blocked(L.ptr(ops[3]), L.ptr(ops[1]), H, H, rows, L.ptr(gW3), L.ptr(gws), gws_b, L.stream()))
        return g_u, g_vt, gW2, gb2, gW3, gb3, gw4
    outs = call(); torch.cuda.synchronize()
    if check:
        d = torch.float64
        t = [x.to(d).requires_grad_(True) for x in (u, v, w2, b2, w3, b3, w4)]
        h1 = torch.relu(t[0][:, None, :] + t[1][None, :, :])
        h2 = torch.relu(h1 @ t[2].t() + t[3])
        h3 = torch.relu(h2 @ t[4].t() + t[5])
        sc = h3 @ t[6]
        (sc * G.to(d)).sum().backward()
        refs = [t[0].grad, t[1].grad.t(), t[2].grad, t[3].grad, t[4].grad, t[5].grad, t[6].grad]
        names = ["g_u", "g_vt", "gW2", "gb2", "gW3", "gb3", "gw4"]
        print(f"bwd n_own={n_own} n_all={n_all}: " + "  ".join(
            f"{n} {((o.to(d) - r).abs().max() / r.abs().max()).item():.2e}" for n, o, r in zip(names, outs, refs)), flush=True)
    if time_it:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for parts in ((1, 0), (0, 1)):
            e0.record()
            for _ in range(3): call(parts)
            e1.record(); torch.cuda.synchronize()
            print("   parts", parts, f"{e0.elapsed_time(e1) / 3:.3f} ms")
        e0.record()
        for _ in range(3): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"bwd n_own={n_own} n_all={n_all} {ms:.3f} ms  {n_own*n_all/ms*1e3:.3e} pairs/s", flush=True)
for n_own, n_all in ((4, 32), (37, 101), (256, 512), (1024, 1024)):
    run_bwd(n_own, n_all)
run_bwd(1024, 4096, True, False)
