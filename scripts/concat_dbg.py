import sys, torch, os
sys.path.insert(0, ".")
from mimrl_b200 import _lib as L
dev="cuda"; H=256; n_own, n_all = 1024, 4096
torch.manual_seed(1)
u = torch.randn(n_own, H, device=dev); v = torch.randn(n_all, H, device=dev)
w2 = torch.randn(H, H, device=dev) / 16; b2 = torch.randn(H, device=dev) * 0.1
w3 = torch.randn(H, H, device=dev) / 16; b3 = torch.randn(H, device=dev) * 0.1
w4 = torch.randn(H, device=dev) / 16
G = torch.randn(n_own, n_all, device=dev).abs()
vt = v.t().contiguous()
rows = L.lib.mimrl_concat_pair_rows(n_own, n_all); sb = L.lib.mimrl_split_bytes(H, rows)
ops = [torch.empty(sb, dtype=torch.uint8, device=dev) for _ in range(4)]
wsb = L.lib.mimrl_concat_workspace_bytes(H); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
g_u = torch.zeros(n_own, H, device=dev); g_vt = torch.zeros(H, n_all, device=dev)
gb = [torch.zeros(H, device=dev) for _ in range(3)]
def call():
    L.check(L.lib.mimrl_concat_grad(L.ptr(u), L.ptr(vt), n_own, n_all, n_all, H, L.ptr(w2), L.ptr(b2), L.ptr(w3), L.ptr(b3),
        L.ptr(w4), L.ptr(G), L.ptr(g_u), L.ptr(g_vt), L.ptr(gb[0]), L.ptr(gb[1]), L.ptr(gb[2]),
        L.ptr(ops[0]), L.ptr(ops[1]), L.ptr(ops[2]), L.ptr(ops[3]), L.ptr(ws), wsb, L.stream()))
for dbg in (0, 1, 2, 3, 7):
    os.environ["MIMRL_CONCAT_DBG"] = str(dbg)
    call(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): call()
    e1.record(); torch.cuda.synchronize()
    print("dbg", dbg, f"{e0.elapsed_time(e1)/3:.3f} ms", flush=True)
