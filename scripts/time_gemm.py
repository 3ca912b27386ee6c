import sys, torch
sys.path.insert(0, ".")
from mimrl_b200 import _lib as L
dev = "cuda"
def t(fn, reps=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def split(x):
    r, c = x.shape
    buf = torch.empty(L.lib.mimrl_split_bytes(r, c), dtype=torch.uint8, device=dev)
    L.check(L.lib.mimrl_split_f32(L.ptr(x), None, r, c, L.ptr(buf), None, L.stream()))
    return buf
for mode, M, N, K in ((0, 65536, 256, 256), (0, 65536, 256, 128), (1, 65536, 256, 256), (2, 256, 256, 65536), (0, 256, 256, 1 << 21),
                      (0, 8192, 256, 256), (0, 1024, 256, 384)):
    A = torch.randn((M, K) if mode != 2 else (K, M), device=dev); B = torch.randn((N, K) if mode == 0 else (K, N), device=dev)
    a, b = split(A), split(B)
    C = torch.empty(M, N, device=dev)
    wsb = L.lib.mimrl_gemm_split_workspace_bytes(mode, M, N, K); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = t(lambda: L.check(L.lib.mimrl_gemm_split(mode, L.ptr(a), L.ptr(b), M, N, K, None, 0, L.ptr(C), L.ptr(ws), wsb, L.stream())))
    ref = (A.double() @ B.double().t()) if mode == 0 else ((A.double() @ B.double()) if mode == 1 else (A.double().t() @ B.double()))
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item() if M * N * K < 1 << 34 else float("nan")
    print(f"mode {mode} M={M} N={N} K={K}: {ms*1e3:.1f} us  {2.0*M*N*K*3/ms*1e-9:.0f} TF/s executed  err {err:.1e}", flush=True)
