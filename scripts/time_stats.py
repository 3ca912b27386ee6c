"""Debug helper: time the stats kernel alone under MIMRL_TC_DEBUG variants (set before import)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimrl_b200 import _lib as L
dev = torch.device("cuda:0")
B = 65536
g = torch.Generator(device="cuda").manual_seed(0)
xe = torch.randn(B, 128, device=dev, generator=g) * 0.3
ye = 0.7 * xe + 0.3 * torch.randn(B, 128, device=dev, generator=g)
ws = torch.empty(L.lib.mimrl_sep_workspace_bytes(B, B, 128) + 16, dtype=torch.uint8, device=dev)
stats = torch.zeros(4, B, device=dev)
def run():
    L.check(L.lib.mimrl_sep_row_stats(L.ptr(ye), L.ptr(xe), B, B, 128, 0, 0, 2, L.ptr(stats[0]), L.ptr(stats[1]),
                                      L.ptr(stats[2]), None, L.ptr(ws), ws.numel(), L.stream()))
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("MIMRL_TC_DEBUG=%s  stats call (prepass+kernel+combine): min %.3f ms  median %.3f ms" % (os.environ.get("MIMRL_TC_DEBUG", "0"), min(ts), sorted(ts)[len(ts)//2]))
