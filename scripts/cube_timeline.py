"""GPU timeline (kernel start / duration / gap) of one graphed CubeMLP forward and forward+backward replay."""
import sys, torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from mimrl_b200.mlp_process import MLPEncoder
from mimrl_b200.graphs import GraphedCallable
dev = "cuda"; bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
enc = MLPEncoder("gelu", [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]], [0.0] * 3, True, False, [True, True]).to(dev)
x = torch.randn(bs, 100, 3, 128, device=dev)
params = list(enc.parameters())
def fwd_only(xx):
    with torch.no_grad():
        return enc(xx)
gf = GraphedCallable(fwd_only, [x])
xs = x.clone().requires_grad_(True)
def fwd_bwd(xx):
    xs.grad = None
    for q in params: q.grad = None
    enc(xs).sum().backward()
    return xs.grad
gb = GraphedCallable(fwd_bwd, [xs.detach()])
for name, g, arg in (("forward", gf, gf.static_in[0]), ("forward+backward", gb, gb.static_in[0])):
    for _ in range(3): g(arg)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        g(arg); torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start; prev_end = t0; busy = 0
    print(f"== {name}: {len(ev)} GPU activities")
    for e in ev:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print(f"{s:9.1f} us  +{d:7.1f}  gap {e.time_range.start - prev_end:6.1f}  {e.name[:70]}")
        prev_end = e.time_range.end; busy += d
    print(f"span {prev_end - t0:.1f} us, busy {busy:.1f} us")
