import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from types import SimpleNamespace
from mimrl_b200.model import MIHeads
from mimrl_b200.train_step import FeaturePool, TwoStageStep
dev = "cuda"
bs, N = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 1284
opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2,
                      radius=1.0, cmi_last_acticate="hardtanh", d_common=128)
heads = MIHeads(opt).to(dev)
enc = torch.nn.Linear(128, 4 * 128).to(dev); cls = torch.nn.Linear(128, 1).to(dev)
def features(batch):
    f = enc(batch).view(-1, 4, 128)
    return cls(f[:, 0]), f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), f[:, 3].contiguous()
main_params = list(enc.parameters()) + list(cls.parameters())
step = TwoStageStep(heads, features, torch.nn.L1Loss(), torch.optim.Adam(main_params, 1e-4),
                    torch.optim.Adam(heads.parameters(), 1e-4), clip_params=main_params + list(heads.parameters()))
pool = FeaturePool()
g = torch.Generator(device="cuda").manual_seed(0)
pool.C = torch.randn(N, 1, device=dev, generator=g).clamp(-3, 3)
pool.F, pool.T, pool.A, pool.V = (torch.randn(N, 128, device=dev, generator=g) for _ in range(4))
batch = torch.randn(bs, 128, device=dev, generator=g); labels = torch.randn(bs, device=dev, generator=g).clamp(-3, 3)
np.random.seed(0)
def one():
    step.stage1(batch, labels, pool); step.stage2(batch, labels, pool)
for _ in range(3): one()
torch.cuda.synchronize(); t = time.time()
for _ in range(5): one()
torch.cuda.synchronize(); print("ms/step", (time.time() - t) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    one(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
ev = prof.key_averages()
print("total cuda ms", sum(e.self_device_time_total for e in ev) / 1e3, "n kernels", sum(e.count for e in ev if e.self_device_time_total > 0))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); one(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
