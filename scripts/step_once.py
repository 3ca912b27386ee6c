"""Two warm-up steps and one more two-stage MI/CMI step at batch bs (eager), for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/step_once.py 128"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from mimrl_b200.model import MIHeads
from mimrl_b200.train_step import FeaturePool, TwoStageStep
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = 1284 if bs <= 256 else 16326
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
for _ in range(3):
    step.stage1(batch, labels, pool); step.stage2(batch, labels, pool)
torch.cuda.synchronize()
