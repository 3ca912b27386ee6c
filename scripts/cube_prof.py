import sys, torch
sys.path.insert(0, ".")
from mimrl_b200.mlp_process import MLPEncoder
dev = "cuda"; bs = 1024
torch.manual_seed(0)
enc = MLPEncoder("gelu", [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]], [0.0] * 3, True, False, [True, True]).to(dev)
x = torch.randn(bs, 100, 3, 128, device=dev, requires_grad=True)
for _ in range(2):
    y = enc(x); y.sum().backward()
torch.cuda.synchronize()
