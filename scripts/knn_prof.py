"""One k-NN search at BASELINE config-4 scale for ncu (usage: knn_prof.py [queries] [k])."""
import sys, torch
sys.path.insert(0, ".")
from mimrl_b200.model import knn_search
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N = 1 << 20
g = torch.Generator(device="cuda").manual_seed(2)
Z = torch.randn(N, 128, device="cuda", generator=g)
ids = torch.randperm(N, device="cuda", generator=g)[:m]
for _ in range(2):
    knn_search(Z, ids, k)
torch.cuda.synchronize()
if len(sys.argv) > 3:          # also one search of the width-1 label pool (sorted route, knn_1d.cu)
    Zl = torch.randn(N, 1, device="cuda", generator=g)
    for _ in range(2):
        knn_search(Zl, ids, k)
    torch.cuda.synchronize()
