import sys, torch
sys.path.insert(0, ".")
from mimrl_b200.model import VMIEstimator
dev = "cuda"
for B in (2048, 4096):
    for chunk in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
        torch.manual_seed(0)
        est = VMIEstimator("concat", "constant", "nwj", 128, 256, 128, 2, "relu", 0, 1).to(dev)
        est.critic_model.pair_chunk = chunk
        x = torch.randn(B, 128, device=dev, requires_grad=True); y = torch.randn(B, 128, device=dev, requires_grad=True)
        for it in range(3):
            if it == 1:
                torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
            mi, loss = est(x, y); loss.backward()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"B={B} chunk={chunk} ms={ms:.1f} pairs/s={B*B/ms*1e3:.3e} mem={torch.cuda.max_memory_allocated()/2**30:.1f}GiB", flush=True)
