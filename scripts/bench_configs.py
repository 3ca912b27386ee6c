"""BASELINE.json configs 3 and 4 on N GPUs of one node (run directly for N=1, under torchrun for N>1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_configs.py [concat] [knn]

config 3: concat critic, all-pairs MLP, NWJ / JS, global batch 16384 sharded by row blocks (fwd + bwd + gradient
          all-reduce);  config 4: k-NN sampler on a 1M x 128 pool, keys row-sharded, batch 8192, k in {2, 16}.
Device-timed (CUDA events), max over ranks, one JSON line per measurement from rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimrl_b200 import rowblock as RB                                    # noqa: E402
from mimrl_b200.model import VMIEstimator, knn_search_sharded, prod_knn_sample_sharded   # noqa: E402

which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["concat", "knn"]
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, warm=1, reps=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def out(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


if "concat" in which:
    B = int(os.environ.get("CONCAT_B", "16384"))
    counts = RB.even_split(B, world)
    rb = RB.from_group(counts[rank], device=dev) if world > 1 else None
    g = torch.Generator().manual_seed(0)
    x_all, y_all = torch.randn(B, 128, generator=g), torch.randn(B, 128, generator=g)
    off = sum(counts[:rank])
    for bound in ("nwj", "js"):
        torch.manual_seed(0)
        est = VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1).to(dev)
        est.rowblock = rb
        x = x_all[off: off + counts[rank]].to(dev).requires_grad_(True)
        y = y_all[off: off + counts[rank]].to(dev).requires_grad_(True)
        params = list(est.parameters())

        def step():
            est.zero_grad(set_to_none=True)
            mi, loss = est(x, y)
            loss.backward()
            if rb is not None:
                RB.all_reduce_param_grads(params, rb)
            return mi
        ms = timed(step)
        flops = 1_181_184.0 * B * B          # reference-dense fwd+bwd flops per pair
        out(config=3, component="concat_fwd_bwd", bound=bound, global_batch=B, n_gpus=world, ms=ms,
            pairs_per_s=B * B / ms * 1e3, reference_dense_tflops=flops / ms * 1e-9, mi=float(step().detach()))

if "knn" in which:
    N, width, bs = int(os.environ.get("KNN_N", str(1 << 20))), 128, 8192
    counts = RB.even_split(N, world)
    rb = RB.from_group(counts[rank], device=dev) if world > 1 else RB.single(N)
    off = sum(counts[:rank])
    gen = torch.Generator(device=dev)
    # every rank generates only its own key block (seeded by block so the pool does not depend on N GPUs' RNG order)
    Z = torch.empty(counts[rank], width, device=dev)
    X = torch.empty(counts[rank], width, device=dev)
    blk = 1 << 16
    for b0 in range(off - off % blk, off + counts[rank], blk):
        gen.manual_seed(1000 + b0 // blk)
        zb = torch.randn(blk, width, device=dev, generator=gen)
        xb = torch.randn(blk, width, device=dev, generator=gen)
        lo, hi = max(b0, off), min(b0 + blk, off + counts[rank])
        Z[lo - off: hi - off] = zb[lo - b0: hi - b0]
        X[lo - off: hi - off] = xb[lo - b0: hi - b0]
    Y = Z[:, :1].contiguous()
    for k in (2, 16):
        m = bs // k
        np.random.seed(0)
        ids = torch.from_numpy(np.random.permutation(N)[:m].astype(np.int64)).to(dev)
        ms_search = timed(lambda: knn_search_sharded(Z, ids, k, rb), warm=1, reps=3)

        def sample():
            np.random.seed(0)
            return prod_knn_sample_sharded(X, Y, Z, bs, k, 1.0, rb)
        ms_sample = timed(sample, warm=1, reps=3)
        nbr, _ = knn_search_sharded(Z, ids, k, rb)
        out(config=4, component="knn", n_keys=N, width=width, k=k, queries=m, n_gpus=world, search_ms=ms_search,
            sampler_ms=ms_sample, key_gbs=N * width * 4 / ms_search * 1e-6, queries_per_s=m / ms_search * 1e3,
            checksum=int(nbr.sum().item()))
if world > 1:
    dist.destroy_process_group()
