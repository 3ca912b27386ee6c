"""Two-GPU parity of the row-block sharded estimator: every rank must obtain
the single-GPU global-batch value and its own rows of the gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, bound, B, q, critic="separate", hidden=64):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mimrl_b200 import rowblock as RB
        from mimrl_b200.model import VMIEstimator
        from oracle import params as P
        torch.backends.cuda.matmul.allow_tf32 = False
        baseline = "unnormalized" if bound in ("tuba", "interpolate") else "constant"
        prm = P.vmi_params(3, critic, baseline, 128, hidden, 128, 2)
        x, y = P.features(4, B, 128, corr=0.6)
        counts = (B // 2 + 3, B - B // 2 - 3)                       # ragged shards
        off = sum(counts[:rank])
        est = VMIEstimator(critic, baseline, bound, 128, hidden, 128, 2, "relu", 0, 1).cuda()
        est.load_state_dict({k: torch.tensor(v) for k, v in P.vmi_state_dict(prm).items()})
        est.rowblock = RB.from_group(counts[rank], device=torch.device("cuda", rank))
        xt = torch.tensor(x[off: off + counts[rank]], device="cuda", requires_grad=True)
        yt = torch.tensor(y[off: off + counts[rank]], device="cuda", requires_grad=True)
        mi, loss = est(xt, yt)
        loss.backward()
        params = list(est.parameters())
        RB.all_reduce_param_grads(params, est.rowblock)
        q.put((rank, float(mi.detach()), xt.grad.cpu().numpy(), yt.grad.cpu().numpy(),
               {n: p.grad.cpu().numpy() for n, p in est.named_parameters()}))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bound", ["infonce", "nwj", "tuba", "js", "interpolate"])
def test_sharded_estimator_matches_oracle(bound):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import params as P
    from oracle import vmi_oracle as O
    B = 700
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bound, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) == 5 for r in res), res
    baseline = "unnormalized" if bound in ("tuba", "interpolate") else "constant"
    prm = P.vmi_params(3, "separate", baseline, 128, 64, 128, 2)
    x, y = P.features(4, B, 128, corr=0.6)
    ref = O.vmi_estimator(prm, "separate", baseline, bound, x, y)
    gx = np.concatenate([r[2] for r in res])
    gy = np.concatenate([r[3] for r in res])
    for r in res:
        assert abs(r[1] - ref["mi"]) <= 1e-4 * max(1.0, abs(ref["mi"]))
    assert np.abs(gx - ref["gx"]).max() <= 1e-4 * np.abs(ref["gx"]).max()
    assert np.abs(gy - ref["gy"]).max() <= 1e-4 * np.abs(ref["gy"]).max()
    for k, v in ref["pg"].items():
        if k.endswith("weight"):
            assert np.abs(res[0][4][k] - v).max() <= 2e-4 * np.abs(v).max(), k
            assert np.array_equal(res[0][4][k], res[1][4][k])


@pytest.mark.parametrize("bound", ["nwj", "js", "tuba"])
def test_sharded_concat_estimator_matches_oracle(bound):
    """BASELINE config 3: the concat critic's score matrix sharded by row blocks (each rank scores its x rows
    against the all-gathered y on the fused tensor-core kernels); global-batch value on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import params as P
    from oracle import vmi_oracle as O
    B = 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bound, B, q, "concat", 256)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) == 5 for r in res), res
    baseline = "unnormalized" if bound in ("tuba", "interpolate") else "constant"
    prm = P.vmi_params(3, "concat", baseline, 128, 256, 128, 2)
    x, y = P.features(4, B, 128, corr=0.6)
    ref = O.vmi_estimator(prm, "concat", baseline, bound, x, y)
    gx = np.concatenate([r[2] for r in res])
    gy = np.concatenate([r[3] for r in res])
    for r in res:
        assert abs(r[1] - ref["mi"]) <= 1e-4 * max(1.0, abs(ref["mi"]))
    # 90k pairs x 512 ReLU units: ~20 pre-activations sit within fp32 rounding of a kink, and the mask of such a
    # unit is decided by rounding in any fp32 implementation (the reference included).  Each flip moves one pair's
    # contribution, i.e. O(1/300) of a sum of 90k mixed-sign terms -- hence the looser bounds here; the 1e-4 parity
    # of the kernels themselves is tests/test_gpu_concat.py (kink pairs masked out).
    assert np.abs(gx - ref["gx"]).max() <= 1e-3 * np.abs(ref["gx"]).max()
    assert np.abs(gy - ref["gy"]).max() <= 1e-3 * np.abs(ref["gy"]).max()
    assert np.linalg.norm(gx - ref["gx"]) <= 5e-4 * np.linalg.norm(ref["gx"])
    assert np.linalg.norm(gy - ref["gy"]) <= 5e-4 * np.linalg.norm(ref["gy"])
    for k, v in ref["pg"].items():
        if k.endswith("weight"):
            assert np.abs(res[0][4][k] - v).max() <= 1e-3 * np.abs(v).max(), k
            assert np.linalg.norm(res[0][4][k] - v) <= 5e-4 * np.linalg.norm(v), k
            assert np.array_equal(res[0][4][k], res[1][4][k])


def _knn_worker(rank, world, port, N, width, m, k, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mimrl_b200 import rowblock as RB
        from mimrl_b200.model import knn_search_sharded, prod_knn_sample_sharded
        rng = np.random.default_rng(11)
        Z = rng.standard_normal((N, width)).astype(np.float32)
        Z[N // 3] = Z[N // 3 + 1]                                    # an exact tie across neighbouring rows
        X = rng.standard_normal((N, 128)).astype(np.float32)
        Y = rng.standard_normal((N, 1)).astype(np.float32)
        ids = rng.permutation(N)[:m].astype(np.int64)
        off = sum(counts[:rank])
        sl = slice(off, off + counts[rank])
        rb = RB.from_group(counts[rank], device=torch.device("cuda", rank))
        T = lambda a: torch.tensor(a, device="cuda")
        nbr, comp, d = knn_search_sharded(T(Z[sl]), T(ids), k, rb, return_distance=True)
        np.random.seed(5 + rank)                                     # differently seeded ranks: rank 0's draw wins
        bx, by, bz = prod_knn_sample_sharded(T(X[sl]), T(Y[sl]), T(Z[sl]), m * k, k, 1.0, rb)
        q.put((rank, nbr.cpu().numpy(), comp.cpu().numpy(), d.cpu().numpy(), bx.detach().cpu().numpy(),
               by.detach().cpu().numpy(), bz.detach().cpu().numpy()))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,width,m,k,counts", [(5000, 128, 64, 4, (2500, 2500)), (3001, 128, 33, 2, (3, 2998)),
                                               (4000, 1, 50, 16, (1000, 3000))])
def test_key_sharded_knn_is_bit_identical(N, width, m, k, counts):
    """BASELINE config 4: keys row-sharded over ranks; merged neighbours equal the float64 oracle's (ties -> lowest
    index), and the sampler returns the same replicated batch on both ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mimrl_b200.model import sklearn_route
    from oracle import knn_oracle as K
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_worker, args=(r, 2, port, N, width, m, k, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) == 7 for r in res), res
    rng = np.random.default_rng(11)
    Z = rng.standard_normal((N, width)).astype(np.float32)
    Z[N // 3] = Z[N // 3 + 1]
    X = rng.standard_normal((N, 128)).astype(np.float32)
    Y = rng.standard_normal((N, 1)).astype(np.float32)
    ids = rng.permutation(N)[:m].astype(np.int64)
    exc = np.zeros(N, dtype=bool)
    exc[ids] = True
    want, wdist = K.knn(Z, Z[ids], k, exc, sklearn_route(width, k, N - m))
    for r in res:
        assert np.array_equal(r[1], want)
        assert np.array_equal(r[2], want - np.searchsorted(np.sort(ids), want))
        assert np.allclose(r[3], wdist, rtol=1e-12, atol=1e-12)
    # the sampler: rank 0's draw (seed 5), identical batches on both ranks, rows consistent with the pools
    np.random.seed(5)
    ids0 = np.random.permutation(N)[:m]
    for a, b in zip(res[0][4:], res[1][4:]):
        assert np.array_equal(a, b)
    assert np.array_equal(res[0][6], np.repeat(np.tile(Z[ids0], (1, 128 // width)), k, axis=0))
    assert np.array_equal(res[0][5], np.repeat(np.tile(Y[ids0], (1, 128)), k, axis=0))
    exc0 = np.zeros(N, dtype=bool)
    exc0[ids0] = True
    nb0, _ = K.knn(Z, Z[ids0], k, exc0, sklearn_route(width, k, N - m))
    assert np.array_equal(res[0][4], X[nb0.reshape(-1)])
