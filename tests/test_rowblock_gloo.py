"""Row-block sharding plumbing on the gloo backend (world_size 2, CPU): the
same functions the NCCL path uses, minus the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mimrl_b200 import rowblock as RB
        n_local = counts[rank]
        rb = RB.from_group(n_local)
        assert rb.counts == tuple(counts) and rb.rank == rank and rb.world == world
        assert rb.offset == sum(counts[:rank]) and rb.n_all == sum(counts) and rb.sharded
        # every rank builds the same global matrix and keeps its block
        g = torch.Generator().manual_seed(0)
        full = torch.randn(sum(counts), 5, generator=g)
        local = full[rb.offset: rb.offset + n_local].clone()
        gathered = RB.all_gather_rows(local, rb)
        assert torch.equal(gathered, full)
        assert torch.equal(RB.own_slice(gathered, rb), local)
        vec = RB.all_gather_rows(local[:, 0].contiguous(), rb)          # 1-D rows (per-row statistics)
        assert torch.equal(vec, full[:, 0])
        # differentiable gather: the backward sums every rank's gradient of the gathered matrix, own rows kept
        from mimrl_b200.vmi import gather_rows
        loc = local.clone().requires_grad_(True)
        got = gather_rows(loc, rb)
        assert torch.equal(got.detach(), full)
        w = torch.arange(full.numel(), dtype=torch.float32).reshape(full.shape)
        (got * w * (rank + 1)).sum().backward()
        assert torch.equal(loc.grad, RB.own_slice(w * sum(range(1, world + 1)), rb))
        # parameter gradients are partial sums over ranks
        p1 = torch.nn.Parameter(torch.zeros(3, 2))
        p2 = torch.nn.Parameter(torch.zeros(4))
        p3 = torch.nn.Parameter(torch.zeros(2))                          # no grad on purpose
        p1.grad = torch.full((3, 2), float(rank + 1))
        p2.grad = torch.arange(4.0) * (rank + 1)
        RB.all_reduce_param_grads([p1, p2, p3], rb)
        tot = sum(range(1, world + 1))
        assert torch.equal(p1.grad, torch.full((3, 2), float(tot)))
        assert torch.equal(p2.grad, torch.arange(4.0) * tot) and p3.grad is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(4, 4), (5, 3)])
def test_rowblock_world2(counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_helpers():
    from mimrl_b200 import rowblock as RB
    rb = RB.single(7)
    assert not rb.sharded and rb.n_own == 7 and rb.offset == 0
    t = torch.randn(7, 3)
    assert RB.all_gather_rows(t, rb) is t
    assert RB.even_split(10, 4) == (3, 3, 2, 2) and sum(RB.even_split(65536, 8)) == 65536
    assert RB.from_group(5).counts == (5,)


def _knn_merge_worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from mimrl_b200 import rowblock as RB
        from mimrl_b200.model import merge_knn_candidates
        from oracle import knn_oracle as K
        rb = RB.from_group(counts[rank])
        N, width, m, k = sum(counts), 24, 17, 5
        rng = np.random.default_rng(3)
        Z = rng.standard_normal((N, width)).astype(np.float32)
        Z[7] = Z[8]                                                  # exact tie inside one shard
        Z[counts[0] - 1] = Z[counts[0]]                              # exact tie ACROSS the shard boundary
        ids = rng.permutation(N)[:m].astype(np.int64)
        ids[0] = 6                                                   # a query next to both tie pairs
        exc = np.zeros(N, dtype=bool)
        exc[ids] = True
        want, wdist = K.knn(Z, Z[ids], k, exc, "brute")
        # this rank's candidates: the float64 oracle on its own key block (the role of mimrl_knn_search_rows),
        # padded with (-1, inf) when the block has fewer than k reachable keys
        off, n_loc = rb.offset, counts[rank]
        k_loc = min(k, int((~exc[off: off + n_loc]).sum()))
        nbr = np.full((m, k), -1, dtype=np.int64)
        dst = np.full((m, k), np.inf)
        if k_loc > 0:
            loc, ld = K.knn(Z[off: off + n_loc], Z[ids], k_loc, exc[off: off + n_loc], "brute")
            nbr[:, :k_loc], dst[:, :k_loc] = loc + off, ld
        got, gd = merge_knn_candidates(torch.from_numpy(nbr), torch.from_numpy(dst), k, N, rb)
        assert np.array_equal(got.numpy(), want), (got.numpy()[:3], want[:3])
        assert np.allclose(gd.numpy(), wdist, rtol=1e-12, atol=1e-12)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(40, 60), (3, 97)])
def test_sharded_knn_merge_world2(counts):
    """Key-sharded k-NN (BASELINE config 4): merging the per-shard candidates by (distance, index) reproduces the
    global float64 search bit for bit, including exact ties inside a shard and across the shard boundary and a
    shard with fewer than k reachable keys."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_merge_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
