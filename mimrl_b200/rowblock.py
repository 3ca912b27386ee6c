"""Row-block sharding of the global batch across ranks.

The reference computes every MI / CMI value over the GLOBAL batch: under
``nn.DataParallel`` the estimator calls sit outside the replicated forward
(Customization.py:99,107; Solver.py:33-35).  Here each rank owns a contiguous
block of rows and all-gathers the small operands it has to sweep against
(critic embeddings, per-row statistics, k-NN candidates), so the global-batch
value is reproduced exactly on every rank.

Only plumbing lives here (no kernels), so it runs unchanged on the gloo
backend in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class RowBlock:
    """Which rows of the global batch this rank owns."""
    rank: int
    world: int
    counts: tuple          # rows per rank
    group: Optional[object] = None

    @property
    def n_own(self) -> int:
        return self.counts[self.rank]

    @property
    def n_all(self) -> int:
        return sum(self.counts)

    @property
    def offset(self) -> int:
        return sum(self.counts[: self.rank])

    @property
    def sharded(self) -> bool:
        return self.world > 1


def single(n: int) -> RowBlock:
    return RowBlock(0, 1, (n,), None)


def from_group(n_local: int, group=None, device=None) -> RowBlock:
    """Build the row-block map by exchanging the local row counts (ragged
    shards are allowed: the last batch of an epoch rarely divides evenly)."""
    if not (dist.is_available() and dist.is_initialized()):
        return single(n_local)
    world = dist.get_world_size(group)
    if world == 1:
        return single(n_local)
    rank = dist.get_rank(group)
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    outs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return RowBlock(rank, world, tuple(int(o.item()) for o in outs), group)


def even_split(n_all: int, world: int) -> tuple:
    """Row counts when a global batch of n_all rows is dealt out in contiguous blocks."""
    base, rem = divmod(n_all, world)
    return tuple(base + (1 if r < rem else 0) for r in range(world))


def all_gather_rows(local: torch.Tensor, rb: RowBlock) -> torch.Tensor:
    """Concatenate every rank's rows in rank order (no autograd; callers own the
    backward).  Ragged shards are padded to the longest block for the collective."""
    if not rb.sharded:
        return local
    local = local.contiguous()
    tail = tuple(local.shape[1:])
    if len(set(rb.counts)) == 1:
        out = local.new_empty((rb.n_all,) + tail)
        dist.all_gather_into_tensor(out, local, group=rb.group)
        return out
    cap = max(rb.counts)
    padded = local.new_zeros((cap,) + tail)
    padded[: rb.n_own] = local
    bufs = [torch.empty_like(padded) for _ in range(rb.world)]
    dist.all_gather(bufs, padded, group=rb.group)
    return torch.cat([b[:c] for b, c in zip(bufs, rb.counts)], dim=0)


def own_slice(full: torch.Tensor, rb: RowBlock) -> torch.Tensor:
    return full[rb.offset: rb.offset + rb.n_own]


def all_reduce_param_grads(params: Sequence[torch.nn.Parameter], rb: RowBlock) -> None:
    """Sum parameter gradients over ranks.  The MI value is a function of the
    global batch and is replicated, so per-rank parameter gradients are partial
    sums (SUM, not mean)."""
    if not rb.sharded:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=rb.group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off: off + n].view_as(g))
        off += n
