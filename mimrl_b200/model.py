"""MI / CMI estimator modules and the k-NN sampler: drop-in for the hot-path
part of the reference's ``Model.py`` (lines 47-225 and the two stage functions
305-386).  Same class names, constructor arguments, ``forward`` signatures and
parameter paths (``critic_model.MLP_g.0.weight``, ``classifier.mlp.0.weight``,
...), so reference ``state_dict``s load and ``Solver.get_optimizer``'s
substring grouping (Solver.py:124-133) keeps working.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import rowblock as RB
from .linear import mlp_apply
from .vmi import (BaselineModel, CriticModel, _scores_bound, concat_bound, gather_rows, get_activation,
                  interp_lower_bound, separable_bound, separable_interp_bound)

FUSED_CONCAT_BOUND = True      # concat critic: bound fused into the all-pairs kernels (no B x B tensor); False = materialise


# --------------------------------------------------------------------------
# k-NN sampler (Model.py:75-106)
# --------------------------------------------------------------------------


def sklearn_route(width, k, n_fit):
    """Which algorithm scikit-learn's 'auto' picks (sklearn/neighbors/_base.py:615-648):
    decides the float64 distance form the re-rank must reproduce."""
    return "brute" if (width > 15 or k >= n_fit // 2) else "kd_tree"


class KnnPool:
    """A key pool fitted once for many searches (``mimrl_knn_fit``): squared norms and the fp16 hi / lo planes the
    tensor-core filter streams, 0.27 ms per 1M x 128 pool that every ``knn_search(Z, ...)`` on the bare tensor pays again.
    The reference refits per call (Model.py:82-85) because it removes the drawn rows from the pool first; here those rows
    are masked per search, so the fit of the whole pool is reusable until its CONTENTS change -- the caller's contract:
    build a new KnnPool when the pool tensor is rewritten (once per epoch in the reference's training loop).
    ``knn_search`` and ``prod_knn_sample`` accept a KnnPool wherever they accept the pool tensor; results are
    bit-identical.  Pools without a preparation pass (width <= 15, fewer than 2048 rows) are simply wrapped."""

    def __init__(self, Z):
        self.Z = L.f32(Z.detach())
        N, width = self.Z.shape
        nbytes = L.lib.mimrl_knn_fit_bytes(N, width)
        self.fitted = None
        if nbytes:
            self.fitted = torch.empty(nbytes, dtype=torch.uint8, device=self.Z.device)
            L.check(L.lib.mimrl_knn_fit(L.ptr(self.Z), N, width, L.ptr(self.fitted), nbytes, L.stream()))

    shape = property(lambda self: self.Z.shape)

    def detach(self):
        return self


def _pool_tensor(P):
    return P.Z if isinstance(P, KnnPool) else P


def knn_search(Z, ids, k, radius=1.0, return_distance=False):
    """Neighbours of ``Z[ids]`` among the rows of ``Z`` not in ``ids``.  ``Z``: the pool tensor or a ``KnnPool``.
    Returns (nbr_orig, nbr_comp[, dist]) int64 [m, k] CUDA tensors."""
    fitted = Z.fitted if isinstance(Z, KnnPool) else None
    Z = L.f32(_pool_tensor(Z))
    N, width = Z.shape
    m = int(ids.numel())
    ids = ids.to(device=Z.device, dtype=torch.int64).contiguous()
    ws_bytes = L.lib.mimrl_knn_workspace_bytes(N, m, width, k)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=Z.device)
    nbr_orig = torch.empty(m, k, dtype=torch.int64, device=Z.device)
    nbr_comp = torch.empty(m, k, dtype=torch.int64, device=Z.device)
    dist = torch.empty(m, k, dtype=torch.float64, device=Z.device) if return_distance else None
    exact = 1 if sklearn_route(width, k, N - m) == "brute" else 0
    if fitted is not None:
        rc = L.lib.mimrl_knn_search_fitted(L.ptr(Z), N, width, L.ptr(fitted), fitted.numel(), L.ptr(ids), m, k, float(radius),
                                           exact, L.ptr(nbr_orig), L.ptr(nbr_comp), L.ptr(dist), L.ptr(ws), ws.numel(),
                                           L.stream())
    else:
        rc = L.lib.mimrl_knn_search(L.ptr(Z), N, width, L.ptr(ids), m, k, float(radius), exact, L.ptr(nbr_orig),
                                    L.ptr(nbr_comp), L.ptr(dist), L.ptr(ws), ws.numel(), L.stream())
    if rc == 2 and b"n_neighbors" in L.lib.mimrl_last_error():
        raise ValueError(L.lib.mimrl_last_error().decode())       # sklearn raises ValueError here
    L.check(rc)
    return (nbr_orig, nbr_comp, dist) if return_distance else (nbr_orig, nbr_comp)


def _gather(src, idx, repeat, out_width):
    src = L.f32(src)
    out = torch.empty(idx.numel() * repeat, out_width, dtype=torch.float32, device=src.device)
    L.check(L.lib.mimrl_gather_rows(L.ptr(src), src.shape[0], src.shape[1], L.ptr(idx), idx.numel(), repeat,
                                    out_width, L.ptr(out), L.stream()))
    return out


_ID_SOURCE = None          # set by graphs.GraphedCallable while it warms up / captures a step


def legacy_permutation_head(N, m, _min_n=4096):
    """``np.random.permutation(N)[:m]`` on numpy's GLOBAL legacy generator -- same values, same generator state afterwards
    (what Model.py:81's ``np.random.choice(range(N), size=m, replace=False)`` draws and leaves behind) -- computed by
    ``mimrl_legacy_permutation_head`` (csrc/host_rng.cu) on a copy of the state, which is then written back.  The
    Fisher-Yates pass over all N elements is inherent to stream parity; this one runs 4.5x faster than numpy's at N = 2^20."""
    import ctypes
    st = np.random.get_state()
    if st[0] != "MT19937" or N < _min_n:          # small pools: numpy's own call is cheaper than the state round trip
        return np.random.permutation(N)[:m].astype(np.int64)
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = ctypes.c_int(int(st[2]))
    out = np.empty(m, np.int64)
    L.check(L.lib.mimrl_legacy_permutation_head(key.ctypes.data, ctypes.addressof(pos), N, m, out.ctypes.data))
    np.random.set_state((st[0], key, pos.value, st[3], st[4]))
    return out


def prod_knn_sample(X, Y, Z, batch_size, k_neighbor, radius):
    """Model.py:75-106.  Draws m = batch_size // k query rows with numpy's GLOBAL
    RNG (same draw and same RNG state afterwards as the reference's
    ``np.random.choice(range(N), m, replace=False)``), finds the k nearest
    neighbours of Z[ids] among the remaining rows, and returns
    ``(X[neighbours], Y[ids] repeated k, Z[ids] repeated k)`` tiled to the widest
    of the three, as new CUDA tensors that require grad and carry no history to
    the pools.  The pools stay on the GPU; only the m ids travel host -> device."""
    pool = Z                                  # tensor or KnnPool (fitted keys)
    X, Y, Z = _pool_tensor(X).detach(), _pool_tensor(Y).detach(), _pool_tensor(Z).detach()
    N = X.shape[0]
    m = batch_size // k_neighbor
    if m > N:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    if _ID_SOURCE is not None:            # CUDA-graph capture (graphs.py): same draw, staged through a pinned buffer
        ids = _ID_SOURCE.next(N, m, Z.device)
    else:
        ids = torch.from_numpy(legacy_permutation_head(N, m)).to(Z.device, non_blocking=True)
    nbr_orig, _ = knn_search(pool if isinstance(pool, KnnPool) else Z, ids, k_neighbor, radius)
    wmax = max(X.shape[1], Y.shape[1], Z.shape[1])
    batch_x = _gather(X, nbr_orig.reshape(-1), 1, wmax)
    batch_y = _gather(Y, ids, k_neighbor, wmax)
    batch_z = _gather(Z, ids, k_neighbor, wmax)
    return batch_x.requires_grad_(True), batch_y.requires_grad_(True), batch_z.requires_grad_(True)


def _sum_over_ranks(t, rb):
    if rb.sharded:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=rb.group)
    return t


def _rows_from_shards(local, ids, rb):
    """rows ``ids`` (global row numbers) of a matrix whose row blocks live on different ranks: every rank
    contributes the rows it owns, one all-reduce assembles them everywhere."""
    local = L.f32(local)
    if local.shape[0] == 0:
        return _sum_over_ranks(torch.zeros(ids.numel(), local.shape[1], dtype=torch.float32, device=local.device), rb)
    mine = (ids >= rb.offset) & (ids < rb.offset + rb.n_own)
    # (clamped gather times the ownership mask: no boolean indexing, i.e. no device-to-host sync)
    out = local[(ids - rb.offset).clamp(0, local.shape[0] - 1)] * mine.unsqueeze(1).to(torch.float32)
    return _sum_over_ranks(out, rb)


def merge_knn_candidates(nbr, dist, k, n_total, rb):
    """All-gather every rank's k candidates per query ([m, k] global indices, -1 = none, and float64 distances) and
    keep the k best by (distance, index) -- the total order of the single-GPU search.  Pure torch + the process
    group, so it runs on gloo in the CPU tests."""
    if not rb.sharded:
        return nbr, dist
    all_nbr = [torch.empty_like(nbr) for _ in range(rb.world)]
    all_dist = [torch.empty_like(dist) for _ in range(rb.world)]
    torch.distributed.all_gather(all_nbr, nbr.contiguous(), group=rb.group)
    torch.distributed.all_gather(all_dist, dist.contiguous(), group=rb.group)
    nbr, dist = torch.cat(all_nbr, dim=1), torch.cat(all_dist, dim=1)
    # order by (distance, index): stable sort by index, then stable sort by distance
    o = torch.sort(torch.where(nbr < 0, torch.full_like(nbr, n_total), nbr), dim=1, stable=True).indices
    nbr, dist = nbr.gather(1, o), dist.gather(1, o)
    o = torch.sort(dist, dim=1, stable=True).indices[:, :k]
    return nbr.gather(1, o).contiguous(), dist.gather(1, o).contiguous()


def knn_search_sharded(Z_local, ids, k, rb, return_distance=False):
    """``knn_search`` with the key pool row-sharded over ranks (BASELINE config 4): each rank searches its own
    key block for every query (mimrl_knn_search_rows returns global indices and float64 distances), the
    world * k candidates per query are all-gathered and merged by (distance, index) -- the same total order the
    single-GPU search uses, so the result is bit-identical to it.  ``ids`` are global row numbers, the same on
    every rank.  Returns (nbr_orig, nbr_comp[, dist]) replicated on every rank."""
    if not rb.sharded:          # one rank holds the whole pool: the plain search (bit-identical, tests/test_gpu_multi.py)
        return knn_search(Z_local, ids, k, return_distance=return_distance)
    Z_local = L.f32(Z_local)
    dev = Z_local.device
    width = Z_local.shape[1]
    N, m = rb.n_all, int(ids.numel())
    if k > N - m:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {N - m}")
    ids = ids.to(device=dev, dtype=torch.int64).contiguous()
    queries = _rows_from_shards(Z_local, ids, rb)
    excluded = torch.sort(ids).values.contiguous()
    n_keys = Z_local.shape[0]
    nbr = torch.full((m, k), -1, dtype=torch.int64, device=dev)
    dist = torch.full((m, k), float("inf"), dtype=torch.float64, device=dev)
    if n_keys > 0:
        ws_bytes = L.lib.mimrl_knn_workspace_bytes(n_keys, m, width, k)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        exact = 1 if sklearn_route(width, k, N - m) == "brute" else 0
        L.check(L.lib.mimrl_knn_search_rows(L.ptr(Z_local), n_keys, width, rb.offset, L.ptr(queries), m, L.ptr(excluded), m,
                                            k, exact, L.ptr(nbr), L.ptr(dist), L.ptr(ws), ws.numel(), L.stream()))
        dist = torch.where(nbr < 0, torch.full_like(dist, float("inf")), dist)
    nbr, dist = merge_knn_candidates(nbr, dist, k, N, rb)
    comp = nbr - torch.searchsorted(excluded, nbr.reshape(-1)).reshape(m, k)      # index with the query rows removed
    return (nbr, comp, dist) if return_distance else (nbr, comp)


def prod_knn_sample_sharded(X_local, Y_local, Z_local, batch_size, k_neighbor, radius, rb):
    """``prod_knn_sample`` (Model.py:75-106) with the pools row-sharded over ranks.  Every rank draws the ids from
    numpy's global RNG exactly as the reference does (rank 0's draw is broadcast so differently seeded ranks
    still agree) and receives the full, replicated product batch."""
    N = rb.n_all
    m = batch_size // k_neighbor
    if m > N:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    dev = Z_local.device
    ids = torch.from_numpy(legacy_permutation_head(N, m)).to(dev)
    if rb.sharded:
        torch.distributed.broadcast(ids, src=torch.distributed.get_global_rank(rb.group, 0) if rb.group is not None else 0,
                                    group=rb.group)
    nbr, _ = knn_search_sharded(Z_local.detach(), ids, k_neighbor, rb)
    wmax = max(X_local.shape[1], Y_local.shape[1], Z_local.shape[1])

    def tiled(rows, repeat):
        width = rows.shape[1]
        rows = rows.repeat(1, wmax // width) if width < wmax else rows          # Model.py:98-104
        return rows.repeat_interleave(repeat, dim=0).contiguous().requires_grad_(True)

    batch_x = tiled(_rows_from_shards(X_local.detach(), nbr.reshape(-1), rb), 1)
    batch_y = tiled(_rows_from_shards(Y_local.detach(), ids, rb), k_neighbor)
    batch_z = tiled(_rows_from_shards(Z_local.detach(), ids, rb), k_neighbor)
    return batch_x, batch_y, batch_z


# --------------------------------------------------------------------------
# variational MI estimator (Model.py:108-148)
# --------------------------------------------------------------------------


class VMIEstimator(nn.Module):
    def __init__(self, critic_type, baseline_type, bound_type, d_common, hidden_dim, embed_dim, layers, activation,
                 mu, rho):
        super().__init__()
        self.critic_type, self.baseline_type, self.bound_type = critic_type, baseline_type, bound_type
        self.critic_model = CriticModel(critic_type, d_common, d_common, hidden_dim=hidden_dim, embed_dim=embed_dim,
                                        layers=layers, activation=activation)
        self.baseline_model = BaselineModel(baseline_type, d_common, hidden_dim=hidden_dim, layers=layers,
                                            activation=activation, mu=mu, rho=rho)
        self.rowblock = None            # set to a rowblock.RowBlock to shard the global batch over ranks
        self.impl = L.IMPL_AUTO

    def forward(self, features_x, features_y):
        alpha_logit = 0.01
        bound = self.bound_type
        if bound not in L.BOUND_IDS:
            raise NotImplementedError
        needs_base = bound in ("tuba", "interpolate")
        if self.critic_type == 'separate' and bound != 'interpolate':
            # fused path: the B x B score matrix never exists
            x_, y_ = self.critic_model.embed(features_x, features_y)
            base = self.baseline_model(features_y) if needs_base else None
            return separable_bound(x_, y_, bound, base, self.rowblock, self.impl)
        if self.critic_type == 'separate' and bound == 'interpolate' and features_x.is_cuda:
            x_, y_ = self.critic_model.embed(features_x, features_y)
            n_own = y_.shape[0]
            n_all = self.rowblock.n_all if self.rowblock is not None else n_own
            if L.lib.mimrl_sep_selected_impl(n_own, n_all, y_.shape[1], self.impl) == L.IMPL_TCGEN05:
                # fused: three statistics sweeps + two weighted-sum sweeps per side over TMEM tiles, row-block shardable
                return separable_interp_bound(x_, y_, self.baseline_model(features_y), alpha_logit, self.rowblock, self.impl)
        if bound == 'interpolate':
            if self.rowblock is not None and self.rowblock.sharded:
                raise NotImplementedError("the interpolated bound over a materialised score matrix (concat critic, or "
                                          "embed_dim > 128) needs the whole matrix on one rank")
            scores = self.critic_model(features_x, features_y)
            mi = interp_lower_bound(scores, self.baseline_model(features_y), alpha_logit)
            return mi, -mi
        if self.critic_type == 'concat' and self.critic_model._fused_pairs() and FUSED_CONCAT_BOUND:
            # fused path: neither the B x B score matrix nor its gradient exists
            base = self.baseline_model(features_y) if needs_base else None
            return concat_bound(self.critic_model, features_x, features_y, bound, base, self.rowblock)
        # materialised scores: this rank's rows (its x against every rank's y, Model.py global-batch semantics)
        scores = self.critic_model(features_x, gather_rows(features_y, self.rowblock))
        base = self.baseline_model(features_y) if needs_base else None
        return _scores_bound(scores, bound, base, self.rowblock)


# --------------------------------------------------------------------------
# classifier-based conditional MI (Model.py:47-72, 150-225)
# --------------------------------------------------------------------------


class _VCMIHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, act_id):
        logits = L.f32(logits)
        n = logits.shape[0] // 2
        result = torch.empty(2, dtype=torch.float32, device=logits.device)
        L.check(L.lib.mimrl_vcmi_head_fwd(L.ptr(logits), n, act_id, L.ptr(result), L.stream()))
        ctx.save_for_backward(logits)
        ctx.act_id = act_id
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(ctx, g_cmi, g_loss):
        (logits,) = ctx.saved_tensors
        n = logits.shape[0] // 2
        z = logits.new_zeros(())
        grad = torch.stack([g_cmi if g_cmi is not None else z, g_loss if g_loss is not None else z]).contiguous()
        gl = torch.empty_like(logits)
        L.check(L.lib.mimrl_vcmi_head_bwd(L.ptr(logits), n, ctx.act_id, L.ptr(grad), L.ptr(gl), L.stream()))
        return gl, None


class MLP_For_CMI(nn.Module):
    """Model.py:47-72.  ``layers`` is accepted and ignored, as in the reference."""

    def __init__(self, dim, hidden_dim, output_dim, layers, activation, last_acticate):
        super().__init__()
        act = get_activation(activation)
        self.mlp = nn.Sequential(
            nn.Linear(dim, hidden_dim), act(),
            nn.Linear(hidden_dim, hidden_dim), act(),
            nn.Linear(hidden_dim, hidden_dim), act(),
            nn.Linear(hidden_dim, output_dim))
        if last_acticate == 'hardtanh':
            self.final_activate = nn.Hardtanh(1e-4, 1 - 1e-4)
        elif last_acticate == 'sigmoid':
            self.final_activate = nn.Sigmoid()
        else:
            raise NotImplementedError
        self.act_id = 0 if last_acticate == 'hardtanh' else 1

    def logits(self, features):
        return mlp_apply(self.mlp, features)

    def forward(self, features):
        return self.final_activate(torch.clamp(self.mlp(features), -10, 10))


class VCMIEstimator(nn.Module):
    def __init__(self, embed_dim, hidden_dim, layers, activation, k_neighbor, radius, last_acticate='hardtanh'):
        super().__init__()
        self.classifier = MLP_For_CMI(embed_dim * 3, hidden_dim, output_dim=2, layers=layers, activation=activation,
                                      last_acticate=last_acticate)
        self.k_neighbor, self.radius = k_neighbor, radius
        self.embed_dim = embed_dim

    def forward(self, features_x, features_y, features_z, random_knn_x, random_knn_y, random_knn_z):
        """I(x;y|z): returns (cmi, loss).  The reference runs the classifier a second
        time inside estimate_cmi on the same batch (Model.py:189,206); both passes
        give identical values, so one pass feeds both heads and the gradients add."""
        e = self.embed_dim

        def widen(f):                                       # Model.py:161-166 (N5)
            return f.repeat(1, e // f.shape[1]) if f.shape[1] != e else f
        joint = torch.cat([widen(features_x), widen(features_y), widen(features_z)], dim=1)
        prod = torch.cat([random_knn_x, random_knn_y, random_knn_z], dim=1)
        if joint.shape[0] != prod.shape[0]:                 # Model.py:180-182 (N4)
            joint = joint[:prod.shape[0]]
        batch = torch.cat([joint, prod], dim=0)
        return _VCMIHead.apply(self.classifier.logits(batch), self.classifier.act_id)

    def estimate_cmi(self, batch, cmi_type='nwj'):
        if cmi_type != 'nwj':
            raise NotImplementedError
        return _VCMIHead.apply(self.classifier.logits(batch), self.classifier.act_id)[0]


# --------------------------------------------------------------------------
# stage functions (Model.py:305-386) as a mixin over any module that owns the
# eleven estimators under the reference's attribute names
# --------------------------------------------------------------------------

_CMI_PLAN = (  # (estimator, X, Y, Z) with C = labels; Z is the searched pool  
    ("ac_t", "A", "C", "T"), ("ta_c", "T", "A", "C"), ("vc_t", "V", "C", "T"),
    ("tv_c", "T", "V", "C"), ("tc_a", "T", "C", "A"), ("tc_v", "T", "C", "V"))


_BRANCH_STREAMS = {}


def _branch_streams(device, n):
    pool = _BRANCH_STREAMS.setdefault(device, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def run_branches(jobs, parallel):
    """``jobs``: list of (key, thunk) of INDEPENDENT pieces of work -> {key: result}.  ``parallel``: every thunk is
    enqueued on its own side stream (forked from and joined back into the current stream), so that the small-batch
    regime -- a stage is eleven estimators of a few latency-bound launches each -- overlaps on the GPU instead of
    running as one chain.  The thunks still RUN on the host in list order (numpy RNG consumption order of the samplers is
    the reference's); autograd replays every branch's backward on the stream of its forward.  Meant for CUDA-graph
    capture (the side streams fork from the capturing stream and rejoin it; GraphedTwoStageStep turns it on).  Eager use is
    correct but not useful: the step is host-bound there, and PyTorch's allocator caches every side stream's workspaces
    separately (at bs = 1024 the main stream then runs into cudaMalloc / cudaFree cycles)."""
    if not parallel or len(jobs) < 2 or not torch.cuda.is_available():
        return {k: f() for k, f in jobs}
    main = torch.cuda.current_stream()
    streams = _branch_streams(main.device, min(len(jobs), 6))
    capturing = torch.cuda.is_current_stream_capturing()
    out = {}
    for i, (k, f) in enumerate(jobs):
        s = streams[i % len(streams)]
        s.wait_stream(main)
        with torch.cuda.stream(s):
            r = f()
        if not capturing:             # results are consumed on the main stream after the join
            for t in (r if isinstance(r, (tuple, list)) else (r,)):
                if torch.is_tensor(t):
                    t.record_stream(main)
        out[k] = r
    for s in streams:
        main.wait_stream(s)
    return out


class MIStageMixin:
    """compute_vmi_loss_stage1 / stage2 with the reference's call order (and
    therefore its numpy RNG consumption order)."""

    parallel_branches = False        # opt-in: the independent estimators of a stage on side streams (run_branches)

    def _mi_terms(self, F_F, T_F, A_F, V_F):
        jobs = []
        for name, (a, b) in (("f_t", (F_F, T_F)), ("f_a", (F_F, A_F)), ("f_v", (F_F, V_F)),
                             ("t_a", (T_F, A_F)), ("t_v", (T_F, V_F))):
            jobs.append((name, (lambda n=name, a=a, b=b: getattr(self, "vmi_estimator_" + n)(a, b))))
        return run_branches(jobs, self.parallel_branches)

    def _cmi_terms(self, labels, T_F, A_F, V_F, C_F_all, T_F_all, A_F_all, V_F_all):
        feats = {"T": T_F, "A": A_F, "V": V_F, "C": labels}
        pools = {"T": T_F_all, "A": A_F_all, "V": V_F_all, "C": C_F_all}
        bs = labels.shape[0]
        # T_F_all is searched twice (Model.py:323,329): fit it once for this call.  A caller who passes KnnPool objects
        # (fitted once per epoch) skips the fit of every pool.
        if not isinstance(pools["T"], KnnPool) and L.lib.mimrl_knn_fit_bytes(*pools["T"].shape):
            pools["T"] = KnnPool(pools["T"])

        def one(name, x, y, z):
            kx, ky, kz = prod_knn_sample(pools[x], pools[y], pools[z], bs, self.k_neighbor, self.radius)
            return getattr(self, "vcmi_estimator_" + name)(feats[x], feats[y], feats[z], kx, ky, kz)
        jobs = [(name, (lambda n=name, x=x, y=y, z=z: one(n, x, y, z))) for name, x, y, z in _CMI_PLAN]
        return run_branches(jobs, self.parallel_branches)

    def compute_vmi_loss_stage1(self, predictions, labels, F_F, T_F, A_F, V_F, C_F_all, F_F_all, T_F_all, A_F_all,
                                V_F_all):
        labels = labels.reshape(-1, 1).repeat((1, self.d_common))
        mi = self._mi_terms(F_F, T_F, A_F, V_F)
        cmi = self._cmi_terms(labels, T_F, A_F, V_F, C_F_all, T_F_all, A_F_all, V_F_all)
        order_mi = ["f_t", "f_a", "f_v", "t_a", "t_v"]
        order_cmi = ["ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"]
        return ([mi[k][0] for k in order_mi] + [cmi[k][0] for k in order_cmi],
                [mi[k][1] for k in order_mi] + [cmi[k][1] for k in order_cmi])

    def compute_vmi_loss_stage2(self, predictions, labels, F_F, T_F, A_F, V_F, C_F_all, F_F_all, T_F_all, A_F_all,
                                V_F_all):
        labels = labels.reshape(-1, 1).repeat((1, self.d_common))
        mi = self._mi_terms(F_F, T_F, A_F, V_F)
        mi_inv = mi["t_a"][0] + mi["t_v"][0]
        c = {k: v[0] for k, v in self._cmi_terms(labels, T_F, A_F, V_F, C_F_all, T_F_all, A_F_all, V_F_all).items()}
        mi_spec_t = c["tc_a"] + c["tc_v"] - c["ta_c"] - c["tv_c"]
        mi_spec_a = c["ac_t"] - c["ta_c"]
        mi_spec_v = c["vc_t"] - c["tv_c"]
        mi_comp = c["ta_c"] + c["tv_c"]
        return ([mi["f_t"][0], mi["f_a"][0], mi["f_v"][0], mi_inv, mi_spec_t, mi_spec_a, mi_spec_v, mi_comp],
                [mi["f_t"][1], mi["f_a"][1], mi["f_v"][1], -mi_inv, -mi_spec_t, -mi_spec_a, -mi_spec_v, -mi_comp])


class MIHeads(nn.Module, MIStageMixin):
    """The MI/CMI part of reference ``Model.__init__`` (Model.py:283-303): five
    VMIEstimators and six VCMIEstimators under the reference's attribute names,
    plus the two stage functions.  A full ``Model`` can inherit ``MIStageMixin``
    instead and keep its own encoders."""

    def __init__(self, opt, d_common=None):
        super().__init__()
        self.d_common = d_common if d_common is not None else opt.d_common
        hidden_dim, embed_dim, layers, activation = 256, 128, 2, 'relu'      # Model.py:285 (hard-coded there)
        hidden_dim = getattr(opt, "mi_hidden_dim", hidden_dim)
        embed_dim = getattr(opt, "mi_embed_dim", embed_dim)
        mu, rho = 0, 1
        self.k_neighbor, self.radius = opt.k_neighbor, opt.radius
        for n in ("f_t", "f_a", "f_v", "t_a", "t_v"):
            setattr(self, "vmi_estimator_" + n,
                    VMIEstimator(opt.critic_type, opt.baseline_type, opt.bound_type, self.d_common, hidden_dim,
                                 embed_dim, layers, activation, mu, rho))
        for n in ("ac_t", "ta_c", "vc_t", "tv_c", "tc_a", "tc_v"):
            setattr(self, "vcmi_estimator_" + n,
                    VCMIEstimator(embed_dim, hidden_dim, layers, activation, opt.k_neighbor, opt.radius,
                                  opt.cmi_last_acticate))
