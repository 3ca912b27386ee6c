#!/usr/bin/env python
"""Headline benchmark: critic pairs/s for the separable-critic InfoNCE
estimator, forward + backward (BASELINE.json configs[1]: global batch 65536,
d_common = 128, hidden 256, fp32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = VMIEstimator.forward + mi_loss.backward() (Model.py:115-148) over one
batch of synthetic features: both critic MLPs forward, the fused score/InfoNCE
sweep, both gradient sweeps, both MLPs backward.  A pair is one (i, j) entry of
the B x B score matrix (never materialised here); pairs per step = B^2.

N > 1 (torchrun, one rank per GPU): the global batch is sharded by row blocks,
embeddings are all-gathered over NCCL, every rank sweeps its rows against the
whole batch, parameter gradients are all-reduced.  Per-GPU work is held fixed
(rows_per_gpu * B_global = 65536^2), so scaling is "weak".

Prints ONE JSON line (see the task contract) including `roofline`,
`cpu_baseline`, `e2e`, `clocks` and `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_SINGLE = 65536
D_COMMON, HIDDEN, EMBED, LAYERS = 128, 256, 128, 2
METRIC = "critic_pairs_per_s_infonce_fwd_bwd"
UNIT = "pairs/s"


def global_batch(n_gpus):
    """B_global with rows_per_gpu * B_global = 65536^2 (per-GPU work fixed)."""
    if n_gpus == 1:
        return B_SINGLE
    q = 128 * n_gpus
    return int(round(B_SINGLE * math.sqrt(n_gpus) / q)) * q


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref, unmodified sources behind import shims) on all host threads;
# the numpy oracle port only when the reference sources did not travel
# --------------------------------------------------------------------------

CPU_SAMPLE_B = 8192          # bounded sample: the reference materialises ~10 B x B fp32 temporaries


def workload_config(B, n_own, world, precision, scaling="weak"):
    """`config` of the JSON line, identical for both arms (the CPU arm times a bounded sample of it)."""
    return {"workload": "separable-critic InfoNCE fwd+bwd (VMIEstimator, Model.py:108-148), d_common=128 "
                        "hidden=256 embed=128 layers=2, BASELINE configs[1] at its largest batch",
            "global_batch": B, "rows_per_gpu": n_own, "d_common": D_COMMON, "parallelism": f"rowblock{world}",
            "l2": "flushed between timed steps (256 MiB write, outside the per-step events)",
            "timing": "CUDA events per step, mean over steps, max over ranks",
            "precision": precision}


GPU_PRECISION = "tcgen05 fp16x3 split (fp32-class); TF32 disabled for every torch GEMM"


def cpu_arm(steps, warmup, cfg1=False):
    """Times the reference's VMIEstimator forward+backward on torch CPU (kind "reference") or, without
    oracle/_ref, the numpy port (kind "port").  Returns (seconds per step, cpu_baseline dict, extras)."""
    from oracle import ref_shim as R
    B = CPU_SAMPLE_B
    cores = os.cpu_count() or 1
    extras = {}
    if R.locate() is not None:
        ref = R.import_reference()
        threads = R.set_threads(cores)
        sec = R.time_fn(R.vmi_step_fn(ref, B), steps, warmup)
        kind = "reference"
        what = (f"B={B} rows ({B * B} pairs per step) of the B={B_SINGLE} workload: the unmodified reference "
                f"Model.VMIEstimator (separate/constant/infonce) forward+backward on torch CPU fp32, {threads} threads "
                f"(set explicitly; torchrun's OMP_NUM_THREADS=1 is overridden)")
        if cfg1:
            t0 = time.perf_counter()
            step = R.cfg1_step_fn(ref)
            t_build = time.perf_counter() - t0
            t0 = time.perf_counter()
            losses = step()
            t_step = time.perf_counter() - t0
            extras["config1_cpu_step"] = {
                "seconds": t_step, "steps_per_s": 1.0 / t_step, "threads": threads, "build_seconds": t_build,
                "losses": [float(v) for v in losses],
                "what": "BASELINE configs[0]: reference Model (random-init bert-base, GRU encoders, CubeMLP 50-3-128=10-3-128, "
                        "separate/constant/infonce, k=2) on a MOSI-shaped synthetic batch bs=128 time_len=100, pools N=1284: "
                        "one stage-1 step + one stage-2 step incl. both Adam updates, CPU, 1 cold step"}
    else:
        from oracle import params as P
        from oracle import vmi_oracle as O
        prm = P.vmi_params(0, "separate", "constant", D_COMMON, HIDDEN, EMBED, LAYERS)
        x, y = P.features(1, B, D_COMMON, corr=0.6)
        fn = lambda: O.separable_infonce_streamed(prm, x, y, dtype=np.float32, block=2048)
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        sec = (time.perf_counter() - t0) / steps
        threads, kind = cores, "port"
        what = f"B={B} rows ({B * B} pairs per step) of the B={B_SINGLE} workload; numpy fp32 oracle port (oracle/_ref absent)"
    cpu = {"value": B * B / sec, "unit": UNIT, "cores": threads, "kind": kind, "sample": what}
    return sec, cpu, extras


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, cpu, extras = cpu_arm(args.steps, args.warmup, cfg1=(args.gpus == 1 and not args.no_cfg1))
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(global_batch(args.gpus), global_batch(args.gpus) // args.gpus, args.gpus, GPU_PRECISION),
        "config_note": f"same workload as the GPU arm; each CPU step is the bounded sample described in cpu_baseline.sample "
                       f"(B={CPU_SAMPLE_B}: the reference materialises the B x B matrix, B=65536 would need ~170 GB)",
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line.update(extras)
    print(json.dumps(line))


# --------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------


class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------
# extras carried in the same JSON line (each guarded: an extra can fail without taking the headline down)
# --------------------------------------------------------------------------


def ncu_record(kernel):
    """dram bytes / tensor-pipe activity of one launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_kernels.json, written by scripts/summarize_ncu.py --json).  {} when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernels.json")) as f:
            return json.load(f).get(kernel, {})
    except Exception:
        return {}


def _rel(a, b):
    import torch
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _ragged(n, world):
    """contiguous row blocks, deliberately uneven (3 rows move from the last rank to the first)"""
    from mimrl_b200 import rowblock as RB
    c = list(RB.even_split(n, world))
    if world > 1 and c[-1] > 3:
        c[0] += 3
        c[-1] -= 3
    return tuple(c)


def extra_parity_check(args, world, rank, dev, flush):
    """N > 1: the row-block sharded estimator (value, own rows of both input gradients, all-reduced parameter
    gradients) and the key-sharded k-NN search against the SAME computation on one rank, which every rank runs itself.
    N = 1: the tcgen05 path against the independent CUDA-core implementation, and the k-NN search against a float64
    brute-force search in torch."""
    import torch
    import torch.distributed as dist
    from mimrl_b200 import _lib as L
    from mimrl_b200 import rowblock as RB
    from mimrl_b200.model import VMIEstimator, knn_search, knn_search_sharded
    B, tol = 4096, 1e-4
    res = {"global_batch": B, "tolerance": tol, "norm": "max|a-b| / max|b| (b = single-rank result)",
           "against": "single-rank run of the same kernels" if world > 1 else "independent CUDA-core (FFMA) kernels"}
    g = torch.Generator().manual_seed(77)
    x_all = torch.randn(B, D_COMMON, generator=g)
    y_all = 0.6 * x_all + 0.8 * torch.randn(B, D_COMMON, generator=g)
    counts = _ragged(B, world)
    off = sum(counts[:rank])
    worst = 0.0
    for bound in ("infonce", "nwj"):
        torch.manual_seed(5)
        est = VMIEstimator("separate", "constant", bound, D_COMMON, HIDDEN, EMBED, LAYERS, "relu", 0, 1).to(dev)
        params = list(est.parameters())

        def run(xa, ya, rb, impl):
            est.rowblock, est.impl = rb, impl
            for p in params:
                p.grad = None
            xt, yt = xa.to(dev).requires_grad_(True), ya.to(dev).requires_grad_(True)
            mi, loss = est(xt, yt)
            loss.backward()
            if rb is not None:
                RB.all_reduce_param_grads(params, rb)
            return mi.detach(), xt.grad, yt.grad, torch.cat([p.grad.reshape(-1) for p in params])
        if world > 1:
            mi0, gx0, gy0, pg0 = run(x_all, y_all, None, L.IMPL_AUTO)
            rb = RB.RowBlock(rank, world, counts, None)
            sl = slice(off, off + counts[rank])
            mi1, gx1, gy1, pg1 = run(x_all[sl], y_all[sl], rb, L.IMPL_AUTO)
            gx0, gy0 = gx0[sl], gy0[sl]
        else:
            mi0, gx0, gy0, pg0 = run(x_all, y_all, None, L.IMPL_FFMA)
            mi1, gx1, gy1, pg1 = run(x_all, y_all, None, L.IMPL_AUTO)
        errs = torch.tensor([float((mi1 - mi0).abs() / mi0.abs().clamp_min(1.0)), _rel(gx1, gx0), _rel(gy1, gy0),
                             _rel(pg1, pg0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        e = [float(v) for v in errs]
        res[bound] = {"mi": float(mi1), "mi_err": e[0], "grad_x_err": e[1], "grad_y_err": e[2], "param_grad_err": e[3]}
        worst = max(worst, e[0], e[1], e[2])
    # k-NN: 64k x 128 pool, 256 queries, k = 4
    N, m, k = 65536, 256, 4
    Z = torch.randn(N, 128, generator=g).to(dev)
    ids = torch.randperm(N, generator=g)[:m].to(dev)
    nbr0, _ = knn_search(Z, ids, k)
    if world > 1:
        kc = _ragged(N, world)
        ko = sum(kc[:rank])
        rbk = RB.RowBlock(rank, world, kc, None)
        nbr1, _ = knn_search_sharded(Z[ko: ko + kc[rank]].contiguous(), ids, k, rbk)
        same = torch.tensor([int(torch.equal(nbr0, nbr1))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        res["knn"] = {"n_keys": N, "queries": m, "k": k, "sharded_equals_single_rank": bool(same.item())}
        knn_ok = bool(same.item())
    else:
        q = Z[ids].double()
        d2 = (q * q).sum(1)[:, None] + (Z.double() ** 2).sum(1)[None, :] - 2.0 * q @ Z.double().t()
        d2[torch.arange(m, device=dev)[:, None].expand(m, m), ids[None, :].expand(m, m)] = float("inf")
        want = torch.topk(d2, k, dim=1, largest=False, sorted=True).indices
        knn_ok = bool(torch.equal(want, nbr0))
        res["knn"] = {"n_keys": N, "queries": m, "k": k, "equals_float64_bruteforce": knn_ok}
    res["ok"] = bool(worst <= tol and knn_ok)
    return res


def extra_strong_scaling(args, world, rank, dev, flush):
    """B = 65536 held fixed and sharded over the ranks (the weak-scaling headline keeps per-GPU work fixed instead),
    with the two collectives of a step timed on their own."""
    import torch
    import torch.distributed as dist
    from mimrl_b200 import rowblock as RB
    from mimrl_b200.model import VMIEstimator
    if world == 1 or args.scaling == "strong":
        return None
    B = B_SINGLE
    counts = RB.even_split(B, world)
    rb = RB.RowBlock(rank, world, counts, None)
    n_own = rb.n_own
    torch.manual_seed(0)
    est = VMIEstimator("separate", "constant", "infonce", D_COMMON, HIDDEN, EMBED, LAYERS, "relu", 0, 1).to(dev)
    est.rowblock = rb
    params = list(est.parameters())
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    x = torch.randn(n_own, D_COMMON, generator=g).to(dev).requires_grad_(True)
    y = torch.randn(n_own, D_COMMON, generator=g).to(dev).requires_grad_(True)

    def step():
        x.grad = y.grad = None
        for p in params:
            p.grad = None
        mi, loss = est(x, y)
        loss.backward()
        RB.all_reduce_param_grads(params, rb)

    def timed(fn, steps):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        t = torch.tensor([tot / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    steps = max(3, min(args.steps, 10))
    ms = timed(step, steps)
    emb = torch.randn(n_own, EMBED, device=dev)
    flat = torch.zeros(sum(p.numel() for p in params), device=dev)
    ms_ag = timed(lambda: RB.all_gather_rows(emb, rb), steps)
    ms_ar = timed(lambda: dist.all_reduce(flat), steps)
    return {"global_batch": B, "rows_per_gpu": n_own, "ms_per_step": ms, "value": float(B) * B / (ms * 1e-3), "unit": UNIT,
            "steps": steps, "collectives_ms": {"all_gather_embeddings_per_call": ms_ag, "calls_per_step": 2,
                                               "all_reduce_param_grads": ms_ar, "param_floats": int(flat.numel())},
            "note": "each rank sweeps B/N rows against all B columns; two embedding all-gathers (forward x, backward y) and "
                    "one flat parameter-gradient all-reduce per step, not overlapped with the sweeps"}


def extra_configs(args, world, rank, dev, flush):
    """BASELINE configs 3, 4, 5 at this N (device-timed, max over ranks): concat critic B = 16384 NWJ/JS sharded by row
    blocks; k-NN on a 1M x 128 pool, keys row-sharded; the full two-stage training step."""
    import torch
    import torch.distributed as dist
    from mimrl_b200 import rowblock as RB
    from mimrl_b200.model import VMIEstimator, knn_search_sharded
    out = {}

    def timed(fn, warm=1, reps=2):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- config 3 ----
    try:
        B = 16384
        counts = RB.even_split(B, world)
        rb = RB.RowBlock(rank, world, counts, None) if world > 1 else None
        off = sum(counts[:rank])
        g = torch.Generator().manual_seed(0)
        x_all, y_all = torch.randn(B, 128, generator=g), torch.randn(B, 128, generator=g)
        c3 = {}
        for bound in ("nwj", "js"):
            torch.manual_seed(0)
            est = VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1).to(dev)
            est.rowblock = rb
            x = x_all[off: off + counts[rank]].to(dev).requires_grad_(True)
            y = y_all[off: off + counts[rank]].to(dev).requires_grad_(True)
            params = list(est.parameters())
            last = {}

            def step():
                est.zero_grad(set_to_none=True)
                mi, loss = est(x, y)
                loss.backward()
                if rb is not None:
                    RB.all_reduce_param_grads(params, rb)
                last["mi"] = mi.detach()
            ms = timed(step, warm=1, reps=2)
            c3[bound] = {"ms": ms, "pairs_per_s": B * B / ms * 1e3, "reference_dense_tflops": 1_181_184.0 * B * B / ms * 1e-9,
                         "mi": float(last["mi"])}
            del est, x, y
            torch.cuda.empty_cache()
        out["config3_concat_B16384"] = c3
    except Exception as e:
        out["config3_concat_B16384"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- config 4 ----
    try:
        N, width, bs = 1 << 20, 128, 8192
        counts = RB.even_split(N, world)
        rb = RB.RowBlock(rank, world, counts, None) if world > 1 else RB.single(N)
        off = sum(counts[:rank])
        gen = torch.Generator(device=dev)
        Z = torch.empty(counts[rank], width, device=dev)
        blk = 1 << 16
        for b0 in range(off - off % blk, off + counts[rank], blk):       # pool seeded per 64k block: independent of N
            gen.manual_seed(1000 + b0 // blk)
            zb = torch.randn(blk, width, device=dev, generator=gen)
            lo, hi = max(b0, off), min(b0 + blk, off + counts[rank])
            Z[lo - off: hi - off] = zb[lo - b0: hi - b0]
        c4 = {}
        for k in (2, 16):
            m = bs // k
            ids = torch.from_numpy(np.random.RandomState(0).permutation(N)[:m].astype(np.int64)).to(dev)
            ms = timed(lambda: knn_search_sharded(Z, ids, k, rb), warm=1, reps=3)
            nbr, _ = knn_search_sharded(Z, ids, k, rb)
            c4[f"k{k}"] = {"queries": m, "search_ms": ms, "key_gbs": N * width * 4 / ms * 1e-6,
                           "algorithmic_tflops": 2.0 * width * m * N / ms * 1e-9, "checksum": int(nbr.sum().item())}
        if world == 1:          # the single-GPU sampler paths around the sharded search above
            import time as _t
            from mimrl_b200.model import KnnPool, knn_search, legacy_permutation_head
            c4["fit_ms"] = timed(lambda: KnnPool(Z), warm=1, reps=3)
            pool = KnnPool(Z)
            for k in (2, 16):
                ids = torch.from_numpy(np.random.RandomState(0).permutation(N)[:bs // k].astype(np.int64)).to(dev)
                c4[f"k{k}"]["unsharded_search_ms"] = timed(lambda: knn_search(Z, ids, k), warm=1, reps=3)
                c4[f"k{k}"]["fitted_pool_search_ms"] = timed(lambda: knn_search(pool, ids, k), warm=1, reps=3)
            gen.manual_seed(7)
            Zl = torch.randn(N, 1, device=dev, generator=gen)
            ids = torch.from_numpy(np.random.RandomState(0).permutation(N)[:bs // 2].astype(np.int64)).to(dev)
            c4["label_pool_width1_k2_search_ms"] = timed(lambda: knn_search(Zl, ids, 2), warm=1, reps=3)
            st = np.random.get_state()
            t0 = _t.perf_counter(); np.random.permutation(N)[:bs // 2]; t1 = _t.perf_counter()
            legacy_permutation_head(N, bs // 2); t2 = _t.perf_counter()
            np.random.set_state(st)
            c4["host_id_draw_ms"] = {"numpy_permutation": (t1 - t0) * 1e3, "mimrl_legacy_permutation_head": (t2 - t1) * 1e3,
                                     "what": "np.random.choice(range(N), m, replace=False) stream of Model.py:81, N = 2^20"}
            del pool, Zl
        out["config4_knn_1Mx128"] = c4
        del Z
        torch.cuda.empty_cache()
    except Exception as e:
        out["config4_knn_1Mx128"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- CubeMLP fusion encoder alone (the third kernel family of the north star), config-5 shape, per GPU ----
    try:
        from mimrl_b200.graphs import GraphedCallable
        from mimrl_b200.mlp_process import MLPEncoder
        torch.manual_seed(0)
        bs = 1024
        enc = MLPEncoder("gelu", [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]], [0.0] * 3, True,
                         False, [True, True]).to(dev)
        xe = torch.randn(bs, 100, 3, 128, device=dev)
        prm = list(enc.parameters())

        def fwd_only(xx):
            with torch.no_grad():
                return enc(xx)
        xs = xe.clone().requires_grad_(True)

        def fwd_bwd(xx):
            xs.grad = None
            for q in prm:
                q.grad = None
            enc(xs).sum().backward()
            return xs.grad
        gf, gb = GraphedCallable(fwd_only, [xe]), GraphedCallable(fwd_bwd, [xs.detach()])
        ms_f = timed(lambda: gf(gf.static_in[0]), warm=2, reps=5)
        ms_b = timed(lambda: gb(gb.static_in[0]), warm=2, reps=5)
        ms_fe = timed(lambda: fwd_only(xe), warm=2, reps=5)
        ms_be = timed(lambda: fwd_bwd(None), warm=2, reps=5)
        alg = bs * 384 * 4.0 * (100 + 2 * 50 + 10)            # one read + one write per block, 330 MB
        peak = peaks()["hbm_gbs"]
        out["cubemlp_encoder"] = {
            "shape": [bs, 100, 3, 128], "blocks": "50-3-128=10-3-128", "fwd_ms_cuda_graph": ms_f, "fwd_bwd_ms_cuda_graph": ms_b,
            "fwd_ms_eager": ms_fe, "fwd_bwd_ms_eager": ms_be, "algorithmic_bytes_fwd": alg,
            "fwd_gbs": alg / ms_f * 1e-6, "hbm_peak_gbs": peak, "fwd_frac_of_hbm_peak": alg / ms_f * 1e-6 / peak,
            "what": "MLPEncoder forward / forward+backward per GPU, CUDA events; graph = the same calls replayed as one CUDA graph"}
        del enc, xe, xs, gf, gb
        torch.cuda.empty_cache()
    except Exception as e:
        out["cubemlp_encoder"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- config 5 ----
    try:
        from mimrl_b200.full_model import bench_config5
        out["config5_train_step"] = bench_config5(world, rank, dev, timed)
    except Exception as e:
        out["config5_train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as G
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        G.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    torch.backends.cuda.matmul.allow_tf32 = False        # parity mode: fp32-class arithmetic everywhere
    torch.backends.cudnn.allow_tf32 = False

    from mimrl_b200 import _lib as L
    from mimrl_b200 import rowblock as RB
    from mimrl_b200.model import VMIEstimator

    B = args.batch or (B_SINGLE if args.scaling == "strong" else global_batch(world))
    counts = RB.even_split(B, world)
    rb = RB.RowBlock(rank, world, counts, None) if world > 1 else RB.single(B)
    n_own = rb.n_own

    torch.manual_seed(0)                                  # random-init weights, identical on every rank
    est = VMIEstimator("separate", "constant", "infonce", D_COMMON, HIDDEN, EMBED, LAYERS, "relu", 0, 1).to(dev)
    est.rowblock = rb if world > 1 else None
    params = [p for p in est.parameters()]

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = torch.randn(n_own, D_COMMON, generator=g).pin_memory()
    y_host = (0.6 * x_host + 0.8 * torch.randn(n_own, D_COMMON, generator=g)).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    y = y_host.to(dev).requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step(xi, yi):
        xi.grad = yi.grad = None
        for p in params:
            p.grad = None
        mi, loss = est(xi, yi)
        loss.backward()
        RB.all_reduce_param_grads(params, rb)
        return mi

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(x, y)
    sync_all()

    # ---- timed region: K steps, device-timed, L2 flushed between steps ------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = L.launch_count()
    with ClockSampler(local_rank) as clk:
        sync_all()
        t_wall = time.perf_counter()
        for k in range(args.steps):
            flush.zero_()
            starts[k].record()
            step(x, y)
            stops[k].record()
        sync_all()
        t_wall = time.perf_counter() - t_wall
    launches = L.launch_count() - launches0
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    pairs = float(B) * float(B)
    value = pairs / (ms * 1e-3)

    # ---- e2e: host buffers in, scalar out, copies inside the timed region ----
    # Every step copies ITS inputs from pinned host memory and reads ITS metric back to the host.  The input copy
    # of step k+1 runs on a copy stream while step k computes (double-buffered device inputs), and the metric of
    # step k is read after step k+1 has been launched -- an ordinary prefetching input pipeline.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(x_host, device=dev).requires_grad_(True),
             torch.empty_like(y_host, device=dev).requires_grad_(True)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    mi_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    mi_done = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[i])                    # the step that used this buffer has finished
            bufs[i][0].copy_(x_host, non_blocking=True)
            bufs[i][1].copy_(y_host, non_blocking=True)
            ready[i].record(copy_stream)

    def e2e_run(n_steps):
        got = []
        for i in range(2):
            consumed[i].record()
        prefetch(0)
        for k in range(n_steps):
            cur = k & 1
            if k + 1 < n_steps:
                prefetch(cur ^ 1)
            torch.cuda.current_stream().wait_event(ready[cur])
            mi = step(*bufs[cur])
            consumed[cur].record()
            mi_host[cur].copy_(mi.detach(), non_blocking=True)   # device -> host read of the metric
            mi_done[cur].record()
            if k > 0:
                mi_done[cur ^ 1].synchronize()
                got.append(float(mi_host[cur ^ 1]))
        mi_done[(n_steps - 1) & 1].synchronize()
        got.append(float(mi_host[(n_steps - 1) & 1]))
        return got

    e2e_run(2)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    e2e_mi = e2e_run(e2e_steps)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    assert len(e2e_mi) == e2e_steps and all(np.isfinite(v) for v in e2e_mi)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # ---- the two sweeps a step launches, each timed alone with CUDA events on its own stream -----
    # (a) mimrl_sep_online_forward: statistics + owned-row weighted sum in one sweep (the step's forward)
    # (b) mimrl_sep_weighted_sum, shift indexed by the swept row: the column-side gradient sweep (the step's backward)
    with torch.no_grad():
        xe_, ye_ = est.critic_model.embed(x, y)
        xe_, ye_ = xe_.contiguous(), ye_.contiguous()
        all_x = RB.all_gather_rows(xe_, rb)
    n_all = all_x.shape[0]
    ws = torch.empty(L.lib.mimrl_sep_workspace_bytes(n_own, n_all, EMBED) + 16, dtype=torch.uint8, device=dev)
    stats = torch.empty(4, n_own, device=dev)
    st = L.stream()

    def k_stats(flags=0):
        L.check(L.lib.mimrl_sep_row_stats(L.ptr(ye_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, flags, 0,
                                          L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(stats[3]), L.ptr(ws),
                                          ws.numel(), st))
    k_stats()
    shift = (stats[0] + torch.log(stats[1])).contiguous()
    ref_pt = stats[0].clone()
    coef = torch.full((1,), -1.0 / n_all, device=dev)
    dcoef = torch.full((n_own,), 1.0 / n_all, device=dev)
    out = torch.empty_like(ye_)
    rsum = torch.empty(n_own, device=dev)

    def k_fused():          # the step's forward: online reference point, statistics and owned-row gradient sum in one sweep
        L.check(L.lib.mimrl_sep_online_forward(L.ptr(ye_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, 1, L.ptr(ref_pt),
                                               L.ptr(out), L.ptr(rsum), None, L.ptr(ws), ws.numel(), st))

    def time_kernel(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        acc = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            acc += a.elapsed_time(b)
        return acc / reps
    shift_all = RB.all_gather_rows(shift, rb)

    def k_wsum_swept():          # the backward sweep: operands swapped, shift indexed by the swept row
        L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(xe_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, 0, 1,
                                             L.ptr(shift_all), 1, L.ptr(coef), L.ptr(dcoef), 0, L.ptr(out), L.ptr(ws),
                                             ws.numel(), st))
    tc_path = L.lib.mimrl_sep_selected_impl(n_own, n_all, EMBED, 0) == L.IMPL_TCGEN05
    ms_fused = time_kernel(k_fused) if tc_path else float("nan")
    ms_wsum_swept = time_kernel(k_wsum_swept)
    ms_wsum = 0.5 * (ms_fused + ms_wsum_swept) if tc_path else ms_wsum_swept
    ms_stats = time_kernel(k_stats)
    pk = peaks()
    flops_wsum = 2.0 * EMBED * n_own * n_all              # algorithmic: one P.X contraction (score recompute not counted)
    achieved = flops_wsum / (ms_wsum * 1e-3) / 1e12
    impl_name = "tcgen05 fp16x3 split (fp32-class)" if tc_path else "fp32 FFMA (CUDA cores)"
    # the committed ncu --set full capture of the two instantiations a step launches (online forward, swept-side sweep)
    ncu = {}
    if world == 1 and B == B_SINGLE:
        recs = [ncu_record("sep_wsum_tc_kernel<0, 1>"), ncu_record("sep_wsum_tc_kernel<0, 0>")]
        if all(r.get("dram_bytes") is not None for r in recs):
            ncu = {"dram_bytes": sum(r["dram_bytes"] for r in recs) / 2,
                   "tensor_pipe_active_pct": sum(r["tensor_pipe_active_pct"] for r in recs) / 2, "source": recs[0].get("source")}
    roofline = {
        "kernel": "sep_wsum_tc_kernel (2 launches per step: mimrl_sep_online_forward = online-softmax forward statistics + "
                  "owned-row gradient sum, and mimrl_sep_weighted_sum = swept-side gradient sweep); ms_per_launch is their mean",
        "bound": "tensor",
        "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
        # dram bytes and tensor-pipe activity of one launch come from the committed ncu --set full capture of this
        # kernel (profiles/ncu_kernels.json, written by scripts/summarize_ncu.py), never from a constant in this file
        "traffic": ncu.get("dram_bytes"), "tensor_pipe_active_pct_ncu": ncu.get("tensor_pipe_active_pct"),
        "ncu_source": ncu.get("source"),
        "peak_source": f"{pk['source']} bf16 burst (MEASURED_PEAKS.json)",
        "algorithmic_flops_per_launch": flops_wsum, "ms_per_launch": ms_wsum, "precision": impl_name,
        "note": "algorithmic fp32 flops (2*E*rows*cols) over a bf16 dense peak; fp32-class accuracy costs 3 split "
                "products plus the score recompute, so executed tensor flops are 6x the algorithmic figure and the "
                "ceiling for frac is 1/6 (weighted sum) or 1/3 (row stats)",
        "executed_tflops": 6.0 * achieved, "executed_frac_of_peak": 6.0 * achieved / pk["bf16_tflops"],
        "ms_per_launch_fused_forward": ms_fused, "ms_per_launch_shift_by_swept": ms_wsum_swept,
        "row_stats_ms_per_launch": ms_stats,
        "sweeps_share_of_step": (ms_fused + ms_wsum_swept) / ms if tc_path else None,
        "row_stats_achieved_tflops": 2.0 * EMBED * n_own * n_all / (ms_stats * 1e-3) / 1e12,
        # one ex2 per score in each sweep; MUFU peak = 16 per clock per SM
        "exp_per_s_in_sweep": float(n_own) * n_all / (ms_wsum * 1e-3),
        "mufu_peak_per_s": 16.0 * 148 * 1.965e9,
        "mufu_frac": float(n_own) * n_all / (ms_wsum * 1e-3) / (16.0 * 148 * 1.965e9),
    }

    # ---- extras: correctness under sharding, strong scaling, the other BASELINE configs ----------------------
    extras = {}
    if not args.no_extras:
        for name, fn in (("parity_check", extra_parity_check), ("strong_scaling", extra_strong_scaling),
                         ("configs", extra_configs)):
            try:
                extras[name] = fn(args, world, rank, dev, flush)
            except Exception as e:                                     # an extra never takes the headline down
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            sync_all()

    # ---- CPU baseline (rank 0, N = 1 only): the reference on a bounded sample, in a child process so that its
    # import shims (identity .cuda(), thread settings) never touch this process ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import subprocess
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                "--warmup", "1", "--no-cfg1"], capture_output=True, text=True, timeout=600)
            cpu = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])["cpu_baseline"]
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.scaling == "strong" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, n_own, world, GPU_PRECISION),
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(2 * n_own * D_COMMON * 4), "d2h_bytes_per_step": 4,
                    "api": "VMIEstimator.forward + backward per step on inputs copied from pinned host memory (copy of step k+1 overlaps step k on a copy stream), mi of every step read back to the host"},
            "gpu_launches": int(launches), "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
            "roofline": roofline, "clocks": clk.summary(),
        }
        line.update(extras)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cfg1", action="store_true", help="reference arm: skip the BASELINE configs[0] CPU step")
    ap.add_argument("--no-extras", action="store_true", help="skip parity_check / strong_scaling / configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU work fixed (rows_per_gpu * B = 65536^2); strong: B = 65536 at every N")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_gpu(args)


if __name__ == "__main__":
    # stdout carries exactly one JSON line: libraries that write to fd 1 themselves (NCCL prints its version
    # there) are sent to stderr for the duration of the run
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*a, **k):          # noqa: A001  (the two JSON prints above resolve this name at call time)
        _buf.append(" ".join(str(x) for x in a))

    try:
        main()
    finally:
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
        for _l in _buf:
            _print(_l, flush=True)
