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
# CPU arm: the oracle port of the reference path (numpy, all host threads)
# --------------------------------------------------------------------------


def cpu_step_fn(B, seed=0):
    from oracle import params as P
    from oracle import vmi_oracle as O
    prm = P.vmi_params(seed, "separate", "constant", D_COMMON, HIDDEN, EMBED, LAYERS)
    x, y = P.features(seed + 1, B, D_COMMON, corr=0.6)
    return lambda: O.separable_infonce_streamed(prm, x, y, dtype=np.float32, block=2048)


def time_cpu(B, steps, warmup):
    fn = cpu_step_fn(B)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The
    reference is pure Python/PyTorch and cannot travel to the GPU box, so this
    arm times the oracle port (numpy restatement pinned to the reference by
    tests/golden) on a bounded sample: B = 8192 rows of the B = 65536 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8192
    sec = time_cpu(B, args.steps, args.warmup)
    value = B * B / sec
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "separable-critic InfoNCE fwd+bwd, d_common=128 hidden=256, CPU sample B=8192 of B=65536",
                   "global_batch": B, "d_common": D_COMMON},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"B={B} rows (B^2 = {B * B} pairs per step) of the B=65536 workload, numpy fp32, row blocks of 2048"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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

    B = args.batch or global_batch(world)
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

    # ---- dominant kernel, timed alone with CUDA events on its own stream -----
    with torch.no_grad():
        xe_, ye_ = est.critic_model.embed(x, y)
        xe_, ye_ = xe_.contiguous(), ye_.contiguous()
        all_x = RB.all_gather_rows(xe_, rb)
    n_all = all_x.shape[0]
    ws = torch.empty(L.lib.mimrl_sep_workspace_bytes(n_own, n_all, EMBED) + 16, dtype=torch.uint8, device=dev)
    stats = torch.empty(4, n_own, device=dev)
    st = L.stream()

    def k_stats():
        L.check(L.lib.mimrl_sep_row_stats(L.ptr(ye_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, 0, 0, L.ptr(stats[0]),
                                          L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(stats[3]), L.ptr(ws), ws.numel(), st))
    k_stats()
    shift = (stats[0] + torch.log(stats[1])).contiguous()
    coef = torch.full((1,), -1.0 / n_all, device=dev)
    dcoef = torch.full((n_own,), 1.0 / n_all, device=dev)
    out = torch.empty_like(ye_)

    def k_wsum():
        L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(ye_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, 0, 1, L.ptr(shift),
                                             0, L.ptr(coef), L.ptr(dcoef), 0, L.ptr(out), L.ptr(ws), ws.numel(), st))

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

    def k_wsum_swept():          # the second backward sweep: operands swapped, shift indexed by the swept row
        L.check(L.lib.mimrl_sep_weighted_sum(L.ptr(xe_), L.ptr(all_x), n_own, n_all, EMBED, rb.offset, 0, 1,
                                             L.ptr(shift_all), 1, L.ptr(coef), L.ptr(dcoef), 0, L.ptr(out), L.ptr(ws),
                                             ws.numel(), st))
    ms_wsum_own = time_kernel(k_wsum)
    ms_wsum_swept = time_kernel(k_wsum_swept)
    ms_wsum = 0.5 * (ms_wsum_own + ms_wsum_swept)
    ms_stats = time_kernel(k_stats)
    pk = peaks()
    flops_wsum = 2.0 * EMBED * n_own * n_all              # algorithmic: one P.X contraction (score recompute not counted)
    achieved = flops_wsum / (ms_wsum * 1e-3) / 1e12
    impl_name = ("tcgen05 fp16x3 split (fp32-class)"
                 if L.lib.mimrl_sep_selected_impl(n_own, n_all, EMBED, 0) == L.IMPL_TCGEN05 else "fp32 FFMA (CUDA cores)")
    roofline = {
        "kernel": "sep_wsum_tc_kernel via mimrl_sep_weighted_sum / mimrl_sep_fused_forward (2 launches per step: the "
                  "fused forward sweep and the swept-side gradient sweep)", "bound": "tensor",
        "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, profiles/sep_kernels_r1.md
        # (operands are L2-resident fp16 hi/lo copies; the 32 MiB partial-output buffer dominates the writes)
        "traffic": 127134976 if (world == 1 and B == B_SINGLE) else None,
        "peak_source": f"{pk['source']} bf16 burst (MEASURED_PEAKS.json)",
        "algorithmic_flops_per_launch": flops_wsum, "ms_per_launch": ms_wsum, "precision": impl_name,
        "note": "algorithmic fp32 flops (2*E*rows*cols) over a bf16 dense peak; fp32-class accuracy costs 3 split "
                "products plus the score recompute, so executed tensor flops are 6x the algorithmic figure and the "
                "ceiling for frac is 1/6 (weighted sum) or 1/3 (row stats)",
        "executed_tflops": 6.0 * achieved, "executed_frac_of_peak": 6.0 * achieved / pk["bf16_tflops"],
        "tensor_pipe_active_pct_ncu": 76.4,           # profiles/sep_kernels_r1.md (74.0 with the row-sum epilogue, 76.4 without)
        "ms_per_launch_shift_by_own": ms_wsum_own, "ms_per_launch_shift_by_swept": ms_wsum_swept,
        "row_stats_ms_per_launch": ms_stats,
        "row_stats_achieved_tflops": 2.0 * EMBED * n_own * n_all / (ms_stats * 1e-3) / 1e12,
        # one ex2 per score in each sweep; MUFU peak = 16 per clock per SM
        "exp_per_s_in_sweep": float(n_own) * n_all / (ms_wsum * 1e-3),
        "mufu_peak_per_s": 16.0 * 148 * 1.965e9,
        "mufu_frac": float(n_own) * n_all / (ms_wsum * 1e-3) / (16.0 * 148 * 1.965e9),
    }

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on a bounded sample ---
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        Bc = 16384
        sec = time_cpu(Bc, 3, 1)
        cpu = {"value": Bc * Bc / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"B={Bc} rows ({Bc * Bc} pairs per step, 3 steps) of the B={B} workload; numpy fp32 oracle port"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "separable-critic InfoNCE fwd+bwd (VMIEstimator, Model.py:108-148), d_common=128 "
                                   "hidden=256 embed=128 layers=2, BASELINE configs[1] at its largest batch",
                       "global_batch": B, "rows_per_gpu": n_own, "d_common": D_COMMON, "parallelism": f"rowblock{world}",
                       "l2": "flushed between timed steps (256 MiB write, outside the per-step events)",
                       "timing": "CUDA events per step, mean over steps, max over ranks",
                       "precision": impl_name + "; TF32 disabled for the torch MLP GEMMs"},
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(2 * n_own * D_COMMON * 4), "d2h_bytes_per_step": 4,
                    "api": "VMIEstimator.forward + backward per step on inputs copied from pinned host memory (copy of step k+1 overlaps step k on a copy stream), mi of every step read back to the host"},
            "gpu_launches": int(launches), "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
            "roofline": roofline, "clocks": clk.summary(),
        }
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
