"""CUDA-graph capture of the small-batch regime.

At the reference's training batch sizes (bs = 128 ... 1024) a stage-1 + stage-2 step is ~4400 kernel launches of a few
microseconds each: launch-bound on the host, not the GPU.  ``GraphedCallable`` captures a whole step (estimators
forward + backward + optimiser) once and replays it with one launch.

Host-side randomness stays on the host: ``prod_knn_sample`` draws its query ids from numpy's GLOBAL RNG exactly as the
reference does (Model.py:81); under capture the draw lands in a pinned buffer that the graph copies to the device,
and ``HostIdSource.refill()`` repeats the draws, in call order, before every replay -- so the RNG stream consumed
per step is the reference's.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


class HostIdSource:
    """Query-id draws of the k-NN samplers inside a captured region (one pinned buffer per sampler call)."""

    def __init__(self):
        self.calls: List[tuple] = []          # (N, m, pinned int64 buffer)
        self.recording = False

    def next(self, N: int, m: int, device) -> torch.Tensor:
        from .model import legacy_permutation_head
        ids = torch.from_numpy(legacy_permutation_head(N, m))
        if not self.recording:
            return ids.to(device, non_blocking=True)
        pinned = torch.empty(m, dtype=torch.int64).pin_memory()
        pinned.copy_(ids)
        self.calls.append((N, m, pinned))
        return pinned.to(device, non_blocking=True)          # a memcpy node of the graph

    def refill(self):
        from .model import legacy_permutation_head
        for N, m, pinned in self.calls:
            pinned.copy_(torch.from_numpy(legacy_permutation_head(N, m)))


class GraphedCallable:
    """``fn(*tensors) -> tensor | tuple of tensors`` captured into one CUDA graph.

    ``fn`` may run backward passes and optimiser steps (optimisers must be built with ``capturable=True``).
    Inputs are copied into static buffers before each replay; outputs are static tensors that the next replay
    overwrites (clone what has to outlive it)."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 id_source: Optional[HostIdSource] = None):
        from . import model as M
        self.fn, self.id_source = fn, id_source
        self.static_in = [t.detach().clone() for t in example_inputs]
        self.done = torch.cuda.Event()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        prev = M._ID_SOURCE
        M._ID_SOURCE = id_source
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fn(*self.static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            if id_source is not None:
                id_source.recording = True
            with torch.cuda.graph(self.graph):
                self.static_out = fn(*self.static_in)
        finally:
            if id_source is not None:
                id_source.recording = False
            M._ID_SOURCE = prev
        self.done.record()

    def __call__(self, *inputs):
        with torch.no_grad():
            for s, t in zip(self.static_in, inputs):
                if s.data_ptr() != t.data_ptr():
                    s.copy_(t, non_blocking=True)
        if self.id_source is not None:
            self.done.synchronize()            # the previous replay has consumed the pinned id buffers
            self.id_source.refill()
        self.graph.replay()
        self.done.record()
        return self.static_out
