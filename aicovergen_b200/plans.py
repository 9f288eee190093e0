"""Per-shape launch plans are cached (LRU).  A 4-minute song cuts into 4 segments of different lengths
(vc_infer_pipeline.py:516-545), so the capacity must exceed the segment count or every segment of every song rebuilds
its plan (buffers + a few hundred prepared launch descriptors)."""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable, Hashable

PLAN_CACHE = int(os.environ.get("B200VC_PLAN_CACHE", "8"))


class PlanCache(OrderedDict):
    def get_or_build(self, key: Hashable, build: Callable[[], object]):
        plan = self.get(key)
        if plan is not None:
            self.move_to_end(key)
            return plan
        while len(self) >= PLAN_CACHE:
            self.popitem(last=False)
        plan = build()
        self[key] = plan
        return plan


# ---------------------------------------------------------------------------------------------------------------
# CUDA-graph replay of a plan's launch sequence
# ---------------------------------------------------------------------------------------------------------------
# A plan is a fixed list of launches over buffers it owns, so after its first (eager) run the whole sequence is captured
# once into a CUDA graph and replayed: no Python, no ctypes crossing, no cuTensorMapEncodeTiled per launch (the tensor
# maps are __grid_constant__ kernel parameters, i.e. part of the captured node).  VC-side kernels of 20-25 us were
# launch-bound before.  B200VC_CUDA_GRAPHS=0 turns it off; `graphs_enabled(False)` suspends replay
# (bench.py's per-launch event profiler needs the eager path).
_GRAPHS = os.environ.get("B200VC_CUDA_GRAPHS", "1") == "1"
_suspended = False


def graphs_enabled(on=None) -> bool:
    global _suspended
    if on is not None:
        _suspended = not on
    return _GRAPHS and not _suspended


class StepGraph:
    """Runs `steps` (callables that only enqueue work on the current stream over fixed buffers): eagerly the first time
    (lazy module loading, cudaFuncSetAttribute), then captured into a torch.cuda.CUDAGraph and replayed."""

    def __init__(self, steps):
        self.steps = steps
        self.graph = None
        self.calls = 0
        self.launches = 0

    def eager(self):
        for st in self.steps:
            st()

    def __call__(self):
        import torch

        from . import _ffi
        if not graphs_enabled():
            return self.eager()
        if self.graph is None:
            self.calls += 1
            if self.calls < 2:
                return self.eager()
            l0 = _ffi.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.eager()
            self.launches = _ffi.launch_count() - l0       # counted while capturing; nothing ran yet
            _ffi.lib().b200vc_count_launches(-self.launches)
            self.graph = g
        self.graph.replay()
        _ffi.lib().b200vc_count_launches(self.launches)
