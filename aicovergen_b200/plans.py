"""Per-shape launch plans are cached (LRU).  A 4-minute song cuts into 4 segments of different lengths
(vc_infer_pipeline.py:516-545), so the capacity must exceed the segment count or every segment of every song rebuilds
its plan (buffers + a few hundred prepared launch descriptors)."""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable, Hashable

PLAN_CACHE = int(os.environ.get("B200VC_PLAN_CACHE", "8"))

# Score scratch of one attention query block.  r02e measured that cutting a 66 s segment (T = 3299) into 512-row blocks sized for
# the L2 made HuBERT 35 % SLOWER (57 -> 77 ms per song: 7x the launches, small GEMMs, and the 81 MB block did not stay
# L2-resident in practice), so by default a segment is ONE block; the blocked path bounds memory for very long segments
# (scores of a 10-min segment would be 10 GB) and is what a fused flash-style kernel would replace.
ATT_SCRATCH_BYTES = int(os.environ.get("B200VC_ATT_SCRATCH_MB", "1024")) * 1024 * 1024



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
    """Runs `steps` (callables that only enqueue library launches on the current stream over fixed buffers).
    First call: the steps are RECORDED into a C-side launch plan (b200vc_plan_begin/_end: every entry point called in
    between is stored with its arguments instead of launched) and the plan is run natively — from then on one forward pass
    is ONE ctypes call (b200vc_plan_run), whatever the host language.  From the second call the plan run is additionally
    captured into a CUDA graph and replayed."""

    def __init__(self, steps):
        self.steps = steps
        self.plan = None
        self.graph = None
        self.calls = 0
        self.launches = 0

    def eager(self):
        for st in self.steps:
            st()

    def _record(self):
        import ctypes as C

        from . import _ffi
        lib = _ffi.lib()
        h = C.c_void_p()
        _ffi.check(lib.b200vc_plan_begin(C.byref(h)), "plan_begin")
        try:
            self.eager()
        finally:
            _ffi.check(lib.b200vc_plan_end(), "plan_end")
        self.plan = h
        self.launches = int(lib.b200vc_plan_size(h))

    def run_plan(self):
        import ctypes as C

        import torch

        from . import _ffi
        _ffi.check(_ffi.lib().b200vc_plan_run(self.plan, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "plan_run")

    def __call__(self):
        import torch

        from . import _ffi
        if not graphs_enabled():
            return self.eager()
        if self.plan is None:
            self._record()
        if self.graph is None:
            self.calls += 1
            if self.calls < 2:
                return self.run_plan()
            l0 = _ffi.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run_plan()
            _ffi.lib().b200vc_count_launches(l0 - _ffi.launch_count())     # counted while capturing; nothing ran yet
            self.graph = g
        self.graph.replay()
        _ffi.lib().b200vc_count_launches(self.launches)

    def __del__(self):
        try:
            if self.plan is not None:
                from . import _ffi
                _ffi.lib().b200vc_plan_destroy(self.plan)
        except Exception:
            pass
