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
