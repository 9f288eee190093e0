"""Wall-clock (synchronised) breakdown of one cover-pipeline step into its stages.
usage (GPU box): python tools/stage_breakdown.py [--seconds 240] > gpurun_out/breakdown.json
Synchronising around every stage removes host/device overlap, so the sum is slightly above bench.py's step time;
use it for SHARES."""
import argparse
import collections
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--fine", action="store_true", help="keep layer indices in the per-GEMM table")
args = ap.parse_args()

eng = bench.build_engine("cuda:0")
song = torch.from_numpy(bench.synth_song(args.seconds, 0)).cuda()
acc = collections.OrderedDict()
enabled = [False]


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        if not enabled[0]:
            return fn(*a, **k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return r

    setattr(obj, name, timed)


import aicovergen_b200.main as M  # noqa: E402
from aicovergen_b200 import ops  # noqa: E402

orig_run = M.run_mdx_device


def run_mdx_timed(mdx, *a, **k):
    if not enabled[0]:
        return orig_run(mdx, *a, **k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = orig_run(mdx, *a, **k)
    torch.cuda.synchronize()
    lab = f"mdx[{mdx.model.dim_f}x{mdx.model.dim_t}]"
    acc[lab] = acc.get(lab, 0.0) + (time.perf_counter() - t0) * 1e3
    return r


M.run_mdx_device = run_mdx_timed
wrap(eng.hubert, "extract_features", "hubert")
wrap(eng.vc.model_rmvpe, "infer_from_audio", "rmvpe")
wrap(eng.net_g, "infer", "synthesizer")
wrap(eng.vc, "pipeline", "vc.pipeline(total)")
wrap(eng, "mix", "mix")
wrap(ops, "resample_sinc_mono", "resample")

# ---- per-GEMM timing by name (CUDA events around every tap-GEMM launch, all backends)
import re  # noqa: E402
from aicovergen_b200 import tapgemm as tg  # noqa: E402

gemm_rec = []
orig_call = tg.TapGemm.__call__


def timed_call(self, stream=None, backend=None):
    if not enabled[0]:
        return orig_call(self, stream, backend)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    orig_call(self, stream, backend)
    e.record()
    be = self.backend if backend is None else backend
    p_ = self.params
    gemm_rec.append((f"{self.name} [M={p_.OW * p_.OH * p_.OB} N={p_.N} K={p_.Kc}x{p_.ntaps}]" if args.fine else self.name, be, self.flops(), s, e))


tg.TapGemm.__call__ = timed_call

from aicovergen_b200 import plans  # noqa: E402

plans.graphs_enabled(False)       # per-launch events need the eager launch path (recorded plans / graphs bypass Python)
for _ in range(args.warmup):
    eng.cover_device(song)
enabled[0] = True
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.cover_device(song)
torch.cuda.synchronize()
total = (time.perf_counter() - t0) * 1e3
acc["vc.pipeline(glue: hpf, index, pad, rms, int16)"] = acc["vc.pipeline(total)"] - acc.get("hubert", 0) - acc.get("rmvpe", 0) - acc.get("synthesizer", 0)
by = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, be, fl, s_, e_ in gemm_rec:
    key = (name if args.fine else re.sub(r"\d+", "#", name)) + ("/simt" if be == tg.BACKEND_SIMT else "")
    by[key][0] += 1
    by[key][1] += s_.elapsed_time(e_)
    by[key][2] += fl
gemms = {k: {"launches": v[0], "ms": round(v[1], 2), "tflops": round(v[2] / v[1] / 1e9, 1) if v[1] > 0 else 0}
         for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])}
print(json.dumps({"seconds": args.seconds, "total_ms": total, "stages_ms": acc,
                  "gemm_ms_total": round(sum(v[1] for v in by.values()), 1), "gemms": gemms}, indent=1))
