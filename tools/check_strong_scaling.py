"""torchrun --nproc-per-node N tools/check_strong_scaling.py [--seconds S] [--full]

ONE song shared by the N ranks of one node (MDX chunk ranges per rank + one all-gather of the stem per
pass; RVC: F0 on rank 0 -> broadcast -> segments round-robin -> all-gather of the PCM) against the unsharded computation of
the same song on the same rank.  Every rank prints the max abs difference of each stem and of the cover (bit-identical is
expected: the same kernels run on the same data, only on different ranks) and the wall time of both forms; exit 1 on
mismatch.  Default: reduced geometry so it runs in seconds; --full = the bench's 3072-bin models and a 4-min song."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from aicovergen_b200.main import MDX_STAGES, CoverEngine  # noqa: E402
from aicovergen_b200.synthetic import (make_hubert_state_dict, make_mdx_trained_like, make_rmvpe_trained_like,  # noqa: E402
                                       make_rvc_checkpoint)
from siggen import song_44k  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=0.0)
ap.add_argument("--full", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = f"cuda:{local}"
if args.full:
    stages, kw, seconds = MDX_STAGES, {}, args.seconds or 240.0
    x_cfg = None
else:
    stages = (dict(name="a", dim_f=512, dim_t=64, n_fft=2048, stem="Vocals", compensate=1.021),
              dict(name="b", dim_f=256, dim_t=64, n_fft=2048, stem="Instrumental", compensate=1.035),
              dict(name="c", dim_f=512, dim_t=128, n_fft=2048, stem="Other", compensate=1.035))
    kw, seconds = dict(g=16, n=3), args.seconds or 45.0
mdx_w = [make_mdx_trained_like(s["dim_f"], s["dim_t"], s["n_fft"], seed=2024 + i, **kw) for i, s in enumerate(stages)]
eng = CoverEngine(mdx_w, make_hubert_state_dict(), make_rmvpe_trained_like(), make_rvc_checkpoint("40k", "v2"), index=None,
                  device=dev, mdx_stages=stages)
if not args.full:       # small segmentation constants so that a 45 s song already cuts into several RVC segments
    for k, v in dict(x_pad=1, x_query=2, x_center=8, x_max=10).items():
        setattr(eng.vc, k, v)
    vc = eng.vc
    vc.t_pad, vc.t_pad_tgt, vc.t_pad2 = 16000 * vc.x_pad, eng.tgt_sr * vc.x_pad, 32000 * vc.x_pad
    vc.t_query, vc.t_center, vc.t_max = 16000 * vc.x_query, 16000 * vc.x_center, 16000 * vc.x_max
song = torch.from_numpy(song_44k(seconds, seed=3)).to(dev)          # the SAME song on every rank


def timed(fn, reps=2):
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t = torch.tensor([best], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


def run(group):
    eng.vc.set_noise_seed(5)
    stems = eng.separate(song, group)
    eng.vc.group = group
    ai = eng.convert(stems["dereverb"], return_device=True)
    eng.vc.group = None
    cover = eng.mix(eng.effects(ai), stems["backup"], stems["instrumental"])
    return dict(stems, converted=ai.float(), cover=cover.float())


single, t1 = timed(lambda: run(None))
shard, tn = timed(lambda: run(dist.group.WORLD))
ok = True
parts = []
for k in ("vocals", "instrumental", "backup", "main", "dereverb", "converted", "cover"):
    same = torch.equal(single[k], shard[k])
    finite = bool(torch.isfinite(single[k]).all())
    d = float((single[k] - shard[k]).abs().max()) if finite else float("nan")
    parts.append(f"{k} {d:.1e}" + ("" if finite else " (non-finite stem!)"))
    ok = ok and same and finite
print(f"rank {rank}/{world}: sharded vs single max abs diff: {', '.join(parts)} | {seconds:.0f} s song: single {t1 * 1e3:.0f} ms, "
      f"sharded over {world} {tn * 1e3:.0f} ms (x{t1 / tn:.2f})", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
