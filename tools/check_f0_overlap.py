"""4-min vocal through CoverEngine.convert with the F0 estimate on the main stream (B200VC_F0_OVERLAP=0 order) and on a side
stream overlapped with the HuBERT / index half of the four segments: the int16 utterances must be bit-identical; prints the
device time of both schedules."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aicovergen_b200.vc_infer_pipeline as vcmod  # noqa: E402
from aicovergen_b200.main import CoverEngine  # noqa: E402

hsd, rsd, cpt, mdx_w = bench.bench_checkpoints()
eng = CoverEngine(mdx_w, hsd, rsd, cpt, index=None, device="cuda:0")
song = torch.from_numpy(bench.synth_song(240.0, 1)).cuda()
vocal = eng.separate(song)["dereverb"]


def run(flag, reps=3):
    vcmod.F0_OVERLAP = flag
    outs, ms = [], []
    for _ in range(reps):
        eng.vc.set_noise_seed(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.convert(vocal, return_device=True)
        e1.record()
        torch.cuda.synchronize()
        outs.append(out.clone())
        ms.append(e0.elapsed_time(e1))
    return outs, ms


base, t0 = run(False)
over, t1 = run(True)
same = all(torch.equal(base[0], o) for o in base[1:] + over)
print(f"F0 on the main stream: {['%.1f' % m for m in t0]} ms; overlapped on a side stream: {['%.1f' % m for m in t1]} ms; "
      f"utterances bit-identical: {same} ({base[0].numel()} samples)", flush=True)
sys.exit(0 if same else 1)
