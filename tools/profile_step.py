"""One cover-pipeline step between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --seconds 30
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tapgemm_tc -c 3 -o gpurun_out/prof \
      python tools/profile_step.py --seconds 30
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--stage", default="all", choices=["all", "mdx", "rvc"])
args = ap.parse_args()

eng = bench.build_engine("cuda:0")
song = torch.from_numpy(bench.synth_song(args.seconds, 0)).cuda()


def step():
    if args.stage == "all":
        eng.cover_device(song)
    elif args.stage == "mdx":
        eng.separate(song)
    else:
        eng.convert(song)


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
