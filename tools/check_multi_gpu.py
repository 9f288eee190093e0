"""torchrun --nproc-per-node N tools/check_multi_gpu.py : the sharded MDX sweep (chunks split over ranks, one NCCL
all-reduce) against the unsharded sweep computed on the same rank.  Prints one line per rank; exit 1 on mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicovergen_b200.mdx import MDX, MDXModel, run_mdx_device  # noqa: E402
from aicovergen_b200.synthetic import make_mdx_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = f"cuda:{local}"
dim_f, dim_t, n_fft = 512, 64, 2048
model = MDXModel(dev, dim_f, dim_t, n_fft, stem_name="Vocals", compensation=1.035)
sess = MDX(make_mdx_state_dict(dim_f=dim_f, dim_t=dim_t, g=16, n=3), model, local)
g = torch.Generator().manual_seed(0)
wave = (torch.randn(2, 44100 * 20, generator=g) * 0.1).to(dev)
ref_main, ref_inv = run_mdx_device(sess, wave, denoise=True)
sh_main, sh_inv = run_mdx_device(sess, wave, denoise=True, group=dist.group.WORLD)
torch.cuda.synchronize()
err = float((sh_main - ref_main).abs().max()), float((sh_inv - ref_inv).abs().max())
scale = float(ref_main.abs().max())
print(f"rank {rank}/{world}: sharded-vs-single max abs diff main {err[0]:.3e} inverse {err[1]:.3e} (signal peak {scale:.3e})", flush=True)
ok = err[0] <= 1e-6 * max(scale, 1.0) + 1e-7 and err[1] <= 1e-5
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
