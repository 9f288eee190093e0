"""A few representative tap-GEMM launches for `ncu --set full` (one GPU):
   ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels python tools/ncu_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicovergen_b200 import tapgemm as tg  # noqa: E402

cases = [
    # name, T, Cin, Cout, k, dil, backends
    ("voc.s3 C64 k7", 1319600, 64, 64, 7, 3, [tg.BACKEND_TC_TILE, tg.BACKEND_TC_WS]),
    ("voc.s1 C256 k11", 65980, 256, 256, 11, 5, [tg.BACKEND_TC_TILE]),
    ("hubert.ffn1", 3299, 768, 3072, 1, 1, [tg.BACKEND_TC_TILE]),
]
for name, T, Ci, Co, k, d, bes in cases:
    x = torch.randn(T, Ci, device="cuda")
    w = torch.randn(k, Co, Ci, device="cuda") / (Ci * k) ** 0.5
    b = torch.randn(Co, device="cuda")
    out = torch.empty(T, Co, device="cuda")
    for be in bes:
        op = tg.conv1d(x, w, out, dilation=d, epi=tg.Epi(bias=b, act_pre=tg.ACT_LRELU, act_pre_p=0.1), backend=be)
        for _ in range(2):
            op()
        torch.cuda.synchronize()
# MDX 2D c=48
x = torch.randn(1, 256, 3072, 48, device="cuda")
w = torch.randn(9, 48, 48, device="cuda") / (9 * 48) ** 0.5
out = torch.empty(1, 256, 3072, 48, device="cuda")
for be in (tg.BACKEND_TC_TILE, tg.BACKEND_TC_WS):
    op = tg.conv2d(x, w, out, 3, 3, (1, 1), tg.Epi(act_pre=tg.ACT_RELU), backend=be)
    for _ in range(2):
        op()
    torch.cuda.synchronize()
print("done")
