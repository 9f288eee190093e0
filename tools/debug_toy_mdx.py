"""Toy-geometry MDX nets (as used by tests/test_call_surface_gpu.py and tools/check_strong_scaling.py) in fp16 and fp32 storage:
finite? parity vs the oracle?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import aicovergen_b200.mdx as bm  # noqa: E402
from aicovergen_b200.synthetic import make_mdx_trained_like  # noqa: E402
from oracle import mdx as om  # noqa: E402
from siggen import song_44k  # noqa: E402

wave = song_44k(12.0, seed=3)
for (dim_f, dim_t, n_fft, g) in ((512, 64, 2048, 16), (512, 128, 2048, 16), (256, 32, 2048, 8), (256, 64, 2048, 8)):
    sd = make_mdx_trained_like(dim_f, dim_t, n_fft, g=g, n=3)
    mp = om.MdxParams(dim_f, dim_t, n_fft)
    ref = om.process_wave(wave.copy(), mp, lambda s: om.convtdfnet(sd, s), 2)
    for half in (True, False):
        bm.MDX_FP16 = half
        sess = bm.MDX(sd, bm.MDXModel("cuda:0", dim_f, dim_t, n_fft), 0)
        got = sess.process_wave(wave.copy(), 2)
        fin = bool(np.isfinite(got).all())
        err = float(np.sqrt(np.nanmean((got - ref) ** 2)))
        print(f"dim_f {dim_f} dim_t {dim_t} n_fft {n_fft} g {g}: half={sess.ort.half} finite={fin} abs rms err {err:.3e} (ref rms {np.sqrt((ref**2).mean()):.3e}, ref finite {np.isfinite(ref).all()})", flush=True)
