"""Run the CoverEngine stages twice on the same song (eager first pass, recorded-plan / CUDA-graph replays after) and report
which stage outputs are not bit-identical between passes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from aicovergen_b200 import ops  # noqa: E402
from aicovergen_b200.main import MDX_STAGES, CoverEngine  # noqa: E402
from aicovergen_b200.synthetic import (make_hubert_state_dict, make_mdx_trained_like, make_rmvpe_trained_like,  # noqa: E402
                                       make_rvc_checkpoint)
from siggen import song_44k  # noqa: E402

mdx_w = [make_mdx_trained_like(s["dim_f"], s["dim_t"], s["n_fft"], seed=2024 + i) for i, s in enumerate(MDX_STAGES)]
eng = CoverEngine(mdx_w, make_hubert_state_dict(), make_rmvpe_trained_like(), make_rvc_checkpoint("40k", "v2"), index=None, device="cuda:0")
song = torch.from_numpy(song_44k(30.0, seed=1)).cuda()
runs = []
for it in range(4):
    stems = eng.separate(song)
    d = stems["dereverb"]
    mono = torch.empty(int(d.shape[1] * 16000 // 44100), device="cuda")
    ops.resample_sinc_mono(d.contiguous(), mono, 44100, 16000)
    feats = eng.hubert.extract_features(source=mono[None], padding_mask=None, output_layer=12)[0].clone()
    f0 = eng.vc.model_rmvpe.infer_from_audio(mono.cpu().numpy(), 0.03)
    eng.vc.set_noise_seed(5)
    ai = eng.convert(d, return_device=True).clone()
    runs.append(dict({k: v.clone() for k, v in stems.items()}, mono=mono.clone(), hubert=feats, f0=torch.from_numpy(f0), converted=ai.float()))
for it in range(1, 4):
    msg = []
    for k in runs[0]:
        a, b = runs[0][k], runs[it][k]
        dd = float((a.double() - b.double()).abs().max())
        msg.append(f"{k} {dd:.2e}")
    print(f"pass {it} vs pass 0: " + ", ".join(msg), flush=True)
# the same dereverb stem fed twice
eng.vc.set_noise_seed(5)
a1 = eng.convert(runs[0]["dereverb"], return_device=True).clone()
eng.vc.set_noise_seed(5)
a2 = eng.convert(runs[0]["dereverb"], return_device=True).clone()
print("convert twice on the SAME stem: max abs diff", int((a1.int() - a2.int()).abs().max()), flush=True)
