"""CUDA path vs the golden vectors generated from the unmodified reference (tools/make_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

from siggen import stereo_tones, vocal_like

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tf32_storage(monkeypatch):
    """Golden fixtures were generated with the round-1 raw random checkpoints: pin TF32 storage for the MDX U-Net (raw
    BatchNorm statistics exceed fp16's range; the default fp16 mode is tested on trained-like checkpoints)."""
    import aicovergen_b200.mdx as bm
    monkeypatch.setattr(bm, "MDX_FP16", False)
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_synth_vs_golden():
    from aicovergen_b200.synth import SynthesizerB200
    from aicovergen_b200.synthetic import make_rvc_checkpoint
    from oracle import synth as osyn

    z = np.load(os.path.join(G, "synth_v2_40k.npz"))
    cpt = make_rvc_checkpoint("40k", "v2", seed=int(z["ckpt_seed"]))
    P = z["phone"].shape[1]
    nz, ns = osyn.draw_noise(int(z["noise_seed"]), P, 192, 400)
    net = SynthesizerB200(cpt, "cuda:0")
    o = net.infer(torch.from_numpy(z["phone"]).cuda(), torch.tensor([P]).cuda(), torch.from_numpy(z["pitch"]).cuda(),
                  torch.from_numpy(z["pitchf"]).cuda(), torch.tensor([0]).cuda(), noise_z=nz.cuda(), noise_src=ns.cuda())[0]
    e = float(np.sqrt(((o.cpu().numpy()[0, 0].astype(np.float64) - z["out"]) ** 2).mean()))
    print(f"[golden synth] abs rms err {e:.3e}")
    assert e < 1e-3


def test_rmvpe_vs_golden():
    from aicovergen_b200.rmvpe import RMVPEB200
    from aicovergen_b200.synthetic import make_rmvpe_state_dict
    from oracle import rmvpe as orm

    z = np.load(os.path.join(G, "rmvpe.npz"))
    x = vocal_like(float(z["seconds"]), seed=int(z["audio_seed"]))
    f0 = RMVPEB200(make_rmvpe_state_dict(seed=int(z["ckpt_seed"])), device="cuda:0").infer_from_audio(x, 0.03)
    assert np.array_equal(orm.coarse_pitch(f0)[0], orm.coarse_pitch(z["f0"])[0])


def test_pipeline_vs_golden():
    from aicovergen_b200.hubert import HubertB200
    from aicovergen_b200.rmvpe import RMVPEB200
    from aicovergen_b200.synth import SynthesizerB200
    from aicovergen_b200.synthetic import make_hubert_state_dict, make_rmvpe_state_dict, make_rvc_checkpoint
    from aicovergen_b200.vc_infer_pipeline import VC

    z = np.load(os.path.join(G, "vc_pipeline.npz"))
    audio = vocal_like(float(z["seconds"]), seed=int(z["audio_seed"]))
    xs = {k: int(z[k]) for k in ("x_pad", "x_query", "x_center", "x_max")}
    vc = VC(40000, types.SimpleNamespace(device="cuda:0", is_half=True, **xs))
    vc.model_rmvpe = RMVPEB200(make_rmvpe_state_dict(seed=4321), device="cuda:0")
    vc.set_noise_seed(int(z["noise_seed"]))
    out = vc.pipeline(HubertB200(make_hubert_state_dict(seed=777), "cuda:0"), SynthesizerB200(make_rvc_checkpoint("40k", "v2", seed=1234), "cuda:0"),
                      0, audio, "x.wav", [0, 0, 0], 0, "rmvpe", "", 0.5, 1, 3, 40000, 0, 0.25, "v2", 0.33, 128)
    d = (out.astype(np.float64) - z["out_int16"].astype(np.float64)) / 32768.0
    e = float(np.sqrt((d ** 2).mean()))
    print(f"[golden pipeline] int16-domain abs rms err {e:.3e}, max {np.abs(d).max():.3e}")
    assert out.shape == z["out_int16"].shape and e < 1e-3


def test_mdx_vs_golden():
    from aicovergen_b200.mdx import MDX, MDXModel
    from aicovergen_b200.synthetic import make_mdx_state_dict

    z = np.load(os.path.join(G, "mdx_small.npz"))
    dim_f, dim_t, n_fft = int(z["dim_f"]), int(z["dim_t"]), int(z["n_fft"])
    sd = make_mdx_state_dict(dim_f=dim_f, dim_t=dim_t, g=8, n=3, seed=int(z["ckpt_seed"]))
    wave = stereo_tones(int(z["n"]), seed=int(z["wave_seed"]))
    got = MDX(sd, MDXModel("cuda:0", dim_f, dim_t, n_fft), 0).process_wave(wave.copy(), 2)
    e = float(np.sqrt(((got.astype(np.float64) - z["processed"]) ** 2).mean()))
    r = float(np.sqrt((z["processed"].astype(np.float64) ** 2).mean()))
    print(f"[golden mdx] abs rms err {e:.3e} (ref rms {r:.3e})")
    assert got.shape == z["processed"].shape and e < 1e-3
