"""GPU parity of the crepe F0 path (`f0_method="mangio-crepe"`, vc_infer_pipeline.py:96-137) against oracle/crepe.py — the
CPU restatement of torchcrepe 0.0.20's published algorithm (third-party, PARITY UNPINNED against torchcrepe itself)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from siggen import vocal_like  # noqa: E402

from aicovergen_b200.synthetic import make_crepe_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu


def test_crepe_activations_parity():
    """Framing + normalisation + the six conv layers + classifier: sigmoid activations [n, 360]."""
    from aicovergen_b200.crepe import CrepeB200
    from oracle import crepe as oc

    sd = make_crepe_state_dict()
    x = vocal_like(2.2, seed=4)
    x = x / np.quantile(np.abs(x), 0.999)
    hop = 128
    ref = torch.cat([oc.model(sd, f) for f in oc.frames_from_audio(x, hop).split(256)])
    net = CrepeB200(sd, "cuda:0", batch_frames=128)          # 276 frames -> 3 batches, the last one partial
    got = net.activations(torch.from_numpy(x).cuda(), hop).cpu()
    err = (got - ref).abs().max().item()
    print(f"[crepe] activations {tuple(got.shape)}: max abs err {err:.3e} (ref max {ref.max().item():.3f}, argmax agreement "
          f"{(got.argmax(1) == ref.argmax(1)).float().mean().item():.4f})")
    assert got.shape == ref.shape == (1 + len(x) // hop, 360)
    assert err < 5e-3


@pytest.mark.parametrize("n", [3, 2500])
def test_viterbi_kernel_matches_restatement(n):
    """b200vc_crepe_logprob + b200vc_viterbi_band against the numpy restatement of torchcrepe.postprocess + decode.viterbi
    (librosa.sequence.viterbi, banded triangular transition), incl. bins masked by fmin / fmax."""
    from aicovergen_b200.crepe import CrepeB200
    from oracle import crepe as oc

    g = torch.Generator().manual_seed(n)
    k = torch.arange(360.0)
    centre = 180 + 120 * torch.sin(torch.arange(n) * 0.01) + 30 * torch.randn(n, generator=g).cumsum(0) / max(n, 1) ** 0.5
    act = torch.sigmoid(6 * torch.exp(-0.5 * ((k[None] - centre[:, None]) / 3.0) ** 2) - 3 + 0.5 * torch.randn(n, 360, generator=g))
    net = CrepeB200(make_crepe_state_dict(calibrate=False), "cuda:0")
    got = net.viterbi_bins(act.cuda(), 50.0, 1100.0).cpu().numpy()
    ref = oc.decode_viterbi(act, 50.0, 1100.0)
    agree = float((got == ref).mean())
    print(f"[viterbi n={n}] bins agree {agree:.4f}; range {ref.min()}..{ref.max()} (allowed {oc.frequency_to_bins(50.0)}..{oc.frequency_to_bins(1100.0, True) - 1})")
    assert got.shape == ref.shape and agree == 1.0


def test_get_f0_mangio_crepe_and_pipeline():
    """`VC.get_f0(..., "mangio-crepe", crepe_hop_length)` against the oracle's get_f0_crepe_computation (dither off: torchcrepe's
    dither is random), then one `VC.pipeline` call with that F0 method."""
    from aicovergen_b200.crepe import CrepeB200
    from aicovergen_b200.hubert import HubertB200
    from aicovergen_b200.synth import SynthesizerB200
    from aicovergen_b200.synthetic import make_hubert_state_dict, make_rvc_checkpoint
    from aicovergen_b200.vc_infer_pipeline import VC
    from oracle import crepe as oc
    from oracle import rmvpe as orm

    sd = make_crepe_state_dict()
    audio = vocal_like(3.0, seed=6)
    vc = VC(40000, types.SimpleNamespace(device="cuda:0", is_half=True, x_pad=1, x_query=1, x_center=2, x_max=3))
    vc.model_crepe = CrepeB200(sd, "cuda:0")
    vc.crepe_dither = False
    pad = np.pad(audio.astype(np.float64), (16000, 16000), mode="reflect")
    p_len = len(pad) // 160
    pitch, pitchf = vc.get_f0("x", pad, p_len, 0, "mangio-crepe", 3, 128)
    f0_ref = oc.get_f0_crepe_computation(sd, pad.copy(), 50, 1100, p_len, 128, rng=None)
    want, wantf = orm.coarse_pitch(f0_ref, 0)
    agree = float((pitch == want).mean())
    print(f"[mangio-crepe] p_len {p_len}: coarse pitch agreement {agree:.4f}; {len(np.unique(want))} distinct levels; max |f0 diff| "
          f"{np.abs(pitchf - wantf).max():.3e} Hz")
    assert pitch.shape == want.shape == (p_len,) and agree >= 0.99
    # dithered output stays within +-20 cents of the bin centres and is reproducible for a seed
    vc.crepe_dither, vc.crepe_dither_seed = True, 3
    _, d1 = vc.get_f0("x", pad, p_len, 0, "mangio-crepe", 3, 128)
    _, d2 = vc.get_f0("x", pad, p_len, 0, "mangio-crepe", 3, 128)
    assert np.array_equal(d1, d2) and np.abs(1200 * np.log2(d1[pitchf > 0] / pitchf[pitchf > 0])).max() <= 20.5
    hub, net = HubertB200(make_hubert_state_dict(), "cuda:0"), SynthesizerB200(make_rvc_checkpoint("40k", "v2"), "cuda:0")
    out = vc.pipeline(hub, net, 0, audio.copy(), "x.wav", [0, 0, 0], 0, "mangio-crepe", "", 0.0, 1, 3, 40000, 0, 0.25, "v2", 0.33, 128)
    assert out.dtype == np.int16 and abs(out.shape[0] - len(audio) * 2.5) < 1000 and np.abs(out).max() > 100
