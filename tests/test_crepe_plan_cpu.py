"""The crepe launch plan (aicovergen_b200/crepe.py) executed on the CPU through the descriptor emulator and compared with the
oracle's restatement of torchcrepe's 'full' CNN: checks, without a GPU, the lowering of conv1 (512 taps, stride 4) to a GEMM over
overlapping frame rows, the k = 64 convolutions with their asymmetric (31, 32) padding, BatchNorm applied AFTER ReLU folded into
the max-pool pass, the position-major flatten in front of the classifier, and the frame buffer layout."""
import numpy as np
import pytest
import torch

from aicovergen_b200 import ops
from aicovergen_b200 import tapgemm as tg
from aicovergen_b200.crepe import FRAME_PITCH, CrepeB200, frequency_to_bins
from aicovergen_b200.synthetic import make_crepe_state_dict
from emu import emulate
from oracle import crepe as oc
from siggen import vocal_like


def _maxpool2_affine(c, s, t, x, rnd):
    from emu import _rn_tf32
    v = c * s + t                                         # BatchNorm (eval) after the ReLU that the conv epilogue applied
    v = torch.maximum(v[:, 0::2], v[:, 1::2])
    x.copy_(_rn_tf32(v) if rnd else v)


@pytest.mark.parametrize("backend", [tg.BACKEND_TC, tg.BACKEND_SIMT])
def test_crepe_plan_matches_oracle_on_cpu(backend, monkeypatch):
    sd = make_crepe_state_dict(calibrate=False)
    x = vocal_like(0.2, seed=2)
    x = x / np.quantile(np.abs(x), 0.999)
    frames = oc.frames_from_audio(x, 640)[:3]             # 3 normalised frames [3, 1024]
    net = CrepeB200(sd, device="cpu", backend=backend, batch_frames=3)
    monkeypatch.setattr(ops, "maxpool2_affine", _maxpool2_affine)
    from aicovergen_b200.crepe import _CrepePlan
    pl = _CrepePlan(net)
    assert pl.frames.shape == (3, FRAME_PITCH)
    pl.frames[:, 254:254 + 1024] = frames                 # what b200vc_crepe_frames writes (zero pads stay zero)
    n_gemm = 0
    for st in pl.steps:
        if isinstance(st, tg.TapGemm):
            n_gemm += 1
            emulate(st)
        else:
            st()
    assert n_gemm == 7                                    # six convolutions + the classifier
    ref = oc.model(sd, frames)
    err = float((pl.act - ref).abs().max())
    # TF32 operand rounding over reductions of up to 65 536 terms on raw random weights; the exact-fp32 plan is tight
    assert err < (2e-3 if backend == tg.BACKEND_TC else 2e-6), err        # measured 2.2e-4 / 3.0e-7
    assert torch.equal(pl.act.argmax(1), ref.argmax(1))


def test_frequency_to_bins_matches_restatement():
    for f in (50.0, 65.4, 440.0, 1100.0, 1975.5):
        for ceil in (False, True):
            assert frequency_to_bins(f, ceil) == oc.frequency_to_bins(f, ceil)
