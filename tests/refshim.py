"""Helpers to run the UNMODIFIED reference (from /root/reference) with duck-typed operators.
Only usable where /root/reference exists (build container). TEST INFRASTRUCTURE ONLY."""
import types

import numpy as np
import torch

from oracle import hubert as ohub
from oracle import ref_import


class HubertShim:
    """Duck-typed fairseq HubertModel backed by the oracle restatement (fairseq itself is not installable)."""

    def __init__(self, sd):
        self.sd = sd

    def extract_features(self, source, padding_mask, output_layer):
        return (ohub.extract_features(self.sd, source.float(), output_layer), padding_mask)

    def final_proj(self, x):
        return ohub.final_proj(self.sd, x)


def ref_net_g(cpt):
    m = ref_import.module("infer_pack.models")
    cls = m.SynthesizerTrnMs768NSFsid if cpt.get("version", "v1") == "v2" else m.SynthesizerTrnMs256NSFsid
    net = cls(*cpt["config"], is_half=False)
    del net.enc_q
    net.load_state_dict(cpt["weight"], strict=False)
    return net.eval().float()


def ref_rmvpe(sd):
    r = ref_import.module("rmvpe")
    model = r.E2E(4, 1, (2, 2))
    model.load_state_dict(sd)
    model.eval()
    rm = r.RMVPE.__new__(r.RMVPE)
    rm.model, rm.is_half, rm.device, rm.resample_kernel = model, False, "cpu", {}
    rm.mel_extractor = r.MelSpectrogram(False, 128, 16000, 1024, 160, None, 30, 8000)
    rm.cents_mapping = np.pad(20 * np.arange(360) + 1997.3794084376191, (4, 4))
    return rm


def ref_vc(tgt_sr, x_pad=3, x_query=10, x_center=60, x_max=65):
    v = ref_import.module("vc_infer_pipeline")
    cfg = types.SimpleNamespace(device="cpu", is_half=False, x_pad=x_pad, x_query=x_query, x_center=x_center, x_max=x_max)
    return v, v.VC(tgt_sr, cfg)


def vocal_like(seconds, sr=16000, seed=7):
    """SURVEY.md §8(d) cfg-3 style synthetic vocal: harmonic stack with vibrato, unvoiced bursts, noise floor."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f0 = 220.0 * 2 ** (0.5 * np.sin(2 * np.pi * 0.2 * t)) * 2 ** (30 / 1200 * np.sin(2 * np.pi * 5.5 * t))
    phase = 2 * np.pi * np.cumsum(f0) / sr
    x = sum(np.sin(k * phase) / k for k in range(1, 9))
    burst = ((t % 3.0) > 2.6)
    x = np.where(burst, rng.standard_normal(n) * 0.7, x)
    x = x + 0.1 * rng.standard_normal(n)
    # a short near-silent gap every ~1.7 s gives the cut-point search something to find
    x = x * np.where((t % 1.7) > 1.62, 0.02, 1.0)
    return (0.5 * x / np.abs(x).max()).astype(np.float32)
