"""B200-native RMVPE F0 estimator — drop-in for the reference `RMVPE` object at the plug point
`vc.model_rmvpe.infer_from_audio(x, thred=0.03)` (vc_infer_pipeline.py:322-329; rmvpe.py:328-409).

log-mel front-end as a windowed-DFT GEMM, the DeepUnet as NHWC tap-GEMM convolutions with
BatchNorm folded into the weights, the BiGRU as a cluster kernel, and the cents decode in the
reference's numpy summation order — all through libb200vc.so.  Default backend is the exact
fp32 SIMT GEMM: the decoded F0 feeds an argmax / coarse-pitch quantiser whose indices must
match the reference bit for bit (BASELINE.json north_star), so TF32 is opt-in here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _ffi, ops
from . import tapgemm as tg
from .plans import PlanCache, StepGraph
from .tapgemm import Epi

BN_EPS = 1e-5
N_FFT, HOP, N_MELS, N_CLASS = 1024, 160, 128, 360


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_basis(sr=16000, n_fft=N_FFT, n_mels=N_MELS, fmin=30.0, fmax=8000.0) -> np.ndarray:
    """librosa.filters.mel(htk=True, norm='slaney') as configured at rmvpe.py:277-284, 343-345."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, sr / 2.0, n_bins)
    mel_f = mel_to_hz_htk(np.linspace(hz_to_mel_htk(fmin), hz_to_mel_htk(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def _fold_bn(w: torch.Tensor, sd, p: str, out_dim: int = 0):
    """conv (no bias) followed by eval BatchNorm -> (w', b')."""
    g, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
    mu, var = sd[p + ".running_mean"].float(), sd[p + ".running_var"].float()
    s = g / torch.sqrt(var + BN_EPS)
    shape = [1] * w.dim()
    shape[out_dim] = -1
    return w.float() * s.reshape(shape), b - mu * s


class RMVPEB200:
    def __init__(self, model, is_half: bool = False, device: str = "cuda:0", backend: int = tg.BACKEND_SIMT,
                 n_blocks: int = 4, n_enc: int = 5, n_inter: int = 4):
        """`model`: path to rmvpe.pt or an already-loaded state dict (E2E(4,1,(2,2)) names)."""
        sd = torch.load(model, map_location="cpu") if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__") else model
        self.device = torch.device(device)
        self.is_half = is_half          # kept for interface parity; arithmetic is fp32 (or TF32 if backend=TC)
        self.backend = backend
        self.n_blocks, self.n_enc, self.n_inter = n_blocks, n_enc, n_inter
        self._plans = PlanCache()
        self._load(sd)

    def _dev(self, t):
        return t.float().contiguous().to(self.device)

    def _block_weights(self, sd, p: str, key: str):
        W = self.W
        w1, b1 = _fold_bn(sd[p + "conv.0.weight"], sd, p + "conv.1")
        w2, b2 = _fold_bn(sd[p + "conv.3.weight"], sd, p + "conv.4")
        W[key + ".w1"], W[key + ".b1"] = self._dev(tg.pack_conv2d(w1)), self._dev(b1)
        W[key + ".w2"], W[key + ".b2"] = self._dev(tg.pack_conv2d(w2)), self._dev(b2)
        if p + "shortcut.weight" in sd:
            W[key + ".ws"] = self._dev(sd[p + "shortcut.weight"][:, :, 0, 0])
            W[key + ".bs"] = self._dev(sd[p + "shortcut.bias"])

    def _load(self, sd):
        self.W = W = {}
        # windowed real-DFT basis: rows [cos(0..512) | -sin(0..512)] * hann(periodic), K = 1024
        n = torch.arange(N_FFT, dtype=torch.float64)
        k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
        win = torch.hann_window(N_FFT, periodic=True, dtype=torch.float64)
        ang = 2 * math.pi * k[:, None] * n[None, :] / N_FFT
        basis = torch.cat([torch.cos(ang) * win, -torch.sin(ang) * win], 0)      # [1026, 1024]
        W["dft"] = self._dev(basis)
        mb = torch.zeros(N_MELS, 516)
        mb[:, :513] = torch.from_numpy(mel_basis())
        W["mel"] = self._dev(mb)
        bn = "unet.encoder.bn"
        s = float(sd[bn + ".weight"][0] / torch.sqrt(sd[bn + ".running_var"][0] + BN_EPS))
        self.bn_a, self.bn_b = s, float(sd[bn + ".bias"][0] - sd[bn + ".running_mean"][0] * s)
        for i in range(self.n_enc):
            for b in range(self.n_blocks):
                self._block_weights(sd, f"unet.encoder.layers.{i}.conv.{b}.", f"enc{i}.{b}")
        for i in range(self.n_inter):
            for b in range(self.n_blocks):
                self._block_weights(sd, f"unet.intermediate.layers.{i}.conv.{b}.", f"mid{i}.{b}")
        for i in range(self.n_enc):
            p = f"unet.decoder.layers.{i}."
            wt, bt = _fold_bn(sd[p + "conv1.0.weight"], sd, p + "conv1.1", out_dim=1)
            W[f"dec{i}.up.w"], W[f"dec{i}.up.b"] = self._dev(tg.pack_convt2d(wt)), self._dev(bt)
            for b in range(self.n_blocks):
                self._block_weights(sd, p + f"conv2.{b}.", f"dec{i}.{b}")
        W["cnn.w"], W["cnn.b"] = self._dev(tg.pack_conv2d(sd["cnn.weight"])), self._dev(sd["cnn.bias"])
        # GRU: input features arrive as (w, c) = w*3 + c (NHWC flatten); the reference uses c*128 + w
        Hh = sd["fc.0.gru.weight_hh_l0"].shape[1]
        self.hidden = Hh
        wi = []
        for suf in ("", "_reverse"):
            w = sd[f"fc.0.gru.weight_ih_l0{suf}"].float()                       # [3H, 3*128] (c-major)
            wi.append(w.view(3 * Hh, 3, N_MELS).permute(0, 2, 1).reshape(3 * Hh, 3 * N_MELS))
        W["gru.wih"] = self._dev(torch.cat(wi, 0))                               # [2*3H, 384]
        W["gru.bih"] = self._dev(torch.cat([sd["fc.0.gru.bias_ih_l0"], sd["fc.0.gru.bias_ih_l0_reverse"]]))
        W["gru.whh"] = self._dev(torch.stack([sd["fc.0.gru.weight_hh_l0"], sd["fc.0.gru.weight_hh_l0_reverse"]]))
        W["gru.bhh"] = self._dev(torch.stack([sd["fc.0.gru.bias_hh_l0"], sd["fc.0.gru.bias_hh_l0_reverse"]]))
        W["fc.w"], W["fc.b"] = self._dev(sd["fc.1.weight"]), self._dev(sd["fc.1.bias"])

    # ------------------------------------------------------------------
    def _plan(self, n_samples: int) -> "_RmvpePlan":
        return self._plans.get_or_build(n_samples, lambda: _RmvpePlan(self, n_samples))

    @_ffi.on_device
    @torch.no_grad()
    def salience_from_audio(self, audio: torch.Tensor) -> torch.Tensor:
        """Device tensor [n_frames, 360] (rmvpe.py:370-373)."""
        pl = self._plan(int(audio.numel()))
        pl.run(audio)
        return pl.sal[:pl.n_frames]

    @_ffi.on_device
    @torch.no_grad()
    def mel2hidden(self, mel: torch.Tensor) -> torch.Tensor:
        """Reference signature (rmvpe.py:350-357): log-mel [1,128,n] (or [128,n]) -> salience [1,n,360] (device tensor)."""
        m2 = mel.reshape(N_MELS, -1)
        n = int(m2.shape[1])
        pl = self._plan((n - 1) * HOP)
        pl.run_from_logmel(m2.t().to(self.device).float().contiguous())
        return pl.sal[:n].unsqueeze(0)

    def decode(self, hidden, thred: float = 0.03) -> np.ndarray:
        """Reference signature (rmvpe.py:359-364): salience [n,360] (numpy or tensor) -> f0 [n] float64."""
        with torch.cuda.device(self.device):
            h = torch.as_tensor(hidden).to(self.device).float().contiguous()
            n = int(h.shape[0])
            cents = torch.empty(n, device=self.device, dtype=torch.float64)
            f0d = torch.empty(n, device=self.device, dtype=torch.float64)
            ops.rmvpe_decode(h, f0d, n, thred, cents=cents)
            c = cents.cpu().numpy()
        f0 = 10 * (2 ** (c / 1200))
        f0[f0 == 10] = 0
        return f0

    @_ffi.on_device
    @torch.no_grad()
    def infer_from_audio_device(self, audio: torch.Tensor, thred: float = 0.03) -> torch.Tensor:
        """f0 [n_frames] float64 on the device (no host sync)."""
        pl = self._plan(int(audio.numel()))
        pl.run(audio)
        ops.rmvpe_decode(pl.sal, pl.f0, pl.n_frames, thred)
        return pl.f0

    @_ffi.on_device
    def infer_from_audio(self, audio: np.ndarray, thred: float = 0.03) -> np.ndarray:
        """Reference signature (rmvpe.py:366-383): np.ndarray[N] -> np.ndarray[1 + N//160] (float64 Hz, 0 = unvoiced)."""
        if isinstance(audio, torch.Tensor):
            a = audio.detach().to(self.device).float().contiguous()      # extension: device tensor in, no H2D
        else:
            a = torch.from_numpy(np.ascontiguousarray(audio)).float().to(self.device)
        pl = self._plan(int(a.numel()))
        pl.run(a)
        ops.rmvpe_decode(pl.sal, pl.f0, pl.n_frames, thred, cents=pl.cents)
        cents_pred = pl.cents.cpu().numpy()
        # the reference's own last two numpy lines (rmvpe.py:361-362), on the host so f0 is bit-identical
        f0 = 10 * (2 ** (cents_pred / 1200))
        f0[f0 == 10] = 0
        return f0


class _RmvpePlan:
    def __init__(self, m: RMVPEB200, n_samples: int):
        dev, W, be = m.device, m.W, m.backend
        R = be == tg.BACKEND_TC
        f32 = dict(device=dev, dtype=torch.float32)
        steps: List = []
        add = steps.append
        self.n_samples = n_samples
        nf = 1 + n_samples // HOP
        self.n_frames = nf
        T = 32 * ((nf - 1) // 32 + 1)
        self.T = T
        pad = N_FFT // 2
        self.audio = torch.empty(n_samples, **f32)
        padded = torch.zeros(n_samples + 2 * pad + 8, **f32)
        add(lambda: ops.reflect_pad_1d(self.audio, padded, pad))
        # ---- log-mel: frames (overlapping rows, stride HOP) x windowed DFT basis, magnitude, mel GEMM
        spec = torch.empty(nf, 2 * (N_FFT // 2 + 1), **f32)
        frames = tg.View(padded, (N_FFT, nf, 1, 1, 1), (1, HOP, 0, 0, 0))
        add(tg.TapGemm(frames, tg.weights(W["dft"]), [(0, 0, 0, 0, 0)], (nf, 1, 1), tg.out_of(spec), None, be, name="stft"))
        mag = torch.empty(nf, 516, **f32)
        add(lambda: ops.magnitude(spec, mag, N_FFT // 2 + 1))
        melp = torch.empty(nf, N_MELS, **f32)
        add(tg.linear(mag, W["mel"], melp, None, be, name="mel"))
        img = torch.empty(1, T, N_MELS, 1, **f32)        # NHWC: H = frames, W = mel bins, C = 1
        self.n_front = len(steps)                        # steps before this one build `melp` from audio
        self.melp, self.mel_is_log = melp, False
        add(lambda: ops.logmel_affine_reflect(melp, img.view(T, N_MELS), nf, -1.0 if self.mel_is_log else 1e-5, m.bn_a, m.bn_b))

        def block(x, key, out, cin, cout, H_, W_):
            t1 = torch.empty(1, H_, W_, cout, **f32)
            add(tg.conv2d(x, W[key + ".w1"], t1, 3, 3, (1, 1), Epi(bias=W[key + ".b1"], act_pre=tg.ACT_RELU, round_out=R), be, name=key + ".c1"))
            if key + ".ws" in W:
                sc = torch.empty(1, H_, W_, cout, **f32)
                add(tg.linear(x, W[key + ".ws"], sc, Epi(bias=W[key + ".bs"]), be, name=key + ".sc"))
            else:
                sc = x
            add(tg.conv2d(t1, W[key + ".w2"], out, 3, 3, (1, 1), Epi(bias=W[key + ".b2"], act_pre=tg.ACT_RELU, res=sc), be, name=key + ".c2"))

        # ---- encoder (rmvpe.py:61-119): level output goes straight into the decoder's concat buffer
        x = img
        H_, W_, cin, cout = T, N_MELS, 1, 16
        cats = []
        for i in range(m.n_enc):
            cat = torch.empty(1, H_, W_, 2 * cout, **f32)
            cats.append((cat, H_, W_, cout))
            for b in range(m.n_blocks):
                last = b == m.n_blocks - 1
                out = cat[..., cout:] if last else torch.empty(1, H_, W_, cout, **f32)
                block(x, f"enc{i}.{b}", out, cin if b == 0 else cout, cout, H_, W_)
                x = out
            pooled = torch.empty(1, H_ // 2, W_ // 2, cout, **f32)
            add(lambda x=x, pooled=pooled: ops.avgpool2x2(x, pooled))
            x = pooled
            H_, W_, cin, cout = H_ // 2, W_ // 2, cout, cout * 2
        # ---- intermediate (rmvpe.py:122-138)
        for i in range(m.n_inter):
            for b in range(m.n_blocks):
                out = torch.empty(1, H_, W_, cout, **f32)
                block(x, f"mid{i}.{b}", out, cin if (i == 0 and b == 0) else cout, cout, H_, W_)
                x = out
        # ---- decoder (rmvpe.py:141-187)
        for i in range(m.n_enc):
            cat, Hc, Wc, cc = cats[-1 - i]
            for op in tg.conv_transpose2d_s2(x, W[f"dec{i}.up.w"], cat[..., :cc], 3, 1,
                                             Epi(bias=W[f"dec{i}.up.b"], act_pre=tg.ACT_RELU), be, name=f"dec{i}.up"):
                add(op)
            x = cat
            for b in range(m.n_blocks):
                out = torch.empty(1, Hc, Wc, cc, **f32)
                block(x, f"dec{i}.{b}", out, 2 * cc if b == 0 else cc, cc, Hc, Wc)
                x = out
        # ---- head (rmvpe.py:241-258)
        feat = torch.empty(1, T, N_MELS, 3, **f32)
        add(tg.conv2d(x, W["cnn.w"], feat, 3, 3, (1, 1), Epi(bias=W["cnn.b"]), be, name="cnn"))
        Hh = m.hidden
        xp = torch.empty(T, 2 * 3 * Hh, **f32)
        add(tg.linear(feat.view(T, 3 * N_MELS), W["gru.wih"], xp, Epi(bias=W["gru.bih"]), be, name="gru.in"))
        hseq = torch.empty(T, 2 * Hh, **f32)
        add(lambda: ops.bigru(xp, W["gru.whh"], W["gru.bhh"], hseq, Hh))
        self.sal = torch.empty(T, N_CLASS, **f32)
        add(tg.linear(hseq, W["fc.w"], self.sal, Epi(bias=W["fc.b"], act_pre=tg.ACT_SIGMOID), be, name="fc"))
        self.f0 = torch.empty(nf, device=dev, dtype=torch.float64)
        self.cents = torch.empty(nf, device=dev, dtype=torch.float64)
        self.steps = steps
        self.graph = StepGraph(steps)

    def run(self, audio: torch.Tensor):
        self.audio.copy_(audio.reshape(-1))
        self.mel_is_log = False
        self.graph()

    def run_from_logmel(self, mel_t: torch.Tensor):
        """mel_t [n_frames, 128] log-mel (what MelSpectrogram.forward returns, transposed): skips the front-end."""
        self.melp.copy_(mel_t)
        self.mel_is_log = True
        for st in self.steps[self.n_front:]:
            st()
        self.mel_is_log = False
