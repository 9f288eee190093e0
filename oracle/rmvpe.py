"""TEST INFRASTRUCTURE ONLY — torch-CPU/numpy restatement of the RMVPE F0 estimator.

Follows rmvpe.py:261-325 (MelSpectrogram), :8-258 (E2E = DeepUnet + Conv2d + BiGRU + Linear + Sigmoid),
:350-409 (mel2hidden / decode / to_local_average_cents) and vc_infer_pipeline.py:346-370 (coarse pitch).
Pinned against the reference's own module by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .dsp import mel_filterbank

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5


def log_mel(audio: torch.Tensor) -> torch.Tensor:
    """MelSpectrogram(is_half=False, 128, 16000, 1024, 160, None, 30, 8000).forward(audio, center=True)
    (rmvpe.py:295-325 with keyshift=0). audio [1,N] -> [1,128,1+N//160]."""
    window = torch.hann_window(1024)
    fft = torch.stft(audio, n_fft=1024, hop_length=160, win_length=1024, window=window, center=True, return_complex=True)
    mag = torch.sqrt(fft.real.pow(2) + fft.imag.pow(2))
    basis = torch.from_numpy(mel_filterbank(16000, 1024, 128, 30, 8000)).float()
    return torch.log(torch.clamp(basis @ mag, min=1e-5))


def _bn(sd: SD, p: str, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _block(sd: SD, p: str, x):
    """ConvBlockRes (rmvpe.py:23-58)."""
    y = F.relu(_bn(sd, p + "conv.1", F.conv2d(x, sd[p + "conv.0.weight"], padding=1)))
    y = F.relu(_bn(sd, p + "conv.4", F.conv2d(y, sd[p + "conv.3.weight"], padding=1)))
    if p + "shortcut.weight" in sd:
        return y + F.conv2d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"])
    return y + x


def e2e(sd: SD, mel: torch.Tensor, n_blocks=4, n_enc=5, n_inter=4) -> torch.Tensor:
    """E2E.forward (rmvpe.py:254-258). mel [1,128,T] with T % 32 == 0 -> salience [1,T,360]."""
    x = mel.transpose(-1, -2).unsqueeze(1)
    x = _bn(sd, "unet.encoder.bn", x)
    skips = []
    for i in range(n_enc):
        for b in range(n_blocks):
            x = _block(sd, f"unet.encoder.layers.{i}.conv.{b}.", x)
        skips.append(x)
        x = F.avg_pool2d(x, 2)
    for i in range(n_inter):
        for b in range(n_blocks):
            x = _block(sd, f"unet.intermediate.layers.{i}.conv.{b}.", x)
    for i in range(n_enc):
        p = f"unet.decoder.layers.{i}."
        x = F.conv_transpose2d(x, sd[p + "conv1.0.weight"], stride=2, padding=1, output_padding=1)
        x = F.relu(_bn(sd, p + "conv1.1", x))
        x = torch.cat((x, skips[-1 - i]), dim=1)
        for b in range(n_blocks):
            x = _block(sd, p + f"conv2.{b}.", x)
    x = F.conv2d(x, sd["cnn.weight"], sd["cnn.bias"], padding=1)
    x = x.transpose(1, 2).flatten(-2)                      # [1,T,3*128], feature = c*128 + w
    h = bigru(sd, x)
    return torch.sigmoid(F.linear(h, sd["fc.1.weight"], sd["fc.1.bias"]))


def bigru(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """nn.GRU(384, 256, 1 layer, batch_first, bidirectional) (rmvpe.py:8-20) — the same torch operator the reference
    calls, loaded with the checkpoint's weights (gate order r, z, n)."""
    Hh = sd["fc.0.gru.weight_hh_l0"].shape[1]
    gru = torch.nn.GRU(x.shape[-1], Hh, num_layers=1, batch_first=True, bidirectional=True)
    gru.load_state_dict({k[len("fc.0.gru."):]: v for k, v in sd.items() if k.startswith("fc.0.gru.")})
    gru.eval()
    with torch.no_grad():
        return gru(x)[0]


def mel2hidden(sd: SD, mel: torch.Tensor) -> torch.Tensor:
    """rmvpe.py:350-357: reflect-pad frames to a multiple of 32, run the net, crop."""
    n = mel.shape[-1]
    mel = F.pad(mel, (0, 32 * ((n - 1) // 32 + 1) - n), mode="reflect")
    with torch.no_grad():
        return e2e(sd, mel)[:, :n]


CENTS = np.pad(20 * np.arange(360) + 1997.3794084376191, (4, 4))


def decode(hidden: np.ndarray, thred: float = 0.03) -> np.ndarray:
    """rmvpe.py:359-364, 385-409 — same numpy ops in the same order (float32 salience, float64 cents)."""
    center = np.argmax(hidden, axis=1)
    sal = np.pad(hidden, ((0, 0), (4, 4)))
    center += 4
    idx = center[:, None] + np.arange(-4, 5)[None, :]
    todo_sal = np.take_along_axis(sal, idx, axis=1)        # float32 [T,9]
    todo_cents = CENTS[idx]                                # float64 [T,9]
    product_sum = np.sum(todo_sal * todo_cents, 1)
    weight_sum = np.sum(todo_sal, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        cents = product_sum / weight_sum
    cents[np.max(sal, axis=1) <= thred] = 0
    f0 = 10 * (2 ** (cents / 1200))
    f0[f0 == 10] = 0
    return f0


def infer_from_audio(sd: SD, audio: np.ndarray, thred: float = 0.03) -> np.ndarray:
    """RMVPE.infer_from_audio (rmvpe.py:366-383), fp32 path."""
    a = torch.from_numpy(audio).float().unsqueeze(0)
    hidden = mel2hidden(sd, log_mel(a)).squeeze(0).numpy()
    return decode(hidden, thred)


def coarse_pitch(f0: np.ndarray, f0_up_key: float = 0):
    """vc_infer_pipeline.py:346-370 (no f0 file): key shift + mel quantisation to 1..255."""
    f0 = f0 * pow(2, f0_up_key / 12)
    f0_mel_min = 1127 * np.log(1 + 50 / 700)
    f0_mel_max = 1127 * np.log(1 + 1100 / 700)
    f0bak = f0.copy()
    f0_mel = 1127 * np.log(1 + f0 / 700)
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - f0_mel_min) * 254 / (f0_mel_max - f0_mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > 255] = 255
    return np.rint(f0_mel).astype(int), f0bak
