"""Seeded synthetic test signals shared by the fixture generator and the tests (no reference needed)."""
import numpy as np


def vocal_like(seconds, sr=16000, seed=7):
    """SURVEY.md §8(d) cfg-3 style synthetic vocal: harmonic stack with vibrato, unvoiced bursts, noise floor,
    short near-silent gaps (so the cut-point search has something to find)."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f0 = 220.0 * 2 ** (0.5 * np.sin(2 * np.pi * 0.2 * t)) * 2 ** (30 / 1200 * np.sin(2 * np.pi * 5.5 * t))
    phase = 2 * np.pi * np.cumsum(f0) / sr
    x = sum(np.sin(k * phase) / k for k in range(1, 9))
    burst = ((t % 3.0) > 2.6)
    x = np.where(burst, rng.standard_normal(n) * 0.7, x)
    x = x + 0.1 * rng.standard_normal(n)
    x = x * np.where((t % 1.7) > 1.62, 0.02, 1.0)
    return (0.5 * x / np.abs(x).max()).astype(np.float32)


def stereo_tones(n, seed=4):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 44100.0
    wave = np.stack([0.3 * np.sin(2 * np.pi * 330 * t), 0.3 * np.sin(2 * np.pi * 440 * t + 1)]) + 0.05 * rng.standard_normal((2, n))
    return wave.astype(np.float32)


def song_44k(seconds, seed=0, sr=44100):
    """SURVEY.md §8(d) cfg-4/5 style stereo song: pink-ish noise bed + chord tones with slow AM + a vocal-like harmonic
    line with vibrato and unvoiced bursts, two decorrelated channels, peak 0.9.  float32 [2, n]."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n, dtype=np.float64) / sr
    out = np.zeros((2, n))
    f0 = 220.0 * 2 ** (0.5 * np.sin(2 * np.pi * 0.2 * t)) * 2 ** (30 / 1200 * np.sin(2 * np.pi * 5.5 * t))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    vocal = sum(np.sin(k * ph) / k for k in range(1, 9))
    vocal = np.where((t % 3.0) > 2.6, rng.standard_normal(n) * 0.5, vocal)
    vocal *= np.where((t % 7.3) > 6.9, 0.02, 1.0)
    for ch in range(2):
        spec = np.fft.rfft(rng.standard_normal(n))
        spec /= np.sqrt(np.maximum(np.arange(len(spec)), 1.0))
        bed = np.fft.irfft(spec, n)
        bed *= 0.25 / np.abs(bed).max()
        chord = sum(np.sin(2 * np.pi * f * (1 + 0.002 * ch) * t + ch) for f in (130.8, 164.8, 196.0))
        chord *= 0.15 * (0.6 + 0.4 * np.sin(2 * np.pi * 0.25 * t + ch))
        out[ch] = bed + chord + 0.35 * vocal * (1.0 - 0.1 * ch)
    out *= 0.9 / np.abs(out).max()
    return out.astype(np.float32)
