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
