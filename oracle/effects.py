"""TEST INFRASTRUCTURE ONLY — Python face of oracle/effects.c (the CPU restatement of `add_audio_effects`, src/main.py:206-226;
PARITY UNPINNED, see the header of effects.c).  `build()` compiles the C file with gcc into oracle/_build/ (git-ignored, shipped
to the GPU box with the tree); `add_audio_effects` runs it."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "effects.c"
OUT = HERE / "_build" / "liboracle_effects.so"
BLOCK = 8192                      # Pedalboard.process hands blocks of this many samples to each plugin


def build(force: bool = False) -> Path:
    stamp = OUT.with_suffix(".sha")
    digest = hashlib.sha256(SRC.read_bytes()).hexdigest()
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(OUT), str(SRC), "-lm"], check=True)
    stamp.write_text(digest)
    return OUT


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        f = C.c_float
        _LIB.oracle_add_audio_effects_mono.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, f, f, f, f, f, f, f, f, f, f,
                                                       C.c_int, C.c_void_p]
        _LIB.oracle_add_audio_effects_mono.restype = C.c_int
    return _LIB


def add_audio_effects(x_i16: np.ndarray, sample_rate: int, reverb_rm_size: float, reverb_wet: float, reverb_dry: float,
                      reverb_damping: float, return_stages: bool = False) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """The int16 samples `add_audio_effects` writes for the int16 mono samples it reads; stages [3, n] float32 = after the
    high-pass, after the compressor, after the reverb (before the 16-bit conversion)."""
    x = np.ascontiguousarray(x_i16, dtype=np.int16)
    if x.ndim != 1:
        raise NotImplementedError("mono only (the RVC output)")
    out = np.empty_like(x)
    stages = np.empty((3, x.size), dtype=np.float32) if return_stages else None
    rc = _lib().oracle_add_audio_effects_mono(x.ctypes.data, out.ctypes.data, x.size, int(sample_rate), 50.0, -15.0, 4.0, 1.0, 100.0,
                                              float(reverb_rm_size), float(reverb_damping), float(reverb_wet), float(reverb_dry), 1.0,
                                              BLOCK, None if stages is None else stages.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle effects: allocation failed")
    return out, stages
