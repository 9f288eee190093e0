"""CPU tests of the effects / mix row: the oracle's mix against CPython's audioop semantics, the host
geometry against the oracle, and the restructured recursions (tests/emu_effects.py mirrors csrc/effects.cu) against the
sequential oracle."""
import numpy as np
import pytest

import emu_effects as emu
from aicovergen_b200 import effects as fx
from oracle import effects as oe
from oracle import mixdown as om


def _vocal(sr, seconds, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * seconds)) / sr
    x = 0.35 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.9 * t)) + 0.02 * rng.standard_normal(t.size)
    x[: sr // 5] = 0.0                                         # leading silence: denormal / snap paths
    return np.clip(np.rint(x * 32767), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("sr_main,ch_main", [(40000, 1), (48000, 1), (32000, 1), (44100, 2)])
def test_mix_closed_form_matches_audioop(sr_main, ch_main):
    rng = np.random.default_rng(sr_main)
    n_main = int(sr_main * 1.3) + 17
    main = rng.integers(-32768, 32768, (n_main, ch_main) if ch_main > 1 else n_main).astype(np.int16)
    backup = rng.integers(-32768, 32768, (int(44100 * 1.1) + 5, 2)).astype(np.int16)      # shorter than the vocal
    inst = rng.integers(-32768, 32768, (int(44100 * 1.6) + 3, 2)).astype(np.int16)        # longer than the vocal
    for gains in ((0, 0, 0), (3, -2, 5), (12, 9, 9)):                                       # the last one clips
        want, rate = om.combine_audio(main, sr_main, backup, 44100, inst, 44100, *gains)
        r, used, n_out = fx.mix_geometry([main.shape[0], backup.shape[0], inst.shape[0]], [sr_main, 44100, 44100])
        assert (r, n_out) == (rate, want.shape[0])
        g = [(fx.db_to_float(b), fx.db_to_float(x)) for b, x in zip((-4, -6, -7), gains)]
        got = emu.pydub_mix([main, backup, inst], [sr_main, 44100, 44100], g, r, used, n_out)
        assert np.array_equal(got, want.reshape(n_out, -1))


def test_mix_geometry_of_the_4min_cover():
    # 4-min vocal at 40 kHz against 44.1 kHz stems: ratecv returns one frame less than 240 s and the millisecond slice pads it
    rate, used, n_out = fx.mix_geometry([9_600_000, 10_584_000, 10_584_000], [40000, 44100, 44100])
    assert (rate, used, n_out) == (44100, [10_583_999, 10_584_000, 10_584_000], 10_584_000)
    with pytest.raises(NotImplementedError):
        fx.mix_geometry([100, 100, 100], [40000, 44100, 48000])


@pytest.mark.parametrize("sr,params", [(40000, (0.15, 0.2, 0.8, 0.7)), (48000, (0.5, 0.33, 0.4, 0.5)), (32000, (0.9, 0.6, 0.3, 1.0))])
def test_restructured_effects_match_sequential_oracle(sr, params):
    x = _vocal(sr, 1.5, seed=sr)
    room, wet, dry, damping = params
    want16, stages = oe.add_audio_effects(x, sr, room, wet, dry, damping, return_stages=True)
    k = fx.effect_constants(sr, room, wet, dry, damping)
    warm = min(k.warm, (x.size + 7) // 8 * 8)
    got16, gotf, comp = emu.effects(x, k, chunk=max(2048, (warm // 4 + 7) // 8 * 8), warm=warm)       # as effects.py sizes them
    assert np.abs(comp - stages[1]).max() < 2e-6                    # high-pass + compressor
    assert np.abs(gotf - stages[2]).max() < 5e-6                    # + reverb, wet/dry
    d = np.abs(got16.astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.02                    # the 16-bit conversion can flip at a rounding boundary


def test_effect_constants_delay_lines():
    k = fx.effect_constants(40000, 0.15, 0.2, 0.8, 0.7)
    assert k.comb_delays == (1012, 1077, 1158, 1229, 1289, 1352, 1412, 1466) and k.allpass_delays == (504, 400, 309, 204)
    assert abs(k.feedback - 0.742) < 1e-6 and abs(k.damp - 0.28) < 1e-6 and abs(k.wet1 - 0.6) < 1e-6 and abs(k.dry - 1.6) < 1e-6
    assert k.warm % 1024 == 0 and 0.9984 ** k.warm < 1e-9


def test_mix_ragged_and_tiny_inputs():
    """Edge cases of the overlay geometry: one-frame operands, vocals shorter / longer than the stems, all four RVC rates, mono
    and stereo stems — closed form == CPython's audioop through the restated pydub glue, for 60 random cases."""
    rng = np.random.default_rng(11)
    for case in range(60):
        sr_main = int(rng.choice([32000, 40000, 44100, 48000]))
        n_main = int(rng.choice([1, 2, 3, 7, 45, 441, 1000, 4411, 20011]))
        n_b = int(rng.choice([1, 2, 44, 999, 4410, 30000]))
        n_i = int(rng.choice([1, 5, 441, 4409, 50000]))
        ch_b = int(rng.choice([1, 2]))
        mk = lambda n, ch: rng.integers(-32768, 32768, (n, ch) if ch > 1 else n).astype(np.int16)
        main, backup, inst = mk(n_main, 1), mk(n_b, ch_b), mk(n_i, 2)
        gains = tuple(float(g) for g in rng.choice([0, -3, 2.5, 6], 3))
        want, rate = om.combine_audio(main, sr_main, backup, 44100, inst, 44100, *gains)
        r, used, n_out = fx.mix_geometry([n_main, n_b, n_i], [sr_main, 44100, 44100])
        assert (r, n_out) == (rate, want.shape[0]), (case, sr_main, n_main, n_b, n_i)
        if n_out == 0:
            continue
        g = [(fx.db_to_float(b), fx.db_to_float(x)) for b, x in zip((-4, -6, -7), gains)]
        got = emu.pydub_mix([main, backup, inst], [sr_main, 44100, 44100], g, r, used, n_out)
        assert np.array_equal(got, want.reshape(n_out, -1)), (case, sr_main, n_main, n_b, n_i)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1013, 4099, 15361])
def test_restructured_effects_short_and_unaligned_lengths(n):
    """Lengths below one comb delay, below the warm-up, and not multiples of the 8-sample group the kernel walks."""
    x = _vocal(40000, 1.0, seed=n)[20000:20000 + n]
    want16, stages = oe.add_audio_effects(x, 40000, 0.15, 0.2, 0.8, 0.7, return_stages=True)
    k = fx.effect_constants(40000, 0.15, 0.2, 0.8, 0.7)
    warm = min(k.warm, (n + 7) // 8 * 8)
    got16, gotf, comp = emu.effects(x, k, chunk=max(2048, (warm // 4 + 7) // 8 * 8), warm=warm)
    assert np.abs(comp - stages[1]).max() < 1e-6 and np.abs(gotf - stages[2]).max() < 2e-6
    assert np.abs(got16.astype(np.int32) - want16.astype(np.int32)).max() <= 1
