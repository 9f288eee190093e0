"""GPU parity of the effects / mix row (reference src/main.py:206-233) through the C ABI:
effects kernels against the sequential oracle (oracle/effects.c: float stages <= 2e-6 abs, int16 within 1 LSB at rounding
boundaries), the pydub mix bit-exact against oracle/mixdown.py (CPython's audioop), PCM_16 quantisation of the stems."""
import numpy as np
import pytest
import torch

from aicovergen_b200 import effects as fx
from oracle import effects as oe
from oracle import mixdown as om
from test_effects_cpu import _vocal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sr,seconds,params", [(40000, 3.0, (0.15, 0.2, 0.8, 0.7)), (48000, 2.0, (0.5, 0.33, 0.4, 0.5)),
                                               (32000, 2.0, (0.9, 0.6, 0.3, 1.0)), (40000, 240.0, (0.15, 0.2, 0.8, 0.7))])
def test_effects_match_sequential_oracle(sr, seconds, params):
    x = _vocal(sr, seconds, seed=sr + int(seconds))
    room, wet, dry, damping = params
    want16, stages = oe.add_audio_effects(x, sr, room, wet, dry, damping, return_stages=True)
    xd = torch.from_numpy(x).cuda()
    got16, gotf = fx.add_audio_effects_device(xd, sr, room, wet, dry, damping, return_float=True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fx.add_audio_effects_device(xd, sr, room, wet, dry, damping)
    ev1.record()
    torch.cuda.synchronize()
    got16, gotf = got16.cpu().numpy(), gotf.cpu().numpy()
    ef = float(np.abs(gotf - stages[2]).max())
    d = np.abs(got16.astype(np.int32) - want16.astype(np.int32))
    print(f"[effects sr {sr} {seconds:.0f} s room {room}] float max abs err {ef:.2e} (rms {float(np.sqrt((stages[2] ** 2).mean())):.3f}); "
          f"int16 max diff {d.max()} on {float((d > 0).mean()):.5f} of the samples; {ev0.elapsed_time(ev1):.2f} ms")
    assert np.isfinite(gotf).all() and ef < 2e-6
    assert d.max() <= 1 and (d > 0).mean() < 0.01


@pytest.mark.parametrize("sr_main,ch_main", [(40000, 1), (48000, 1), (32000, 1), (44100, 2), (44100, 1)])
def test_pydub_mix_bit_exact(sr_main, ch_main):
    rng = np.random.default_rng(sr_main + ch_main)
    n_main = int(sr_main * 2.3) + 17
    main = rng.integers(-32768, 32768, (n_main, ch_main) if ch_main > 1 else n_main).astype(np.int16)
    backup = rng.integers(-32768, 32768, (int(44100 * 2.1) + 5, 2)).astype(np.int16)      # shorter than the vocal
    inst = rng.integers(-32768, 32768, (int(44100 * 2.6) + 3, 2)).astype(np.int16)        # longer than the vocal
    for gains in ((0, 0, 0), (3, -2, 5), (12, 9, 9)):                                       # the last one clips
        want, rate = om.combine_audio(main, sr_main, backup, 44100, inst, 44100, *gains)
        got, r = fx.combine_audio_device([(torch.from_numpy(main).cuda(), sr_main), (torch.from_numpy(backup).cuda(), 44100),
                                          (torch.from_numpy(inst).cuda(), 44100)], *gains)
        assert r == rate and tuple(got.shape) == (want.shape[0], 2)
        assert np.array_equal(got.cpu().numpy(), want.reshape(want.shape[0], -1))


def test_pydub_mix_full_size_and_refusals():
    # 4-min cover: 9.6 M vocal frames at 40 kHz against 10.584 M-frame stems
    rng = np.random.default_rng(7)
    main = rng.integers(-20000, 20000, 9_600_000).astype(np.int16)
    backup = rng.integers(-20000, 20000, (10_584_000, 2)).astype(np.int16)
    inst = rng.integers(-20000, 20000, (10_584_000, 2)).astype(np.int16)
    want, rate = om.combine_audio(main, 40000, backup, 44100, inst, 44100, 1, 2, 3)
    got, r = fx.combine_audio_device([(torch.from_numpy(main).cuda(), 40000), (torch.from_numpy(backup).cuda(), 44100),
                                      (torch.from_numpy(inst).cuda(), 44100)], 1, 2, 3)
    assert r == rate == 44100 and got.shape[0] == 10_584_000 and np.array_equal(got.cpu().numpy(), want)
    with pytest.raises(ValueError):
        fx.combine_audio_device([(torch.zeros(4, device="cuda"), 40000)] * 3)               # not int16
    with pytest.raises(NotImplementedError):
        fx.add_audio_effects_device(torch.zeros(8, 2, dtype=torch.int16, device="cuda"), 40000, 0.15, 0.2, 0.8, 0.7)


def test_pcm16_from_planar():
    from oracle import dsp as odsp
    rng = np.random.default_rng(3)
    x = rng.uniform(-1.2, 1.2, (2, 100_003)).astype(np.float32)                # beyond +-1: saturates
    x[0, :8] = [0.5 / 32768, 0.99999 / 32768, 1.0 / 32768, -0.5 / 32768, 1.0, -1.0, 0.99999994, 65535.7 / 2 ** 31]
    got = fx.pcm16_from_planar(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, odsp.pcm16_soundfile(x).T)
