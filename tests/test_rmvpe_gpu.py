"""GPU parity: B200 RMVPE (infer_from_audio plug point) vs the CPU oracle restatement (oracle/rmvpe.py,
pinned against src/rmvpe.py).  BASELINE.json config 1: 10 s 100->1000 Hz sine sweep at 16 kHz."""
import numpy as np
import pytest
import torch

from aicovergen_b200 import tapgemm as tg
from aicovergen_b200.synthetic import make_rmvpe_state_dict

pytestmark = pytest.mark.gpu


def sweep(seconds):
    t = np.arange(int(16000 * seconds)) / 16000.0
    return (0.5 * np.sin(2 * np.pi * (100 * t + 45 * t * t))).astype(np.float32)


@pytest.mark.parametrize("seconds", [2.93, 10.0])
def test_rmvpe_f0_parity(seconds):
    from aicovergen_b200.rmvpe import RMVPEB200
    from oracle import rmvpe as orm

    sd = make_rmvpe_state_dict()
    x = sweep(seconds)
    a = torch.from_numpy(x)[None]
    hid_ref = orm.mel2hidden(sd, orm.log_mel(a))[0]
    f0_ref = orm.decode(hid_ref.numpy().copy(), 0.03)
    pitch_ref, pitchf_ref = orm.coarse_pitch(f0_ref, 0)

    net = RMVPEB200(sd, device="cuda:0", backend=tg.BACKEND_SIMT)
    sal = net.salience_from_audio(torch.from_numpy(x).cuda()).cpu()
    err = (sal - hid_ref).abs().max().item()
    print(f"[rmvpe {seconds}s] salience max abs err {err:.3e} over {tuple(sal.shape)}")
    f0 = net.infer_from_audio(x, thred=0.03)
    assert f0.shape == f0_ref.shape == (1 + len(x) // 160,)
    pitch, pitchf = orm.coarse_pitch(f0, 0)
    mism = int((pitch != pitch_ref).sum())
    rel = np.abs(f0 - f0_ref) / np.maximum(f0_ref, 1e-9)
    print(f"[rmvpe {seconds}s] coarse-pitch mismatches {mism}/{len(pitch)}; f0 max rel diff {rel.max():.3e}")
    assert err < 3e-4
    assert mism == 0, "coarse pitch indices must match the reference bit for bit"
    assert rel.max() < 1e-4


def test_rmvpe_decode_kernel_exact():
    """The decode kernel alone reproduces numpy's float32/float64 summation order bit for bit."""
    from aicovergen_b200 import ops
    from oracle import rmvpe as orm

    g = torch.Generator().manual_seed(3)
    sal = torch.rand(2000, 360, generator=g)
    sal[5] = 0.01                      # below threshold
    sal[6, :] = 0.5                    # full tie -> first index
    sal[7, 359] = 2.0                  # edge window
    sal[8, 0] = 2.0
    ref = orm.decode(sal.numpy().copy(), 0.03)
    f0 = torch.empty(2000, dtype=torch.float64, device="cuda")
    cents = torch.empty(2000, dtype=torch.float64, device="cuda")
    ops.rmvpe_decode(sal.cuda(), f0, 2000, 0.03, cents=cents)
    got = f0.cpu().numpy()
    ulp = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    print(f"[rmvpe decode] device f0 max rel diff {ulp.max():.3e}, exact {(got == ref).mean():.4f}")
    assert ulp.max() < 1e-14
    # host finish (what infer_from_audio does) must be bit-identical to numpy
    c = cents.cpu().numpy()
    f0h = 10 * (2 ** (c / 1200))
    f0h[f0h == 10] = 0
    print(f"[rmvpe decode] host-finished f0 exact {(f0h == ref).mean():.4f}")
    assert np.array_equal(f0h, ref)
    assert (orm.coarse_pitch(got)[0] == orm.coarse_pitch(ref)[0]).all()
