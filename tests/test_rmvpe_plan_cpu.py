"""The rmvpe U-Net launch plan in its tensor-core form (3xTF32 split operands: activations stored [hi | lo | hi], weights
[W_hi | W_hi | W_lo]) executed on CPU through the descriptor emulator and compared with the oracle's fp32 U-Net: checks the
K-concatenation, the split residuals, the shared concat buffers of the decoder and the split average pooling without a GPU."""
import pytest
import torch
import torch.nn.functional as F

from aicovergen_b200 import ops
from aicovergen_b200 import tapgemm as tg
from aicovergen_b200.rmvpe import RMVPEB200, split3_weights
from aicovergen_b200.synthetic import make_rmvpe_state_dict
from emu import emulate
from oracle import rmvpe as orm


def _avgpool_split(x, in_split, out, out_split):
    from emu import _flat, _rn_tf32
    B, H, W, C = x.shape
    full = x + torch.as_strided(x, x.shape, x.stride(), x.storage_offset() + in_split)
    v = 0.25 * (full[:, 0::2, 0::2] + full[:, 0::2, 1::2] + full[:, 1::2, 0::2] + full[:, 1::2, 1::2])
    hi = _rn_tf32(v)
    lo = _rn_tf32(v - hi)
    for plane, val in ((0, hi), (1, lo), (2, hi)):
        torch.as_strided(out, out.shape, out.stride(), out.storage_offset() + plane * out_split).copy_(val)


def test_split3_weights_reconstruct():
    w = torch.randn(9, 8, 12, generator=torch.Generator().manual_seed(0))
    s = split3_weights(w)
    assert s.shape == (9, 8, 36) and torch.equal(s[..., :12], s[..., 12:24])
    assert (s[..., :12] + s[..., 24:] - w).abs().max() < 2 ** -21 * w.abs().max()


@pytest.mark.parametrize("backend", [tg.BACKEND_TC, tg.BACKEND_SIMT])
def test_rmvpe_unet_plan_matches_oracle_on_cpu(backend, monkeypatch):
    cfg = dict(n_blocks=2, en_de_layers=2, inter_layers=1, en_out_channels=8)
    sd = make_rmvpe_state_dict(seed=5, **cfg)
    net = RMVPEB200(sd, device="cpu", backend=backend, n_blocks=2, n_enc=2, n_inter=1)
    monkeypatch.setattr(ops, "avgpool2x2_split", _avgpool_split)
    monkeypatch.setattr(ops, "avgpool2x2", lambda x, out: out.copy_(F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)))
    pl = net._plan(31 * 160)                       # 32 frames
    T = pl.T
    g = torch.Generator().manual_seed(1)
    mel = torch.randn(1, 128, T, generator=g)      # stands for the BatchNorm'd log-mel image
    pl.img.copy_(mel[0].t().reshape(1, T, 128, 1))
    n_tc = 0
    for st in pl.steps[pl.n_front + 1: pl.n_unet_end]:
        if isinstance(st, tg.TapGemm):
            n_tc += st.backend == tg.BACKEND_TC
            if st.backend == tg.BACKEND_TC:
                assert st.params.a_stride[1] % 4 == 0 and st.params.Kc % 4 == 0, st.name      # TMA-addressable
            emulate(st)
        else:
            st()
    assert (n_tc > 10) == (backend == tg.BACKEND_TC)
    # oracle U-Net on the same image (rmvpe.py:190-258 without the input BatchNorm, which the plan folds into the log-mel step)
    x = mel.transpose(-1, -2).unsqueeze(1)
    skips = []
    with torch.no_grad():
        for i in range(2):
            for b in range(2):
                x = orm._block(sd, f"unet.encoder.layers.{i}.conv.{b}.", x)
            skips.append(x)
            x = F.avg_pool2d(x, 2)
        for b in range(2):
            x = orm._block(sd, f"unet.intermediate.layers.0.conv.{b}.", x)
        for i in range(2):
            p = f"unet.decoder.layers.{i}."
            x = F.relu(orm._bn(sd, p + "conv1.1", F.conv_transpose2d(x, sd[p + "conv1.0.weight"], stride=2, padding=1, output_padding=1)))
            x = torch.cat((x, skips[-1 - i]), dim=1)
            for b in range(2):
                x = orm._block(sd, p + f"conv2.{b}.", x)
        ref = F.conv2d(x, sd["cnn.weight"], sd["cnn.bias"], padding=1)             # [1,3,T,128]
    got = pl.feat[0].permute(2, 0, 1)
    err = float((got - ref[0]).abs().max() / ref.abs().max())
    assert err < 2e-5, err
