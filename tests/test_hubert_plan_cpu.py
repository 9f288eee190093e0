"""The HuBERT launch plan (aicovergen_b200/hubert.py) executed on the CPU through the descriptor emulator and compared with the
oracle (oracle/hubert.py, pinned against transformers.HubertModel): checks, without a GPU, the lowering of the 7-layer conv
front-end (strided Conv1d as GEMMs over [T/s, s*C] views), the grouped positional convolution, the fused q|k projection with the
1/sqrt(d_k) scaling folded in, V computed transposed, the batched attention products and the residual / LayerNorm wiring."""
import pytest
import torch
import torch.nn.functional as F

from aicovergen_b200 import ops
from aicovergen_b200 import tapgemm as tg
from aicovergen_b200.hubert import HubertB200, _HubertPlan
from aicovergen_b200.synthetic import make_hubert_state_dict
from emu import _act, _rn_tf32, emulate
from oracle import hubert as ohub


def _rnd(v, on):
    return _rn_tf32(v) if on else v


def _conv1d_from1(src, w, out, stride, src_off, bias=None, res=None, out2=None, act2=0, act2_p=0.0, round_out2=False):
    T, C = out.shape
    K = w.shape[1]
    frames = torch.as_strided(src, (T, K), (stride, 1), src_off)
    v = frames @ w.t()
    if bias is not None:
        v = v + bias
    if res is not None:
        v = v + res
    out.copy_(v)
    if out2 is not None:
        out2.copy_(_rnd(_act(v, act2, act2_p), round_out2))


def _groupnorm_time(x, gamma, beta, out, stats, eps=1e-5, act_code=0, round_out=False):
    mu = x.double().mean(0)
    var = x.double().var(0, unbiased=False)
    v = ((x.double() - mu) / torch.sqrt(var + eps)).float() * gamma + beta
    out.copy_(_rnd(_act(v, act_code, 0.0), round_out))


def _layernorm(x, gamma, beta, out, res=None, eps=1e-5, round_out=False):
    v = x if res is None else x + res
    out.copy_(_rnd(F.layer_norm(v, (v.shape[1],), gamma, beta, eps), round_out))


def _softmax_rows(S, T, q=None, emb_rel_k=None, window=0, round_out=False, row0=0):
    assert q is None
    S[:, :, :T] = _rnd(torch.softmax(S[:, :, :T], dim=-1), round_out)


@pytest.mark.parametrize("backend,layers", [(tg.BACKEND_SIMT, 12), (tg.BACKEND_TC, 9)])
def test_hubert_plan_matches_oracle_on_cpu(backend, layers, monkeypatch):
    sd = make_hubert_state_dict()
    g = torch.Generator().manual_seed(3)
    L = 6000                                           # 18 frames
    wav = 0.3 * torch.randn(1, L, generator=g)
    net = HubertB200(sd, device="cpu", backend=backend)
    for name, fn in (("conv1d_from1", _conv1d_from1), ("groupnorm_time", _groupnorm_time), ("layernorm", _layernorm),
                     ("softmax_rows", _softmax_rows)):
        monkeypatch.setattr(ops, name, fn)
    pl = _HubertPlan(net, L, layers)
    pl.wav[:L].copy_(wav.reshape(-1))
    n_gemm = 0
    for st in pl.steps:
        if isinstance(st, tg.TapGemm):
            n_gemm += 1
            emulate(st)
        else:
            st()
    assert n_gemm >= 6 + 1 + 16 + 7 * layers           # front-end, projection, positional conv groups, 7 GEMMs per layer
    ref = ohub.extract_features(sd, wav, layers)[0]
    got = pl.x.view(pl.T, -1)
    assert got.shape == ref.shape
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < (1e-5 if backend == tg.BACKEND_SIMT else 3e-3), err          # measured 1.1e-6 / 7.6e-4
