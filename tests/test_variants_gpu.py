"""GPU parity of the model / option variants the reference supports besides v2-768 / 40k / f0 (VERDICT r1 "untested
variants"): v1 (256-dim features + HuBERT layer 9 + final_proj), the no-pitch `_nono` synthesizers, the 32k / 48k / 32k_v2 /
48k_v2 upsample sets (src/configs/*.json), `f0_file`, `exact_hpf`; index edge cases (short and empty inverted lists)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from siggen import vocal_like  # noqa: E402

from aicovergen_b200.synthetic import (make_hubert_state_dict, make_ivf_index_data, make_rmvpe_trained_like,  # noqa: E402
                                       make_rvc_checkpoint)

pytestmark = pytest.mark.gpu


def rms(x):
    return float(np.sqrt(np.mean(np.asarray(x, dtype=np.float64) ** 2)))


@pytest.mark.parametrize("sr_key,version,f0", [("40k", "v1", 1), ("40k", "v2", 0), ("48k", "v1", 0), ("32k", "v2", 1),
                                               ("48k", "v2", 1), ("32k_v2", "v2", 1), ("48k_v2", "v2", 1)])
def test_synth_variants(sr_key, version, f0):
    """SynthesizerTrnMs{256,768}NSFsid[_nono].infer (models.py:532-751, 847-955) for every shipped upsample set."""
    from aicovergen_b200.synth import SynthesizerB200
    from oracle import synth as osyn

    cpt = make_rvc_checkpoint(sr_key, version, f0=f0)
    upp = int(np.prod(cpt["config"][12]))
    P = 96
    g = torch.Generator().manual_seed(3)
    phone = torch.randn(1, P, 768 if version == "v2" else 256, generator=g)
    pitch = torch.randint(1, 255, (1, P), generator=g)
    pitchf = (220.0 * 2 ** (0.5 * torch.sin(torch.arange(P) * 0.05)))[None].float()
    pitchf[:, 20:31] = 0
    sid = torch.tensor([0])
    nz, ns = osyn.draw_noise(7, P, 192, upp)
    ref = osyn.infer(cpt, phone, pitch if f0 else None, pitchf if f0 else None, sid, nz, ns if f0 else None)
    net = SynthesizerB200(cpt, "cuda:0")
    if f0:
        o = net.infer(phone.cuda(), torch.tensor([P]).cuda(), pitch.cuda(), pitchf.cuda(), sid.cuda(), noise_z=nz.cuda(), noise_src=ns.cuda())[0]
    else:
        o = net.infer(phone.cuda(), torch.tensor([P]).cuda(), sid.cuda(), noise_z=nz.cuda())[0]          # _nono call form
    e = rms(o.cpu().numpy() - ref.numpy())
    print(f"[synth {sr_key} {version} f0={f0}] upp {upp}: waveform abs rms err {e:.3e} (ref rms {rms(ref.numpy()):.3e})")
    assert o.shape == ref.shape == (1, 1, P * upp) and torch.isfinite(o).all()
    assert e < 1e-3


@pytest.mark.parametrize("version,if_f0", [("v1", 1), ("v2", 0)])
def test_vc_pipeline_v1_and_nono(version, if_f0):
    """v1: HuBERT layer 9 + final_proj -> 256-dim features, 256-dim index (vc_infer_pipeline.py:401-406); if_f0 = 0: no F0,
    3-argument net_g.infer (:462-465)."""
    from aicovergen_b200.hubert import HubertB200
    from aicovergen_b200.index import write_index_npz
    from aicovergen_b200.rmvpe import RMVPEB200
    from aicovergen_b200.synth import SynthesizerB200
    from aicovergen_b200.vc_infer_pipeline import VC
    from oracle import hubert as ohub
    from oracle import pipeline as opipe
    from oracle.index import IvfFlatIndex
    import tempfile

    hsd, rsd, cpt = make_hubert_state_dict(), make_rmvpe_trained_like(), make_rvc_checkpoint("40k", version, f0=if_f0)
    audio = vocal_like(5.1, seed=11)
    xs = dict(x_pad=1, x_query=1, x_center=2, x_max=3)
    base = ohub.extract_features(hsd, torch.from_numpy(vocal_like(3.0, seed=3))[None], 9 if version == "v1" else 12)
    base = (ohub.final_proj(hsd, base) if version == "v1" else base)[0]
    cent, vecs = make_ivf_index_data(base, n_total=3000, nlist=32, lloyd=False)
    index = IvfFlatIndex(cent, vecs)
    tmp = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
    tmp.close()
    write_index_npz(tmp.name, cent, vecs)
    ref_i16, info = opipe.pipeline(hsd, cpt, rsd, audio.copy(), index=index, seed=5, return_all=True, version=version, if_f0=if_f0, **xs)
    vc = VC(40000, types.SimpleNamespace(device="cuda:0", is_half=True, **xs))
    vc.model_rmvpe = RMVPEB200(rsd, device="cuda:0")
    vc.set_noise_seed(5)
    vc.keep_float = True
    out = vc.pipeline(HubertB200(hsd, "cuda:0"), SynthesizerB200(cpt, "cuda:0"), 0, audio.copy(), "x.wav", [0, 0, 0], 0, "rmvpe",
                      tmp.name, 0.5, if_f0, 3, 40000, 0, 0.25, version, 0.33, 128)
    os.unlink(tmp.name)
    e = rms(vc.last_float_output.astype(np.float64) - info["float_out"])
    print(f"[pipeline {version} if_f0={if_f0}] cuts {info['opt_ts']}; float waveform abs rms err {e:.3e} (ref rms {rms(info['float_out']):.3e})")
    assert out.shape == ref_i16.shape and len(info["opt_ts"]) >= 1
    assert e < 1e-3


def test_f0_file_and_exact_hpf():
    """`f0_file` overrides the estimated F0 from x_pad seconds on (vc_infer_pipeline.py:347-357); `exact_hpf` runs the
    reference's ba-form scipy filtfilt on the host instead of the device sos cascade."""
    from aicovergen_b200.hubert import HubertB200
    from aicovergen_b200.rmvpe import RMVPEB200
    from aicovergen_b200.synth import SynthesizerB200
    from aicovergen_b200.vc_infer_pipeline import VC
    from oracle import rmvpe as orm
    import tempfile

    hsd, rsd, cpt = make_hubert_state_dict(), make_rmvpe_trained_like(), make_rvc_checkpoint("40k", "v2")
    audio = vocal_like(2.5, seed=5)
    vc = VC(40000, types.SimpleNamespace(device="cuda:0", is_half=True, x_pad=1, x_query=1, x_center=2, x_max=3))
    vc.model_rmvpe = RMVPEB200(rsd, device="cuda:0")
    pad = np.pad(audio.astype(np.float64), (16000, 16000), mode="reflect")
    p_len = len(pad) // 160
    # f0 file: "time,f0" lines (the reference parses them at :536-546)
    tf = np.arange(0, 1.0, 0.01)
    curve = 150.0 + 100.0 * tf
    inp = np.stack([tf, curve], 1).astype("float32")
    pitch, pitchf = vc.get_f0("x", pad, p_len, 0, "rmvpe", 3, 128, inp_f0=inp)
    f0 = orm.infer_from_audio(rsd, pad.astype(np.float32), 0.03)
    delta_t = np.round((inp[:, 0].max() - inp[:, 0].min()) * 100 + 1).astype("int16")
    rep = np.interp(list(range(delta_t)), inp[:, 0] * 100, inp[:, 1])
    f0[100:100 + len(rep)] = rep[:f0[100:100 + len(rep)].shape[0]]
    want, wantf = orm.coarse_pitch(f0, 0)
    assert np.array_equal(pitch[100:100 + len(rep)], want[100:100 + len(rep)]) and np.allclose(pitchf[100:200], wantf[100:200])
    f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
    f.write("\n".join(f"{a:.2f},{b:.3f}" for a, b in zip(tf, curve)) + "\n")
    f.close()
    hub, net = HubertB200(hsd, "cuda:0"), SynthesizerB200(cpt, "cuda:0")
    outs = {}
    for mode in ("device", "exact"):
        vc.exact_hpf = mode == "exact"
        vc.set_noise_seed(3)
        vc.keep_float = True
        vc.pipeline(hub, net, 0, audio.copy(), "x.wav", [0, 0, 0], 0, "rmvpe", "", 0.0, 1, 3, 40000, 0, 1, "v2", 0.33, 128,
                    f0_file=types.SimpleNamespace(name=f.name))
        outs[mode] = vc.last_float_output.copy()
    os.unlink(f.name)
    e = rms(outs["device"] - outs["exact"])
    print(f"[exact_hpf] device sos cascade vs host ba-form filtfilt: waveform abs rms diff {e:.3e}")
    assert np.isfinite(outs["exact"]).all() and e < 1e-3


def test_ivf_index_short_and_empty_lists():
    """faiss pads a short inverted list with +inf distances / label -1 and the reference's weights (vc_infer_pipeline.py:
    424-427) then give those slots weight 0; an EMPTY list makes every weight 0/0 = NaN (a reference quirk that must not be
    papered over).  Crafted index: list sizes 0, 3, 8, 50."""
    from aicovergen_b200.index import IvfIndexB200
    from oracle.index import IvfFlatIndex, blend

    rng = np.random.default_rng(0)
    d = 64
    cent = np.stack([np.full(d, v, dtype=np.float32) for v in (-30.0, -10.0, 10.0, 30.0)])
    sizes = [0, 3, 8, 50]
    vecs = np.concatenate([cent[i][None] + 0.5 * rng.standard_normal((n, d)).astype(np.float32) for i, n in enumerate(sizes)])
    oidx = IvfFlatIndex(cent, vecs)
    assert [len(l) for l in oidx.lists] == sizes
    q = np.concatenate([cent[i][None] + 0.5 * rng.standard_normal((5, d)).astype(np.float32) for i in range(4)])
    D0, I0 = oidx.search(q, 8)
    gidx = IvfIndexB200(cent, vecs, "cuda:0")
    D1, I1 = gidx.search(q, 8)
    assert np.array_equal(I0, I1), "ids (incl. the -1 padding of short lists)"
    assert np.array_equal(np.isinf(D0), np.isinf(D1)) and np.allclose(D0[np.isfinite(D0)], D1[np.isfinite(D1)], rtol=1e-5)
    assert (I1[:5] == -1).all() and (I1[5:10, 3:] == -1).all() and (I1[10:] >= 0).all()
    ref = blend(oidx, vecs, q, 0.5)
    got = gidx.search_blend(torch.from_numpy(q).cuda(), 0.5).cpu().numpy()
    assert np.isnan(ref[:5]).all() and np.isnan(got[:5]).all(), "empty list -> NaN row on both sides"
    assert np.allclose(got[5:], ref[5:], rtol=1e-4, atol=1e-5)
    assert gidx.ntotal == 61 and np.array_equal(gidx.reconstruct_n(0, 61), vecs)


def test_blocked_attention_matches_single_block(monkeypatch):
    """Attention in query blocks (plans.ATT_SCRATCH_BYTES bounds the score scratch; relative-position band terms take the
    block's first row) gives the same result as one block per segment: HuBERT features and the synthesizer's latents."""
    import aicovergen_b200.hubert as bh
    import aicovergen_b200.synth as bs
    from oracle import synth as osyn

    hsd, cpt = make_hubert_state_dict(), make_rvc_checkpoint("40k", "v2")
    x = torch.from_numpy(vocal_like(9.0, seed=2))[None].cuda()
    P = 700
    g = torch.Generator().manual_seed(3)
    phone, pitch = torch.randn(1, P, 768, generator=g).cuda(), torch.randint(1, 255, (1, P), generator=g).cuda()
    pitchf = (220.0 * 2 ** (0.5 * torch.sin(torch.arange(P) * 0.05)))[None].float().cuda()
    nz, ns = osyn.draw_noise(7, P, 192, 400)
    outs = []
    for scratch in (1 << 30, 1 << 20):          # one block / blocks of 128 rows
        monkeypatch.setattr(bh, "ATT_SCRATCH_BYTES", scratch)
        monkeypatch.setattr(bs, "ATT_SCRATCH_BYTES", scratch)
        feats = bh.HubertB200(hsd, "cuda:0").extract_features(source=x, padding_mask=None, output_layer=12)[0].clone()
        o, _, lat = bs.SynthesizerB200(cpt, "cuda:0").infer(phone, torch.tensor([P]).cuda(), pitch, pitchf, torch.tensor([0]).cuda(),
                                                           noise_z=nz.cuda(), noise_src=ns.cuda())
        outs.append((feats, lat[2].clone(), o.clone()))
    for a, b, name in zip(outs[0], outs[1], ("hubert features", "enc_p m_p", "waveform")):
        d = float((a - b).abs().max())
        print(f"[blocked attention] {name}: max abs diff vs single block {d:.2e}")
        assert d <= 1e-6 * float(a.abs().max()) + 1e-7
