"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference with the stub
recipe of oracle/ref_import.py) on seeded synthetic checkpoints.  Only runs in the build container.

    python tools/make_golden.py
The fixtures pin (a) the oracle restatements on any machine (tests/test_golden_cpu.py) and (b) the CUDA path on
the GPU box (tests/test_golden_gpu.py), where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aicovergen_b200.synthetic import (make_hubert_state_dict, make_mdx_state_dict, make_rmvpe_state_dict,  # noqa: E402
                                       make_rvc_checkpoint)
from oracle import hubert as ohub  # noqa: E402
from oracle import mdx as om  # noqa: E402
from oracle import ref_import  # noqa: E402
from refshim import HubertShim, ref_net_g, ref_rmvpe, ref_vc  # noqa: E402
from siggen import stereo_tones, vocal_like  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def synth():
    cpt = make_rvc_checkpoint("40k", "v2", seed=1234)
    net = ref_net_g(cpt)
    P = 40
    g = torch.Generator().manual_seed(21)
    phone = torch.randn(1, P, 768, generator=g)
    pitch = torch.randint(1, 255, (1, P), generator=g)
    pitchf = (180 + 60 * torch.sin(torch.arange(P) * 0.2))[None].float()
    pitchf[:, 11:17] = 0
    torch.manual_seed(77)
    with torch.no_grad():
        o, _, (z, z_p, m_p, logs_p) = net.infer(phone, torch.tensor([P]), pitch, pitchf, torch.tensor([0]))
    np.savez_compressed(os.path.join(OUT, "synth_v2_40k.npz"), phone=phone.numpy(), pitch=pitch.numpy(), pitchf=pitchf.numpy(),
                        noise_seed=77, ckpt_seed=1234, out=o.numpy()[0, 0], m_p=m_p.numpy(), z=z.numpy())


def rmvpe():
    sd = make_rmvpe_state_dict(seed=4321)
    rm = ref_rmvpe(sd)
    x = vocal_like(1.6, seed=11)
    f0 = rm.infer_from_audio(x, 0.03)
    with torch.no_grad():
        hid = rm.mel2hidden(rm.mel_extractor(torch.from_numpy(x)[None]))[0].numpy()
    np.savez_compressed(os.path.join(OUT, "rmvpe.npz"), seconds=1.6, audio_seed=11, f0=f0, salience_argmax=hid.argmax(1), ckpt_seed=4321,
                        salience_max=hid.max(1))


def pipeline():
    hsd, rsd, cpt = make_hubert_state_dict(seed=777), make_rmvpe_state_dict(seed=4321), make_rvc_checkpoint("40k", "v2", seed=1234)
    audio = vocal_like(3.4, seed=13)
    xs = dict(x_pad=1, x_query=1, x_center=1, x_max=2)
    _, vc = ref_vc(40000, **xs)
    vc.model_rmvpe = ref_rmvpe(rsd)
    net = ref_net_g(cpt)
    torch.manual_seed(9)
    out = vc.pipeline(HubertShim(hsd), net, 0, audio.copy(), "x.wav", [0, 0, 0], 0, "rmvpe", "", 0.5, 1, 3, 40000, 0, 0.25,
                      "v2", 0.33, 128)
    np.savez_compressed(os.path.join(OUT, "vc_pipeline.npz"), seconds=3.4, audio_seed=13, out_int16=out, noise_seed=9, **{k: np.int64(v) for k, v in xs.items()})


def mdx():
    ref = ref_import.module("mdx")
    dim_f, dim_t, n_fft = 256, 16, 2048
    sd = make_mdx_state_dict(dim_f=dim_f, dim_t=dim_t, g=8, n=3, seed=2024)

    class FakeSession:
        def __init__(self, path, providers=None):
            pass

        def run(self, _, feed):
            return [om.convtdfnet(sd, torch.from_numpy(feed["input"])).numpy()]

    sys.modules["onnxruntime"].InferenceSession = FakeSession
    ref.ort.InferenceSession = FakeSession
    model = ref.MDXModel(torch.device("cpu"), dim_f=dim_f, dim_t=dim_t, n_fft=n_fft, stem_name="Vocals", compensation=1.035)
    sess = ref.MDX("fake.onnx", model, processor=-1)
    n = 44100 * 2 + 4321        # > 2 margins so both halves keep something
    wave = stereo_tones(n, seed=4)
    out = sess.process_wave(wave.copy(), 2)
    spec = model.stft(torch.from_numpy(wave[:, :model.chunk_size].copy())[None])
    np.savez_compressed(os.path.join(OUT, "mdx_small.npz"), n=n, wave_seed=4, processed=out.astype(np.float32), dim_f=dim_f, dim_t=dim_t,
                        n_fft=n_fft, ckpt_seed=2024, stft_abs_sum=float(spec.abs().sum()), spec_slice=spec[0, :, :8, :4].numpy())


if __name__ == "__main__":
    synth()
    rmvpe()
    pipeline()
    mdx()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
