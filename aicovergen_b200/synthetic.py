"""Seeded synthetic checkpoints in the reference's on-disk formats.

No model files exist offline, so tests and bench.py run on
random weights with the exact names/shapes the reference loaders expect:
  * RVC `.pth`  : {"config": [...18 positional...], "weight": state_dict, "f0": 1, "version": "v2"}
                  (rvc.py:113-134, infer_pack/models.py:643-664)
  * rmvpe.pt    : E2E(4, 1, (2, 2)).state_dict()           (rmvpe.py:331-333)
  * hubert_base : fairseq HubertModel names  (rvc.py:99)
  * MDX-Net     : restated TFC-TDF net ("ConvTDFNet") parameter dict (mdx.py:74)
  * IVF index   : centroids / list assignment / vectors     (vc_infer_pipeline.py:505-507)
Deterministic given the seed on any machine (CPU generator, fp32).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

# Reference training configs (src/configs/*.json model blocks) as the positional cpt["config"] list
# (infer_pack/models.py:643-664): spec_channels, segment_size, inter, hidden, filter, heads, layers,
# kernel, p_dropout, resblock, rb_kernels, rb_dilations, up_rates, up_init, up_kernels, spk_embed, gin, sr
RVC_CONFIGS = {
    "40k": [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
            [10, 10, 2, 2], 512, [16, 16, 4, 4], 109, 256, 40000],
    "32k": [513, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
            [10, 4, 2, 2, 2], 512, [16, 16, 4, 4, 4], 109, 256, 32000],
    "48k": [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
            [10, 6, 2, 2, 2], 512, [16, 16, 4, 4, 4], 109, 256, 48000],
    "32k_v2": [513, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
               [10, 8, 2, 2], 512, [20, 16, 4, 4], 109, 256, 32000],
    "48k_v2": [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
               [12, 10, 2, 2], 512, [24, 20, 4, 4], 109, 256, 48000],
}


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def normal(self, shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=self.g) * std + mean

    def uniform(self, shape, lo, hi):
        return torch.rand(*shape, generator=self.g) * (hi - lo) + lo

    def conv(self, shape, gain=1.0):
        """weight [out, in, *k] ~ N(0, gain^2 / fan_in)."""
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return self.normal(shape, gain / math.sqrt(fan_in))


def _wn(sd: Dict[str, torch.Tensor], name: str, v: torch.Tensor):
    """Old-style torch.nn.utils.weight_norm parametrisation (dim=0): weight = g * v / ||v||."""
    sd[name + ".weight_v"] = v
    sd[name + ".weight_g"] = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1))).clone()


def make_rvc_checkpoint(sr_key: str = "40k", version: str = "v2", seed: int = 1234, f0: int = 1) -> dict:
    """Synthetic RVC voice model in the `.pth` dict format consumed by rvc.get_vc (rvc.py:112-143)."""
    cfg = [c if not isinstance(c, list) else [list(x) if isinstance(x, list) else x for x in c]
           for c in RVC_CONFIGS[sr_key]]
    (_, _, inter, hidden, filt, heads, layers, ksz, _, _, rb_k, rb_d, up_r, up_init, up_k, spk, gin, sr) = cfg
    in_dim = 768 if version == "v2" else 256
    g = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    # ---- enc_p (models.py:64-108, attentions.py:13-73)
    sd["enc_p.emb_phone.weight"] = g.conv((hidden, in_dim))
    sd["enc_p.emb_phone.bias"] = g.normal((hidden,), 0.05)
    if f0:
        sd["enc_p.emb_pitch.weight"] = g.normal((256, hidden), 0.3)
    dk = hidden // heads
    for i in range(layers):
        p = f"enc_p.encoder.attn_layers.{i}."
        sd[p + "emb_rel_k"] = g.normal((1, 21, dk), dk ** -0.5)
        sd[p + "emb_rel_v"] = g.normal((1, 21, dk), dk ** -0.5)
        for n in ("q", "k", "v", "o"):
            sd[p + f"conv_{n}.weight"] = g.conv((hidden, hidden, 1), 1.0 if n in "qk" else 0.7)
            sd[p + f"conv_{n}.bias"] = g.normal((hidden,), 0.05)
        sd[f"enc_p.encoder.norm_layers_1.{i}.gamma"] = g.uniform((hidden,), 0.8, 1.2)
        sd[f"enc_p.encoder.norm_layers_1.{i}.beta"] = g.normal((hidden,), 0.05)
        p = f"enc_p.encoder.ffn_layers.{i}."
        sd[p + "conv_1.weight"] = g.conv((filt, hidden, ksz))
        sd[p + "conv_1.bias"] = g.normal((filt,), 0.05)
        sd[p + "conv_2.weight"] = g.conv((hidden, filt, ksz), 0.7)
        sd[p + "conv_2.bias"] = g.normal((hidden,), 0.05)
        sd[f"enc_p.encoder.norm_layers_2.{i}.gamma"] = g.uniform((hidden,), 0.8, 1.2)
        sd[f"enc_p.encoder.norm_layers_2.{i}.beta"] = g.normal((hidden,), 0.05)
    sd["enc_p.proj.weight"] = g.conv((inter * 2, hidden, 1), 0.5)
    sd["enc_p.proj.bias"] = g.normal((inter * 2,), 0.05)
    # ---- flow (models.py:111-157, modules.py:136-221, 405-462): 4 coupling layers at even indices
    half = inter // 2
    for f in range(4):
        p = f"flow.flows.{2 * f}."
        sd[p + "pre.weight"] = g.conv((hidden, half, 1))
        sd[p + "pre.bias"] = g.normal((hidden,), 0.05)
        for j in range(3):
            _wn(sd, p + f"enc.in_layers.{j}", g.conv((2 * hidden, hidden, 5)))
            sd[p + f"enc.in_layers.{j}.bias"] = g.normal((2 * hidden,), 0.05)
            rs = 2 * hidden if j < 2 else hidden
            _wn(sd, p + f"enc.res_skip_layers.{j}", g.conv((rs, hidden, 1), 0.7))
            sd[p + f"enc.res_skip_layers.{j}.bias"] = g.normal((rs,), 0.05)
        _wn(sd, p + "enc.cond_layer", g.conv((2 * hidden * 3, gin, 1), 0.5))
        sd[p + "enc.cond_layer.bias"] = g.normal((2 * hidden * 3,), 0.05)
        # the reference zero-inits `post` (modules.py:437-438); redrawn so the flow is non-trivial
        sd[p + "post.weight"] = g.conv((half, hidden, 1), 0.3)
        sd[p + "post.bias"] = g.normal((half,), 0.02)
    # ---- dec (models.py:422-522)
    if f0:
        sd["dec.m_source.l_linear.weight"] = torch.tensor([[0.9]])
        sd["dec.m_source.l_linear.bias"] = torch.tensor([0.01])
    sd["dec.conv_pre.weight"] = g.conv((up_init, inter, 7))
    sd["dec.conv_pre.bias"] = g.normal((up_init,), 0.05)
    sd["dec.cond.weight"] = g.conv((up_init, gin, 1), 0.5)
    sd["dec.cond.bias"] = g.normal((up_init,), 0.05)
    ch = up_init
    for i, (u, k) in enumerate(zip(up_r, up_k)):
        cin, cout = up_init // (2 ** i), up_init // (2 ** (i + 1))
        # ConvTranspose1d weight [Cin, Cout, k]; each output sums ~k/u taps x Cin inputs
        v = g.normal((cin, cout, k), 1.0 / math.sqrt(cin * k / u))
        _wn(sd, f"dec.ups.{i}", v)
        sd[f"dec.ups.{i}.bias"] = g.normal((cout,), 0.05)
        if f0:
            if i + 1 < len(up_r):
                s = 1
                for r in up_r[i + 1:]:
                    s *= r
                sd[f"dec.noise_convs.{i}.weight"] = g.normal((cout, 1, 2 * s), 1.0 / math.sqrt(2 * s) * 3.0)
            else:
                sd[f"dec.noise_convs.{i}.weight"] = g.normal((cout, 1, 1), 3.0)
            sd[f"dec.noise_convs.{i}.bias"] = g.normal((cout,), 0.05)
        for j, (kk, dd) in enumerate(zip(rb_k, rb_d)):
            n = i * len(rb_k) + j
            for m in range(len(dd)):
                _wn(sd, f"dec.resblocks.{n}.convs1.{m}", g.conv((cout, cout, kk), 1.0))
                sd[f"dec.resblocks.{n}.convs1.{m}.bias"] = g.normal((cout,), 0.05)
                _wn(sd, f"dec.resblocks.{n}.convs2.{m}", g.conv((cout, cout, kk), 0.5))
                sd[f"dec.resblocks.{n}.convs2.{m}.bias"] = g.normal((cout,), 0.05)
        ch = cout
    sd["dec.conv_post.weight"] = g.conv((1, ch, 7), 0.5)
    sd["emb_g.weight"] = g.normal((spk, gin), 1.0)
    return {"config": cfg, "weight": sd, "f0": f0, "version": version, "info": "synthetic", "sr": sr_key}


# ---------------------------------------------------------------------------
# rmvpe.pt  (E2E(4, 1, (2, 2)).state_dict(), rmvpe.py:221-258, 331-333)
# ---------------------------------------------------------------------------
def _bn(sd, g: _Gen, name: str, c: int, gamma=(0.5, 1.5), beta=(0.0, 0.1)):
    """BatchNorm2d eval statistics per SURVEY.md §8(d) weights policy."""
    sd[name + ".weight"] = g.uniform((c,), gamma[0], gamma[1])
    sd[name + ".bias"] = g.normal((c,), beta[1], beta[0])
    sd[name + ".running_mean"] = g.normal((c,), 0.1)
    sd[name + ".running_var"] = g.uniform((c,), 0.5, 1.5)
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv_block_res(sd, g: _Gen, p: str, cin: int, cout: int, **bn):
    """ConvBlockRes (rmvpe.py:23-58): conv.0 / bn conv.1 / conv.3 / bn conv.4 (+ 1x1 shortcut if cin != cout)."""
    sd[p + "conv.0.weight"] = g.conv((cout, cin, 3, 3), 1.3)
    _bn(sd, g, p + "conv.1", cout, **bn)
    sd[p + "conv.3.weight"] = g.conv((cout, cout, 3, 3), 0.45)
    _bn(sd, g, p + "conv.4", cout, **bn)
    if cin != cout:
        sd[p + "shortcut.weight"] = g.conv((cout, cin, 1, 1), 0.8)
        sd[p + "shortcut.bias"] = g.normal((cout,), 0.05)


def make_rmvpe_state_dict(seed: int = 4321, n_blocks: int = 4, en_de_layers: int = 5, inter_layers: int = 4,
                          en_out_channels: int = 16, n_mels: int = 128, n_gru_hidden: int = 256,
                          n_class: int = 360, bn_gamma=(0.5, 1.5), bn_beta=(0.0, 0.1)) -> Dict[str, torch.Tensor]:
    g = _Gen(seed)
    bn = dict(gamma=bn_gamma, beta=bn_beta)
    sd: Dict[str, torch.Tensor] = {}
    # input BN over the single mel "channel": keep log-mel roughly standardised
    sd["unet.encoder.bn.weight"] = torch.tensor([0.35])
    sd["unet.encoder.bn.bias"] = torch.tensor([0.9])
    sd["unet.encoder.bn.running_mean"] = torch.tensor([-4.0])
    sd["unet.encoder.bn.running_var"] = torch.tensor([1.3])
    sd["unet.encoder.bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    cin, cout = 1, en_out_channels
    for i in range(en_de_layers):
        for b in range(n_blocks):
            _conv_block_res(sd, g, f"unet.encoder.layers.{i}.conv.{b}.", cin if b == 0 else cout, cout, **bn)
        cin, cout = cout, cout * 2
    # intermediate: first 256 -> 512, then 512 -> 512
    ci, co = cin, cout
    for i in range(inter_layers):
        for b in range(n_blocks):
            _conv_block_res(sd, g, f"unet.intermediate.layers.{i}.conv.{b}.", ci if b == 0 else co, co, **bn)
        ci = co
    dc = co
    for i in range(en_de_layers):
        do = dc // 2
        # ConvTranspose2d weight [Cin, Cout, 3, 3]; stride 2 -> ~2.25 taps per output
        sd[f"unet.decoder.layers.{i}.conv1.0.weight"] = g.normal((dc, do, 3, 3), 1.2 / math.sqrt(dc * 2.25))
        _bn(sd, g, f"unet.decoder.layers.{i}.conv1.1", do, **bn)
        for b in range(n_blocks):
            _conv_block_res(sd, g, f"unet.decoder.layers.{i}.conv2.{b}.", do * 2 if b == 0 else do, do, **bn)
        dc = do
    sd["cnn.weight"] = g.conv((3, en_out_channels, 3, 3), 1.0)
    sd["cnn.bias"] = g.normal((3,), 0.05)
    Hh = n_gru_hidden
    for suf in ("", "_reverse"):
        sd[f"fc.0.gru.weight_ih_l0{suf}"] = g.conv((3 * Hh, 3 * n_mels), 1.0)
        sd[f"fc.0.gru.weight_hh_l0{suf}"] = g.conv((3 * Hh, Hh), 1.0)
        sd[f"fc.0.gru.bias_ih_l0{suf}"] = g.normal((3 * Hh,), 0.1)
        sd[f"fc.0.gru.bias_hh_l0{suf}"] = g.normal((3 * Hh,), 0.1)
    sd["fc.1.weight"] = g.conv((n_class, 2 * Hh), 1.0)
    sd["fc.1.bias"] = g.normal((n_class,), 0.3, -1.5)
    return sd


# ---------------------------------------------------------------------------
# hubert_base.pt — fairseq 0.12.2 HubertModel state-dict names (rvc.py:98-109)
# ---------------------------------------------------------------------------
HUBERT_CONV = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2


def make_hubert_state_dict(seed: int = 777, layers: int = 12, dim: int = 768, ffn: int = 3072,
                           conv=HUBERT_CONV, pos_k: int = 128, pos_groups: int = 16) -> Dict[str, torch.Tensor]:
    g = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    cin = 1
    for i, (c, k, s) in enumerate(conv):
        sd[f"feature_extractor.conv_layers.{i}.0.weight"] = g.conv((c, cin, k), 1.6)
        cin = c
    sd["feature_extractor.conv_layers.0.2.weight"] = g.uniform((conv[0][0],), 0.7, 1.3)
    sd["feature_extractor.conv_layers.0.2.bias"] = g.normal((conv[0][0],), 0.1)
    sd["layer_norm.weight"] = g.uniform((cin,), 0.8, 1.2)
    sd["layer_norm.bias"] = g.normal((cin,), 0.05)
    sd["post_extract_proj.weight"] = g.conv((dim, cin), 1.0)
    sd["post_extract_proj.bias"] = g.normal((dim,), 0.05)
    v = g.conv((dim, dim // pos_groups, pos_k), 1.0)
    sd["encoder.pos_conv.0.weight_v"] = v
    sd["encoder.pos_conv.0.weight_g"] = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt().clone()   # weight_norm dim=2
    sd["encoder.pos_conv.0.bias"] = g.normal((dim,), 0.05)
    sd["encoder.layer_norm.weight"] = g.uniform((dim,), 0.8, 1.2)
    sd["encoder.layer_norm.bias"] = g.normal((dim,), 0.05)
    for i in range(layers):
        p = f"encoder.layers.{i}."
        for n in ("q", "k", "v", "out"):
            sd[p + f"self_attn.{n}_proj.weight"] = g.conv((dim, dim), 1.0 if n in ("q", "k") else 0.7)
            sd[p + f"self_attn.{n}_proj.bias"] = g.normal((dim,), 0.05)
        sd[p + "self_attn_layer_norm.weight"] = g.uniform((dim,), 0.8, 1.2)
        sd[p + "self_attn_layer_norm.bias"] = g.normal((dim,), 0.05)
        sd[p + "fc1.weight"] = g.conv((ffn, dim), 1.0)
        sd[p + "fc1.bias"] = g.normal((ffn,), 0.05)
        sd[p + "fc2.weight"] = g.conv((dim, ffn), 0.7)
        sd[p + "fc2.bias"] = g.normal((dim,), 0.05)
        sd[p + "final_layer_norm.weight"] = g.uniform((dim,), 0.8, 1.2)
        sd[p + "final_layer_norm.bias"] = g.normal((dim,), 0.05)
    sd["final_proj.weight"] = g.conv((256, dim), 1.0)
    sd["final_proj.bias"] = g.normal((256,), 0.05)
    sd["mask_emb"] = g.uniform((dim,), 0.0, 1.0)
    sd["label_embs_concat"] = g.normal((504, 256), 1.0)
    return sd


# ---------------------------------------------------------------------------
# IVF-Flat index contents (faiss `added_IVF{nlist}_Flat_nprobe_1_*.index`, vc_infer_pipeline.py:505-507)
# ---------------------------------------------------------------------------
def make_ivf_index_data(base_feats: torch.Tensor, n_total: int = 87243, nlist: int = 2237, seed: int = 99,
                        jitter: float = 0.05, lloyd: bool = True):
    """Database = rows of `base_feats` (HuBERT features of a seeded clip) cycled to n_total + N(0, jitter);
    centroids = one k-means iteration from a seeded sample (SURVEY.md §8(d) cfg 3). Returns (centroids, vectors)."""
    g = torch.Generator().manual_seed(seed)
    base = base_feats.float().cpu()
    reps = (n_total + base.shape[0] - 1) // base.shape[0]
    vecs = base.repeat(reps, 1)[:n_total].clone()
    vecs += torch.randn(vecs.shape, generator=g) * jitter
    perm = torch.randperm(n_total, generator=g)[:nlist]
    cent = vecs[perm].clone()
    if not lloyd:
        return cent.numpy(), vecs.numpy()
    # one Lloyd iteration
    c2 = (cent.double() ** 2).sum(1)
    assign = torch.empty(n_total, dtype=torch.long)
    for s in range(0, n_total, 8192):
        d = c2[None, :] - 2.0 * vecs[s:s + 8192].double() @ cent.double().t()
        assign[s:s + 8192] = d.argmin(1)
    sums = torch.zeros_like(cent, dtype=torch.float64).index_add_(0, assign, vecs.double())
    cnt = torch.bincount(assign, minlength=nlist).clamp(min=1)[:, None]
    cent = (sums / cnt).float()
    return cent.numpy(), vecs.numpy()


# ---------------------------------------------------------------------------
# MDX-Net TFC-TDF U-Net ("ConvTDFNet") — the architecture inside the UVR-MDX-NET ONNX files (mdx.py:74).
# The .onnx graphs are not in the reference repo: names follow the public KUIELab/UVR module layout.
# ---------------------------------------------------------------------------
def make_mdx_state_dict(dim_f: int = 3072, dim_t: int = 256, g: int = 48, l: int = 3, n: int = 5, bn: int = 8,
                        k: int = 3, dim_c: int = 4, seed: int = 2024, bn_gamma=(0.7, 1.3), bn_beta=(0.0, 0.1)) -> Dict[str, torch.Tensor]:
    """bn_gamma = U(lo, hi) range of the BatchNorm scales, bn_beta = (mean, std) of the BatchNorm shifts."""
    gen = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bnorm(name, c):
        sd[name + ".weight"] = gen.uniform((c,), bn_gamma[0], bn_gamma[1])
        sd[name + ".bias"] = gen.normal((c,), bn_beta[1], bn_beta[0])
        sd[name + ".running_mean"] = gen.normal((c,), 0.1)
        sd[name + ".running_var"] = gen.uniform((c,), 0.6, 1.4)

    def tfc_tdf(p, c, f):
        for j in range(l):
            sd[f"{p}.tfc.H.{j}.0.weight"] = gen.conv((c, c, k, k), 1.3)
            sd[f"{p}.tfc.H.{j}.0.bias"] = gen.normal((c,), 0.05)
            bnorm(f"{p}.tfc.H.{j}.1", c)
        sd[f"{p}.tdf.0.weight"] = gen.conv((f // bn, f), 1.0)
        bnorm(f"{p}.tdf.1", c)
        sd[f"{p}.tdf.3.weight"] = gen.conv((f, f // bn), 1.0)
        bnorm(f"{p}.tdf.4", c)

    sd["first_conv.0.weight"] = gen.conv((g, dim_c, 1, 1), 1.5)
    sd["first_conv.0.bias"] = gen.normal((g,), 0.05)
    bnorm("first_conv.1", g)
    f, c = dim_f, g
    for i in range(n):
        tfc_tdf(f"encoding_blocks.{i}", c, f)
        sd[f"ds.{i}.0.weight"] = gen.conv((c + g, c, 2, 2), 1.3)
        sd[f"ds.{i}.0.bias"] = gen.normal((c + g,), 0.05)
        bnorm(f"ds.{i}.1", c + g)
        f, c = f // 2, c + g
    tfc_tdf("bottleneck_block", c, f)
    for i in range(n):
        sd[f"us.{i}.0.weight"] = gen.normal((c, c - g, 2, 2), 1.3 / math.sqrt(c))     # ConvTranspose2d [Cin, Cout, 2, 2]
        sd[f"us.{i}.0.bias"] = gen.normal((c - g,), 0.05)
        bnorm(f"us.{i}.1", c - g)
        f, c = f * 2, c - g
        tfc_tdf(f"decoding_blocks.{i}", c, f)
    sd["final_conv.0.weight"] = gen.conv((dim_c, c, 1, 1), 0.1)
    sd["final_conv.0.bias"] = gen.normal((dim_c,), 0.02)
    sd["_meta"] = torch.tensor([dim_f, dim_t, g, l, n, bn, k, dim_c])
    return sd


def calibrate_mdx_batchnorm(sd: Dict[str, torch.Tensor], x: torch.Tensor, eps: float = 1e-5,
                            out_gain: float = 0.0) -> Dict[str, torch.Tensor]:
    """Gives a synthetic ConvTDFNet the activation statistics of a TRAINED one: one forward pass over `x`
    ([B,4,dim_f,dim_t] spectrogram-like input) sets every BatchNorm's running_mean / running_var to the statistics of
    its own input (what training converges to), so each normalised layer emits O(1) values and the multiplicative skips
    stay bounded (the default synthetic checkpoints reach 1e5..1e10 at full size: harmless in fp32/TF32, beyond fp16).
    Returns a NEW state dict; weights, affine parameters and `_meta` are untouched.  Plain torch on the CPU — a weight
    generator for tests/bench, not part of the inference path."""
    import torch.nn.functional as F

    out = {k_: v.clone() for k_, v in sd.items()}
    dim_f, dim_t, g, l, n, bnf, k, dim_c = [int(v) for v in sd["_meta"]]

    def bn_fit(name, t):
        mean = t.mean(dim=(0, 2, 3))
        var = t.var(dim=(0, 2, 3), unbiased=False)
        out[name + ".running_mean"], out[name + ".running_var"] = mean, var.clamp_min(1e-6)
        return F.batch_norm(t, mean, var.clamp_min(1e-6), out[name + ".weight"], out[name + ".bias"], False, 0.0, eps)

    def tfc_tdf(p, t):
        for j in range(l):
            t = F.relu(bn_fit(f"{p}.tfc.H.{j}.1", F.conv2d(t, out[f"{p}.tfc.H.{j}.0.weight"], out[f"{p}.tfc.H.{j}.0.bias"], padding=1)))
        h = F.relu(bn_fit(f"{p}.tdf.1", F.linear(t, out[f"{p}.tdf.0.weight"])))
        h = F.relu(bn_fit(f"{p}.tdf.4", F.linear(h, out[f"{p}.tdf.3.weight"])))
        return t + h

    with torch.no_grad():
        t = F.relu(bn_fit("first_conv.1", F.conv2d(x.float(), out["first_conv.0.weight"], out["first_conv.0.bias"]))).transpose(-1, -2)
        skips = []
        for i in range(n):
            t = tfc_tdf(f"encoding_blocks.{i}", t)
            skips.append(t)
            t = F.relu(bn_fit(f"ds.{i}.1", F.conv2d(t, out[f"ds.{i}.0.weight"], out[f"ds.{i}.0.bias"], stride=2)))
        t = tfc_tdf("bottleneck_block", t)
        for i in range(n):
            t = F.relu(bn_fit(f"us.{i}.1", F.conv_transpose2d(t, out[f"us.{i}.0.weight"], out[f"us.{i}.0.bias"], stride=2)))
            t = t * skips[-i - 1]
            t = tfc_tdf(f"decoding_blocks.{i}", t)
        if out_gain > 0:
            # a separator's output has the level of its input: rescale the final 1x1 conv so that the FLUCTUATING part of the
            # output spectrogram (what survives the iSTFT as audio; a per-bin constant is only a click per frame) has
            # out_gain x the input's level, and drop most of its constant part
            y = F.conv2d(t.transpose(-1, -2), out["final_conv.0.weight"], out["final_conv.0.bias"])
            fl = y - y.mean(dim=(0, 2, 3), keepdim=True)
            sc = out_gain * x.float().pow(2).mean().sqrt() / fl.pow(2).mean().sqrt().clamp_min(1e-12)
            out["final_conv.0.weight"] = out["final_conv.0.weight"] * sc
            out["final_conv.0.bias"] = (out["final_conv.0.bias"] - y.mean(dim=(0, 2, 3))) * sc
    return out


def calibrate_rmvpe(sd: Dict[str, torch.Tensor], audio: torch.Tensor, peak_sigma_bins: float = 1.3,
                    logit_gain: float = 4.0, logit_bias: float = -4.8, voicing_sigma: float = 2.0,
                    eps: float = 1e-5) -> Dict[str, torch.Tensor]:
    """Gives a synthetic rmvpe checkpoint the two properties of a TRAINED one that the F0 decode relies on:
      * every BatchNorm's running statistics equal the statistics of its own input on a calibration clip (what training
        converges to), so activations stay O(1) through the ~70 convolutions instead of drifting with depth;
      * the salience head emits ONE smooth peak per frame over neighbouring 20-cent bins (a trained rmvpe's output is a
        narrow bump around the pitch, rmvpe.py:385-409 decodes it with a +-4-bin local average): `fc.1.weight` rows are
        made smooth across bins (Gaussian, `peak_sigma_bins`) and scaled so that 3 sigma of the per-bin logits is
        `logit_gain` around `logit_bias` (most bins ~0), plus a per-frame "voicing" term common to all bins (std
        `voicing_sigma`) so that frames split into clearly voiced (peak 0.3 .. 0.99) and unvoiced (max <= 0.03) ones —
        the `protect` / f0 == 0 branches of VC.pipeline need both.  With a smooth peak an argmax flip between adjacent
        near-equal bins moves the decoded f0 by a fraction of a cent, as it does for real weights.
    `audio` [N] float32 @16 kHz.  Plain torch on the CPU: a weight generator for tests / bench, not an inference path.
    Returns a NEW state dict with the same keys."""
    import torch.nn.functional as F

    from .rmvpe import HOP, N_FFT, N_MELS, mel_basis

    out = {k_: v.clone() for k_, v in sd.items()}

    def bn_fit(name, t):
        mean = t.mean(dim=(0, 2, 3))
        var = t.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6)
        out[name + ".running_mean"], out[name + ".running_var"] = mean, var
        return F.batch_norm(t, mean, var, out[name + ".weight"], out[name + ".bias"], False, 0.0, eps)

    def block(p, x):
        y = F.relu(bn_fit(p + "conv.1", F.conv2d(x, out[p + "conv.0.weight"], padding=1)))
        y = F.relu(bn_fit(p + "conv.4", F.conv2d(y, out[p + "conv.3.weight"], padding=1)))
        if p + "shortcut.weight" in out:
            return y + F.conv2d(x, out[p + "shortcut.weight"], out[p + "shortcut.bias"])
        return y + x

    n_enc = 1 + max(int(k_.split(".")[3]) for k_ in sd if k_.startswith("unet.encoder.layers."))
    n_blocks = 1 + max(int(k_.split(".")[5]) for k_ in sd if k_.startswith("unet.encoder.layers.0.conv."))
    n_inter = 1 + max(int(k_.split(".")[3]) for k_ in sd if k_.startswith("unet.intermediate.layers."))
    with torch.no_grad():
        a = audio.float().reshape(1, -1)
        spec = torch.stft(a, N_FFT, HOP, N_FFT, torch.hann_window(N_FFT), center=True, return_complex=True).abs()
        mel = torch.log(torch.clamp(torch.from_numpy(mel_basis()) @ spec, min=1e-5))            # [1, 128, n]
        n = mel.shape[-1]
        mel = F.pad(mel, (0, 32 * ((n - 1) // 32 + 1) - n), mode="reflect")
        x = bn_fit("unet.encoder.bn", mel.transpose(-1, -2).unsqueeze(1))
        skips = []
        for i in range(n_enc):
            for b in range(n_blocks):
                x = block(f"unet.encoder.layers.{i}.conv.{b}.", x)
            skips.append(x)
            x = F.avg_pool2d(x, 2)
        for i in range(n_inter):
            for b in range(n_blocks):
                x = block(f"unet.intermediate.layers.{i}.conv.{b}.", x)
        for i in range(n_enc):
            p = f"unet.decoder.layers.{i}."
            x = F.relu(bn_fit(p + "conv1.1", F.conv_transpose2d(x, out[p + "conv1.0.weight"], stride=2, padding=1, output_padding=1)))
            x = torch.cat((x, skips[-1 - i]), dim=1)
            for b in range(n_blocks):
                x = block(p + f"conv2.{b}.", x)
        x = F.conv2d(x, out["cnn.weight"], out["cnn.bias"], padding=1)
        x = x.transpose(1, 2).flatten(-2)
        # standardise the GRU input so the gates work in their active range
        mu, sdv = x.mean(), x.std().clamp_min(1e-6)
        out["cnn.weight"] = out["cnn.weight"] / sdv
        out["cnn.bias"] = (out["cnn.bias"] - mu) / sdv
        x = (x - mu) / sdv
        Hh = out["fc.0.gru.weight_hh_l0"].shape[1]
        gru = torch.nn.GRU(x.shape[-1], Hh, num_layers=1, batch_first=True, bidirectional=True)
        gru.load_state_dict({k_[len("fc.0.gru."):]: v for k_, v in out.items() if k_.startswith("fc.0.gru.")})
        h = gru.eval()(x)[0][0]                                                                   # [T, 2H]
        # smooth-across-bins head: W = G W0 with a Gaussian G over the 360 pitch bins, logits standardised on the clip
        w0 = out["fc.1.weight"]
        nb = w0.shape[0]
        k = torch.arange(nb, dtype=torch.float32)
        G = torch.exp(-0.5 * ((k[:, None] - k[None, :]) / peak_sigma_bins) ** 2)
        G = G / G.sum(1, keepdim=True)
        w = G @ w0
        z = h @ w.t()
        w = w * (logit_gain / 3.0 / z.std().clamp_min(1e-6))           # 3 sigma of the logits ~ logit_gain
        wv = w0[nb // 2:nb // 2 + 8].mean(0)                            # a direction of the GRU output: per-frame voicing term
        wv = wv * (voicing_sigma / (h @ wv).std().clamp_min(1e-6))
        w = w + wv[None, :]
        out["fc.1.weight"] = w
        out["fc.1.bias"] = torch.full((nb,), float(logit_bias)) - (h @ w.t()).mean(0)
    return out


_TRAINED_LIKE_CACHE: Dict[tuple, Dict[str, torch.Tensor]] = {}


def calibration_clip(seconds: float = 6.0, sr: int = 16000, seed: int = 11) -> torch.Tensor:
    """Seeded vocal-like clip (harmonic stack with vibrato + unvoiced bursts + noise floor) used to fit the statistics."""
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * sr)
    t = torch.arange(n, dtype=torch.float64) / sr
    f0 = 220.0 * 2 ** (0.5 * torch.sin(2 * math.pi * 0.23 * t)) * 2 ** (30 / 1200 * torch.sin(2 * math.pi * 5.5 * t))
    ph = 2 * math.pi * torch.cumsum(f0, 0) / sr
    x = sum(torch.sin(k * ph) / k for k in range(1, 9))
    noise = torch.randn(n, generator=g, dtype=torch.float64)
    x = torch.where((t % 2.0) > 1.7, 0.7 * noise, x) + 0.05 * noise
    return (0.5 * x / x.abs().max()).float()


def make_rmvpe_trained_like(seed: int = 4321) -> Dict[str, torch.Tensor]:
    """make_rmvpe_state_dict + calibrate_rmvpe on the seeded calibration clip (cached per process)."""
    key = ("rmvpe", seed)
    if key not in _TRAINED_LIKE_CACHE:
        # BatchNorm shifts well above zero / small scales: the ReLUs work mostly in their linear range, which keeps a RANDOM
        # deep BatchNorm-ReLU stack from being chaotic (such stacks amplify any perturbation ~1.3x per layer at init; trained
        # networks do not).  Without it fp32 rounding noise grows to 1e-5 of salience, TF32 noise to 10 % of an MDX stem.
        _TRAINED_LIKE_CACHE[key] = calibrate_rmvpe(make_rmvpe_state_dict(seed, bn_gamma=(0.3, 0.5), bn_beta=(0.6, 0.1)),
                                                   calibration_clip())
    return _TRAINED_LIKE_CACHE[key]


def calibration_song(seconds: float = 3.0, sr: int = 44100, seed: int = 4242) -> torch.Tensor:
    """Seeded stereo test song [2, n] (pink-ish bed + chord tones + a vocal-like line), peak-normalised."""
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * sr)
    t = torch.arange(n, dtype=torch.float64) / sr
    f0 = 220.0 * 2 ** (0.5 * torch.sin(2 * math.pi * 0.2 * t))
    ph = 2 * math.pi * torch.cumsum(f0, 0) / sr
    vocal = sum(torch.sin(k * ph) / k for k in range(1, 9))
    out = []
    for ch in range(2):
        white = torch.randn(n, generator=g, dtype=torch.float64)
        spec = torch.fft.rfft(white)
        spec = spec / torch.sqrt(torch.arange(spec.numel(), dtype=torch.float64).clamp_min(1.0))
        bed = torch.fft.irfft(spec, n)
        bed = bed * (0.25 / bed.abs().max())
        chord = sum(torch.sin(2 * math.pi * f * (1 + 0.002 * ch) * t + ch) for f in (130.8, 164.8, 196.0))
        out.append(bed + 0.15 * chord + 0.35 * vocal * (1.0 - 0.1 * ch))
    x = torch.stack(out)
    return (x / x.abs().max()).float()


def make_mdx_trained_like(dim_f: int = 3072, dim_t: int = 256, n_fft: int = 7680, seed: int = 2024, cal_frames: int = 64,
                          **kw) -> Dict[str, torch.Tensor]:
    """make_mdx_state_dict + calibrate_mdx_batchnorm on the STFT of the seeded calibration song (the BatchNorm statistics a
    trained network would carry), cached per process.  The calibration runs on `cal_frames` time frames (the statistics
    are per channel over batch x time x frequency, and the TDF linears act along frequency only), which keeps it to a few
    seconds of CPU time at the full 3072-bin geometry."""
    key = ("mdx", dim_f, dim_t, n_fft, seed, cal_frames, tuple(sorted(kw.items())))
    if key not in _TRAINED_LIKE_CACHE:
        kw.setdefault("bn_gamma", (0.3, 0.5))          # near-linear BatchNorm-ReLU operating point: see make_rmvpe_trained_like
        kw.setdefault("bn_beta", (0.6, 0.1))
        sd = make_mdx_state_dict(dim_f=dim_f, dim_t=dim_t, seed=seed, **kw)
        n_lvl = int(sd["_meta"][4])
        T = max(cal_frames, 2 ** n_lvl)
        T = min(T, dim_t)
        hop = 1024
        song = calibration_song(max(3.0, (hop * (T - 1) + n_fft) / 44100.0 + 1.5))
        seg = song[:, 22050:22050 + hop * (T - 1)]
        z = torch.stft(seg, n_fft=n_fft, hop_length=hop, window=torch.hann_window(n_fft), center=True, return_complex=True)
        z = torch.view_as_real(z).permute(0, 3, 1, 2)[:, :, :dim_f]                 # [ch, ri, F, T]
        _TRAINED_LIKE_CACHE[key] = calibrate_mdx_batchnorm(sd, z.reshape(1, 4, dim_f, T), out_gain=0.3)
    return _TRAINED_LIKE_CACHE[key]


# ---------------------------------------------------------------------------
# torchcrepe 'full' weights (torchcrepe/assets/full.pth layout; requirements.txt:19, vc_infer_pipeline.py:116-126)
# ---------------------------------------------------------------------------
CREPE_CHANNELS = [1024, 128, 128, 128, 256, 512]


def make_crepe_state_dict(seed: int = 606, calibrate: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded synthetic Crepe('full') checkpoint: conv{i}.weight [Cout, Cin, k, 1] (k = 512 then 64), conv{i}.bias,
    conv{i}_BN.{weight,bias,running_mean,running_var,num_batches_tracked}, classifier.{weight [360, 2048], bias}.
    calibrate=True fits the BatchNorm statistics on normalised frames of the seeded calibration clip and gives the classifier
    a smooth-across-bins structure (one activation bump per frame), like make_rmvpe_trained_like."""
    import torch.nn.functional as F

    g = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    cin = 1
    for i, co in enumerate(CREPE_CHANNELS):
        k = 512 if i == 0 else 64
        sd[f"conv{i + 1}.weight"] = g.conv((co, cin, k, 1), 1.4)
        sd[f"conv{i + 1}.bias"] = g.normal((co,), 0.05, 0.1)
        p = f"conv{i + 1}_BN"
        sd[p + ".weight"] = g.uniform((co,), 0.6, 1.4)
        sd[p + ".bias"] = g.normal((co,), 0.1, 0.2)
        sd[p + ".running_mean"] = g.normal((co,), 0.1, 0.3)
        sd[p + ".running_var"] = g.uniform((co,), 0.5, 1.5)
        sd[p + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        cin = co
    sd["classifier.weight"] = g.conv((360, 2048), 1.0)
    sd["classifier.bias"] = g.normal((360,), 0.2, -1.0)
    if not calibrate:
        return sd
    clip = calibration_clip(4.0)
    hop = 160
    a = F.pad(clip[None], (512, 512))
    frames = F.unfold(a[:, None, None, :], kernel_size=(1, 1024), stride=(1, hop)).transpose(1, 2).reshape(-1, 1024)
    frames = frames - frames.mean(dim=1, keepdim=True)
    frames = frames / torch.max(torch.tensor(1e-10), frames.std(dim=1, keepdim=True))
    eps = 0.0010000000474974513
    with torch.no_grad():
        x = frames[:, None, :, None]
        for i in range(6):
            x = F.pad(x, (0, 0, 254, 254) if i == 0 else (0, 0, 31, 32))
            x = F.relu(F.conv2d(x, sd[f"conv{i + 1}.weight"], sd[f"conv{i + 1}.bias"], stride=(4, 1) if i == 0 else (1, 1)))
            p = f"conv{i + 1}_BN"
            sd[p + ".running_mean"] = x.mean(dim=(0, 2, 3))
            sd[p + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-4)
            x = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)
            x = F.max_pool2d(x, (2, 1), (2, 1))
        h = x.permute(0, 2, 1, 3).reshape(x.shape[0], -1)
        k = torch.arange(360, dtype=torch.float32)
        G = torch.exp(-0.5 * ((k[:, None] - k[None, :]) / 2.0) ** 2)
        w = (G / G.sum(1, keepdim=True)) @ sd["classifier.weight"]
        z = h @ w.t()
        w = w * (2.5 / z.std().clamp_min(1e-6))
        sd["classifier.weight"] = w
        sd["classifier.bias"] = torch.full((360,), -3.0) - (h @ w.t()).mean(0)
    return sd
