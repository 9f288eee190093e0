"""TEST INFRASTRUCTURE ONLY — torch-CPU fp32 restatement of the RVC synthesizer inference path.

Follows infer_pack/models.py:745-751 (SynthesizerTrnMs768NSFsid.infer; 256 variant :634-640) and
its callees; pinned against the reference's own modules by tests/test_oracle_vs_reference.py.
Takes the checkpoint's raw state dict (weight_g / weight_v pairs included).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def wn_weight(sd: SD, name: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm over all dims but 0
    (modules.py:163,175,185,229-296; models.py:454)."""
    if name + ".weight" in sd:
        return sd[name + ".weight"].float()
    v, g = sd[name + ".weight_v"].float(), sd[name + ".weight_g"].float()
    nrm = v.flatten(1).norm(dim=1).reshape(g.shape)
    return v * (g / nrm)


def layer_norm_c(x, gamma, beta, eps=1e-5):
    """modules.py:29-32 — LayerNorm over the channel dim of [B, C, T]."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


# ---------------------------------------------------------------------------
# relative-position attention encoder (attentions.py:13-73, 216-348, 391-417)
# ---------------------------------------------------------------------------
def rel_attention(sd: SD, p: str, x, n_heads: int, window: int = 10):
    b, d, t = x.shape
    dk = d // n_heads
    q = F.conv1d(x, sd[p + "conv_q.weight"], sd[p + "conv_q.bias"])
    k = F.conv1d(x, sd[p + "conv_k.weight"], sd[p + "conv_k.bias"])
    v = F.conv1d(x, sd[p + "conv_v.weight"], sd[p + "conv_v.bias"])
    q = q.view(b, n_heads, dk, t).transpose(2, 3) / math.sqrt(dk)
    k = k.view(b, n_heads, dk, t).transpose(2, 3)
    v = v.view(b, n_heads, dk, t).transpose(2, 3)
    scores = q @ k.transpose(-2, -1)
    # banded relative-key bias: score[i, j] += q_i . emb_rel_k[j - i + W] for |j - i| <= W  (attentions.py:238-243)
    ek, ev = sd[p + "emb_rel_k"][0], sd[p + "emb_rel_v"][0]            # [2W+1, dk]
    rel = q @ ek.t()                                                    # [b, h, t, 2W+1]
    idx = torch.arange(t)
    diff = idx[None, :] - idx[:, None]                                  # j - i
    band = diff.abs() <= window
    gather = (diff + window).clamp(0, 2 * window)
    scores = scores + torch.where(band, rel.gather(-1, gather.expand(b, n_heads, t, t)), torch.zeros(()))
    # mask is all ones in inference (vc_infer_pipeline.py:453) -> masked_fill(-1e4) is a no-op
    pa = F.softmax(scores, dim=-1)
    out = pa @ v
    # banded relative-value term (attentions.py:264-271)
    pband = torch.where(band, pa, torch.zeros(()))
    relw = torch.zeros(b, n_heads, t, 2 * window + 1)
    relw.scatter_add_(-1, gather.expand(b, n_heads, t, t), pband)
    out = out + relw @ ev
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return F.conv1d(out, sd[p + "conv_o.weight"], sd[p + "conv_o.bias"])


def ffn(sd: SD, p: str, x, k: int):
    """attentions.py:391-399 (activation=None -> relu, same padding)."""
    pl, pr = (k - 1) // 2, k // 2
    y = F.conv1d(F.pad(x, (pl, pr)), sd[p + "conv_1.weight"], sd[p + "conv_1.bias"])
    y = torch.relu(y)
    return F.conv1d(F.pad(y, (pl, pr)), sd[p + "conv_2.weight"], sd[p + "conv_2.bias"])


def text_encoder(sd: SD, phone, pitch, n_heads, n_layers, ksz, hidden):
    """models.py:93-108 (mask all ones for a single full-length item)."""
    x = F.linear(phone, sd["enc_p.emb_phone.weight"], sd["enc_p.emb_phone.bias"])
    if pitch is not None:
        x = x + F.embedding(pitch, sd["enc_p.emb_pitch.weight"])
    x = x * math.sqrt(hidden)
    x = F.leaky_relu(x, 0.1)
    x = x.transpose(1, -1)
    for i in range(n_layers):
        y = rel_attention(sd, f"enc_p.encoder.attn_layers.{i}.", x, n_heads)
        x = layer_norm_c(x + y, sd[f"enc_p.encoder.norm_layers_1.{i}.gamma"], sd[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        y = ffn(sd, f"enc_p.encoder.ffn_layers.{i}.", x, ksz)
        x = layer_norm_c(x + y, sd[f"enc_p.encoder.norm_layers_2.{i}.gamma"], sd[f"enc_p.encoder.norm_layers_2.{i}.beta"])
    stats = F.conv1d(x, sd["enc_p.proj.weight"], sd["enc_p.proj.bias"])
    m, logs = torch.split(stats, stats.shape[1] // 2, dim=1)
    return m, logs


# ---------------------------------------------------------------------------
# reverse flow (models.py:146-153; modules.py:440-459, 188-213, 377-384)
# ---------------------------------------------------------------------------
def wn_stack(sd: SD, p: str, h, g, hidden: int, n_layers: int = 3, k: int = 5):
    out = torch.zeros_like(h)
    cond = F.conv1d(g, wn_weight(sd, p + "cond_layer"), sd[p + "cond_layer.bias"])
    for i in range(n_layers):
        x_in = F.conv1d(h, wn_weight(sd, p + f"in_layers.{i}"), sd[p + f"in_layers.{i}.bias"], padding=(k - 1) // 2)
        a = x_in + cond[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])      # commons.py:105-112
        rs = F.conv1d(acts, wn_weight(sd, p + f"res_skip_layers.{i}"), sd[p + f"res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            h = h + rs[:, :hidden]
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out


def flow_reverse(sd: SD, z, g, hidden: int, n_flows: int = 4):
    x = z
    half = x.shape[1] // 2
    for f in reversed(range(n_flows)):
        x = torch.flip(x, [1])                                               # Flip comes first in reverse order
        p = f"flow.flows.{2 * f}."
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[p + "pre.weight"], sd[p + "pre.bias"])
        h = wn_stack(sd, p + "enc.", h, g, hidden)
        m = F.conv1d(h, sd[p + "post.weight"], sd[p + "post.bias"])
        x = torch.cat([x0, x1 - m], 1)                                       # mean_only: logs = 0
    return x


# ---------------------------------------------------------------------------
# NSF source + HiFi-GAN generator (models.py:281-419, 494-516; modules.py:299-312)
# ---------------------------------------------------------------------------
def sine_source(f0, upp: int, sr: int, noise, rand_ini: Optional[torch.Tensor] = None):
    """SineGen.forward for harmonic_num=0 (models.py:320-370). f0 [1,T]; noise [1,T*upp,1] ~ N(0,1)
    is the randn_like draw of line 368. Returns sine_waves [1, T*upp, 1]."""
    f0 = f0[:, None].transpose(1, 2)                                         # [1,T,1]
    rad = (f0 / sr) % 1
    # rand_ini is drawn but zeroed for the fundamental (models.py:337-341)
    tmp = torch.cumsum(rad, 1) * upp
    tmp = F.interpolate(tmp.transpose(2, 1), scale_factor=float(upp), mode="linear", align_corners=True).transpose(2, 1)
    rad_up = F.interpolate(rad.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    tmp = tmp % 1
    wrap = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
    shift = torch.zeros_like(rad_up)
    shift[:, 1:, :] = wrap * -1.0
    sine = torch.sin(torch.cumsum(rad_up + shift, dim=1) * 2 * np.pi) * 0.1
    uv = (f0 > 0).float()
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    return sine * uv + noise_amp * noise


def resblock1(sd: SD, p: str, x, k: int, dil):
    for m, d in enumerate(dil):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, wn_weight(sd, p + f"convs1.{m}"), sd[p + f"convs1.{m}.bias"], dilation=d, padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, wn_weight(sd, p + f"convs2.{m}"), sd[p + f"convs2.{m}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def generator(sd: SD, z, f0, g, cfg, noise):
    up_r, up_k, rb_k, rb_d, sr = cfg[12], cfg[14], cfg[10], cfg[11], cfg[17]
    upp = int(np.prod(up_r))
    nsf = f0 is not None          # f0=None: the plain `Generator` of the *_nono models (models.py:188-250), no source branch
    if nsf:
        sine = sine_source(f0, upp, sr, noise)
        har = torch.tanh(F.linear(sine, sd["dec.m_source.l_linear.weight"], sd["dec.m_source.l_linear.bias"])).transpose(1, 2)
    x = F.conv1d(z, sd["dec.conv_pre.weight"], sd["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, sd["dec.cond.weight"], sd["dec.cond.bias"])
    nk = len(rb_k)
    for i, (u, k) in enumerate(zip(up_r, up_k)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, wn_weight(sd, f"dec.ups.{i}"), sd[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if nsf:
            if i + 1 < len(up_r):
                s = int(np.prod(up_r[i + 1:]))
                xs = F.conv1d(har, sd[f"dec.noise_convs.{i}.weight"], sd[f"dec.noise_convs.{i}.bias"], stride=s, padding=s // 2)
            else:
                xs = F.conv1d(har, sd[f"dec.noise_convs.{i}.weight"], sd[f"dec.noise_convs.{i}.bias"])
            x = x + xs
        acc = None
        for j in range(nk):
            r = resblock1(sd, f"dec.resblocks.{i * nk + j}.", x, rb_k[j], rb_d[j])
            acc = r if acc is None else acc + r
        x = acc / nk
    x = F.leaky_relu(x)                                                      # default slope 0.01 (models.py:513)
    x = F.conv1d(x, sd["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


def infer(cpt: dict, phone, pitch, nsff0, sid, noise_z, noise_src, return_all: bool = False):
    """SynthesizerTrnMs{256,768}NSFsid.infer (models.py:634-640, 745-751) and, with pitch = nsff0 = None, the
    `_nono` variants (models.py:847-853, 949-955: no pitch embedding, plain HiFi-GAN `Generator`).
    phone [1,P,768|256] f32, pitch [1,P] i64, nsff0 [1,P] f32, sid [1] i64,
    noise_z [1,192,P] (the randn_like of line 748), noise_src [1,P*upp,1] (line 368). -> [1,1,P*upp]."""
    sd = {k: v.float() if v.is_floating_point() else v for k, v in cpt["weight"].items()}
    cfg = cpt["config"]
    hidden, n_heads, n_layers, ksz = cfg[3], cfg[5], cfg[6], cfg[7]
    with torch.no_grad():
        g = F.embedding(sid, sd["emb_g.weight"]).unsqueeze(-1)
        m_p, logs_p = text_encoder(sd, phone, pitch, n_heads, n_layers, ksz, hidden)
        z_p = m_p + torch.exp(logs_p) * noise_z * 0.66666
        z = flow_reverse(sd, z_p, g, hidden)
        o = generator(sd, z, nsff0, g, cfg, noise_src)
    if return_all:
        return o, dict(m_p=m_p, logs_p=logs_p, z_p=z_p, z=z, g=g)
    return o


def draw_noise(seed: int, P: int, inter: int, upp: int):
    """Replays the reference's RNG consumption inside net_g.infer after torch.manual_seed(seed):
    randn_like(m_p) [1,inter,P] (models.py:748), rand(1,1) (:337), randn_like(sine) [1,P*upp,1] (:368)."""
    torch.manual_seed(seed)
    nz = torch.randn(1, inter, P)
    _ = torch.rand(1, 1)
    ns = torch.randn(1, P * upp, 1)
    return nz, ns
