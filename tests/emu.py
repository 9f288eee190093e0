"""CPU emulator of the tap-GEMM descriptor semantics (TEST INFRASTRUCTURE ONLY).

Executes a `TapGemm` built over CPU tensors exactly as csrc/tapgemm.cuh defines
it, in float64, so the host-side lowerings in aicovergen_b200/tapgemm.py can be
validated against torch.nn.functional without a GPU.  Never imported by the
product path.
"""
from __future__ import annotations

import math

import torch

from aicovergen_b200._ffi import (ACT_EXP, ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID,
                                  ACT_TANH)


def _flat(t: torch.Tensor) -> torch.Tensor:
    """1-D alias of the whole storage of t."""
    n = t.untyped_storage().nbytes() // t.element_size()
    return torch.as_strided(t, (n,), (1,), 0)


def _act(v, code, p):
    if code == ACT_NONE:
        return v
    if code == ACT_RELU:
        return torch.relu(v)
    if code == ACT_LRELU:
        return torch.where(v > 0, v, v * p)
    if code == ACT_GELU:
        return 0.5 * v * (1 + torch.erf(v / math.sqrt(2.0)))
    if code == ACT_TANH:
        return torch.tanh(v)
    if code == ACT_SIGMOID:
        return torch.sigmoid(v)
    if code == ACT_EXP:
        return torch.exp(v)
    raise ValueError(code)


def _rn_tf32(t: torch.Tensor) -> torch.Tensor:
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def emulate(op) -> None:
    """Run op on its (CPU) tensors, writing op.out (and out2) in place."""
    p = op.params
    a, w, out, epi = op.a, op.w, op.out, op.epi
    A = _flat(a.t).double()
    a_base = a.t.storage_offset() + a.off
    Wf = _flat(w.t).double()
    w_base = w.t.storage_offset() + w.off
    OW, OH, OB, N, Kc = p.OW, p.OH, p.OB, p.N, p.Kc
    bb, hh, ww = torch.meshgrid(torch.arange(OB), torch.arange(OH), torch.arange(OW), indexing="ij")
    bb, hh, ww = bb.reshape(-1), hh.reshape(-1), ww.reshape(-1)
    M = bb.numel()
    acc = torch.zeros(M, N, dtype=torch.float64)
    kk = torch.arange(Kc)
    for (c_off, dw, dh, dp, widx) in op.taps:
        iw, ih = ww + dw, hh + dh
        pix_ok = (iw >= 0) & (iw < a.dims[1]) & (ih >= 0) & (ih < a.dims[2]) & (bb < a.dims[3])
        c = c_off + kk
        c_ok = c < a.dims[0]
        idx = (a_base + iw.clamp(0, max(a.dims[1] - 1, 0)) * a.strides[1]
               + ih.clamp(0, max(a.dims[2] - 1, 0)) * a.strides[2] + bb * a.strides[3] + dp * a.strides[4])
        idx2 = idx[:, None] + c.clamp(max=max(a.dims[0] - 1, 0))[None, :]
        vals = A[idx2] * (pix_ok[:, None] & c_ok[None, :])
        # weights: per batch slice
        wsl = widx + bb * p.w_batch_step
        for ws in wsl.unique().tolist():
            sel = wsl == ws
            widx2 = (w_base + ws * w.wstride + torch.arange(N)[:, None] * w.ldw + kk[None, :])
            Wm = Wf[widx2]
            acc[sel] += vals[sel] @ Wm.t()
    # ---- epilogue
    v = acc
    if getattr(epi, "acc_in", None) is not None:
        ai = epi.acc_in
        st = list(ai.stride())
        while len(st) < 4:
            st.insert(0, 0)
        aif = _flat(ai).double()
        v = v + aif[(ai.storage_offset() + bb * st[0] + hh * st[1] + ww * st[2])[:, None] + torch.arange(N)[None, :]]
    if epi.row_scale_pre is not None:
        v = v * epi.row_scale_pre.double()[(hh * OW + ww)][:, None]
    if epi.bias is not None:
        bias = epi.bias.double()
        if epi.bias_per_row:
            v = v + bias[(hh * OW + ww)][:, None]
        else:
            v = v + bias[None, :N]
    v = _act(v, epi.act_pre, epi.act_pre_p)
    if epi.row_scale is not None:
        v = v * epi.row_scale.double()[(hh * OW + ww)][:, None]
    n_idx = torch.arange(N)
    if epi.res is not None:
        r = epi.res
        if epi.res_strides is not None:
            st = list(epi.res_strides)
        else:
            st = list(r.stride())
            while len(st) < 4:
                st.insert(0, 0)
            st = st[:3] + [1]
        rf = _flat(r).double()
        if epi.res_mapped:
            ridx = r.storage_offset() + bb * st[0] + (hh * out.osh + out.ooh).clamp(0, out.fh - 1) * st[1] + (ww * out.osw + out.oow).clamp(0, out.fw - 1) * st[2]
        else:
            ridx = r.storage_offset() + bb * st[0] + hh * st[1] + ww * st[2]
        rv = rf[ridx[:, None] + n_idx[None, :] * st[3]]
        if getattr(epi, "res_split", 0):
            rv = rv + rf[ridx[:, None] + n_idx[None, :] * st[3] + epi.res_split]
        v = v * rv if epi.res_mul else v + rv
    v = v * epi.scale
    mh, mw = hh * out.osh + out.ooh, ww * out.osw + out.oow
    valid = (mh >= 0) & (mh < out.fh) & (mw >= 0) & (mw < out.fw)
    o_rel = bb * out.sb + mh * out.sh + mw * out.sw
    o_idx = (out.t.storage_offset() + out.off + o_rel)[:, None] + n_idx[None, :] * out.sn
    if epi.res2 is not None:
        r2 = _flat(epi.res2).double()
        r2_idx = (epi.res2.storage_offset() + out.off + o_rel)[:, None] + n_idx[None, :] * out.sn
        v = v + torch.where(valid[:, None], r2[r2_idx.clamp(0, r2.numel() - 1)], torch.zeros((), dtype=torch.float64))
    v = _act(v, epi.act_post, epi.act_post_p)
    of = _flat(out.t)
    sel = valid
    if getattr(epi, "split_out", 0):        # 3xTF32 planes hi | lo | hi
        v32 = v.float()
        hi = _rn_tf32(v32)
        lo = _rn_tf32(v32 - hi)
        for plane, val in ((0, hi), (1, lo), (2, hi)):
            of[(o_idx[sel] + plane * epi.split_out).reshape(-1)] = val[sel].reshape(-1)
        return
    of[o_idx[sel].reshape(-1)] = v[sel].reshape(-1).to(of.dtype)
    if epi.out2 is not None and hasattr(epi.out2, "sn"):       # tapgemm.Out: its own layout, addressed by the mapped pixel
        q = epi.out2
        o2 = _flat(q.t)
        o2_idx = (q.t.storage_offset() + q.off + bb * q.sb + mh * q.sh + mw * q.sw)[:, None] + n_idx[None, :] * q.sn
        o2[o2_idx[sel].reshape(-1)] = _act(v, epi.act2, epi.act2_p)[sel].reshape(-1).to(o2.dtype)
    elif epi.out2 is not None:
        o2 = _flat(epi.out2)
        o2_idx = (epi.out2.storage_offset() + out.off + o_rel)[:, None] + n_idx[None, :] * out.sn
        o2[o2_idx[sel].reshape(-1)] = _act(v, epi.act2, epi.act2_p)[sel].reshape(-1).to(o2.dtype)
