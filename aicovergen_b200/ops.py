"""Thin torch-tensor wrappers over the non-GEMM C-ABI kernels (include/b200vc.h).

Every function enqueues on the current torch CUDA stream and returns immediately.
Tensors are passed by data_ptr(); nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _ffi

_P = C.c_void_p
_i64, _i32, _f32 = C.c_int64, C.c_int, C.c_float

_ffi.declare("b200vc_layernorm", [_P, _P, _P, _P, _P, _i64, _i32, _i64, _i64, _i64, _f32, _i32, _P])
_ffi.declare("b200vc_softmax_rows", [_P, _i32, _i32, _i32, _i64, _i64, _P, _i32, _P, _i32, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_relpos_value_add", [_P, _i32, _P, _i32, _i32, _i64, _i64, _P, _i32, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_gather_rows", [_P, _P, _P, _i64, _i32, _P])
_ffi.declare("b200vc_gate_tanh_sigmoid", [_P, _P, _i64, _i32, _i32, _P])
_ffi.declare("b200vc_zp_sample", [_P, _P, _P, _i64, _i32, _f32, _P])
_ffi.declare("b200vc_axpby", [_P, _P, _P, _i64, _f32, _f32, _P])
_ffi.declare("b200vc_act", [_P, _P, _i64, _i32, _f32, _i32, _P])
_ffi.declare("b200vc_nsf_source", [_P, _P, _P, _P, _i32, _i32, _f32, _f32, _f32, _P])
_ffi.declare("b200vc_conv1d_to1", [_P, _P, _P, _i64, _i32, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_reflect_pad_1d", [_P, _P, _i64, _i64, _P])
_ffi.declare("b200vc_magnitude", [_P, _P, _i64, _i32, _i64, _i64, _P])
_ffi.declare("b200vc_logmel_affine_reflect", [_P, _P, _i32, _i32, _i32, _f32, _f32, _f32, _P])
_ffi.declare("b200vc_avgpool2x2", [_P, _P, _i32, _i32, _i32, _i32, _i64, _P])
_ffi.declare("b200vc_avgpool2x2_split", [_P, _P, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _P])
_ffi.declare("b200vc_crepe_frames", [_P, _i64, _i64, _i32, _i32, _P, _i64, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_maxpool2_affine", [_P, _P, _P, _P, _i64, _i32, _i32, _P])
_ffi.declare("b200vc_crepe_logprob", [_P, _P, _i32, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_viterbi_band", [_P, _P, C.c_double, C.c_double, _P, _P, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_bigru", [_P, _P, _P, _P, _i32, _i32, _P])
_ffi.declare("b200vc_rmvpe_decode", [_P, _P, _P, _i32, _i32, _i64, _f32, _P])
_ffi.declare("b200vc_groupnorm_time", [_P, _P, _P, _P, _P, _i64, _i32, _f32, _i32, _i32, _P])
_ffi.declare("b200vc_argmin_rows", [_P, _P, _i32, _i32, _i64, _P])
_ffi.declare("b200vc_ivf_scan_blend", [_P, _i64, _P, _P, _P, _P, _P, _i64, _i32, _i32, _f32, _P, _P, _P])
_ffi.declare("b200vc_upsample2_protect", [_P, _P, _P, _P, _i64, _i32, _f32, _i32, _P])
_ffi.declare("b200vc_boxsum_f64", [_P, _P, _i64, _i32, _P])
_ffi.declare("b200vc_mdx_gather_chunks", [_P, _i64, _P, _P, _P, _P, _i32, _i32, _i32, _f32, _i32, _P])
_ffi.declare("b200vc_nhwc_to_nhcw", [_P, _P, _P, _i64, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_nhcw_to_nhwc_add", [_P, _P, _P, _i64, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_mdx_ola_store", [_P, _P, _P, _P, _P, _P, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _P])
_ffi.declare("b200vc_mdx_finalize", [_P, _P, _P, _i64, _f32, _f32, _P])
_ffi.declare("b200vc_resample_sinc_mono", [_P, _i64, _i32, _P, _i64, C.c_double, _i32, _P])
_ffi.declare("b200vc_change_rms", [_P, _i64, _i32, _P, _i64, _i32, C.c_double, _P, _P])
_ffi.declare("b200vc_to_int16_peak_guard", [_P, _i64, _P, _P, _P])
_ffi.declare("b200vc_sosfiltfilt_f64", [_P, _i64, _P, _i32, _P, _P, _P, _i32, _i32, _P, _P, _P])


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor):
    assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)
    # launches go to the current device's stream (see _ffi.on_device): refuse cross-device calls instead of faulting
    assert t.device.index == torch.cuda.current_device(), f"tensor on {t.device} but cuda:{torch.cuda.current_device()} is current"
    return t


def layernorm(x, gamma, beta, out, res=None, eps=1e-5, round_out=False):
    """out[r,:] = LN(x[r,:] + res[r,:]); x/res/out are [rows, C] with unit channel stride."""
    rows, Cc = x.shape
    assert x.stride(1) == 1 and out.stride(1) == 1 and (res is None or res.stride(1) == 1)
    _ffi.check(_ffi.lib().b200vc_layernorm(_p(_f32c(x)), _p(res), _p(gamma), _p(beta), _p(out), rows, Cc,
                                           x.stride(0), 0 if res is None else res.stride(0), out.stride(0),
                                           eps, int(round_out), _s()), "layernorm")


def softmax_rows(S, T, q=None, emb_rel_k=None, window=0, round_out=False, row0=0):
    """In-place softmax over S[heads, rows, :T]; optional relative-key bias (q [T_q, heads*dk] the FULL query matrix,
    emb_rel_k [2W+1, dk]); S holds query rows row0 .. row0 + rows - 1."""
    heads, rows, _ = S.shape
    assert S.stride(2) == 1
    dk = 0 if emb_rel_k is None else emb_rel_k.shape[-1]
    _ffi.check(_ffi.lib().b200vc_softmax_rows(_p(_f32c(S)), heads, rows, T, S.stride(1), S.stride(0), _p(q),
                                              0 if q is None else q.stride(0), _p(emb_rel_k), window, dk,
                                              int(round_out), int(row0), _s()), "softmax_rows")


def relpos_value_add(out, P, T, emb_rel_v, window, heads, row0=0):
    """out (full [T_q, heads*dk]) rows row0.. += banded P (block [heads, rows, >=T]) x emb_rel_v."""
    dk = emb_rel_v.shape[-1]
    _ffi.check(_ffi.lib().b200vc_relpos_value_add(_p(_f32c(out)), out.stride(0), _p(P), P.shape[1], T, P.stride(1), P.stride(0),
                                                  _p(emb_rel_v), window, dk, heads, int(row0), _s()), "relpos_value_add")


def gather_rows(table, idx, out):
    assert idx.dtype == torch.int64 and idx.is_cuda and table.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_gather_rows(_p(_f32c(table)), _p(idx), _p(out), idx.numel(), table.shape[1], _s()),
               "gather_rows")


def gate_tanh_sigmoid(a, out, round_out=False):
    rows, C2 = a.shape
    assert a.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_gate_tanh_sigmoid(_p(_f32c(a)), _p(out), rows, C2 // 2, int(round_out), _s()), "gate")


def zp_sample(stats, noise, z, scale=0.66666):
    P, C2 = stats.shape
    assert stats.is_contiguous() and noise.is_contiguous() and z.is_contiguous() and noise.numel() == P * (C2 // 2)
    _ffi.check(_ffi.lib().b200vc_zp_sample(_p(_f32c(stats)), _p(noise), _p(z), P, C2 // 2, scale, _s()), "zp_sample")


def axpby(a, b, out, alpha=1.0, beta=1.0):
    assert a.is_contiguous() and out.is_contiguous() and (b is None or b.is_contiguous())
    _ffi.check(_ffi.lib().b200vc_axpby(_p(_f32c(a)), _p(b), _p(out), a.numel(), alpha, beta, _s()), "axpby")


def act(x, out, code, p=0.0, round_out=False):
    assert x.is_contiguous() and out.is_contiguous()
    flags = 2 if out.dtype == torch.float16 else int(round_out)          # bit1: `out` holds fp16
    _ffi.check(_ffi.lib().b200vc_act(_p(_f32c(x)), _p(out), x.numel(), code, p, flags, _s()), "act")


def nsf_source(f0, noise, har, scratch, upp, sr, lin_w, lin_b):
    """har[:T*upp] = tanh(lin_w * sine_source(f0) + lin_b); f0 [T] f32, noise [T*upp] f32, scratch [T] f64."""
    T = f0.numel()
    assert f0.is_contiguous() and noise.is_contiguous() and noise.numel() >= T * upp and scratch.dtype == torch.float64
    _ffi.check(_ffi.lib().b200vc_nsf_source(_p(_f32c(f0)), _p(noise), _p(har), _p(scratch), T, upp, float(sr),
                                            float(lin_w), float(lin_b), _s()), "nsf_source")


_ffi.declare("b200vc_conv1d_from1", [_P, _i64, _P, _P, _P, _P, _P, _i64, _i32, _i32, _i32, _i64, _i32, _f32, _i32, _P])


def conv1d_from1(src, w, out, stride, src_off, bias=None, res=None, out2=None, act2=0, act2_p=0.0, round_out2=False):
    """out[t,c] = res[t,c] + bias[c] + sum_j w[c,j] * src[src_off + t*stride + j]; out2 = act2(out).  src 1-D, w [C,K], out [T,C]."""
    T, Cc = out.shape
    assert src.dim() == 1 and src.is_contiguous() and w.shape[0] == Cc and w.is_contiguous() and out.is_contiguous()
    assert res is None or (res.shape == out.shape and res.is_contiguous())
    assert out2 is None or (out2.shape == out.shape and out2.is_contiguous())
    _ffi.check(_ffi.lib().b200vc_conv1d_from1(_p(_f32c(src)), src.numel(), _p(w), _p(bias), _p(res), _p(out), _p(out2), T, Cc,
                                              w.shape[1], int(stride), int(src_off), int(act2), float(act2_p),
                                              2 if (out2 is not None and out2.dtype == torch.float16) else int(round_out2), _s()),
               "conv1d_from1")


def conv1d_to1(x, w, out, pad, act_code):
    """out[t] = act(sum_{k,c} w[k,c] x[t+k-pad,c]); x [T,C] contiguous, w [K,C]."""
    T, Cc = x.shape
    assert x.is_contiguous() and w.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_conv1d_to1(_p(_f32c(x)), _p(w), _p(out), T, Cc, w.shape[0], pad, act_code, _s()),
               "conv1d_to1")


def reflect_pad_1d(x, out, pad):
    N = x.numel()
    assert x.is_contiguous() and out.numel() >= N + 2 * pad
    _ffi.check(_ffi.lib().b200vc_reflect_pad_1d(_p(_f32c(x)), _p(out), N, pad, _s()), "reflect_pad_1d")


def magnitude(spec, mag, nb):
    rows = spec.shape[0]
    _ffi.check(_ffi.lib().b200vc_magnitude(_p(_f32c(spec)), _p(mag), rows, nb, spec.stride(0), mag.stride(0), _s()),
               "magnitude")


def logmel_affine_reflect(x, out, rows, clampv, a, b):
    rows_total, Cc = out.shape
    assert x.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_logmel_affine_reflect(_p(_f32c(x)), _p(out), rows, rows_total, Cc, clampv, a, b, _s()),
               "logmel_affine_reflect")


def avgpool2x2(x, out):
    """x [B,H,W,C] (channel slice allowed: stride(-1)==1, pixel pitch x.stride(2)), out [B,H/2,W/2,C] contiguous."""
    B, H, W, Cc = x.shape
    assert x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1) and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_avgpool2x2(_p(_f32c(x)), _p(out), B, H, W, Cc, x.stride(2), _s()), "avgpool2x2")


def bigru(xp, whh, bhh, out, hidden):
    T = xp.shape[0]
    assert xp.is_contiguous() and whh.is_contiguous() and bhh.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_bigru(_p(_f32c(xp)), _p(whh), _p(bhh), _p(out), T, hidden, _s()), "bigru")


def avgpool2x2_split(x, in_split, out, out_split):
    """3xTF32 split tensors: x / out are the plane-0 views [B,H,W,C] / [B,H/2,W/2,C] (pixel pitches from the strides)."""
    B, H, W, Cc = x.shape
    assert x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and out.stride(3) == 1 and out.stride(1) == (W // 2) * out.stride(2)
    _ffi.check(_ffi.lib().b200vc_avgpool2x2_split(_p(_f32c(x)), _p(out), B, H, W, Cc, x.stride(2), in_split, out.stride(2),
                                                  out_split, _s()), "avgpool2x2_split")


def crepe_frames(audio, first_frame, hop, win, out, off, nframes, round_out=False):
    """audio [N] f32; out [B, ld] (frame f written at out[f - first_frame, off:off+win])."""
    assert audio.is_contiguous() and out.stride(1) == 1 and nframes <= out.shape[0]
    _ffi.check(_ffi.lib().b200vc_crepe_frames(_p(_f32c(audio)), audio.numel(), first_frame, hop, win, _p(out), out.stride(0), off,
                                              nframes, int(round_out), _s()), "crepe_frames")


def maxpool2_affine(x, scale, shift, out, round_out=False):
    """x [B, L, C] -> out [B, L/2, C] = max over position pairs of scale[c] * x + shift[c]."""
    assert x.is_contiguous() and out.is_contiguous() and x.shape[1] % 2 == 0
    _ffi.check(_ffi.lib().b200vc_maxpool2_affine(_p(_f32c(x)), _p(scale), _p(shift), _p(out), out.shape[0] * out.shape[1], x.shape[2],
                                                 int(round_out), _s()), "maxpool2_affine")


def crepe_logprob(act, logp, lo, hi):
    n, nb = act.shape
    assert act.is_contiguous() and logp.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_crepe_logprob(_p(_f32c(act)), _p(logp), n, nb, lo, hi, _s()), "crepe_logprob")


def viterbi_band(logp, log_band, log_out, log_init, ptr, states, band):
    n, ns = logp.shape
    assert logp.is_contiguous() and log_band.dtype == torch.float64 and log_band.is_contiguous() and ptr.element_size() == 2
    assert states.dtype == torch.int32 and tuple(log_band.shape) == (ns, 2 * band - 1)
    _ffi.check(_ffi.lib().b200vc_viterbi_band(_p(_f32c(logp)), _p(log_band), float(log_out), float(log_init), _p(ptr), _p(states), n, ns,
                                              band, _s()), "viterbi_band")


def rmvpe_decode(sal, f0, T, thred, cents=None):
    assert f0.dtype == torch.float64 and sal.stride(1) == 1 and (cents is None or cents.dtype == torch.float64)
    _ffi.check(_ffi.lib().b200vc_rmvpe_decode(_p(_f32c(sal)), _p(f0), _p(cents), T, sal.shape[1], sal.stride(0), thred,
                                              _s()), "rmvpe_decode")


def groupnorm_time(x, gamma, beta, out, stats, eps=1e-5, act_code=0, round_out=False):
    """Per-channel normalisation over rows of x [rows, C] + activation; stats: [2*C] float64 scratch."""
    rows, Cc = x.shape
    assert x.is_contiguous() and out.is_contiguous() and stats.dtype == torch.float64 and stats.numel() >= 2 * Cc
    _ffi.check(_ffi.lib().b200vc_groupnorm_time(_p(_f32c(x)), _p(gamma), _p(beta), _p(out), _p(stats), rows, Cc, eps,
                                                act_code, int(round_out), _s()), "groupnorm_time")


def argmin_rows(S, out):
    rows, n = S.shape
    assert S.stride(1) == 1 and out.dtype == torch.int32
    _ffi.check(_ffi.lib().b200vc_argmin_rows(_p(_f32c(S)), _p(out), rows, n, S.stride(0), _s()), "argmin_rows")


def ivf_scan_blend(q, assign, offsets, ids, vecs, out, rate, out_score=None, out_ids=None):
    T, d = q.shape
    assert q.stride(1) == 1 and out.stride(1) == 1 and vecs.is_contiguous()
    assert assign.dtype == torch.int32 and offsets.dtype == torch.int32 and ids.dtype == torch.int64
    _ffi.check(_ffi.lib().b200vc_ivf_scan_blend(_p(_f32c(q)), q.stride(0), _p(assign), _p(offsets), _p(ids), _p(vecs),
                                                _p(out), out.stride(0), T, d, float(rate), _p(out_score), _p(out_ids),
                                                _s()), "ivf_scan_blend")


def upsample2_protect(feats, feats0, pitchf, out, protect, do_protect):
    P, Cc = out.shape
    assert feats.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_upsample2_protect(_p(_f32c(feats)), _p(feats0), _p(pitchf), _p(out), P, Cc,
                                                   float(protect), int(do_protect), _s()), "upsample2_protect")


def boxsum_f64(x, out, n, window):
    assert x.dtype == torch.float64 and out.dtype == torch.float64 and x.numel() >= n + window - 1
    _ffi.check(_ffi.lib().b200vc_boxsum_f64(_p(x), _p(out), n, window, _s()), "boxsum_f64")


def mdx_gather_chunks(wave, src_start, lo, hi, out, chunk, half, sign, round_out):
    """wave [2, n_song] contiguous; src_start/lo/hi int64 [B]; out [B, 2, chunk + 2*half]."""
    B = src_start.numel()
    assert wave.is_contiguous() and out.is_contiguous() and src_start.dtype == torch.int64
    _ffi.check(_ffi.lib().b200vc_mdx_gather_chunks(_p(_f32c(wave)), wave.shape[1], _p(src_start), _p(lo), _p(hi), _p(out),
                                                   B, chunk, half, float(sign), int(round_out), _s()), "mdx_gather_chunks")


_ffi.declare("b200vc_mdx_first_conv", [_P, _P, _P, _P, _i32, _i32, _i32, _i32, _i32, _i32, _P])
_ffi.declare("b200vc_mdx_final_conv", [_P, _P, _P, _P, _i32, _i32, _i32, _i32, _i32, _i32, _P])


def mdx_first_conv(spec, w4, bias, out, round_out=False):
    """spec [B,2,T,2F] (f, ri interleaved) -> out [B,T,F,g] = relu(W4 . (ch0re, ch0im, ch1re, ch1im) + bias)."""
    B, T, F, g = out.shape
    assert spec.shape == (B, 2, T, 2 * F) and spec.is_contiguous() and out.is_contiguous() and w4.shape == (g, 4)
    _ffi.check(_ffi.lib().b200vc_mdx_first_conv(_p(_f32c(spec)), _p(w4), _p(bias), _p(out), B, T, F, g, int(round_out),
                                                int(out.dtype == torch.float16), _s()), "mdx_first_conv")


def mdx_final_conv(x, w, bias, spec, round_out=False):
    """x [B,T,F,c] -> spec [B,2,T,2F] = W[4,c] . x + bias (rows of W: ch0re, ch0im, ch1re, ch1im)."""
    B, T, F, c = x.shape
    assert spec.shape == (B, 2, T, 2 * F) and spec.is_contiguous() and x.is_contiguous() and w.shape == (4, c)
    _ffi.check(_ffi.lib().b200vc_mdx_final_conv(_p(x), _p(w), _p(bias), _p(spec), B, T, F, c, int(round_out),
                                                int(x.dtype == torch.float16), _s()), "mdx_final_conv")


def nhwc_to_nhcw(x, scale, out, round_out=False):
    """x [..., W, C] contiguous -> out [..., C, W] (* scale[c])."""
    W, Cc = x.shape[-2], x.shape[-1]
    R = x.numel() // (W * Cc)
    assert x.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_nhwc_to_nhcw(_p(_f32c(x)), _p(scale), _p(out), R, W, Cc, int(round_out), _s()), "nhwc_to_nhcw")


def nhcw_to_nhwc_add(t, x, out, round_out=False):
    """out[..., w, c] = x[..., w, c] + t[..., c, w]."""
    W, Cc = x.shape[-2], x.shape[-1]
    R = x.numel() // (W * Cc)
    assert x.is_contiguous() and out.is_contiguous() and t.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_nhcw_to_nhwc_add(_p(_f32c(t)), _p(x), _p(out), R, W, Cc, int(round_out), _s()),
               "nhcw_to_nhwc_add")


def mdx_ola_store(frames, env, dst_start, keep_lo, keep_hi, song, T, n_fft, hop, chunk, trim, coef, accumulate):
    B = dst_start.numel()
    assert frames.is_contiguous() and song.is_contiguous() and env.numel() >= chunk + n_fft
    _ffi.check(_ffi.lib().b200vc_mdx_ola_store(_p(_f32c(frames)), _p(env), _p(dst_start), _p(keep_lo), _p(keep_hi), _p(song),
                                               song.shape[1], B, T, n_fft, hop, chunk, trim, float(coef), int(accumulate),
                                               _s()), "mdx_ola_store")


def mdx_finalize(proc, wave_norm, inverse, peak, compensation):
    assert proc.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_mdx_finalize(_p(_f32c(proc)), _p(wave_norm), _p(inverse), proc.numel(), float(peak),
                                              float(compensation), _s()), "mdx_finalize")


def resample_sinc_mono(x, out, rate_in, rate_out, zero_crossings=16):
    """x [channels, n_in] -> out [n_out] mono at rate_out."""
    ch, n_in = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    _ffi.check(_ffi.lib().b200vc_resample_sinc_mono(_p(_f32c(x)), n_in, ch, _p(out), out.numel(), float(rate_in) / float(rate_out),
                                                    zero_crossings, _s()), "resample_sinc_mono")


def change_rms(data1, sr1, data2, sr2, rate):
    """In-place loudness-envelope mix of data2 (fp32 device) towards data1 (fp64 device)."""
    assert data1.dtype == torch.float64 and data2.dtype == torch.float32 and data1.is_contiguous() and data2.is_contiguous()
    n1, n2 = data1.numel(), data2.numel()
    scratch = torch.empty(4 + n1 // (sr1 // 2) + n2 // (sr2 // 2), dtype=torch.float64, device=data2.device)
    _ffi.check(_ffi.lib().b200vc_change_rms(_p(data1), n1, sr1, _p(data2), n2, sr2, float(rate), _p(scratch), _s()), "change_rms")


def to_int16_peak_guard(x):
    """float32 device waveform -> int16 device tensor with the reference's peak guard."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
    scratch = torch.zeros(1, dtype=torch.float32, device=x.device)
    _ffi.check(_ffi.lib().b200vc_to_int16_peak_guard(_p(x), x.numel(), _p(scratch), _p(out), _s()), "to_int16_peak_guard")
    return out


_FILTFILT_CACHE = {}


def _sos_tables(b, a, L, device):
    """tf2sos sections, sosfilt_zi, and per section the zero-input response table / L-step state transition (float64)."""
    import numpy as np
    from scipy import signal

    key = (tuple(np.asarray(b).tolist()), tuple(np.asarray(a).tolist()), L, str(device))
    if key not in _FILTFILT_CACHE:
        sos = np.ascontiguousarray(signal.tf2sos(b, a), dtype=np.float64)
        zi = signal.sosfilt_zi(sos)
        nsec = sos.shape[0]
        H = np.zeros((nsec, L, 2))
        ML = np.zeros((nsec, 2, 2))
        for s_ in range(nsec):
            a1, a2 = sos[s_, 4] / sos[s_, 3], sos[s_, 5] / sos[s_, 3]
            for q in range(2):
                z = np.zeros(2)
                z[q] = 1.0
                for j in range(L):
                    y = z[0]
                    H[s_, j, q] = y
                    z = np.array([z[1] - a1 * y, -a2 * y])
                ML[s_, :, q] = z
        t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(device)
        _FILTFILT_CACHE[key] = (sos, t(zi), t(H), t(ML))
    return _FILTFILT_CACHE[key]


def filtfilt(x, b, a, L=512):
    """Zero-phase IIR of a float32 device vector -> float64 device vector (scipy filtfilt defaults: odd extension,
    padlen = 3*max(len(a), len(b)); evaluated as a second-order-section cascade)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n = x.numel()
    sos, zi, H, ML = _sos_tables(b, a, L, x.device)
    padlen = 3 * max(len(a), len(b))
    ne = n + 2 * padlen
    nblk = (ne + L - 1) // L
    work = torch.empty(2 * ne + 2 + 4 * nblk + 16, dtype=torch.float64, device=x.device)
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    _ffi.check(_ffi.lib().b200vc_sosfiltfilt_f64(_p(x), n, sos.ctypes.data_as(C.c_void_p), sos.shape[0], _p(zi), _p(H), _p(ML), L,
                                                 padlen, _p(work), _p(out), _s()), "sosfiltfilt")
    return out
