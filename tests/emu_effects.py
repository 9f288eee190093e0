"""numpy emulation of the kernels in aicovergen_b200/csrc/effects.cu — the SAME restructuring of the recursions (chunks with a
warm-up, K-term Horner form of the comb's damping low-pass walked in blocks of the delay, all-pass as a finite sum, closed
form of audioop.ratecv), in float32 / int64, so that the restructuring itself is checked on the CPU against the sequential
oracle (oracle/effects.c, oracle/mixdown.py); the GPU tests then check the kernels."""
import math

import numpy as np

f32 = np.float32


def hpf_comp(x16, k, chunk, warm):
    n = x16.size
    x = x16.astype(f32) * f32(1.0 / 32768.0)
    starts = np.arange(0, n, chunk)
    y = np.zeros(n, f32)
    lv1 = np.zeros(starts.size, f32)
    yold = np.zeros(starts.size, f32)
    b0, b1, a1, cat, crl = f32(k.b0), f32(k.b1), f32(k.a1), f32(k.cte_at), f32(k.cte_rl)
    thr, thr_inv, expo = f32(k.thr), f32(k.thr_inv), f32(k.expo)
    for step in range(-warm, chunk):                      # all chunks in lockstep
        t = starts + step
        live = (t >= 0) & (t < n) & (t < starts + chunk)
        xin = np.where(live, x[np.clip(t, 0, n - 1)], f32(0))
        out = xin * b0 + lv1
        nlv = xin * b1 - out * a1
        a = np.abs(out)
        cte = np.where(a > yold, cat, crl)
        env = a + cte * (yold - a)
        lv1 = np.where(live, nlv, lv1)
        yold = np.where(live, env, yold)
        if step >= 0:
            with np.errstate(all="ignore"):
                g = np.where(env < thr, f32(1), np.power(env * thr_inv, expo, dtype=f32))
            w = live
            y[t[w]] = (g * out)[w]
    return y


def combs(x, k):
    n = x.size
    damp, omd, fb, gain = f32(k.damp), f32(1.0) - f32(k.damp), f32(k.feedback), f32(k.gain)
    total = np.zeros(n, f32)
    for D in k.comb_delays:
        Y = np.zeros(n + D + k.comb_terms, f32)                # Y[off + t]; everything before t = 0 is zero
        off = D + k.comb_terms
        for s in range(0, n, D):
            t = np.arange(s, min(s + D, n))
            last = np.zeros(t.size, f32)
            for kk in range(k.comb_terms - 1, -1, -1):
                last = Y[off + t - D - kk] * omd + last * damp
            Y[off + t] = x[t] * gain + last * fb
        delayed = np.zeros(n, f32)
        delayed[D:] = Y[off: off + n - D]
        total = total + delayed
    return total


def allpass(x, D, M=32):
    n = x.size
    pad = np.concatenate([np.zeros((M + 1) * D, f32), x])
    off = (M + 1) * D
    t = np.arange(n)
    acc = np.zeros(n, f32)
    for m in range(M - 1, -1, -1):
        acc = pad[off + t - D - m * D] + acc * f32(0.5)
    return acc - x


def finish(rev, x, k):
    v = rev * f32(k.wet1) + x * f32(k.dry)
    d = v.astype(np.float64)
    i32 = np.where(d <= -1.0, -2 ** 31, np.where(d >= 1.0, 2 ** 31 - 1, np.rint(2147483647.0 * np.clip(d, -1, 1)))).astype(np.int64)
    return (i32 >> 16).astype(np.int16), v


def effects(x16, k, chunk, warm):
    c = hpf_comp(x16, k, chunk, warm)
    r = combs(c, k)
    for D in k.allpass_delays:
        r = allpass(r, D)
    return finish(r, c, k) + (c,)


# ---- pydub mix
def _mul16(v, f):
    r = v.astype(np.float64) * f
    return np.where(r > 32767.0, 32767.0, np.where(r < -32767.0, -32768.0, np.floor(r))).astype(np.int64)


def _src(x, ch_out, inr, outr, g1, g2, j):
    """samples [len(j), ch_out] of the gained, rate-converted operand at output frames j (closed form of audioop.ratecv)."""
    x2 = x[:, None] if x.ndim == 1 else x
    cols = [min(c, x2.shape[1] - 1) for c in range(ch_out)]
    g = _mul16(_mul16(x2[:, cols].astype(np.int64), g1), g2)
    if inr == outr:
        return g[j]
    d0 = math.gcd(inr, outr)
    inr, outr = inr // d0, outr // d0
    kk = -((-j * inr) // outr)
    d = (kk * outr - j * inr)[:, None]
    cur = g[kk]
    prev = np.where((kk > 0)[:, None], g[np.maximum(kk - 1, 0)], 0)
    N = 65536 * (prev * d + cur * (outr - d))
    q = np.where(N >= 0, N // outr, -((-N) // outr))
    return q >> 16


def pydub_mix(xs, rates, gains, rate, used, n_out):
    ch = max(1 if x.ndim == 1 else x.shape[1] for x in xs)
    out = np.zeros((n_out, ch), np.int64)
    for i in range(3):
        j = np.arange(used[i], dtype=np.int64)
        s = _src(xs[i], ch, rates[i], rate, gains[i][0], gains[i][1], j)
        out[:used[i]] = np.clip(out[:used[i]] + s, -32768, 32767) if i else s
    return out.astype(np.int16)
