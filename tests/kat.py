"""Known-answer blocks with hand-computable spectra (KATs 1-7).

The reference has no vectors for this path, so these closed forms are what pin
the oracle.  Each builder returns (block_uint8, expected_uint64_sums).
Layout: block[idf][chunk][t][ch][pol][re,im] int16 (capture.c:540-542).
"""
from __future__ import annotations

import numpy as np


def _shape(ndf, nchunk, nch, nsamp):
    return (ndf, nchunk, nsamp, nch, 4)


def _pack(x: np.ndarray, big_endian=True) -> np.ndarray:
    return np.ascontiguousarray(x.astype(">i2" if big_endian else "<i2")).view(np.uint8).reshape(-1)


def kat_zero(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    x = np.zeros(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    return _pack(x, big_endian), np.zeros(nchunk * nch, dtype=np.uint64)


def kat_ones(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    """every component = 1 -> each channel = 4*ndf*nsamp (read with the wrong
    byte order each component would be 256 -> 65536x larger)."""
    x = np.ones(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    return _pack(x, big_endian), np.full(nchunk * nch, 4 * ndf * nsamp, dtype=np.uint64)


def kat_min(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    """every component = -32768 -> each word contributes exactly 2^32 (overflow/sign trap)."""
    x = np.full(_shape(ndf, nchunk, nch, nsamp), -32768, dtype=np.int16)
    return _pack(x, big_endian), np.full(nchunk * nch, (1 << 32) * ndf * nsamp, dtype=np.uint64)


def kat_channel_ramp(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    """component value = channel index + 1 -> channel k gives 4(k+1)^2*ndf*nsamp."""
    x = np.empty(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    k = (np.arange(nchunk)[:, None] * nch + np.arange(nch)[None, :] + 1).astype(np.int16)
    x[...] = k[None, :, None, :, None]
    kk = np.arange(1, nchunk * nch + 1, dtype=np.uint64)
    return _pack(x, big_endian), 4 * kk * kk * np.uint64(ndf * nsamp)


def kat_single_word(idf, chunk, t, ch, ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True,
                    value=(3, -4, 5, -6)):
    """one non-zero (t,ch) word -> exactly one non-zero channel (stride trap)."""
    x = np.zeros(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    x[idf, chunk, t, ch, :] = value
    want = np.zeros(nchunk * nch, dtype=np.uint64)
    want[chunk * nch + ch] = sum(int(v) * int(v) for v in value)
    return _pack(x, big_endian), want


def kat_time_pattern(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    """value depends on (idf,t) only: ((idf*nsamp+t) mod 251) - 125 on all four components
    -> all channels equal one closed-form sum (dropped/duplicated sample trap)."""
    n = np.arange(ndf * nsamp, dtype=np.int64)
    v = (n % 251) - 125
    x = np.empty(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    x[...] = v.reshape(ndf, 1, nsamp, 1, 1)
    total = int((4 * v * v).sum())
    return _pack(x, big_endian), np.full(nchunk * nch, total, dtype=np.uint64)


def kat_pol_pattern(ndf=2, nchunk=48, nch=7, nsamp=128, big_endian=True):
    """distinct value per component slot (Xre=1, Xim=2, Yre=3, Yim=-4) -> 30 per word."""
    x = np.empty(_shape(ndf, nchunk, nch, nsamp), dtype=np.int16)
    x[...] = np.array([1, 2, 3, -4], dtype=np.int16)
    return _pack(x, big_endian), np.full(nchunk * nch, 30 * ndf * nsamp, dtype=np.uint64)


def all_kats(ndf=4, nchunk=48, nch=7, nsamp=128, big_endian=True):
    kw = dict(ndf=ndf, nchunk=nchunk, nch=nch, nsamp=nsamp, big_endian=big_endian)
    out = {
        "zero": kat_zero(**kw),
        "ones": kat_ones(**kw),
        "min": kat_min(**kw),
        "channel_ramp": kat_channel_ramp(**kw),
        "time_pattern": kat_time_pattern(**kw),
        "pol_pattern": kat_pol_pattern(**kw),
    }
    corners = [(0, 0, 0, 0), (ndf - 1, nchunk - 1, nsamp - 1, nch - 1), (ndf // 2, nchunk // 2, 1, nch // 2),
               (0, nchunk - 1, nsamp - 1, 0), (ndf - 1, 0, 0, nch - 1)]
    for i, (a, b, c, d) in enumerate(corners):
        out[f"single_{i}"] = kat_single_word(a, b, c, d, **kw)
    return out


def hdr_kv(text):
    """DADA header text -> {key: value} (comments and padding dropped)."""
    out = {}
    for line in text.splitlines():
        line = line.split("#")[0].strip()
        if line:
            p = line.split(None, 1)
            out[p[0]] = p[1].strip() if len(p) > 1 else ""
    return out
