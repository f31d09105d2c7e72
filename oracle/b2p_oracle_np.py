"""Independent numpy restatement of the baseband->power specification.

TEST INFRASTRUCTURE ONLY — see oracle/b2p_oracle.c for the rules (who may
import this) and for the PARITY UNPINNED statement: the reference has no
implementation of this path (kernel.cu:1-7, baseband2power.cu:1-16) and no
golden vectors, so this file follows the same specification as the C oracle
but is written independently (array reshapes instead of pointer loops) so the
two can cross-check each other.

Spec sources: layout capture.c:540-542 / sync.c:157; geometry
paf-baseband2power.conf:2-5,9,24-25 and capture.h:28; output format
header_baseband2power.txt:39-42; semantics README.md:2 and
paf_baseband2power.cu:20.
"""
from __future__ import annotations

import numpy as np

NCHUNK = 48          # NCHK_NIC, paf-baseband2power.conf:5
NCH_PER_CHUNK = 7    # NCHAN 336 / 48, paf-baseband2power.conf:24
NSAMP_DF = 128       # paf-baseband2power.conf:2
NDF_BLOCK = 8192     # paf-baseband2power.conf:9
PKT_BYTES = 7168     # capture.h:28


def frame_bytes(nchunk=NCHUNK, nch=NCH_PER_CHUNK, nsamp=NSAMP_DF) -> int:
    return nchunk * nsamp * nch * 8


def channel_sums(block, ndf=None, nchunk=NCHUNK, nch=NCH_PER_CHUNK, nsamp=NSAMP_DF,
                 big_endian=True) -> np.ndarray:
    """Exact per-channel sum of Xre^2+Xim^2+Yre^2+Yim^2 over all (idf, t).

    `block` is a bytes-like / uint8 array in [idf][chunk][t][ch][pol][re,im]
    order.  Returns uint64[nchunk*nch], channel = chunk*nch + ch.
    """
    raw = np.frombuffer(block, dtype=np.uint8) if not isinstance(block, np.ndarray) else block
    raw = raw.view(np.uint8).reshape(-1)
    fb = frame_bytes(nchunk, nch, nsamp)
    if ndf is None:
        if raw.size % fb:
            raise ValueError("block is not a whole number of data frames")
        ndf = raw.size // fb
    raw = raw[: ndf * fb]
    dt = np.dtype(">i2") if big_endian else np.dtype("<i2")
    x = raw.view(dt).reshape(ndf, nchunk, nsamp, nch, 4).astype(np.int64)
    p = (x * x).sum(axis=4)                 # (idf, chunk, t, ch), each <= 2^32
    s = p.sum(axis=(0, 2))                  # (chunk, ch)
    return s.reshape(nchunk * nch).astype(np.uint64)


def finish(sums: np.ndarray, scale: float = 1.0) -> np.ndarray:
    """float32 output block: one round-to-nearest conversion, then an fp32 multiply."""
    return (sums.astype(np.float32) * np.float32(scale)).astype(np.float32)


# ---------------------------------------------------------------------------
# Synthetic stream, restated from include/b2p_synth.h with numpy uint64 math.
# ---------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _ih4(u: np.ndarray) -> np.ndarray:
    u = u.astype(np.int64)
    return (u & 0xFF) + ((u >> 8) & 0xFF) + ((u >> 16) & 0xFF) + ((u >> 24) & 0xFF) - 510


def synth_block(ndf, seed, first_word=0, mode=1, nchunk=NCHUNK, nch=NCH_PER_CHUNK,
                nsamp=NSAMP_DF, big_endian=True) -> np.ndarray:
    """uint8 block of `ndf` frames of the counter-based synthetic stream."""
    wpf = nchunk * nsamp * nch
    nchan = nchunk * nch
    w = np.arange(ndf * wpf, dtype=np.uint64)
    wpp = np.uint64(nsamp * nch)
    chan = ((w // wpp) % np.uint64(nchunk)).astype(np.int64) * nch + ((w % wpp) % np.uint64(nch)).astype(np.int64)
    with np.errstate(over="ignore"):
        h1 = _mix64(np.uint64(seed) ^ ((w + np.uint64(first_word)) * np.uint64(0xD1342543DE82EF95)))
    v = np.empty((w.size, 4), dtype=np.int16)
    if mode == 0:
        for k in range(4):
            v[:, k] = ((h1 >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
    else:
        h2 = _mix64(h1 ^ np.uint64(0xA0761D6478BD642F))
        g = 222 + (222 * chan) // nchan
        lanes = [h1 & np.uint64(0xFFFFFFFF), h1 >> np.uint64(32),
                 h2 & np.uint64(0xFFFFFFFF), h2 >> np.uint64(32)]
        for k, lane in enumerate(lanes):
            prod = _ih4(lane) * g
            # C integer division truncates toward zero
            q = np.where(prod >= 0, prod // 64, -((-prod) // 64))
            v[:, k] = q.astype(np.int16)
    out = v.astype(">i2" if big_endian else "<i2")
    return out.view(np.uint8).reshape(-1)
