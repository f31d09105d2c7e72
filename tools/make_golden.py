#!/usr/bin/env python
"""Generate tests/golden/*.json (run in the build container, where /root/reference exists).

  reference_artifacts.json  sha256 + key/value content of the reference's header template and conf
                            (header_baseband2power.txt, paf-baseband2power.conf) — the
                            drop-in artefacts, which this repo ships byte for byte.
  bmf_hdr_vectors.json      BMF packet-header decode vectors produced by the REFERENCE's own
                            hdr.c (the one source file that compiles standalone), built by
                            oracle/Makefile into oracle/_ref/libpafhdr_ref.so.
  bswap64_vectors.json      payload words through the REFERENCE's BSWAP_64 macro (cudautil.cuh:118-125),
                            the only byte-order statement the reference makes for the GPU side.
  oracle_vectors.json       spectra of seeded synthetic blocks: sha256 of the block, exact
                            uint64 sums, float32 bit patterns.  Computed with the numpy
                            restatement and cross-checked against the C oracle here.  The
                            reference holds no vectors for this path (PARITY UNPINNED), so
                            these pin the oracle against regressions, not against the reference.
  tiny_block.json           a 576-byte block of a reduced geometry with sums from pure-Python loops.
"""
import configparser
import ctypes
import hashlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

import oracle  # noqa: E402
from oracle import b2p_oracle_np as onp  # noqa: E402


def header_kv(path):
    out = []
    for line in open(path):
        line = line.split("#")[0].strip()
        if line:
            p = line.split(None, 1)
            out.append([p[0], p[1].strip() if len(p) > 1 else ""])
    return out


def conf_dict(path):
    c = configparser.ConfigParser()
    c.read(path)
    return {s: dict(c[s]) for s in c.sections()}


def _sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def reference_artifacts():
    hdr, conf = os.path.join(REF, "header_baseband2power.txt"), os.path.join(REF, "paf-baseband2power.conf")
    return {"source": "xinpingdeng/paf-baseband2power: header_baseband2power.txt, paf-baseband2power.conf",
            # the two drop-in artefacts are shipped byte for byte
            "sha256": {"header_baseband2power.txt": _sha256(hdr), "paf-baseband2power.conf": _sha256(conf)},
            "bytes": {"header_baseband2power.txt": os.path.getsize(hdr), "paf-baseband2power.conf": os.path.getsize(conf)},
            "header_kv": header_kv(hdr),
            "conf": conf_dict(conf)}


class HdrT(ctypes.Structure):  # hdr.h:6-14
    _fields_ = [("valid", ctypes.c_int), ("idf", ctypes.c_uint64), ("sec", ctypes.c_uint64),
                ("epoch", ctypes.c_int), ("beam", ctypes.c_int), ("freq", ctypes.c_double)]


def bmf_hdr_vectors(n=64):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpafhdr_ref.so"))
    lib.hdr_keys.argtypes = [ctypes.c_char_p, ctypes.POINTER(HdrT)]
    rng = np.random.default_rng(20240517)
    vec = []
    for i in range(n):
        raw = rng.integers(0, 256, size=64, dtype=np.uint8).tobytes()
        if i == 0:
            raw = bytes(64)
        if i == 1:
            raw = bytes([0xFF] * 64)
        h = HdrT()
        lib.hdr_keys(raw, ctypes.byref(h))
        vec.append({"raw": raw.hex(), "valid": h.valid, "idf": int(h.idf), "sec": int(h.sec),
                    "epoch": h.epoch, "beam": h.beam, "freq": h.freq})
    return {"source": "reference hdr.c:10-28 (hdr_keys) compiled into oracle/_ref/libpafhdr_ref.so",
            "vectors": vec}


def bswap64_vectors(n=96):
    """The reference's only statement about sample byte order is the unused BSWAP_64 macro in
    the GPU utility header (cudautil.cuh:118-125): record what it does to payload words."""
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbswap_ref.so"))
    lib.ref_bswap_64.restype = ctypes.c_uint64
    lib.ref_bswap_64.argtypes = [ctypes.c_uint64]
    rng = np.random.default_rng(118125)
    words = [int(x) for x in rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)]
    words[:4] = [0, 2 ** 64 - 1, 0x0001000200030004, 0x8000800080008000]
    return {"source": "reference cudautil.cuh:118-125 (BSWAP_64) via oracle/ref_bswap_wrap.cpp -> oracle/_ref/libbswap_ref.so",
            "vectors": [{"word": "%016x" % w, "bswap": "%016x" % lib.ref_bswap_64(w)} for w in words]}


def oracle_vectors():
    cases = [(1, 0, 1, 0), (2, 1, 1, 0), (3, 0, 3, 12345), (4, 1, 5, 352321536), (20240517, 1, 8, 0),
             (7, 0, 2, 2 ** 40)]
    out = []
    for seed, mode, ndf, fw in cases:
        blk = onp.synth_block(ndf, seed, fw, mode)
        blk_c = oracle.synth_fill(ndf, seed, fw, mode)
        assert np.array_equal(blk, blk_c)
        sums = onp.channel_sums(blk)
        assert np.array_equal(sums, oracle.accumulate(blk_c))
        f1 = onp.finish(sums, 1.0)
        fm = onp.finish(sums, 2.0 ** -20)
        assert np.array_equal(f1.view(np.uint32), oracle.finish(sums, 1.0).view(np.uint32))
        out.append({"seed": seed, "mode": mode, "ndf": ndf, "first_word": fw,
                    "sha256": hashlib.sha256(blk.tobytes()).hexdigest(),
                    "sums": [str(int(x)) for x in sums],
                    "f32_sum_bits": [int(x) for x in f1.view(np.uint32)],
                    "f32_mean_bits": [int(x) for x in fm.view(np.uint32)]})
    return {"geometry": {"nchunk": 48, "nch_per_chunk": 7, "nsamp_df": 128, "big_endian": True},
            "cases": out}


def tiny_block():
    nchunk, nch, nsamp, ndf = 2, 3, 4, 3
    rng = np.random.default_rng(7)
    vals = rng.integers(-32768, 32768, size=ndf * nchunk * nsamp * nch * 4).tolist()
    vals[0], vals[1], vals[2], vals[3] = -32768, -32768, -32768, -32768
    raw = b"".join(struct.pack(">h", v) for v in vals)
    sums = [0] * (nchunk * nch)
    i = 0
    for idf in range(ndf):           # pure-Python loops in the layout's own order
        for c in range(nchunk):
            for t in range(nsamp):
                for ch in range(nch):
                    for _ in range(4):
                        sums[c * nch + ch] += vals[i] * vals[i]
                        i += 1
    return {"nchunk": nchunk, "nch_per_chunk": nch, "nsamp_df": nsamp, "ndf": ndf, "big_endian": True,
            "raw_hex": raw.hex(), "sums": [str(s) for s in sums]}


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    oracle.build()
    for name, fn in [("reference_artifacts", reference_artifacts), ("bmf_hdr_vectors", bmf_hdr_vectors),
                     ("bswap64_vectors", bswap64_vectors),
                     ("oracle_vectors", oracle_vectors), ("tiny_block", tiny_block)]:
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
