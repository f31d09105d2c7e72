"""ctypes front end of oracle/libb2p_oracle.so (TEST INFRASTRUCTURE ONLY).

See oracle/b2p_oracle.c for what the oracle restates (reference file:line) and
for the PARITY UNPINNED statement.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Geom(ctypes.Structure):
    _fields_ = [("nchunk", ctypes.c_int), ("nch_per_chunk", ctypes.c_int),
                ("nsamp_df", ctypes.c_int), ("big_endian", ctypes.c_int)]


@dataclass(frozen=True)
class Geometry:
    """Block geometry; defaults are paf-baseband2power.conf:2-5,24 and capture.h:28."""
    nchunk: int = 48
    nch_per_chunk: int = 7
    nsamp_df: int = 128
    big_endian: bool = True

    @property
    def nchan(self) -> int:
        return self.nchunk * self.nch_per_chunk

    @property
    def frame_bytes(self) -> int:
        return self.nchunk * self.nsamp_df * self.nch_per_chunk * 8

    def c(self) -> _Geom:
        return _Geom(self.nchunk, self.nch_per_chunk, self.nsamp_df, int(self.big_endian))


def build(native: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if native:
        out = os.path.join(_HERE, "_native")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libb2p_oracle.so")
        subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-std=gnu11", "-shared",
                        "-o", so, os.path.join(_HERE, "b2p_oracle.c")], check=True)
        return so
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return os.path.join(_HERE, "libb2p_oracle.so")


def lib(path: str | None = None) -> ctypes.CDLL:
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    so = path or os.path.join(_HERE, "libb2p_oracle.so")
    if not os.path.exists(so):
        build()
    L = ctypes.CDLL(so)
    u8p, u64p = ctypes.c_void_p, ctypes.c_void_p
    gp = ctypes.POINTER(_Geom)
    L.b2p_oracle_accumulate.argtypes = [u8p, ctypes.c_uint64, gp, u64p]
    L.b2p_oracle_accumulate.restype = ctypes.c_int
    L.b2p_oracle_accumulate_omp.argtypes = [u8p, ctypes.c_uint64, gp, u64p, ctypes.c_int]
    L.b2p_oracle_accumulate_omp.restype = ctypes.c_int
    L.b2p_oracle_accumulate_f64.argtypes = [u8p, ctypes.c_uint64, gp, ctypes.c_void_p]
    L.b2p_oracle_accumulate_f64.restype = ctypes.c_int
    L.b2p_oracle_accumulate_f32_naive.argtypes = [u8p, ctypes.c_uint64, gp, ctypes.c_void_p]
    L.b2p_oracle_accumulate_f32_naive.restype = ctypes.c_int
    L.b2p_oracle_finish.argtypes = [u64p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    L.b2p_oracle_finish.restype = None
    L.b2p_oracle_synth_fill.argtypes = [u8p, ctypes.c_uint64, gp, ctypes.c_uint64, ctypes.c_uint64,
                                        ctypes.c_int]
    L.b2p_oracle_synth_fill.restype = ctypes.c_int
    L.b2p_oracle_max_threads.restype = ctypes.c_int
    if path is None:
        _LIB = L
    return L


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _check_block(block: np.ndarray, ndf: int | None, g: Geometry) -> tuple[np.ndarray, int]:
    b = np.ascontiguousarray(block).view(np.uint8).reshape(-1)
    if ndf is None:
        if b.size % g.frame_bytes:
            raise ValueError("block is not a whole number of data frames")
        ndf = b.size // g.frame_bytes
    if b.size < ndf * g.frame_bytes:
        raise ValueError("block shorter than ndf frames")
    return b, ndf


def accumulate(block, ndf=None, g: Geometry = Geometry(), sums: np.ndarray | None = None,
               L: ctypes.CDLL | None = None) -> np.ndarray:
    b, ndf = _check_block(block, ndf, g)
    if sums is None:
        sums = np.zeros(g.nchan, dtype=np.uint64)
    gc = g.c()
    rc = (L or lib()).b2p_oracle_accumulate(_ptr(b), ndf, ctypes.byref(gc), _ptr(sums))
    if rc:
        raise RuntimeError(f"b2p_oracle_accumulate failed rc={rc}")
    return sums


def accumulate_omp(block, ndf=None, g: Geometry = Geometry(), sums: np.ndarray | None = None,
                   nthreads: int = 0, L: ctypes.CDLL | None = None) -> np.ndarray:
    b, ndf = _check_block(block, ndf, g)
    if sums is None:
        sums = np.zeros(g.nchan, dtype=np.uint64)
    gc = g.c()
    rc = (L or lib()).b2p_oracle_accumulate_omp(_ptr(b), ndf, ctypes.byref(gc), _ptr(sums), nthreads)
    if rc:
        raise RuntimeError(f"b2p_oracle_accumulate_omp failed rc={rc}")
    return sums


def accumulate_f64(block, ndf=None, g: Geometry = Geometry()) -> np.ndarray:
    b, ndf = _check_block(block, ndf, g)
    sums = np.zeros(g.nchan, dtype=np.float64)
    gc = g.c()
    rc = lib().b2p_oracle_accumulate_f64(_ptr(b), ndf, ctypes.byref(gc), _ptr(sums))
    if rc:
        raise RuntimeError(f"b2p_oracle_accumulate_f64 failed rc={rc}")
    return sums


def accumulate_f32_naive(block, ndf=None, g: Geometry = Geometry()) -> np.ndarray:
    b, ndf = _check_block(block, ndf, g)
    sums = np.zeros(g.nchan, dtype=np.float32)
    gc = g.c()
    rc = lib().b2p_oracle_accumulate_f32_naive(_ptr(b), ndf, ctypes.byref(gc), _ptr(sums))
    if rc:
        raise RuntimeError(f"b2p_oracle_accumulate_f32_naive failed rc={rc}")
    return sums


def finish(sums: np.ndarray, scale: float = 1.0) -> np.ndarray:
    s = np.ascontiguousarray(sums, dtype=np.uint64)
    out = np.empty(s.size, dtype=np.float32)
    lib().b2p_oracle_finish(_ptr(s), s.size, ctypes.c_float(scale), _ptr(out))
    return out


def synth_fill(ndf: int, seed: int, first_word: int = 0, mode: int = 1, g: Geometry = Geometry(),
               out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty(ndf * g.frame_bytes, dtype=np.uint8)
    gc = g.c()
    rc = lib().b2p_oracle_synth_fill(_ptr(out), ndf, ctypes.byref(gc), seed, first_word, mode)
    if rc:
        raise RuntimeError(f"b2p_oracle_synth_fill failed rc={rc}")
    return out


def max_threads() -> int:
    return lib().b2p_oracle_max_threads()
