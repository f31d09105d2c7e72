"""Host-side mirror of the baseband2power stage over the C ABI (include/b2p.h).

`Baseband2Power` plays the role of the stage the reference declared but never
wrote — init_/do_/destroy_baseband2power behind conf_t {device_id, dir, key_in,
key_out} (baseband2power.cuh:18-23; sibling convention diskdb.cuh:32-34): it is
created for a GPU, fed ring-buffer blocks, and emits one float32 spectrum of
NCHAN values per integration (header_baseband2power.txt:39-42).

Everything here is plumbing over libb2p.so; no arithmetic happens in Python and
nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_uint64, c_void_p
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import (B2pParams, KERNEL_AUTO, KERNEL_LDG, KERNEL_TMA, MODE_EXACT, MODE_FLOAT)

_KERNELS = {"auto": KERNEL_AUTO, "ldg": KERNEL_LDG, "tma": KERNEL_TMA}
_MODES = {"exact": MODE_EXACT, "float": MODE_FLOAT}


class B2pError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b2p error {code}: {msg}")
        self.code = code


def _check(rc: int, ctx=None):
    if rc:
        msg = _lib.load().b2p_last_error(ctx)
        raise B2pError(rc, msg.decode() if msg else "")


def device_count() -> int:
    return _lib.load().b2p_device_count()


def device_info(device: int = 0) -> dict:
    lib = _lib.load()
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr, mem = c_int(), c_int(), c_int(), c_uint64()
    _check(lib.b2p_device_info(device, name, 256, byref(sm), byref(maj), byref(mnr), byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value),
            "mem_bytes": mem.value}


def device_sync(device: int = 0):
    """cudaDeviceSynchronize on `device` (the context streams are non-blocking streams)."""
    _check(_lib.load().b2p_device_sync(device))


class PinnedBuffer:
    """Pinned, device-mapped host memory (the ring block a reader would borrow)."""

    def __init__(self, nbytes: int):
        self._lib = _lib.load()
        p = c_void_p()
        _check(self._lib.b2p_host_alloc(byref(p), nbytes))
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            _check(self._lib.b2p_host_free(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation made through the C ABI (tests and bench only)."""

    def __init__(self, nbytes: int, device: int = 0):
        self._lib = _lib.load()
        p = c_void_p()
        _check(self._lib.b2p_device_alloc(device, byref(p), nbytes))
        self.ptr, self.nbytes, self.device = p.value, nbytes, device

    def upload(self, a: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(a)
        if offset + a.nbytes > self.nbytes:
            raise ValueError("upload past the end of the device buffer")
        _check(self._lib.b2p_memcpy_h2d(self.device, self.ptr + offset, a.ctypes.data, a.nbytes))

    def download(self, nbytes: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        _check(self._lib.b2p_memcpy_d2h(self.device, out.ctypes.data, self.ptr + offset, n))
        return out

    def synth_fill(self, ndf: int, seed: int, first_word: int = 0, mode: int = 1, nchunk: int = 48,
                   nch_per_chunk: int = 7, nsamp_df: int = 128, big_endian: bool = True,
                   offset: int = 0, stream: int | None = None):
        need = ndf * nchunk * nch_per_chunk * nsamp_df * 8
        if offset + need > self.nbytes:
            raise ValueError("synth_fill past the end of the device buffer")
        _check(self._lib.b2p_synth_fill_device(self.device, self.ptr + offset, ndf, nchunk,
                                               nch_per_chunk, nsamp_df, int(big_endian), seed,
                                               first_word, mode, stream))

    def free(self):
        if self.ptr:
            _check(self._lib.b2p_device_free(self.device, self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _host_ptr(x) -> int:
    if isinstance(x, int):
        return x
    if isinstance(x, PinnedBuffer):
        return x.ptr
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("host block must be C-contiguous")
        return x.ctypes.data
    raise TypeError(f"cannot take a host pointer from {type(x)}")


def _dev_ptr(x) -> int:
    if isinstance(x, int):
        return x
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if hasattr(x, "data_ptr"):  # torch tensor on the device
        return int(x.data_ptr())
    raise TypeError(f"cannot take a device pointer from {type(x)}")


class Baseband2Power:
    """One baseband->power stage instance on one GPU for `nbeam` beam streams."""

    def __init__(self, device_id: int = 0, nbeam: int = 1, nchunk: int = 48, nch_per_chunk: int = 7,
                 nsamp_df: int = 128, big_endian: bool = True, scale: float = 1.0,
                 mode: str = "exact", kernel: str = "auto", nsplit: int = 0, stage_ndf: int = 0,
                 nstage_bufs: int = 0, first_chunk: int = 0, nchunk_total: int = 0, resizable: bool = False):
        self._lib = _lib.load()
        p = B2pParams()
        self._lib.b2p_default_params(byref(p))
        p.device_id, p.nbeam = device_id, nbeam
        p.nchunk, p.nch_per_chunk, p.nsamp_df = nchunk, nch_per_chunk, nsamp_df
        p.big_endian, p.scale = int(big_endian), scale
        p.mode, p.kernel = _MODES[mode], _KERNELS[kernel]
        p.nsplit, p.stage_ndf, p.nstage_bufs = nsplit, stage_ndf, nstage_bufs
        p.first_chunk, p.nchunk_total, p.resizable = first_chunk, nchunk_total, int(resizable)
        ctx = c_void_p()
        _check(self._lib.b2p_create(byref(ctx), byref(p)))
        self._ctx = ctx
        self.nbeam = nbeam
        self.nchan = self._lib.b2p_nchan(ctx)
        self.frame_bytes = self._lib.b2p_frame_bytes(ctx)
        self.source_frame_bytes = self._lib.b2p_source_frame_bytes(ctx)
        self.mode = mode

    # -- properties ---------------------------------------------------------
    @property
    def kernel(self) -> str:
        k = self._lib.b2p_kernel_in_use(self._ctx)
        return {KERNEL_LDG: "ldg", KERNEL_TMA: "tma"}.get(k, "?")

    @property
    def nsplit(self) -> int:
        return self._lib.b2p_nsplit_in_use(self._ctx)

    @property
    def launch_count(self) -> int:
        return self._lib.b2p_launch_count(self._ctx)

    @property
    def stream(self) -> int:
        return self._lib.b2p_stream(self._ctx)

    # -- the hot path ---------------------------------------------------------
    def _ptr_array(self, items: Iterable, conv) -> ctypes.Array:
        ptrs = [conv(x) for x in items]
        if len(ptrs) != self.nbeam:
            raise ValueError(f"expected {self.nbeam} beam pointers, got {len(ptrs)}")
        return (c_void_p * self.nbeam)(*ptrs)

    def accumulate_device(self, dptrs: Sequence, ndf: int, stream: int | None = None):
        arr = self._ptr_array(dptrs, _dev_ptr)
        _check(self._lib.b2p_accumulate_device(self._ctx, arr, ndf, stream), self._ctx)

    def _host_ndf(self, blocks: Sequence, ndf: int | None) -> int:
        if ndf is not None:
            return ndf
        sizes = set()
        for b in blocks:
            n = b.nbytes if hasattr(b, "nbytes") else None
            if n is None:
                raise ValueError("ndf is required when passing raw pointers")
            if n % self.source_frame_bytes:
                raise ValueError("host block is not a whole number of data frames")
            sizes.add(n // self.source_frame_bytes)
        if len(sizes) != 1:
            raise ValueError("all beams must supply the same number of frames")
        return sizes.pop()

    def accumulate_host(self, blocks: Sequence, ndf: int | None = None):
        ndf = self._host_ndf(blocks, ndf)
        arr = self._ptr_array(blocks, _host_ptr)
        _check(self._lib.b2p_accumulate_host(self._ctx, arr, ndf), self._ctx)

    def integrate_device(self, dptrs: Sequence, ndf: int, out_dev, stream: int | None = None):
        """accumulate_device + finish_device in one kernel launch."""
        arr = self._ptr_array(dptrs, _dev_ptr)
        _check(self._lib.b2p_integrate_device(self._ctx, arr, ndf, _dev_ptr(out_dev), stream), self._ctx)

    def integrate_host(self, blocks: Sequence, ndf: int | None = None) -> np.ndarray:
        """accumulate_host + finish: the last staging piece's kernel closes the integration."""
        ndf = self._host_ndf(blocks, ndf)
        arr = self._ptr_array(blocks, _host_ptr)
        out = np.empty((self.nbeam, self.nchan), dtype=np.float32)
        _check(self._lib.b2p_integrate_host(self._ctx, arr, ndf, out.ctypes.data), self._ctx)
        return out

    def accumulate_host_async(self, blocks: Sequence, ndf: int | None = None, finish: bool = False):
        ndf = self._host_ndf(blocks, ndf)
        arr = self._ptr_array(blocks, _host_ptr)
        _check(self._lib.b2p_accumulate_host_async(self._ctx, arr, ndf, int(finish)), self._ctx)

    def wait_input(self):
        _check(self._lib.b2p_wait_input(self._ctx), self._ctx)

    def wait_output(self) -> np.ndarray:
        out = np.empty((self.nbeam, self.nchan), dtype=np.float32)
        _check(self._lib.b2p_wait_output(self._ctx, out.ctypes.data), self._ctx)
        return out

    def last_h2d_ms(self) -> float:
        ms = c_double()
        _check(self._lib.b2p_last_h2d_ms(self._ctx, byref(ms)), self._ctx)
        return ms.value

    def accumulate_host_mapped(self, blocks: Sequence, ndf: int | None = None):
        ndf = self._host_ndf(blocks, ndf)
        arr = self._ptr_array(blocks, _host_ptr)
        _check(self._lib.b2p_accumulate_host_mapped(self._ctx, arr, ndf), self._ctx)

    def finish(self) -> np.ndarray:
        out = np.empty((self.nbeam, self.nchan), dtype=np.float32)
        _check(self._lib.b2p_finish(self._ctx, out.ctypes.data), self._ctx)
        return out

    def finish_device(self, out_dev, stream: int | None = None):
        _check(self._lib.b2p_finish_device(self._ctx, _dev_ptr(out_dev), stream), self._ctx)

    def set_chunk_range(self, first_chunk: int, nchunk: int):
        """Move a resizable shard context to another chunk range (between integrations)."""
        _check(self._lib.b2p_set_chunk_range(self._ctx, first_chunk, nchunk), self._ctx)
        self.nchan = self._lib.b2p_nchan(self._ctx)
        self.frame_bytes = self._lib.b2p_frame_bytes(self._ctx)

    def read_sums(self) -> np.ndarray:
        out = np.empty((self.nbeam, self.nchan), dtype=np.uint64)
        _check(self._lib.b2p_read_sums(self._ctx, out.ctypes.data), self._ctx)
        return out

    def reset(self):
        _check(self._lib.b2p_reset(self._ctx), self._ctx)

    # -- timing ---------------------------------------------------------------
    def set_timing(self, enabled: bool):
        _check(self._lib.b2p_set_timing(self._ctx, int(enabled)), self._ctx)

    def fused_time_ms(self) -> tuple[float, int]:
        ms, n = c_double(), c_uint64()
        _check(self._lib.b2p_fused_time_ms(self._ctx, byref(ms), byref(n)), self._ctx)
        return ms.value, n.value

    # -- lifetime -------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.b2p_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def probe_h2d(devices: Sequence[int], nbytes: int = 256 << 20, reps: int = 4) -> list[float]:
    """Pinned host->device GB/s of every listed GPU, all copying at once."""
    n = len(devices)
    out = (c_double * n)()
    _check(_lib.load().b2p_probe_h2d((c_int * n)(*devices), n, nbytes, reps, out))
    return list(out)


def split_chunks(weights: Sequence[float] | None, n: int, nchunk: int = 48) -> list[int]:
    """nchunk chunks over n GPUs in proportion to `weights` (largest remainder)."""
    counts = (c_int * n)()
    w = (c_double * n)(*weights) if weights is not None else None
    _check(_lib.load().b2p_split_chunks(w, n, nchunk, counts))
    return list(counts)


class ShardGroup:
    """Beam streams spread over several GPUs by channel group (include/b2p.h b2p_group_*):
    GPU devices[i] covers nchunks[i] consecutive chunks of every data frame."""

    def __init__(self, devices: Sequence[int], nchunks: Sequence[int], nbeam: int = 1, nchunk: int = 48,
                 nch_per_chunk: int = 7, nsamp_df: int = 128, big_endian: bool = True,
                 scale: float = 1.0, mode: str = "exact", kernel: str = "auto", nstage_bufs: int = 0):
        self._lib = _lib.load()
        p = B2pParams()
        self._lib.b2p_default_params(byref(p))
        p.nbeam, p.nchunk, p.nch_per_chunk, p.nsamp_df = nbeam, nchunk, nch_per_chunk, nsamp_df
        p.big_endian, p.scale = int(big_endian), scale
        p.mode, p.kernel, p.nstage_bufs = _MODES[mode], _KERNELS[kernel], nstage_bufs
        n = len(devices)
        if len(nchunks) != n:
            raise ValueError("one chunk count per device")
        g = c_void_p()
        rc = self._lib.b2p_group_create(byref(g), byref(p), (c_int * n)(*devices), (c_int * n)(*nchunks), n)
        _check(rc)
        self._g = g
        self.nbeam, self.nchan = nbeam, nchunk * nch_per_chunk
        self.source_frame_bytes = nchunk * nch_per_chunk * nsamp_df * 8

    def _check(self, rc):
        if rc:
            msg = self._lib.b2p_group_last_error(self._g)
            raise B2pError(rc, msg.decode() if msg else "")

    @property
    def shards(self) -> list[tuple[int, int, int]]:
        """(device, first_chunk, nchunk) of every shard that holds at least one chunk."""
        out = []
        for i in range(self._lib.b2p_group_size(self._g)):
            d, f, n = c_int(), c_int(), c_int()
            self._lib.b2p_group_shard(self._g, i, byref(d), byref(f), byref(n))
            out.append((d.value, f.value, n.value))
        return out

    @property
    def launch_count(self) -> int:
        return sum(self._lib.b2p_launch_count(self._lib.b2p_group_ctx(self._g, i))
                   for i in range(self._lib.b2p_group_size(self._g)))

    def _ptrs(self, blocks):
        ptrs = [_host_ptr(x) for x in blocks]
        if len(ptrs) != self.nbeam:
            raise ValueError(f"expected {self.nbeam} beam pointers, got {len(ptrs)}")
        return (c_void_p * self.nbeam)(*ptrs)

    def accumulate_host(self, blocks: Sequence, ndf: int):
        self._check(self._lib.b2p_group_accumulate_host(self._g, self._ptrs(blocks), ndf))

    def integrate_host(self, blocks: Sequence, ndf: int) -> np.ndarray:
        out = np.empty((self.nbeam, self.nchan), dtype=np.float32)
        self._check(self._lib.b2p_group_integrate_host(self._g, self._ptrs(blocks), ndf, out.ctypes.data))
        return out

    def finish(self) -> np.ndarray:
        out = np.empty((self.nbeam, self.nchan), dtype=np.float32)
        self._check(self._lib.b2p_group_finish(self._g, out.ctypes.data))
        return out

    def reset(self):
        self._check(self._lib.b2p_group_reset(self._g))

    def rebalance(self) -> bool:
        """Between integrations: shift chunks towards the faster links (measured); True if moved."""
        ch = c_int()
        self._check(self._lib.b2p_group_rebalance(self._g, byref(ch)))
        return bool(ch.value)

    def close(self):
        if getattr(self, "_g", None):
            self._lib.b2p_group_destroy(self._g)
            self._g = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def selftest_unpack(device: int = 0, big_endian: bool = True) -> np.ndarray:
    out = np.empty(65536, dtype=np.int32)
    _check(_lib.load().b2p_selftest_unpack(device, int(big_endian),
                                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
    return out
