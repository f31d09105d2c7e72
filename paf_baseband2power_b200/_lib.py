"""ctypes loader for libb2p.so — the C ABI of include/b2p.h.

This is the reference-side binding a Python maintainer would write (the
reference's own launcher is Python, paf-baseband2power.py).  It fails loudly:
there is no CPU fallback anywhere in the product path, so a missing library is
an ImportError-grade condition, never a silent detour.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_size_t,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2p.so")

B2P_OK, B2P_EINVAL, B2P_ECUDA, B2P_ENOMEM, B2P_ESTATE = 0, 1, 2, 3, 4
MODE_EXACT, MODE_FLOAT = 0, 1
KERNEL_AUTO, KERNEL_LDG, KERNEL_TMA = 0, 1, 2
MAX_BEAMS = 64
MAX_GROUP = 16


class B2pParams(Structure):
    """struct b2p_params (include/b2p.h)."""
    _fields_ = [
        ("device_id", c_int), ("nchunk", c_int), ("nch_per_chunk", c_int), ("nsamp_df", c_int),
        ("big_endian", c_int), ("scale", c_float), ("mode", c_int), ("nbeam", c_int),
        ("kernel", c_int), ("nsplit", c_int), ("stage_ndf", c_uint64), ("nstage_bufs", c_int),
        ("first_chunk", c_int), ("nchunk_total", c_int), ("resizable", c_int),
    ]


# every symbol include/b2p.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b2p_default_params": (None, [POINTER(B2pParams)]),
    "b2p_create": (c_int, [POINTER(c_void_p), POINTER(B2pParams)]),
    "b2p_destroy": (None, [c_void_p]),
    "b2p_last_error": (c_char_p, [c_void_p]),
    "b2p_accumulate_device": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_void_p]),
    "b2p_integrate_device": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_void_p, c_void_p]),
    "b2p_accumulate_host": (c_int, [c_void_p, POINTER(c_void_p), c_uint64]),
    "b2p_integrate_host": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_void_p]),
    "b2p_accumulate_host_async": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_int]),
    "b2p_wait_input": (c_int, [c_void_p]),
    "b2p_wait_output": (c_int, [c_void_p, c_void_p]),
    "b2p_last_h2d_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "b2p_accumulate_host_mapped": (c_int, [c_void_p, POINTER(c_void_p), c_uint64]),
    "b2p_finish": (c_int, [c_void_p, c_void_p]),
    "b2p_finish_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b2p_set_chunk_range": (c_int, [c_void_p, c_int, c_int]),
    "b2p_read_sums": (c_int, [c_void_p, c_void_p]),
    "b2p_reset": (c_int, [c_void_p]),
    "b2p_nchan": (c_int, [c_void_p]),
    "b2p_frame_bytes": (c_uint64, [c_void_p]),
    "b2p_source_frame_bytes": (c_uint64, [c_void_p]),
    "b2p_first_chunk": (c_int, [c_void_p]),
    "b2p_kernel_in_use": (c_int, [c_void_p]),
    "b2p_nsplit_in_use": (c_int, [c_void_p]),
    "b2p_launch_count": (c_uint64, [c_void_p]),
    "b2p_stream": (c_void_p, [c_void_p]),
    "b2p_version": (c_char_p, []),
    "b2p_device_count": (c_int, []),
    "b2p_device_info": (c_int, [c_int, c_char_p, c_size_t, POINTER(c_int), POINTER(c_int),
                                POINTER(c_int), POINTER(c_uint64)]),
    "b2p_set_timing": (c_int, [c_void_p, c_int]),
    "b2p_fused_time_ms": (c_int, [c_void_p, POINTER(c_double), POINTER(c_uint64)]),
    "b2p_group_create": (c_int, [POINTER(c_void_p), POINTER(B2pParams), POINTER(c_int), POINTER(c_int), c_int]),
    "b2p_group_destroy": (None, [c_void_p]),
    "b2p_group_accumulate_host": (c_int, [c_void_p, POINTER(c_void_p), c_uint64]),
    "b2p_group_integrate_host": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_void_p]),
    "b2p_group_finish": (c_int, [c_void_p, c_void_p]),
    "b2p_group_issue_host": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_int]),
    "b2p_group_wait_input": (c_int, [c_void_p]),
    "b2p_group_wait_output": (c_int, [c_void_p, c_void_p]),
    "b2p_group_reset": (c_int, [c_void_p]),
    "b2p_group_rebalance": (c_int, [c_void_p, POINTER(c_int)]),
    "b2p_group_size": (c_int, [c_void_p]),
    "b2p_group_ctx": (c_void_p, [c_void_p, c_int]),
    "b2p_group_shard": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "b2p_group_last_error": (c_char_p, [c_void_p]),
    "b2p_probe_h2d": (c_int, [POINTER(c_int), c_int, c_size_t, c_int, POINTER(c_double)]),
    "b2p_split_chunks": (c_int, [POINTER(c_double), c_int, c_int, POINTER(c_int)]),
    "b2p_host_alloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "b2p_host_free": (c_int, [c_void_p]),
    "b2p_host_register": (c_int, [c_void_p, c_size_t]),
    "b2p_host_unregister": (c_int, [c_void_p]),
    "b2p_device_alloc": (c_int, [c_int, POINTER(c_void_p), c_size_t]),
    "b2p_device_free": (c_int, [c_int, c_void_p]),
    "b2p_memcpy_h2d": (c_int, [c_int, c_void_p, c_void_p, c_size_t]),
    "b2p_memcpy_d2h": (c_int, [c_int, c_void_p, c_void_p, c_size_t]),
    "b2p_device_sync": (c_int, [c_int]),
    "b2p_synth_fill_device": (c_int, [c_int, c_void_p, c_uint64, c_int, c_int, c_int, c_int,
                                      c_uint64, c_uint64, c_int, c_void_p]),
    "b2p_selftest_unpack": (c_int, [c_int, c_int, POINTER(c_int32)]),
}

_LIB = None


class B2pLibraryMissing(ImportError):
    pass


def load() -> ctypes.CDLL:
    """Load libb2p.so and bind every declared symbol; raise if anything is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise B2pLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C paf_baseband2power_b200/csrc`. There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
