"""paf_baseband2power_b200 — B200-native baseband -> power stage (hot path only).

Drop-in for the unpack / detect / integrate path that paf-baseband2power
declared (baseband2power.cuh:18-23, paf_baseband2power.cu:17-28) and never
implemented (kernel.cu:1-7).  The product is libb2p.so (hand-written sm_100a
kernels behind the C ABI of include/b2p.h) plus the host executables; this
package is the thin ctypes mirror of that ABI.
"""
from .geometry import Geometry, BMF  # noqa: F401

__all__ = ["Geometry", "BMF", "Baseband2Power", "ShardGroup", "PinnedBuffer", "DeviceBuffer"]


def __getattr__(name):
    # the ABI binding is imported lazily so that `import paf_baseband2power_b200`
    # works for host-only helpers; using the stage without libb2p.so raises.
    if name in ("Baseband2Power", "PinnedBuffer", "DeviceBuffer", "B2pError", "device_count",
                "device_info", "device_sync", "selftest_unpack", "ShardGroup", "probe_h2d",
                "split_chunks"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
