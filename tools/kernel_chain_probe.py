#!/usr/bin/env python
"""Diagnostic: per-step time of back-to-back integrations on the context stream for each
kernel, (a) one launch per integration, (b) accumulate + separate finish launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paf_baseband2power_b200 import BMF, Baseband2Power, _lib  # noqa: E402

lib = _lib.load()
blk = BMF.block_bytes
nrot = 4
dev = torch.empty(nrot * blk, dtype=torch.uint8, device="cuda")
for r in range(nrot):
    lib.b2p_synth_fill_device(0, dev.data_ptr() + r * blk, 8192, 48, 7, 128, 1, 7 + r, 0, 1, None)
out = torch.empty(336, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for kernel in ("ldg", "tma"):
    st = Baseband2Power(kernel=kernel)
    xs = torch.cuda.ExternalStream(st.stream)
    for mode in ("one_launch", "two_launches"):
        def step(i):
            p = [dev.data_ptr() + (i % nrot) * blk]
            if mode == "one_launch":
                st.integrate_device(p, 8192, out)
            else:
                st.accumulate_device(p, 8192)
                st.finish_device(out)
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(xs)
            for i in range(64):
                step(i)
            e1.record(xs)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 64)
        print(f"{kernel} {mode} early={os.environ.get('B2P_NO_EARLY','0')=='0'} variant={os.environ.get('B2P_VARIANT','0')}: {best:.4f} ms/step {blk / best / 1e6:.0f} GB/s")
    st.close()
