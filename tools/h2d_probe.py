"""Development probe: host->device strategies for one 2.8 GB ring block (GPU box)."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paf_baseband2power_b200 import BMF, Baseband2Power, PinnedBuffer  # noqa: E402

blk = BMF.block_bytes
pb = PinnedBuffer(blk)
pb.array[:] = 1
host = torch.from_numpy(pb.array)
dev = torch.empty(blk, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return blk * reps / (time.perf_counter() - t0) / 1e9


print("single memcpy            %.2f GB/s" % timeit(lambda: dev.copy_(host, non_blocking=True)))
for nstream in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    part = blk // nstream

    def multi():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[i * part:(i + 1) * part].copy_(host[i * part:(i + 1) * part], non_blocking=True)
    print("%d concurrent streams     %.2f GB/s" % (nstream, timeit(multi)))
for piece in (64, 256, 1024, 4096):
    st = Baseband2Power(stage_ndf=piece, nstage_bufs=3)
    def staged():
        st.accumulate_host([pb], 8192)
        st.finish()
    print("staged path, %4d-frame pieces  %.2f GB/s" % (piece, timeit(staged)))
    st.close()
st = Baseband2Power()
def mapped():
    st.accumulate_host_mapped([pb], 8192)
    st.finish()
print("zero-copy mapped kernel  %.2f GB/s" % timeit(mapped, reps=3))
st.close()
