"""Development probe (GPU box, one GPU): what does QUEUEING a channel-group shard's copies cost the
host thread?  A group stage drives all GPUs of a beam from one process; if the strided
(2-D) H2D calls are expensive to issue, one thread cannot keep eight links busy.

For shard widths 48 (whole frames, 1-D copies), 24, 12, 6, 5 chunks it prints the host time spent
inside b2p_accumulate_host_async (issue), the time until the copies have drained, and the
resulting link rate; then the same for the zero-copy mapped kernel."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paf_baseband2power_b200 import BMF, Baseband2Power, PinnedBuffer  # noqa: E402

ndf = 8192
blk = BMF.block_bytes
pb = PinnedBuffer(blk)
pb.array[:] = 3
reps = 4
for nchunk in (48, 24, 12, 6, 5):
    kw = {} if nchunk == 48 else dict(nchunk=nchunk, first_chunk=7 if nchunk < 40 else 0, nchunk_total=48)
    with Baseband2Power(**kw) as st:
        st.accumulate_host_async([pb], ndf, finish=True)
        st.wait_input()
        st.wait_output()
        t_issue = t_all = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            st.accumulate_host_async([pb], ndf, finish=True)
            t1 = time.perf_counter()
            st.wait_input()
            t2 = time.perf_counter()
            st.wait_output()
            t_issue += t1 - t0
            t_all += t2 - t0
        nbytes = ndf * nchunk * 7168
        print("shard of %2d chunks: issue %7.3f ms  copies drained after %7.3f ms  -> %6.2f GB/s  (h2d events %.3f ms)"
              % (nchunk, 1e3 * t_issue / reps, 1e3 * t_all / reps, nbytes * reps / t_all / 1e9, st.last_h2d_ms()), flush=True)
for nchunk in (48, 6):
    kw = {} if nchunk == 48 else dict(nchunk=nchunk, first_chunk=7, nchunk_total=48)
    with Baseband2Power(**kw) as st:
        st.accumulate_host_mapped([pb], ndf)
        st.finish()
        t0 = time.perf_counter()
        for _ in range(reps):
            st.accumulate_host_mapped([pb], ndf)
            st.finish()
        dt = time.perf_counter() - t0
        print("zero-copy mapped kernel, %2d chunks: %6.2f GB/s" % (nchunk, ndf * nchunk * 7168 * reps / dt / 1e9), flush=True)
