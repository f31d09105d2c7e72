#!/usr/bin/env python
"""BASELINE.json configs[1] through the real process surface: a memory-resident producer
(paf_memdb) -> input ring -> paf_baseband2power (GPU) -> output ring -> paf_dbdisk.
Prints one JSON object with the stage's own throughput (blocks * block bytes / busy seconds,
from its log) and the wall-clock rate.  Measurement only; parity of this path is checked by
tests/test_host_ring.py."""
import argparse
import json
import os
import random
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "paf_baseband2power_b200")
BIN = os.path.join(PKG, "bin")
HDR = os.path.join(PKG, "conf", "header_baseband2power.txt")
FRAME = 48 * 7168


def run(ndf=8192, nbufs=4, nblocks=16, gpu=0, kernel="auto", pin=1, timeout=300, keys=None,
        ndf_integration=0, producer_threads=0, seed=1, start_barrier=None):
    """`keys` = (key_in, key_out) integers when several pipelines run side by side (one per GPU,
    own ring pair each — paf-baseband2power.py:114-115); `gpu` may be a list "0,1,2" (channel
    groups over several GPUs); `ndf_integration` > ndf lets an integration span ring blocks;
    `start_barrier()` is called once the rings exist and the stage is up, right before the
    producer starts (so that side-by-side pipelines stream at the same time)."""
    if keys is None:
        kin = "%x" % (random.randint(0x2000, 0x6FFF) & 0xFFF0)
        kout = "%x" % (random.randint(0x7000, 0xDFFF) & 0xFFF0)
    else:
        kin, kout = "%x" % keys[0], "%x" % keys[1]
    env = dict(os.environ)
    if producer_threads:
        env["OMP_NUM_THREADS"] = str(producer_threads)   # torchrun exports OMP_NUM_THREADS=1
    blk = ndf * FRAME
    d = tempfile.mkdtemp(prefix="b2p_ring_")
    q = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=timeout)
    q(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(blk), "-n", str(nbufs))
    q(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "8")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", d, "-f", "spectra.dada", "-W"],
                                stderr=subprocess.DEVNULL)
        cmd = [os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", d, "-d", str(gpu),
               "-k", kernel, "-p", str(pin)]
        if ndf_integration:
            cmd += ["-n", str(ndf_integration)]
        stage = subprocess.Popen(cmd, stderr=subprocess.PIPE)
        if start_barrier is not None:
            start_barrier()
        t0 = time.perf_counter()
        prod = subprocess.run([os.path.join(BIN, "paf_memdb"), "-k", kin, "-n", str(nblocks), "-s", str(seed), "-H", HDR],
                              check=True, capture_output=True, text=True, timeout=timeout, env=env)
        rc = stage.wait(timeout=timeout)
        wall = time.perf_counter() - t0
        sink.wait(timeout=timeout)
        if rc != 0:
            raise RuntimeError(stage.stderr.read().decode())
    finally:
        subprocess.run([os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin], capture_output=True)
        subprocess.run([os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout], capture_output=True)
    log = open(os.path.join(d, "paf_baseband2power.log")).read()
    m = re.search(r"END: (\d+) blocks in, (\d+) spectra out, ([0-9.]+) s busy", log)
    nin, nout, busy = int(m.group(1)), int(m.group(2)), float(m.group(3))
    st = re.search(r"STEADY: (\d+) blocks after the split settled, ([0-9.]+) s busy, slowest block ([0-9.]+) s", log)
    steady = None
    if st and float(st.group(2)) > 0:
        steady = {"blocks": int(st.group(1)), "busy_s": float(st.group(2)), "slowest_block_s": float(st.group(3)),
                  "GBps": round(int(st.group(1)) * blk / float(st.group(2)) / 1e9, 3)}
    gen = re.search(r"published .* in ([0-9.]+) s", prod.stderr)
    size = os.path.getsize(os.path.join(d, "spectra.dada"))
    t_int = ndf * 128 * 27.0 / 32.0 * 1e-6
    spectra = None
    try:
        import numpy as np
        spectra = np.fromfile(os.path.join(d, "spectra.dada"), dtype=np.float32, offset=4096).reshape(-1, 336)
    except Exception:
        pass
    groups = re.findall(r"gpu (\d+): host link ([0-9.]+) GB/s -> (\d+) chunks", log)
    moves = re.findall(r"rebalanced chunks per gpu:((?: gpu\d+:\d+)+)", log)
    final = [int(x.split(":")[1]) for x in moves[-1].split()] if moves else None
    return {"path": "paf_memdb -> ring -> paf_baseband2power -> ring -> paf_dbdisk",
            "gpu": gpu, "keys": [kin, kout], "_spectra": spectra,
            "channel_groups": [{"gpu": int(a), "link_GBps": float(b), "chunks": int(c)} for a, b, c in groups] or None,
            "rebalanced": len(moves), "chunks_after_rebalancing": final,
            "ndf_per_block": ndf, "ring_blocks": nbufs, "blocks": nin, "spectra": nout,
            "spectra_file_bytes": size, "ring_pinned": "ring pinned" in log,
            "stage_busy_s": busy, "stage_GBps": round(nin * blk / busy / 1e9, 3),
            "stage_realtime_factor": round(nin * t_int / busy, 2), "steady": steady,
            "wall_s": round(wall, 3), "producer_s": float(gen.group(1)) if gen else None}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndf", type=int, default=8192)
    ap.add_argument("--nbufs", type=int, default=4)
    ap.add_argument("--nblocks", type=int, default=16)
    ap.add_argument("--gpu", default="0", help="GPU index, or a list 0,1,2,3 (channel groups)")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--pin", type=int, default=1)
    a = ap.parse_args()
    res = run(a.ndf, a.nbufs, a.nblocks, a.gpu, a.kernel, a.pin)
    res.pop("_spectra", None)
    print(json.dumps(res))
