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


def run(ndf=8192, nbufs=4, nblocks=16, gpu=0, kernel="auto", pin=1, timeout=300):
    kin = "%x" % (random.randint(0x2000, 0x6FFF) & 0xFFF0)
    kout = "%x" % (random.randint(0x7000, 0xDFFF) & 0xFFF0)
    blk = ndf * FRAME
    d = tempfile.mkdtemp(prefix="b2p_ring_")
    q = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=timeout)
    q(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(blk), "-n", str(nbufs))
    q(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "8")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", d, "-f", "spectra.dada", "-W"],
                                stderr=subprocess.DEVNULL)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", d, "-d", str(gpu),
                                  "-k", kernel, "-p", str(pin)], stderr=subprocess.PIPE)
        t0 = time.perf_counter()
        prod = q(os.path.join(BIN, "paf_memdb"), "-k", kin, "-n", str(nblocks), "-s", "1", "-H", HDR)
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
    gen = re.search(r"published .* in ([0-9.]+) s", prod.stderr)
    size = os.path.getsize(os.path.join(d, "spectra.dada"))
    t_int = ndf * 128 * 27.0 / 32.0 * 1e-6
    return {"path": "paf_memdb -> ring -> paf_baseband2power -> ring -> paf_dbdisk",
            "ndf_per_block": ndf, "ring_blocks": nbufs, "blocks": nin, "spectra": nout,
            "spectra_file_bytes": size, "ring_pinned": "ring pinned" in log,
            "stage_busy_s": busy, "stage_GBps": round(nin * blk / busy / 1e9, 3),
            "stage_realtime_factor": round(nin * t_int / busy, 2),
            "wall_s": round(wall, 3), "producer_s": float(gen.group(1)) if gen else None}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndf", type=int, default=8192)
    ap.add_argument("--nbufs", type=int, default=4)
    ap.add_argument("--nblocks", type=int, default=16)
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--pin", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(run(a.ndf, a.nbufs, a.nblocks, a.gpu, a.kernel, a.pin)))
