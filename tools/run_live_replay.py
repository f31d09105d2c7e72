#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU over loopback: bmf_replay (synthetic BMF packets at a
chosen fraction of line rate) -> paf_capture -> ring -> paf_baseband2power -> ring -> paf_dbdisk.
Prints one JSON object: packets sent / received / zero-filled, capture wall time, stage busy
time.  Loopback UDP costs two kernel copies per packet on the host CPUs, so this measures the
host more than anything else; a real deployment receives from NICs."""
import argparse
import json
import os
import random
import re
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "paf_baseband2power_b200")
BIN = os.path.join(PKG, "bin")
HDR = os.path.join(PKG, "conf", "header_baseband2power.txt")
FRAME = 48 * 7168
LINE_FPS = 1.0 / 1.08e-4


def run(ndf=8192, nblocks=3, rate_frac=1.0, threads=6, gpu=0, timeout=300, keys=None, port=None,
        nbufs=4, ndf_integration=0, start_barrier=None, settle_s=3.0, gso=8, capture="paf_capture",
        capture_args=()):
    """One beam: bmf_replay -> UDP loopback -> paf_capture -> ring -> paf_baseband2power -> ring ->
    paf_dbdisk.  `keys`/`port` keep side-by-side beams (one per GPU) apart: own ring pair, own six
    UDP ports (capture.h:22-24 has one port set per NIC; here one per beam on loopback)."""
    if keys is None:
        kin = "%x" % (random.randint(0x2000, 0x6FFF) & 0xFFF0)
        kout = "%x" % (random.randint(0x7000, 0xDFFF) & 0xFFF0)
    else:
        kin, kout = "%x" % keys[0], "%x" % keys[1]
    if port is None:
        port = random.randint(20000, 40000)
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads)
    d = tempfile.mkdtemp(prefix="b2p_live_")
    q = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=timeout, env=env)
    q(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf * FRAME), "-n", str(nbufs))
    q(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "8")
    nframes = ndf * nblocks
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", d, "-f", "spectra.dada", "-W"], stderr=subprocess.DEVNULL)
        scmd = [os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", d, "-d", str(gpu)]
        if ndf_integration:
            scmd += ["-n", str(ndf_integration)]
        stage = subprocess.Popen(scmd, stderr=subprocess.PIPE)
        cap = subprocess.Popen([os.path.join(BIN, capture), *capture_args, "-a", kin, "-b", "1", "-c", str(ndf), "-d", "0", "-f", HDR, "-g", "none",
                                "-i", "1340.5", "-j", repr(nframes * 1.08e-4), "-k", d, "-I", "127.0.0.1", "-p", str(port), "-t", "5"],
                               stderr=subprocess.PIPE)
        time.sleep(settle_s)   # the stage page-locks the ring (GBs) before it reads
        if start_barrier is not None:
            start_barrier()
        rep = q(os.path.join(BIN, "bmf_replay"), "-D", "127.0.0.1", "-p", str(port), "-n", str(nframes + 64), "-s", "5",
                "-r", repr(LINE_FPS * rate_frac), "-C", "512", "-T", str(threads), "-G", str(gso))
        cap.wait(timeout=timeout)
        rc = stage.wait(timeout=timeout)
        sink.wait(timeout=timeout)
        if rc != 0:
            raise RuntimeError(stage.stderr.read().decode())
    finally:
        subprocess.run([os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin], capture_output=True)
        subprocess.run([os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout], capture_output=True)
    clog = open(os.path.join(d, "paf_capture.log")).read()
    slog = open(os.path.join(d, "paf_baseband2power.log")).read()
    m = re.search(r"blocks (\d+)\s+frames received (\d+)\s+expected (\d+)\s+missing\(zero-filled\) (\d+)\s+late (\d+).*in ([0-9.]+) s", clog)
    s = re.search(r"END: (\d+) blocks in, (\d+) spectra out, ([0-9.]+) s busy", slog)
    gro = re.search(r"udp_gro (on|off)\s+messages (\d+) \(([0-9.]+) frames per message\)\s+blocked on a full ring ([0-9.]+) s", clog)
    r = re.search(r"(\d+) packets sent.*in ([0-9.]+) s \(([0-9.]+) frames/s, ([0-9.]+) GB/s, ([0-9.]+)x line rate", rep.stdout)
    out = {"path": "bmf_replay -> UDP loopback -> paf_capture -> ring -> paf_baseband2power -> ring -> paf_dbdisk",
           "gpu": gpu, "udp_port_base": port, "capture": capture, "sender_gso_frames": gso, "ndf_per_block": ndf, "blocks_requested": nblocks, "rate_frac_requested": rate_frac, "sender_threads": threads}
    if r:
        out.update({"packets_sent": int(r.group(1)), "replay_s": float(r.group(2)), "replay_GBps": float(r.group(4)),
                    "replay_x_line_rate": float(r.group(5))})
    if m:
        out.update({"capture_blocks": int(m.group(1)), "packets_received": int(m.group(2)), "packets_expected": int(m.group(3)),
                    "packets_zero_filled": int(m.group(4)), "packets_late": int(m.group(5)), "capture_s": float(m.group(6)),
                    "received_frac": round(int(m.group(2)) / max(1, int(m.group(3))), 4)})
    if gro:
        out.update({"udp_gro": gro.group(1), "frames_per_recv_message": float(gro.group(3)),
                    "capture_blocked_on_ring_s": float(gro.group(4))})
    if s:
        out.update({"stage_blocks": int(s.group(1)), "spectra": int(s.group(2)), "stage_busy_s": float(s.group(3))})
        t_data = int(s.group(1)) * ndf * 1.08e-4
        if float(s.group(3)) > 0:
            out["stage_headroom_x_realtime"] = round(t_data / float(s.group(3)), 2)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndf", type=int, default=8192)
    ap.add_argument("--nblocks", type=int, default=3)
    ap.add_argument("--rate", type=float, default=1.0, help="fraction of the BMF line rate (9259 frames/s)")
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--gso", type=int, default=8, help="frames per send call of the replayer (1 = plain sendmsg)")
    ap.add_argument("--capture", default="paf_capture", help="paf_capture | paf_capture_stock")
    ap.add_argument("--no-gro", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.ndf, a.nblocks, a.rate, a.threads, gso=a.gso, capture=a.capture,
                         capture_args=("-G", "0") if a.no_gro else ())))
