"""Differential checks against the REFERENCE's own producer stages.

oracle/_ref/ref_paf_diskdb and ref_paf_capture are the reference's diskdb.cu + paf_diskdb.cu
and capture.c + sync.c + hdr.c + paf_capture.c, compiled unmodified from /root/reference
against this repo's PSRDADA-named shim (oracle/Makefile, ref-producers).  That they build at all
shows the shim is source-compatible with every PSRDADA call the reference makes; running the
reference paf_diskdb against the shim's rings pins host/paf_diskdb.c (and the ring semantics:
short last block, empty terminating block, header from the template) to the reference's
behaviour.  On the GPU box the same binary feeds the B200 stage: BASELINE.json configs[0]
with the reference's own producer.
"""
import os
import random
import subprocess
import time

import numpy as np
import pytest

from tests.kat import hdr_kv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "paf_baseband2power_b200")
BIN = os.path.join(PKG, "bin")
REF = os.path.join(ROOT, "oracle", "_ref")
HDR = os.path.join(PKG, "conf", "header_baseband2power.txt")
FRAME = 48 * 7168

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_paf_diskdb")),
                               reason="oracle/_ref producers not built (no /root/reference at build time)")


def _key():
    return "%x" % (random.randint(0x1000, 0xEFFF) & 0xFFF0)


def run(*cmd, **kw):
    return subprocess.run(list(cmd), check=True, capture_output=True, text=True, timeout=120, **kw)


def _through_ring(producer, tmp_path, src_name, out_name, ndf_block, nbufs=3):
    key = _key()
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(ndf_block * FRAME), "-n", str(nbufs))
    try:
        reader = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", out_name, "-W"],
                                  stderr=subprocess.PIPE)
        time.sleep(0.2)
        run(producer, "-a", key, "-b", str(tmp_path), "-c", src_name, "-d", HDR, "-e", "1")
        assert reader.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
    return (tmp_path / out_name).read_bytes()


@needs_ref
def test_reference_producers_link_against_the_shim():
    for exe in ("ref_paf_diskdb", "ref_paf_capture"):
        r = subprocess.run([os.path.join(REF, exe), "-h", "x"], capture_output=True, text=True, timeout=30)
        assert r.returncode != 0 and "Usage" in r.stdout       # usage() then EXIT_FAILURE, paf_diskdb.cu:34-36


@needs_ref
@pytest.mark.parametrize("ndf_file,ndf_block", [(10, 3), (6, 3), (2, 5)])
def test_paf_diskdb_matches_the_reference_paf_diskdb(tmp_path, ndf_file, ndf_block):
    """Same file, same ring geometry, reference producer vs this repo's: the bytes that come out
    of the ring are identical (payload and the header the ring carried)."""
    src = tmp_path / "in.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_file), "-s", "19", "-H", HDR)
    ref_out = _through_ring(os.path.join(REF, "ref_paf_diskdb"), tmp_path, "in.dada", "ref.dada", ndf_block)
    our_out = _through_ring(os.path.join(BIN, "paf_diskdb"), tmp_path, "in.dada", "our.dada", ndf_block)
    assert len(ref_out) == len(our_out) == 4096 + ndf_file * FRAME
    assert ref_out[4096:] == our_out[4096:] == src.read_bytes()[4096:]
    assert hdr_kv(ref_out[:4096].rstrip(b"\0").decode()) == hdr_kv(our_out[:4096].rstrip(b"\0").decode())


@needs_ref
@pytest.mark.gpu
def test_reference_paf_diskdb_feeds_the_b200_stage(tmp_path, oracle_mod, b2p):
    """BASELINE.json configs[0]: the reference's paf_diskdb -> (this repo's) paf_baseband2power."""
    ndf_block, nblk = 64, 3
    kin, kout = _key(), "1" + _key()
    src = tmp_path / "in.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_block * nblk), "-s", "23", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf_block * FRAME), "-n", "4")
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "s.dada", "-W"], stderr=subprocess.PIPE)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path), "-d", "0"], stderr=subprocess.PIPE)
        time.sleep(0.3)
        run(os.path.join(REF, "ref_paf_diskdb"), "-a", kin, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR, "-e", "1")
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    spectra = np.frombuffer((tmp_path / "s.dada").read_bytes()[4096:], dtype=np.float32).reshape(-1, 336)
    assert spectra.shape[0] == nblk
    payload = np.fromfile(src, dtype=np.uint8)[4096:]
    per = ndf_block * FRAME
    for i in range(nblk):
        want = oracle_mod.finish(oracle_mod.accumulate_omp(payload[i * per:(i + 1) * per]))
        assert np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)), i


# ---------------------------------------------------------------------------------------------
# The reference's own paf_capture, run on the shim over a loopback alias, against this repo's.
# It only binds to 10.17.<last hostname digit>.<nic> (paf_capture.c:115-118), so the test needs
# a UTS namespace (hostname "pacifix0") and the alias 10.17.0.1 on lo; it skips where the
# container does not allow that.
# ---------------------------------------------------------------------------------------------
def _loopback_alias(ip="10.17.0.1"):
    import fcntl
    import socket
    import struct
    try:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        ifr = struct.pack("16sH2s4s8s", b"lo:1", socket.AF_INET, b"\0\0", socket.inet_aton(ip), b"\0" * 8)
        fcntl.ioctl(s, 0x8916, ifr)                                   # SIOCSIFADDR
        t = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        t.bind((ip, 0))
        t.close()
        return True
    except OSError:
        return False


def _can_unshare_uts():
    try:
        r = subprocess.run(["unshare", "--uts", "sh", "-c", "hostname pacifix0 && hostname"], capture_output=True, text=True, timeout=10)
        return r.returncode == 0 and r.stdout.strip() == "pacifix0"
    except (OSError, subprocess.TimeoutExpired):
        return False


def _captured_stream(data, seed, oracle_mod, slack=6):
    """Classify every packet slot of a capture file against the generator.

    The header gives the first frame (PICOSECONDS = idf*108 us for a start less than 1 s into
    the period).  Per port (8 chunks) the frame offset -slack..+slack that explains the slots
    best is looked up — the reference keeps one reference header PER PORT (hdr_ref[ithread],
    capture.c:431-434), so its ports may sit a frame or two apart; this repo's capture uses one
    reference for all ports.  Returns (header kv, idf_start, per-port offsets, matched mask,
    zero mask)."""
    kv = hdr_kv(bytes(data[:4096]).rstrip(b"\0").decode())
    idf_start = int(round(int(kv["PICOSECONDS"]) * 1e-12 / 1.08e-4))
    PKT = 7168
    a = data[4096:].reshape(-1, 48, PKT)                       # [frame][chunk][bytes]
    nfr = a.shape[0]
    wpf = FRAME // 8
    first = max(0, idf_start - slack)
    gen = oracle_mod.synth_fill(nfr + 2 * slack, seed=seed, first_word=first * wpf, mode=1).reshape(-1, 48, PKT)
    zero = ~a.any(axis=2)
    matched = np.zeros((nfr, 48), dtype=bool)
    offsets = []
    for port in range(6):
        cs = slice(8 * port, 8 * port + 8)
        best, best_o = None, 0
        for o in range(-slack, slack + 1):
            lo = idf_start + o - first
            if lo < 0:
                continue
            m = (a[:, cs] == gen[lo:lo + nfr, cs]).all(axis=2)
            if best is None or m.sum() > best.sum():
                best, best_o = m, o
        matched[:, cs] = best
        offsets.append(best_o)
    return kv, idf_start, offsets, matched, zero


@needs_ref
def test_paf_capture_agrees_with_the_reference_paf_capture(tmp_path, oracle_mod):
    if not _loopback_alias() or not _can_unshare_uts():
        pytest.skip("needs CAP_NET_ADMIN (loopback alias) and a UTS namespace")
    ndf_block, seed = 16, 77
    (tmp_path / "epoch.txt").write_text("# epoch  days since 1970-01-01\n37 17714.0 2018-07-02\n")
    def one_run(who, attempt):
        """-> capture file as bytes, or None when the capture process hung"""
        key = _key()
        d = tmp_path / f"{who}{attempt}"
        d.mkdir()
        run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(ndf_block * FRAME), "-n", "16")   # more blocks than the capture fills: a missed slot stays zero
        sink = cap = None
        hung = False
        try:
            sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(d), "-f", "cap.dada", "-W"], stderr=subprocess.PIPE)
            common = ["-a", key, "-b", "1", "-c", str(ndf_block), "-d", "0", "-e", "1", "-f", HDR, "-g", str(tmp_path / "epoch.txt"),
                      "-i", "1340.5", "-j", "0.02", "-k", str(d)]
            if who == "ref":
                cmd = ["unshare", "--uts", "sh", "-c", "hostname pacifix0; exec \"$0\" \"$@\"", os.path.join(REF, "ref_paf_capture")] + common
            else:
                cmd = [os.path.join(BIN, who_exe[who])] + common + ["-I", "10.17.0.1", "-t", "3"]
            cap = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            time.sleep(0.7)
            run(os.path.join(BIN, "bmf_replay"), "-D", "10.17.0.1", "-p", "17100", "-n", "2500", "-s", str(seed), "-r", "1500", "-C", "2500")
            try:
                rc = cap.wait(timeout=25)
            except subprocess.TimeoutExpired:
                hung = True          # the reference has unsynchronised shared state (sync.c:109 vs capture.c:542)
                rc = None
            if not hung:
                assert rc == 0, cap.stderr.read().decode()
                assert sink.wait(timeout=60) == 0
        finally:
            for p_ in (cap, sink):
                if p_ is not None and p_.poll() is None:
                    p_.kill()
                    p_.wait()
            run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
        return None if hung else np.fromfile(d / "cap.dada", dtype=np.uint8)

    who_exe = {"our": "paf_capture", "our_stock": "paf_capture_stock"}
    results, attempts = {}, {}
    for who in ("our", "our_stock", "ref"):
        # The reference binary hangs now and then (its own races) and, on a busy machine, drops
        # most of the stream (capture.c:20-29 says so itself): it gets up to 5 tries, the best
        # capture is kept and the number of tries is recorded.  This repo's captures get one try
        # and must not need another.
        best = None
        for attempt in range(1, (5 if who == "ref" else 1) + 1):
            data = one_run(who, attempt)
            if data is None:
                continue
            cls = _captured_stream(data, seed, oracle_mod)
            if best is None or cls[3].mean() > best[3].mean():
                best = cls
                attempts[who] = attempt
            if who != "ref" or cls[3].mean() > 0.4:
                break
        assert best is not None or who == "ref", "this repo's paf_capture must never hang"
        if best is not None:
            results[who] = best
    print("attempts needed:", attempts,
          "reference matched fraction: %.3f" % results["ref"][3].mean() if "ref" in results else "reference hung 5 times")
    (tmp_path / "attempts.json").write_text(__import__("json").dumps(attempts))

    for who, (kv, idf_start, offsets, matched, zero) in results.items():
        # UTC_START / PICOSECONDS are the same function of the first frame in both programs
        # (capture.c:791-843): day 17714, sec 27000 (07:30:00), idf_start*108 us into the second
        assert kv["UTC_START"] == "2018-07-02-07:30:00" and kv["FREQ"] == "1340.5", (who, kv)
        assert int(kv["PICOSECONDS"]) == idf_start * 108000000, who
        # every packet that is in the ring sits at (idf*48 + chunk)*7168 (capture.c:540-542):
        # a slot holds the generator's packet for its own (frame, chunk), or nothing
        unexplained = ~(matched | zero)
        if who.startswith("our"):
            assert not unexplained.any()
            assert offsets == [0] * 6          # one reference frame for all ports
            assert matched.mean() > 0.9
        else:
            # the reference clobbers the first payload byte of frames that went through its side
            # buffer (`tbuf[tbuf_loc + 1] = 'N'`, sync.c:162) and keeps a reference header per
            # port, so a few per cent of its slots differ and its ports may sit frames apart
            # (2-6 % observed from run to run).  How much of the stream it catches at all depends
            # on the machine (28-90 % seen; it loses packets by design, capture.c:20-29): the
            # bounds below are about the reference, not about this repo — what is compared is
            # WHERE the packets it did catch were put.
            assert unexplained.mean() < 0.12, (offsets, int(unexplained.sum()))
            assert matched.mean() > 0.05, matched.mean()
    if "ref" not in results:    # this repo's two captures were checked above; the comparison could not be made
        pytest.skip("the reference's paf_capture hung in 5 of 5 attempts (its own races, sync.c:109 vs capture.c:542)")
