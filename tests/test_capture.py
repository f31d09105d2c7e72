"""paf_capture over loopback UDP fed by the synthetic BMF replayer (CPU only), and the BMF
packet header against vectors produced by the reference's own hdr.c."""
import ctypes
import json
import os
import random
import re
import subprocess
import time

import numpy as np
import pytest

from tests.kat import hdr_kv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "paf_baseband2power_b200")
BIN = os.path.join(PKG, "bin")
HDR = os.path.join(PKG, "conf", "header_baseband2power.txt")
FRAME, PKT = 48 * 7168, 7168


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(os.path.join(BIN, "paf_capture")):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG, "host"), "all"], check=True)


def test_header_decode_matches_reference_hdr_c(tmp_path):
    """bmf_hdr_decode (host/bmf_packet.h) against tests/golden/bmf_hdr_vectors.json, which was
    produced by the reference's hdr.c:10-28 (oracle/_ref)."""
    src = tmp_path / "t.c"
    src.write_text('#include "%s"\n'
                   'void dec(const void *p, bmf_hdr_t *h) { bmf_hdr_decode(p, h); }\n'
                   'void enc(void *p, const bmf_hdr_t *h) { bmf_hdr_encode(p, h); }\n'
                   'long since(unsigned long s, unsigned long i, unsigned long s0, unsigned long i0) { return bmf_frames_since(s, i, s0, i0); }\n'
                   'int chunk(unsigned char x, unsigned char y) { return bmf_chunk_of_source(x, y); }\n'
                   % os.path.join(PKG, "host", "bmf_packet.h"))
    so = tmp_path / "t.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))

    class H(ctypes.Structure):
        _fields_ = [("valid", ctypes.c_int), ("idf", ctypes.c_uint64), ("sec", ctypes.c_uint64),
                    ("epoch", ctypes.c_int), ("beam", ctypes.c_int), ("freq", ctypes.c_double)]
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bmf_hdr_vectors.json")))
    for v in gold["vectors"]:
        raw = bytes.fromhex(v["raw"])
        h = H()
        lib.dec(raw, ctypes.byref(h))
        assert (h.valid, h.idf, h.sec, h.epoch, h.beam, h.freq) == (v["valid"], v["idf"], v["sec"], v["epoch"], v["beam"], v["freq"])
        out = ctypes.create_string_buffer(64)      # encode(decode(x)) decodes to the same fields
        lib.enc(out, ctypes.byref(h))
        h2 = H()
        lib.dec(out.raw, ctypes.byref(h2))
        assert (h2.valid, h2.idf, h2.sec, h2.epoch, h2.beam, h2.freq) == (h.valid, h.idf, h.sec, h.epoch, h.beam, h.freq)
    lib.since.restype = ctypes.c_long
    lib.since.argtypes = [ctypes.c_ulong] * 4
    assert lib.since(27000, 10, 27000, 3) == 7
    assert lib.since(27027, 2, 27000, 249999) == 3                  # across a period boundary
    assert lib.since(27000, 0, 27027, 0) == -250000
    # source address -> chunk, capture.c:570-584: (X-1)*6 + ceil(Y/2) - 1
    assert [lib.chunk(1, 1), lib.chunk(1, 2), lib.chunk(1, 11), lib.chunk(2, 1), lib.chunk(8, 12)] == [0, 0, 5, 6, 47]


# (capture binary, extra capture flags, extra replay flags): the ring-shim back-end with two open
# blocks and the stock-PSRDADA back-end (open/close_block_write + spill window), each with plain
# datagrams and with UDP segmentation offload on the sender / UDP_GRO on the receiver
VARIANTS = {
    "ahead-gro": ("paf_capture", [], ["-G", "8"]),
    "ahead-plain": ("paf_capture", ["-G", "0"], []),
    "stock-gro": ("paf_capture_stock", ["-w", "4"], ["-G", "5"]),
    "stock-plain": ("paf_capture_stock", ["-G", "0", "-w", "2"], []),
}


def _capture_run(tmp_path, ndf_block, nframes_capture, nframes_sent, seed, drop_every=0, rate=800, variant="ahead-plain"):
    exe, cap_extra, rep_extra = VARIANTS[variant]
    key = "%x" % (random.randint(0x2000, 0xDFFF) & 0xFFF0)
    port = random.randint(20000, 40000)
    run = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=120)
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(ndf_block * FRAME), "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", "cap.dada", "-W"],
                                stderr=subprocess.PIPE)
        cap = subprocess.Popen([os.path.join(BIN, exe), "-a", key, "-b", "1", "-c", str(ndf_block), "-d", "0",
                                "-f", HDR, "-g", "none", "-i", "1340.5", "-j", repr(nframes_capture * 1.08e-4),
                                "-k", str(tmp_path), "-I", "127.0.0.1", "-p", str(port), "-t", "3"] + cap_extra, stderr=subprocess.PIPE)
        time.sleep(0.5)
        args = [os.path.join(BIN, "bmf_replay"), "-D", "127.0.0.1", "-p", str(port), "-n", str(nframes_sent),
                "-s", str(seed), "-r", str(rate), "-i", "249990"] + rep_extra      # the frame counter wraps mid-run
        if drop_every:
            args += ["-L", str(drop_every)]
        rep = run(*args)
        assert cap.wait(timeout=60) == 0, cap.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
    data = np.fromfile(tmp_path / "cap.dada", dtype=np.uint8)
    log = (tmp_path / "paf_capture.log").read_text()
    return data, log, rep.stdout


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_capture_assembles_the_generator_block(tmp_path, oracle_mod, variant):
    """40 frames captured into 16-frame blocks: every packet that arrived sits at
    (idf*48 + chunk)*7168 (capture.c:540-542) and equals the generator; what did not arrive
    (UDP may drop, and the tail of the last block was never sent) is zero and is counted."""
    ndf_block, ncap, nsent = 16, 40, 44
    data, log, _ = _capture_run(tmp_path, ndf_block, ncap, nsent, seed=9, variant=variant)
    nblk = (ncap + ndf_block - 1) // ndf_block
    assert data.size == 4096 + nblk * ndf_block * FRAME
    pay = data[4096:]
    want = oracle_mod.synth_fill(ncap, seed=9, mode=1)
    npk_cap, npk_tot = ncap * 48, nblk * ndf_block * 48
    equal = zero = 0
    for i in range(npk_tot):
        got = pay[i * PKT:(i + 1) * PKT]
        if i < npk_cap and np.array_equal(got, want[i * PKT:(i + 1) * PKT]):
            equal += 1
        else:
            assert not got.any(), f"packet {i} is neither the generator's nor zero"
            zero += 1
    m = re.search(r"blocks (\d+)\s+frames received (\d+)\s+expected (\d+)\s+missing\(zero-filled\) (\d+)", log)
    assert m, log
    blocks, recv, expected, missing = map(int, m.groups())
    assert (blocks, expected) == (nblk, npk_tot)
    assert recv == equal and missing == zero and recv + missing == expected
    assert equal >= 0.9 * npk_cap                       # loopback at 800 frames/s should lose ~nothing
    hdr = bytes(data[:4096]).rstrip(b"\0").decode()
    kv = hdr_kv(hdr)
    assert kv["UTC_START"] == "2018-07-02-10:30:26" and kv["FREQ"] == "1340.5"   # 27000 s + 249990*108us
    assert kv["PICOSECONDS"] == "998920000000" and kv["INSTRUMENT"] == "PAF-BMF"
    gro = re.search(r"udp_gro (on|off)\s+messages (\d+) \(([0-9.]+) frames per message\)", log)
    assert gro, log
    if variant.endswith("plain"):
        assert gro.group(1) == "off" and float(gro.group(3)) <= 1.0
    elif gro.group(1) == "on":                          # a kernel without UDP_GRO falls back to "off"
        assert float(gro.group(3)) > 1.5                # several frames per message did arrive coalesced


def test_stock_capture_uses_only_the_calls_the_reference_capture_makes(tmp_path):
    """-DB2P_STOCK_PSRDADA: the object refers to no ring symbol beyond stock PSRDADA's (the block
    calls are the reference's two, capture.c:316 and sync.c:101-109) — in particular not to this
    repo's ipcbuf_get_write_ahead — so it links against -lpsrdada as it stands."""
    obj = tmp_path / "cap_stock.o"
    subprocess.run(["gcc", "-O1", "-std=gnu11", "-DB2P_STOCK_PSRDADA", "-DB2P_NO_SHIM_EXTENSIONS", "-c", "-o", str(obj),
                    os.path.join(PKG, "host", "paf_capture.c")], check=True)
    und = subprocess.run(["nm", "-u", str(obj)], check=True, capture_output=True, text=True).stdout.split()
    ring = sorted(x for x in und if x.startswith(("ipcbuf_", "ipcio_", "dada_hdu_", "ascii_header_", "multilog", "fileread")))
    stock = {"ipcio_open_block_write", "ipcio_close_block_write", "ipcbuf_get_bufsz", "ipcbuf_get_next_write",
             "ipcbuf_mark_filled", "ipcbuf_enable_sod", "ipcbuf_disable_sod", "dada_hdu_create", "dada_hdu_set_key",
             "dada_hdu_connect", "dada_hdu_lock_write", "dada_hdu_unlock_write", "dada_hdu_disconnect",
             "dada_hdu_destroy", "ascii_header_set", "multilog", "multilog_open", "multilog_add", "multilog_close",
             "fileread"}
    assert "ipcbuf_get_write_ahead" not in ring
    assert set(ring) <= stock, sorted(set(ring) - stock)
    assert {"ipcio_open_block_write", "ipcio_close_block_write"} <= set(ring)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_capture_zero_fills_injected_loss(tmp_path, oracle_mod, variant):
    ndf_block, ncap = 8, 24
    data, log, rep = _capture_run(tmp_path, ndf_block, ncap, ncap + 2, seed=3, drop_every=7, variant=variant)
    pay = data[4096:]
    want = oracle_mod.synth_fill(ncap, seed=3, mode=1)
    dropped = 0
    for i in range(ncap * 48):
        got = pay[i * PKT:(i + 1) * PKT]
        if (i + 1) % 7 == 0:                            # the replayer withheld every 7th packet
            assert not got.any()
            dropped += 1
        else:
            assert np.array_equal(got, want[i * PKT:(i + 1) * PKT]) or not got.any()
    assert dropped == (ncap * 48) // 7
    assert "dropped on purpose" in rep


def test_capture_refuses_a_mis_sized_ring(tmp_path):
    key = "%x" % (random.randint(0x2000, 0xDFFF) & 0xFFF0)
    subprocess.run([os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(3 * FRAME), "-n", "4"], check=True, capture_output=True)
    try:
        r = subprocess.run([os.path.join(BIN, "paf_capture"), "-a", key, "-c", "4", "-f", HDR, "-k", str(tmp_path),
                            "-I", "127.0.0.1", "-p", "31999", "-t", "1"], capture_output=True, text=True, timeout=30)
        assert r.returncode != 0 and "Buffer size mismatch" in r.stderr
    finally:
        subprocess.run([os.path.join(BIN, "paf_dada_db"), "-d", "-k", key], check=True, capture_output=True)


@pytest.mark.gpu
def test_live_pipeline_replay_capture_baseband2power(tmp_path, oracle_mod, b2p):
    """BASELINE.json configs[4] in miniature: synthetic BMF packets -> paf_capture -> ring ->
    paf_baseband2power (GPU) -> ring -> paf_dbdisk.  Spectra of blocks that arrived complete
    are bit-identical to the oracle on the generator's block."""
    ndf_block, nblk = 32, 3
    kin = "%x" % (random.randint(0x2000, 0x6FFF) & 0xFFF0)
    kout = "%x" % (random.randint(0x7000, 0xDFFF) & 0xFFF0)
    port = random.randint(20000, 40000)
    run = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=120)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf_block * FRAME), "-n", "4")
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "spectra.dada", "-W"], stderr=subprocess.PIPE)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path), "-d", "0"], stderr=subprocess.PIPE)
        cap = subprocess.Popen([os.path.join(BIN, "paf_capture"), "-a", kin, "-b", "1", "-c", str(ndf_block), "-d", "0", "-f", HDR,
                                "-g", "none", "-i", "1340.5", "-j", repr(ndf_block * nblk * 1.08e-4), "-k", str(tmp_path),
                                "-I", "127.0.0.1", "-p", str(port), "-t", "3"], stderr=subprocess.PIPE)
        time.sleep(1.0)
        run(os.path.join(BIN, "bmf_replay"), "-D", "127.0.0.1", "-p", str(port), "-n", str(ndf_block * nblk + 2), "-s", "21", "-r", "800")
        assert cap.wait(timeout=60) == 0
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    out = (tmp_path / "spectra.dada").read_bytes()
    spectra = np.frombuffer(out[4096:], dtype=np.float32).reshape(-1, 336)
    assert spectra.shape[0] == nblk
    log = (tmp_path / "paf_capture.log").read_text()
    missing = int(re.search(r"missing\(zero-filled\) (\d+)", log).group(1))
    want_all = oracle_mod.synth_fill(ndf_block * nblk, seed=21, mode=1)
    per = ndf_block * FRAME
    exact = 0
    for i in range(nblk):
        want = oracle_mod.finish(oracle_mod.accumulate_omp(want_all[i * per:(i + 1) * per]))
        if np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)):
            exact += 1
        else:
            assert missing > 0 and np.all(spectra[i] <= want)   # lost packets only remove power
    assert exact == nblk or missing > 0
    assert exact >= 1


def test_capture_can_keep_the_frame_headers(tmp_path, oracle_mod):
    """-d 1 (debug mode of the reference, capture.c:216,222): the ring holds whole 7232-byte frames."""
    ndf_block, ncap = 8, 16
    DF = 7232
    key = "%x" % (random.randint(0x2000, 0xDFFF) & 0xFFF0)
    port = random.randint(20000, 40000)
    run = lambda *c: subprocess.run(list(c), check=True, capture_output=True, text=True, timeout=120)
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(ndf_block * 48 * DF), "-n", "4")
    sink = cap = None
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", "cap.dada", "-W"], stderr=subprocess.PIPE)
        cap = subprocess.Popen([os.path.join(BIN, "paf_capture"), "-a", key, "-b", "1", "-c", str(ndf_block), "-d", "1", "-f", HDR, "-g", "none",
                                "-i", "1340.5", "-j", repr(ncap * 1.08e-4), "-k", str(tmp_path), "-I", "127.0.0.1", "-p", str(port), "-t", "3"],
                               stderr=subprocess.PIPE)
        time.sleep(0.5)
        run(os.path.join(BIN, "bmf_replay"), "-D", "127.0.0.1", "-p", str(port), "-n", str(ncap + 2), "-s", "4", "-r", "500", "-i", "100", "-b", "5")
        assert cap.wait(timeout=60) == 0, cap.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        for p_ in (cap, sink):
            if p_ is not None and p_.poll() is None:
                p_.kill()
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
    frames = np.fromfile(tmp_path / "cap.dada", dtype=np.uint8)[4096:].reshape(-1, 48, DF)
    want = oracle_mod.synth_fill(ncap, seed=4, mode=1).reshape(ncap, 48, PKT)
    seen = 0
    for f in range(ncap):
        for c in range(48):
            fr = frames[f, c]
            if not fr.any():
                continue
            seen += 1
            w0 = int.from_bytes(bytes(fr[0:8]), "big")
            w2 = int.from_bytes(bytes(fr[16:24]), "big")
            assert (w0 & 0xFFFFFFFF) == 100 + f and (w0 >> 63) == 1          # idf, valid (hdr.c:15-18)
            assert (w2 & 0xFFFF) == 5 and ((w2 >> 16) & 0xFFFF) == 1173 + 7 * c   # beam, chunk frequency (hdr.c:23-25)
            assert np.array_equal(fr[64:], want[f, c])
    assert seen >= 0.9 * ncap * 48
