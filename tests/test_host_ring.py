"""Host-side logic on CPU: the PSRDADA-named ring shim, the DADA header helpers and the
producer/consumer executables (paf_dada_db, b2p_gen, paf_diskdb, paf_dbdisk)."""
import ctypes
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
HDR = os.path.join(PKG, "conf", "header_baseband2power.txt")
FRAME = 48 * 7168


def _key():
    return random.randint(0x1000, 0xEFFF) & 0xFFFE


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(os.path.join(BIN, "paf_diskdb")):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG, "host"), "all"], check=True)


def run(*cmd, **kw):
    return subprocess.run(list(cmd), check=True, capture_output=True, text=True, timeout=120, **kw)


@pytest.fixture
def dada():
    lib = ctypes.CDLL(os.path.join(PKG, "host", "libb2p_dada.so"))
    lib.ascii_header_get.restype = ctypes.c_int
    lib.ascii_header_set.restype = ctypes.c_int
    return lib


def test_ascii_header_set_get(dada):
    buf = ctypes.create_string_buffer(open(HDR, "rb").read(), 4096)
    val = ctypes.create_string_buffer(64)
    assert dada.ascii_header_get(buf, b"INSTRUMENT", b"%63s", val) == 1 and val.value == b"PAF-BMF"
    assert dada.ascii_header_get(buf, b"UTC_START", b"%63s", val) == 1 and val.value == b"unset"
    n = ctypes.c_int()
    assert dada.ascii_header_get(buf, b"NCHAN", b"%d", ctypes.byref(n)) == 1 and n.value == 336
    assert dada.ascii_header_set(buf, b"UTC_START", b"%s", b"2026-10-18-12:00:00") == 0
    assert dada.ascii_header_get(buf, b"UTC_START", b"%63s", val) == 1 and val.value == b"2026-10-18-12:00:00"
    assert dada.ascii_header_set(buf, b"NEWKEY", b"%d", 42) == 0
    assert dada.ascii_header_get(buf, b"NEWKEY", b"%d", ctypes.byref(n)) == 1 and n.value == 42
    assert dada.ascii_header_get(buf, b"MISSING", b"%d", ctypes.byref(n)) < 0
    # NCHAN must not be confused with a key it prefixes / is prefixed by
    assert dada.ascii_header_set(buf, b"NCHAN_EXTRA", b"%d", 7) == 0
    assert dada.ascii_header_get(buf, b"NCHAN", b"%d", ctypes.byref(n)) == 1 and n.value == 336


@pytest.mark.parametrize("ndf_file,ndf_block,nbufs", [(10, 3, 3), (6, 3, 2), (1, 4, 2), (17, 1, 4)])
def test_diskdb_ring_dbdisk_roundtrip(tmp_path, ndf_file, ndf_block, nbufs):
    """file -> paf_diskdb -> ring -> paf_dbdisk -> file: payload identical, incl. a short last
    block (10/3), an exact multiple (6/3: empty terminating block) and a ring that wraps."""
    key = "%x" % _key()
    src, dst = tmp_path / "in.dada", tmp_path / "out.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_file), "-s", "11", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(ndf_block * FRAME), "-n", str(nbufs))
    try:
        reader = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", "out.dada", "-W"],
                                  stderr=subprocess.PIPE)
        time.sleep(0.2)
        run(os.path.join(BIN, "paf_diskdb"), "-a", key, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR, "-e", "1")
        assert reader.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
    a, b = src.read_bytes(), dst.read_bytes()
    assert len(b) == 4096 + ndf_file * FRAME
    assert a[4096:] == b[4096:]
    hdr = b[:4096].rstrip(b"\0").decode()
    assert hdr_kv(hdr)["FILE_SIZE"] == str(ndf_file * FRAME)
    assert hdr_kv(hdr)["INSTRUMENT"] == "PAF-BMF"   # ring header came from the template (diskdb.cu:79-85)


def test_second_writer_is_refused(tmp_path):
    key = "%x" % _key()
    run(os.path.join(BIN, "b2p_gen"), "-o", str(tmp_path / "in.dada"), "-n", "2", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(FRAME), "-n", "2")
    try:
        # no reader: the first writer fills both buffers and blocks holding the write lock
        w1 = subprocess.Popen([os.path.join(BIN, "paf_diskdb"), "-a", key, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR],
                              stderr=subprocess.PIPE)
        time.sleep(0.5)
        w2 = subprocess.run([os.path.join(BIN, "paf_diskdb"), "-a", key, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR],
                            capture_output=True, text=True, timeout=30)
        assert w2.returncode != 0 and "already has a writer" in w2.stderr
        drain = subprocess.run([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", "o.dada", "-W"],
                               capture_output=True, timeout=60)
        assert drain.returncode == 0 and w1.wait(timeout=30) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)


def test_missing_ring_and_bad_key_fail_loudly(tmp_path):
    r = subprocess.run([os.path.join(BIN, "paf_diskdb"), "-a", "zzzz"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not parse key" in r.stderr
    (tmp_path / "x.dada").write_bytes(b"\0" * 5000)
    r = subprocess.run([os.path.join(BIN, "paf_diskdb"), "-a", "%x" % _key(), "-b", str(tmp_path), "-c", "x.dada", "-d", HDR],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Can not connect to hdu" in r.stderr


def test_b2p_gen_matches_oracle_generator(tmp_path, oracle_mod):
    run(os.path.join(BIN, "b2p_gen"), "-o", str(tmp_path / "g.dada"), "-n", "3", "-s", "77", "-m", "0", "-H", HDR)
    data = np.fromfile(tmp_path / "g.dada", dtype=np.uint8)
    assert np.array_equal(data[4096:], oracle_mod.synth_fill(3, seed=77, mode=0))
    hdr = bytes(data[:4096]).rstrip(b"\0").decode()
    assert hdr_kv(hdr)["UTC_START"] == "2026-10-18-00:00:00" and hdr_kv(hdr)["NBIT"] == "16"


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["ldg", "tma"])
def test_pipeline_diskdb_baseband2power_dbdisk(tmp_path, oracle_mod, b2p, kernel):
    """BASELINE.json configs[0]: paf_diskdb -> paf_baseband2power -> (dada_dbdisk role), one
    synthetic DADA file, spectra compared bit-for-bit with the oracle.  Blocks are 64 frames
    (22 MB) instead of 8192 so the test stays small; 5 blocks + a short tail that is dropped."""
    ndf_block, nblk, tail = 64, 5, 10
    kin, kout = "%x" % _key(), "%x" % (_key() | 0x10000)
    src = tmp_path / "in.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_block * nblk + tail), "-s", "5", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf_block * FRAME), "-n", "4")
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "spectra.dada", "-W"],
                                stderr=subprocess.PIPE)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path),
                                  "-d", "0", "-k", kernel], stderr=subprocess.PIPE)
        time.sleep(0.3)
        run(os.path.join(BIN, "paf_diskdb"), "-a", kin, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR, "-e", "1")
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    out = (tmp_path / "spectra.dada").read_bytes()
    assert len(out) == 4096 + nblk * 1344
    spectra = np.frombuffer(out[4096:], dtype=np.float32).reshape(nblk, 336)
    payload = np.fromfile(src, dtype=np.uint8)[4096:]
    for i in range(nblk):
        blk = payload[i * ndf_block * FRAME:(i + 1) * ndf_block * FRAME]
        want = oracle_mod.finish(oracle_mod.accumulate_omp(blk))
        assert np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)), i
    hdr = out[:4096].rstrip(b"\0").decode()
    tsamp = ndf_block * 128 * 27.0 / 32.0
    kv = hdr_kv(hdr)
    assert kv["TSAMP"] == "%.4f" % tsamp and kv["NBIT"] == "32" and kv["NCHAN"] == "336"
    log = (tmp_path / "paf_baseband2power.log").read_text()
    assert "START PAF_PROCESS" in log and "5 spectra out" in log and "partial integration of 10 frames" in log


@pytest.mark.gpu
def test_memory_resident_stream_through_the_rings(tmp_path, oracle_mod, b2p):
    """BASELINE.json configs[1] in miniature: paf_memdb publishes 11 blocks (3 distinct, generated in
    place in the ring) -> paf_baseband2power -> paf_dbdisk; spectrum i is the oracle's for block i % 3."""
    ndf, nbufs, nblk = 48, 3, 11
    kin, kout = "%x" % _key(), "%x" % (_key() | 0x10000)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf * FRAME), "-n", str(nbufs))
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "s.dada", "-W"], stderr=subprocess.PIPE)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path), "-d", "0"], stderr=subprocess.PIPE)
        time.sleep(0.3)
        run(os.path.join(BIN, "paf_memdb"), "-k", kin, "-n", str(nblk), "-s", "33", "-H", HDR)
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    spectra = np.frombuffer((tmp_path / "s.dada").read_bytes()[4096:], dtype=np.float32).reshape(nblk, 336)
    wpb = ndf * FRAME // 8
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(oracle_mod.synth_fill(ndf, seed=33, first_word=b * wpb, mode=1)))
            for b in range(nbufs)]
    for i in range(nblk):
        assert np.array_equal(spectra[i].view(np.uint32), want[i % nbufs].view(np.uint32)), i


@pytest.mark.gpu
def test_stage_options_average_and_multi_block_integration(tmp_path, oracle_mod, b2p):
    """-s 1 (time average: scale 1/(frames*128)) and -n (an integration spanning two ring blocks)."""
    ndf_block, nblk = 32, 5
    kin, kout = "%x" % _key(), "%x" % (_key() | 0x10000)
    src = tmp_path / "in.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_block * nblk), "-s", "8", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf_block * FRAME), "-n", "3")
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "s.dada", "-W"], stderr=subprocess.PIPE)
        stage = subprocess.Popen([os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path), "-d", "0",
                                  "-s", "1", "-n", str(2 * ndf_block), "-p", "0"], stderr=subprocess.PIPE)
        time.sleep(0.3)
        run(os.path.join(BIN, "paf_diskdb"), "-a", kin, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR, "-e", "1")
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    out = (tmp_path / "s.dada").read_bytes()
    spectra = np.frombuffer(out[4096:], dtype=np.float32).reshape(-1, 336)
    assert spectra.shape[0] == 2                                   # 5 blocks -> 2 integrations of 2 blocks, 1 block left over
    payload = np.fromfile(src, dtype=np.uint8)[4096:]
    per = 2 * ndf_block * FRAME
    scale = 1.0 / (2 * ndf_block * 128)                            # 2^-13
    for i in range(2):
        want = oracle_mod.finish(oracle_mod.accumulate_omp(payload[i * per:(i + 1) * per]), scale)
        assert np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)), i
    hdr = out[:4096].rstrip(b"\0").decode()
    assert hdr_kv(hdr)["TSAMP"] == "%.4f" % (2 * ndf_block * 128 * 27.0 / 32.0)
    assert "partial integration of 32 frames" in (tmp_path / "paf_baseband2power.log").read_text()


@pytest.mark.gpu
@pytest.mark.parametrize("gpus,chunks", [("0,0,0", None), ("0,0,0,0", "5,9,11,23")])
def test_stage_spreads_channel_groups_over_a_gpu_list(tmp_path, oracle_mod, b2p, gpus, chunks):
    """-d with a list: one beam's chunks are split over the listed GPUs (measured-link split, or
    -g counts); every GPU reads its columns of the same ring block and the 1344-byte output
    block is the oracle's spectrum.  On a one-GPU box all shards land on GPU 0."""
    ndf_block, nblk = 40, 4
    kin, kout = "%x" % _key(), "%x" % (_key() | 0x10000)
    src = tmp_path / "in.dada"
    run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(ndf_block * nblk), "-s", "15", "-H", HDR)
    run(os.path.join(BIN, "paf_dada_db"), "-k", kin, "-b", str(ndf_block * FRAME), "-n", "3")
    run(os.path.join(BIN, "paf_dada_db"), "-k", kout, "-b", "1344", "-n", "4")
    try:
        sink = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", kout, "-D", str(tmp_path), "-f", "s.dada", "-W"], stderr=subprocess.PIPE)
        cmd = [os.path.join(BIN, "paf_baseband2power"), "-a", kin, "-b", kout, "-c", str(tmp_path), "-d", gpus, "-n", str(2 * ndf_block)]
        if chunks:
            cmd += ["-g", chunks]
        stage = subprocess.Popen(cmd, stderr=subprocess.PIPE)
        time.sleep(0.3)
        run(os.path.join(BIN, "paf_diskdb"), "-a", kin, "-b", str(tmp_path), "-c", "in.dada", "-d", HDR, "-e", "1")
        assert stage.wait(timeout=120) == 0, stage.stderr.read().decode()
        assert sink.wait(timeout=60) == 0
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kin)
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", kout)
    spectra = np.frombuffer((tmp_path / "s.dada").read_bytes()[4096:], dtype=np.float32).reshape(-1, 336)
    assert spectra.shape[0] == 2
    payload = np.fromfile(src, dtype=np.uint8)[4096:]
    per = 2 * ndf_block * FRAME
    for i in range(2):
        want = oracle_mod.finish(oracle_mod.accumulate_omp(payload[i * per:(i + 1) * per]))
        assert np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)), i
    log = (tmp_path / "paf_baseband2power.log").read_text()
    assert "channel groups" in log
    if not chunks:
        assert "host link" in log                     # the split came from the link probe


def test_ring_is_reusable_for_a_second_observation(tmp_path):
    """Rings persist across runs in the reference's design (dada_db -p, paf-baseband2power.py:114):
    after end-of-data has been consumed, a new writer/reader pair starts a fresh observation."""
    key = "%x" % _key()
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(2 * FRAME), "-n", "3")
    try:
        for obs, (nfr, seed) in enumerate([(5, 1), (4, 2)]):
            src, name = tmp_path / f"in{obs}.dada", f"out{obs}.dada"
            run(os.path.join(BIN, "b2p_gen"), "-o", str(src), "-n", str(nfr), "-s", str(seed), "-H", HDR)
            reader = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", name, "-W"],
                                      stderr=subprocess.PIPE)
            time.sleep(0.2)
            run(os.path.join(BIN, "paf_diskdb"), "-a", key, "-b", str(tmp_path), "-c", src.name, "-d", HDR, "-e", "1")
            assert reader.wait(timeout=60) == 0
            assert src.read_bytes()[4096:] == (tmp_path / name).read_bytes()[4096:]
    finally:
        run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)


def test_waiters_notice_a_destroyed_ring(tmp_path):
    """A reader blocked on an empty ring must end when the ring is destroyed under it
    (no orphan processes after a crashed or torn-down pipeline)."""
    key = "%x" % _key()
    run(os.path.join(BIN, "paf_dada_db"), "-k", key, "-b", str(FRAME), "-n", "2")
    reader = subprocess.Popen([os.path.join(BIN, "paf_dbdisk"), "-k", key, "-D", str(tmp_path), "-f", "o.dada", "-W"],
                              stderr=subprocess.PIPE)
    time.sleep(0.3)
    assert reader.poll() is None                      # blocked waiting for a header
    run(os.path.join(BIN, "paf_dada_db"), "-d", "-k", key)
    assert reader.wait(timeout=10) != 0               # woke up, reported "no header", exited


def test_ascii_header_property_roundtrip(dada):
    """Random set/get sequences: the last value set for a key is the value read, other keys keep theirs."""
    from hypothesis import given, settings, strategies as st

    keys = st.sampled_from(["UTC_START", "FREQ", "BW", "NCHAN", "TSAMP", "SOURCE", "XKEY", "Y2", "BYTES_PER_SECOND"])
    vals = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFXYZ0123456789.-:+_", min_size=1, max_size=40)

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(keys, vals), min_size=1, max_size=25))
    def check(ops):
        buf = ctypes.create_string_buffer(open(HDR, "rb").read(), 4096)
        model = {}
        for k, v in ops:
            assert dada.ascii_header_set(buf, k.encode(), b"%s", v.encode()) == 0
            model[k] = v
        out = ctypes.create_string_buffer(128)
        for k, v in model.items():
            assert dada.ascii_header_get(buf, k.encode(), b"%127s", out) == 1
            assert out.value.decode() == v
        assert dada.ascii_header_get(buf, b"INSTRUMENT", b"%127s", out) == 1 and out.value == b"PAF-BMF"
        assert len(buf.value) < 4096

    check()
