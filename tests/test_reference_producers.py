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
