"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2p.h declares.

No compute call is made here (there is no GPU); compute entry points must fail
loudly instead of falling back.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "b2p.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(b2p_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_what_the_binding_binds():
    from paf_baseband2power_b200 import _lib
    declared = _declared_functions()
    assert declared, "no functions parsed from include/b2p.h"
    assert sorted(_lib.SYMBOLS) == declared


def test_library_exports_every_declared_symbol():
    from paf_baseband2power_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libb2p.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in b2p.h but not exported"
    hdr = open(os.path.join(ROOT, "include", "b2p.h")).read()
    assert _lib.load().b2p_version() == re.search(r'#define B2P_VERSION "([^"]+)"', hdr).group(1).encode()


def test_params_struct_matches_header_defaults():
    from paf_baseband2power_b200 import _lib
    lib = _lib.load()
    p = _lib.B2pParams()
    lib.b2p_default_params(ctypes.byref(p))
    assert (p.nchunk, p.nch_per_chunk, p.nsamp_df, p.big_endian) == (48, 7, 128, 1)
    assert (p.scale, p.mode, p.nbeam, p.kernel, p.nsplit) == (1.0, 0, 1, 0, 0)
    assert (p.stage_ndf, p.nstage_bufs, p.device_id) == (0, 0, 0)
    assert (p.first_chunk, p.nchunk_total, p.resizable) == (0, 0, 0)
    # the ctypes mirror and the C struct must agree on the size (a drifted field would shift all)
    assert ctypes.sizeof(_lib.B2pParams) == 64


def test_no_batched_memcpy_names_in_the_artefact():
    """the shipped library is linked against the shared CUDA runtime, so it carries
    no runtime symbol names beyond the calls this project makes."""
    from paf_baseband2power_b200 import _lib
    blob = open(_lib.LIB_PATH, "rb").read()
    for name in (b"MemcpyBatchAsync", b"Memcpy3DBatchAsync"):
        assert name not in blob


def test_no_cpu_fallback_without_gpu():
    from paf_baseband2power_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.B2pError) as ei:
        api.Baseband2Power()
    assert ei.value.code == 2  # B2P_ECUDA


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "paf_baseband2power_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "b2p_oracle" not in text and "import oracle" not in text, f


def test_stage_library_exports_the_stage_driver():
    """include/baseband2power.h: init_/do_/destroy_baseband2power (+ default_) from libb2p_stage.so,
    together with the PSRDADA-named shim entry points the stage and the producers use."""
    path = os.path.join(ROOT, "paf_baseband2power_b200", "libb2p_stage.so")
    assert os.path.exists(path), "libb2p_stage.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    src = open(os.path.join(ROOT, "include", "baseband2power.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b([a-z_]*baseband2power)\s*\(", src)))
    assert names == ["default_baseband2power", "destroy_baseband2power", "do_baseband2power", "init_baseband2power"]
    for n in names:
        assert hasattr(lib, n), n
    for n in ["dada_hdu_create", "dada_hdu_set_key", "dada_hdu_connect", "dada_hdu_lock_read", "dada_hdu_lock_write",
              "dada_hdu_unlock_read", "dada_hdu_unlock_write", "dada_hdu_disconnect", "dada_hdu_destroy",
              "ipcbuf_get_bufsz", "ipcbuf_enable_sod", "ipcbuf_disable_sod", "ipcbuf_get_next_write", "ipcbuf_mark_filled",
              "ipcbuf_get_next_read", "ipcbuf_mark_cleared", "ipcbuf_eod", "ipcio_open_block_write", "ipcio_close_block_write",
              "ipcio_open_block_read", "ipcio_close_block_read", "ascii_header_set", "ascii_header_get", "fileread",
              "multilog_open", "multilog_add", "multilog", "multilog_close", "dada_cuda_dbregister", "dada_cuda_dbunregister"]:
        assert hasattr(lib, n), n   # the PSRDADA calls the reference uses + the reader side


def test_stage_defaults_match_the_conf():
    lib = ctypes.CDLL(os.path.join(ROOT, "paf_baseband2power_b200", "libb2p_stage.so"))

    class Conf(ctypes.Structure):   # leading fields of conf_t, include/baseband2power.h
        _fields_ = [("device_id", ctypes.c_int), ("dir", ctypes.c_char * 512), ("key_in", ctypes.c_int),
                    ("key_out", ctypes.c_int), ("nchunk", ctypes.c_int), ("nch_per_chunk", ctypes.c_int),
                    ("nsamp_df", ctypes.c_int), ("big_endian", ctypes.c_int), ("average", ctypes.c_int),
                    ("ndf_integration", ctypes.c_uint64), ("kernel", ctypes.c_int), ("pin_ring", ctypes.c_int),
                    ("rest", ctypes.c_char * 256)]
    c = Conf()
    lib.default_baseband2power(ctypes.byref(c))
    assert (c.key_in, c.key_out) == (0xDADA, 0xADAD)          # paf-baseband2power.conf:13,20
    assert (c.nchunk, c.nch_per_chunk, c.nsamp_df, c.big_endian, c.average, c.pin_ring) == (48, 7, 128, 1, 0, 1)
    # connecting to rings that do not exist fails with EXIT_FAILURE, it does not crash
    c.key_in, c.key_out = 0x7E57, 0x7E59
    assert lib.init_baseband2power(ctypes.byref(c)) == 1
    lib.destroy_baseband2power(ctypes.byref(c))
