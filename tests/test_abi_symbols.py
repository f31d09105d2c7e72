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
    assert _lib.load().b2p_version() == b"0.1.0"


def test_params_struct_matches_header_defaults():
    from paf_baseband2power_b200 import _lib
    lib = _lib.load()
    p = _lib.B2pParams()
    lib.b2p_default_params(ctypes.byref(p))
    assert (p.nchunk, p.nch_per_chunk, p.nsamp_df, p.big_endian) == (48, 7, 128, 1)
    assert (p.scale, p.mode, p.nbeam, p.kernel, p.nsplit) == (1.0, 0, 1, 0, 0)
    assert (p.stage_ndf, p.nstage_bufs, p.device_id) == (0, 0, 0)


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
