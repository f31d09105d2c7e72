"""Row a6 pinned to the reference: host/paf_baseband2power.c against the reference's own
`main` (paf_baseband2power.cu:32-93), compiled unmodified into oracle/_ref/ref_paf_baseband2power
(oracle/Makefile: g++ -x c++ on the file where it lies, this repo's PSRDADA-named shim, libcudart).

The reference main parses -a/-b/-c/-d/-h, opens <dir>/paf_baseband2power.log for appending,
logs START PAF_PROCESS, asks for the device count and returns.  Everything it does before the
return is compared here on the same argument lists: exit codes, the error texts, the usage text,
the log file's name, mode and first line.  Where the two differ by design the test says so.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "paf_baseband2power_b200", "bin", "paf_baseband2power")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_paf_baseband2power")

pytestmark = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)),
                                reason="oracle/_ref/ref_paf_baseband2power not built (needs /root/reference at build time)")


def _run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("args", [("-a", "zzzz"), ("-a", "dada", "-b", "qq"), ("-b", "xyz", "-a", "dada")])
def test_unparsable_key_same_exit_code_and_message(args):
    r, o = _run(REF, *args), _run(OURS, *args)
    assert r.returncode == o.returncode == 1                      # EXIT_FAILURE, paf_baseband2power.cu:52,60
    pat = r"^Could not parse key from (\S+), which happens at \".*\", line \[\d+\]\.$"
    mr, mo = re.match(pat, r.stderr.strip()), re.match(pat, o.stderr.strip())
    assert mr and mo and mr.group(1) == mo.group(1)
    assert r.stdout == o.stdout == ""


def test_help_same_exit_code_and_reference_text_first():
    r, o = _run(REF, "-h", "x"), _run(OURS, "-h", "x")           # the reference declares "h:" (:40)
    assert r.returncode == o.returncode == 1                      # usage() then EXIT_FAILURE (:46-47)
    ref_lines = r.stdout.splitlines()
    assert len(ref_lines) == 8
    assert o.stdout.splitlines()[:8] == ref_lines                 # the reference's text, verbatim, first
    assert _run(OURS, "-h").returncode == 1                       # and without the stray argument too


def test_log_directory_that_does_not_exist(tmp_path):
    bad = str(tmp_path / "nope")
    args = ("-a", "dada", "-b", "adad", "-c", bad, "-d", "0")
    r, o = _run(REF, *args), _run(OURS, *args)
    assert r.returncode == o.returncode == 1
    want = f"Can not open log file {bad}/paf_baseband2power.log"  # paf_baseband2power.cu:79
    assert r.stderr.strip() == want
    assert o.stderr.strip().splitlines()[0] == want


def test_log_file_name_append_mode_and_first_line(tmp_path):
    """Same file name, opened "ab+" (a second run appends), same first line.  Exit codes after that
    point differ by design: the reference returns right after cudaGetDeviceCount — EXIT_SUCCESS
    with a GPU, exit(-1) from CudaSafeCall without (cudautil.cuh:29-41) — while this repo's main
    goes on to connect to the rings and fails (EXIT_FAILURE) because none exist here."""
    dr, do = tmp_path / "ref", tmp_path / "ours"
    dr.mkdir()
    do.mkdir()
    for _ in range(2):
        r = _run(REF, "-a", "dada", "-b", "adad", "-c", str(dr), "-d", "0")
        o = _run(OURS, "-a", "7e57", "-b", "7e59", "-c", str(do), "-d", "0")
        assert r.returncode in (0, 255)
        assert o.returncode == 1
    strip = lambda s: re.sub(r"^\[[^\]]*\]\s*", "", s)
    lr = (dr / "paf_baseband2power.log").read_text().splitlines()
    lo = (do / "paf_baseband2power.log").read_text().splitlines()
    assert [strip(x) for x in lr] == ["paf_baseband2power INFO: START PAF_PROCESS"] * 2 or \
        all("START PAF_PROCESS" in x for x in lr) and len(lr) == 2
    starts = [strip(x) for x in lo if "START PAF_PROCESS" in x]
    assert len(starts) == 2                                      # appended, not truncated
    assert strip(lo[0]) == strip(lr[0])                          # first line identical but for the time stamp
    assert sorted(os.listdir(dr)) == ["paf_baseband2power.log"]
    assert "paf_baseband2power.log" in os.listdir(do)


def test_missing_rings_fail_loudly_not_silently(tmp_path):
    o = _run(OURS, "-a", "7e57", "-b", "7e59", "-c", str(tmp_path), "-d", "0")
    assert o.returncode == 1
    assert "Can not connect to input hdu 7e57" in o.stderr
