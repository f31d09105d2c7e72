"""The Python-3 launcher: conf parsing and command construction on CPU, a two-beam run on GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from paf_baseband2power_b200 import launcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "paf_baseband2power_b200")
CONF = os.path.join(PKG, "conf", "paf-baseband2power.conf")
BIN = os.path.join(PKG, "bin")


def test_conf_sizes_follow_the_reference_formulas():
    c = launcher.read_conf(CONF)
    assert c.diskdb_rbufsz == 8192 * 48 * 7168 == 2818572288      # paf-baseband2power.py:67
    assert c.b2p_rbufsz == 336 * 4                                 # paf-baseband2power.py:79
    assert (c.diskdb_key, c.b2p_key) == (0xDADA, 0xADAD)
    assert (c.diskdb_nbuf, c.b2p_nbuf, c.diskdb_nreader, c.diskdb_sod) == (8, 4, 1, 1)
    assert c.diskdb_kfname == "diskdb.key" and c.b2p_kfname == "baseband2power.key"


def test_single_beam_plan_uses_the_conf_keys(tmp_path):
    c = launcher.read_conf(CONF)
    (bp,) = launcher.plan(c, str(tmp_path), ["obs.dada"], [3], pin=False)
    assert bp.create[0][1:] == ["-l", "-p", "-k", "dada", "-b", "2818572288", "-n", "8", "-r", "1"]
    assert bp.create[1][1:] == ["-l", "-p", "-k", "adad", "-b", "1344", "-n", "4", "-r", "1"]
    diskdb, stage, dbdisk = bp.stages
    assert diskdb[1:9] == ["-a", "dada", "-b", str(tmp_path), "-c", "obs.dada", "-d", diskdb[8]]
    assert diskdb[8].endswith("header_baseband2power.txt") and diskdb[-2:] == ["-e", "1"]
    assert stage[1:] == ["-a", "dada", "-b", "adad", "-c", str(tmp_path), "-d", "3"]
    assert dbdisk[1:3] == ["-k", "adad"] and "-W" in dbdisk
    assert bp.destroy == [[bp.create[0][0], "-d", "-k", "dada"], [bp.create[0][0], "-d", "-k", "adad"]]


def test_multibeam_plan_shards_beams_over_gpus(tmp_path):
    c = launcher.read_conf(CONF)
    plans = launcher.plan(c, str(tmp_path), [f"b{i}.dada" for i in range(5)], [0, 1], pin=False)
    assert [p.gpu for p in plans] == [0, 1, 0, 1, 0]
    keys = [k for p in plans for k in (p.key_in, p.key_in + 1, p.key_out, p.key_out + 1)]
    assert len(set(keys)) == len(keys)
    mem = launcher.plan(c, str(tmp_path), ["a.dada"], [0], memcheck=True, pin=False)[0]
    assert mem.stages[1][:3] == ["compute-sanitizer", "--tool", "memcheck"]


def test_dry_run_cli(tmp_path):
    script = os.path.join(PKG, "scripts", "paf-baseband2power.py")
    r = subprocess.run([sys.executable, script, "-a", CONF, "-b", str(tmp_path), "-c", "0", "-d", "0", "-e", "0",
                        "-f", "x.dada", "--dry-run"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 7 and "paf_dada_db -l -p -k dada -b 2818572288 -n 8 -r 1" in lines[0]
    assert any("paf_baseband2power -a dada -b adad" in ln for ln in lines)


@pytest.mark.gpu
def test_two_beam_pipeline_through_the_launcher(tmp_path, oracle_mod, b2p):
    ndf_block, nblk = 32, 3
    hdr = os.path.join(PKG, "conf", "header_baseband2power.txt")
    if not os.path.exists(os.path.join(BIN, "b2p_gen")):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG, "host"), "all"], check=True)
    names = []
    for b in range(2):
        names.append(f"beam{b}.dada")
        subprocess.run([os.path.join(BIN, "b2p_gen"), "-o", str(tmp_path / names[-1]), "-n", str(ndf_block * nblk),
                        "-s", str(40 + b), "-H", hdr], check=True, capture_output=True)
    gpus = ["0", "1"] if b2p.device_count() >= 2 else ["0"]     # beams round-robin over the GPUs present
    rc = launcher.main(["-a", CONF, "-b", str(tmp_path), "-c", *gpus, "-d", "", "-e", "0", "-f", *names,
                        "--ndf", str(ndf_block), "--nblk", "3", "--timeout", "120"])
    assert rc == 0
    assert (tmp_path / "diskdb.key.beam00").read_text().startswith("DADA INFO:\nkey ")
    for b in range(2):
        out = (tmp_path / f"beam{b:02d}_spectra.dada").read_bytes()
        spectra = np.frombuffer(out[4096:], dtype=np.float32).reshape(nblk, 336)
        payload = np.fromfile(tmp_path / names[b], dtype=np.uint8)[4096:]
        per = ndf_block * 48 * 7168
        for i in range(nblk):
            want = oracle_mod.finish(oracle_mod.accumulate_omp(payload[i * per:(i + 1) * per]))
            assert np.array_equal(spectra[i].view(np.uint32), want.view(np.uint32)), (b, i)


def test_launcher_does_not_hang_when_the_stage_dies(tmp_path):
    """paf_baseband2power exits at once (no usable GPU: CUDA_VISIBLE_DEVICES=-1 via -d);
    paf_diskdb would then block for ever on a ring nobody reads and paf_dbdisk on a header that
    never comes.  The launcher must notice, stop the siblings, destroy the rings, return non-zero."""
    import time
    ndf_block = 4
    hdr = os.path.join(PKG, "conf", "header_baseband2power.txt")
    subprocess.run([os.path.join(BIN, "b2p_gen"), "-o", str(tmp_path / "x.dada"), "-n", str(ndf_block * 12),
                    "-s", "1", "-H", hdr], check=True, capture_output=True)
    script = os.path.join(PKG, "scripts", "paf-baseband2power.py")
    t0 = time.monotonic()
    r = subprocess.run([sys.executable, script, "-a", CONF, "-b", str(tmp_path), "-c", "0", "-d", "-1", "-e", "0",
                        "-f", "x.dada", "--ndf", str(ndf_block), "--nblk", "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert time.monotonic() - t0 < 60
    # the rings are gone: connecting to the input key fails
    chk = subprocess.run([os.path.join(BIN, "paf_diskdb"), "-a", "dada", "-b", str(tmp_path), "-c", "x.dada", "-d", hdr],
                         capture_output=True, text=True, timeout=30)
    assert chk.returncode != 0 and "Can not connect to hdu" in chk.stderr


def test_spread_plan_places_beams_across_the_box(tmp_path):
    c = launcher.read_conf(CONF)
    plans = launcher.plan(c, str(tmp_path), [f"b{i}.dada" for i in range(4)], [], pin=False, ngpus_box=8)
    assert [p.gpu for p in plans] == [0, 2, 4, 6]


def test_gpus_per_beam_plan_gives_the_stage_a_gpu_list(tmp_path):
    """--gpus-per-beam K: beam b's stage gets `-d g0,...` (channel groups over K GPUs, one ring pair)."""
    c = launcher.read_conf(CONF)
    plans = launcher.plan(c, str(tmp_path), ["a.dada", "b.dada"], [0, 2, 4, 6], pin=False, gpus_per_beam=2)
    assert [p.stages[1][p.stages[1].index("-d") + 1] for p in plans] == ["0,2", "4,6"]
    assert [p.gpu for p in plans] == [0, 4]
    (one,) = launcher.plan(c, str(tmp_path), ["a.dada"], [0, 1, 2, 3, 4, 5, 6, 7], pin=False, gpus_per_beam=8)
    assert one.stages[1][-1] == "0,1,2,3,4,5,6,7" and (one.key_in, one.key_out) == (0xDADA, 0xADAD)
    with pytest.raises(ValueError):
        launcher.plan(c, str(tmp_path), ["a.dada"], [0], pin=False, gpus_per_beam=2)

