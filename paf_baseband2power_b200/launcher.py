#!/usr/bin/env python3
"""Pipeline launcher — a working Python-3 restatement of paf-baseband2power.py.

The reference launcher cannot run (Python 2, a syntax error at
paf-baseband2power.py:90,92, undefined args.psrname/args.dfname at :37-38).
What it set out to do is kept: read paf-baseband2power.conf (:49-80), size the
two rings (input NDF*NCHK_NIC*7168 :67, output NCHAN*NBYTE :79), write the key
files (:101-112), create the rings (`dada_db -l -p -k -b -n -r`, :114-115), run
paf_diskdb, paf_baseband2power and the ring-to-disk writer pinned to cores 0, 1,
2 (:68,80,83,85-95), wait for the three, destroy the rings (:129-130).

Same flags (-a conf, -b directory, -c gpu, -d visible gpu, -e memcheck) plus the
data-file name the reference forgot to declare (-f).  Multi-beam: --beams N runs
N independent pipelines (own ring keys per beam, beam b on GPU gpus[b % len]) —
the sharding of DESIGN.md §7; no process talks to another beam's process.
--gpus-per-beam K gives every beam K GPUs of the -c list: its stage then runs
`paf_baseband2power -d g0,g1,...` and spreads the beam's chunks over them as channel
groups (one ring pair, one stage process, every one of the K host links in use).
"""
from __future__ import annotations

import argparse
import configparser
import os
import shutil
import subprocess
import sys
import threading
from dataclasses import dataclass, field
from typing import List

from .sharding import gpu_for_rank, ring_keys_for_beam

PKG = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(PKG, "bin")
PKT_BYTES = 7168  # DT_SIZE, capture.h:28 — the literal the reference launcher uses (:67)


@dataclass
class PipelineConf:
    nsamp_df: int
    npol_samp: int
    ndim_pol: int
    nchk_nic: int
    diskdb_ndf: int
    diskdb_nbuf: int
    diskdb_key: int
    diskdb_kfname: str
    diskdb_hfname: str
    diskdb_nreader: int
    diskdb_sod: int
    b2p_key: int
    b2p_kfname: str
    b2p_nreader: int
    b2p_sod: int
    b2p_nbuf: int
    b2p_nchan: int
    b2p_nbyte: int

    @property
    def diskdb_rbufsz(self) -> int:
        return self.diskdb_ndf * self.nchk_nic * PKT_BYTES

    @property
    def b2p_rbufsz(self) -> int:
        return self.b2p_nchan * self.b2p_nbyte


def read_conf(path: str) -> PipelineConf:
    cp = configparser.ConfigParser()
    if not cp.read(path):
        raise FileNotFoundError(path)
    b, d, p = cp["BasicConf"], cp["DiskdbConf"], cp["Baseband2powerConf"]
    return PipelineConf(
        nsamp_df=int(b["nsamp_df"]), npol_samp=int(b["npol_samp"]), ndim_pol=int(b["ndim_pol"]),
        nchk_nic=int(b["nchk_nic"]),
        diskdb_ndf=int(d["ndf"]), diskdb_nbuf=int(d["nblk"]), diskdb_key=int(d["key"], 16),
        diskdb_kfname=d["kfname_prefix"] + ".key", diskdb_hfname=d["hfname"],
        diskdb_nreader=int(d["nreader"]), diskdb_sod=int(d["sod"]),
        b2p_key=int(p["key"], 16), b2p_kfname=p["kfname_prefix"] + ".key",
        b2p_nreader=int(p["nreader"]), b2p_sod=int(p["sod"]), b2p_nbuf=int(p["nblk"]),
        b2p_nchan=int(p["nchan"]), b2p_nbyte=int(p["nbyte"]))


@dataclass
class BeamPlan:
    beam: int
    gpu: int
    key_in: int
    key_out: int
    create: List[List[str]] = field(default_factory=list)
    stages: List[List[str]] = field(default_factory=list)
    destroy: List[List[str]] = field(default_factory=list)


def _pin(cmd: List[str], cpu: int | None) -> List[str]:
    if cpu is None or shutil.which("taskset") is None or cpu >= (os.cpu_count() or 1):
        return cmd
    return ["taskset", "-c", str(cpu)] + cmd


def plan(conf: PipelineConf, directory: str, dfnames: List[str], gpus: List[int], memcheck: bool = False,
         hfname: str | None = None, extra_stage_args: List[str] | None = None, pin: bool = True,
         ngpus_box: int = 0, gpus_per_beam: int = 1) -> List[BeamPlan]:
    """The exact commands for every beam; nothing is executed here.  `gpus` empty: beams are
    spread over the `ngpus_box` GPUs of the box (sharding.gpu_for_rank).  `gpus_per_beam` K > 1:
    beam b gets GPUs gpus[b*K : b*K+K] (wrapping around the list) as channel groups."""
    if gpus_per_beam > 1 and len(gpus) < gpus_per_beam:
        raise ValueError(f"--gpus-per-beam {gpus_per_beam} needs at least that many GPUs in -c")
    plans = []
    hf = hfname or conf.diskdb_hfname
    if not os.path.isabs(hf):
        cand = os.path.join(directory, hf)
        hf = cand if os.path.exists(cand) else os.path.join(PKG, "conf", os.path.basename(hf))
    for beam, dfname in enumerate(dfnames):
        if len(dfnames) == 1:
            kin, kout = conf.diskdb_key, conf.b2p_key
        else:
            kin, kout = ring_keys_for_beam(beam, conf.diskdb_key, conf.b2p_key)
        gpu = gpus[beam % len(gpus)] if gpus else gpu_for_rank(beam, len(dfnames), ngpus_box, "spread")
        bp = BeamPlan(beam, gpu, kin, kout)
        db = os.path.join(BIN, "paf_dada_db")
        bp.create = [
            [db, "-l", "-p", "-k", f"{kin:x}", "-b", str(conf.diskdb_rbufsz), "-n", str(conf.diskdb_nbuf), "-r", str(conf.diskdb_nreader)],
            [db, "-l", "-p", "-k", f"{kout:x}", "-b", str(conf.b2p_rbufsz), "-n", str(conf.b2p_nbuf), "-r", str(conf.b2p_nreader)],
        ]
        bp.destroy = [[db, "-d", "-k", f"{kin:x}"], [db, "-d", "-k", f"{kout:x}"]]
        cpu0 = 3 * beam if pin else None
        dev = str(gpu)
        if gpus_per_beam > 1:
            mine = [gpus[(beam * gpus_per_beam + i) % len(gpus)] for i in range(gpus_per_beam)]
            bp.gpu, dev = mine[0], ",".join(str(x) for x in mine)
        stage = [os.path.join(BIN, "paf_baseband2power"), "-a", f"{kin:x}", "-b", f"{kout:x}", "-c", directory, "-d", dev]
        stage += extra_stage_args or []
        if memcheck:  # the reference wrapped the stage in cuda-memcheck (:89-90); its successor:
            stage = ["compute-sanitizer", "--tool", "memcheck"] + stage
        out_name = f"beam{beam:02d}_spectra.dada"
        bp.stages = [
            _pin([os.path.join(BIN, "paf_diskdb"), "-a", f"{kin:x}", "-b", directory, "-c", dfname, "-d", hf, "-e", str(conf.diskdb_sod)], cpu0),
            _pin(stage, None if cpu0 is None else cpu0 + 1),
            _pin([os.path.join(BIN, "paf_dbdisk"), "-k", f"{kout:x}", "-D", directory, "-f", out_name, "-W"], None if cpu0 is None else cpu0 + 2),
        ]
        plans.append(bp)
    return plans


def write_key_files(conf: PipelineConf, directory: str, plans: List[BeamPlan]):
    for bp in plans:
        suffix = "" if len(plans) == 1 else f".beam{bp.beam:02d}"
        for fname, key in ((conf.diskdb_kfname, bp.key_in), (conf.b2p_kfname, bp.key_out)):
            with open(os.path.join(directory, fname + suffix), "w") as f:
                f.write("DADA INFO:\n")
                f.write(f"key {key:x}\n")


def run(plans: List[BeamPlan], timeout: float | None = None, poll_s: float = 0.2) -> int:
    """Create rings, run the three stages of every beam concurrently, destroy rings.

    A stage that dies (no GPU, size mismatch, ...) would leave its siblings blocked on a ring
    for ever — the producer in ipcbuf_get_next_write, the writer waiting for a header.  So the
    processes are watched together: the first non-zero exit (or the timeout) terminates what
    is still running of that beam's pipeline, and the rings are destroyed in any case."""
    import time
    rc = 0
    created: List[BeamPlan] = []
    procs: List[tuple] = []          # (beam index, Popen)
    try:
        for bp in plans:
            for cmd in bp.create:
                subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
            created.append(bp)
        for bi, bp in enumerate(plans):
            # consumer first, producer last: nobody blocks on a ring without a reader
            for cmd in (bp.stages[2], bp.stages[1], bp.stages[0]):
                procs.append((bi, subprocess.Popen(cmd)))
        deadline = None if timeout is None else time.monotonic() + timeout
        failed_beams = set()
        while any(p.poll() is None for _, p in procs):
            for bi, p in procs:
                code = p.poll()
                if code not in (None, 0) and bi not in failed_beams:
                    failed_beams.add(bi)
                    for bj, q in procs:          # its siblings can only hang now
                        if bj == bi and q.poll() is None:
                            q.terminate()
            if deadline is not None and time.monotonic() > deadline:
                for _, p in procs:
                    if p.poll() is None:
                        p.kill()
                rc = max(rc, 9)
                break
            time.sleep(poll_s)
        for _, p in procs:
            try:
                code = p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
                code = p.wait()
            rc = max(rc, abs(code))
    finally:
        for _, p in procs:
            if p.poll() is None:
                p.kill()
        for bp in created:
            for cmd in bp.destroy:
                subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return rc


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Detect PAF BMF baseband from DADA files and integrate it (B200)")
    ap.add_argument("-a", "--cfname", required=True, help="The name of configuration file")
    ap.add_argument("-b", "--directory", required=True, help="Directory with the data files; spectra and logs are written there")
    ap.add_argument("-c", "--gpu", type=int, nargs="+", default=[0], help="The index of GPU (several: beams round-robin)")
    ap.add_argument("-d", "--visiblegpu", default="", help="Visible GPU(s) inside a container; sets CUDA_VISIBLE_DEVICES unless '' or 'all'")
    ap.add_argument("-e", "--memcheck", type=int, default=0, help="Run the stage under compute-sanitizer memcheck")
    ap.add_argument("-f", "--dfname", nargs="+", required=True, help="DADA data file name(s), one per beam")
    ap.add_argument("--ndf", type=int, default=0, help="override NDF (frames per ring block) of the conf")
    ap.add_argument("--nblk", type=int, default=0, help="override NBLK (input ring blocks) of the conf")
    ap.add_argument("--average", type=int, default=0, help="1: time average instead of integral")
    ap.add_argument("--spread", type=int, default=0, metavar="NGPU",
                    help="ignore -c and spread the beams over the box's NGPU GPUs (beam i -> GPU floor(i*NGPU/nbeams)): "
                         "neighbouring GPUs tend to share a host bridge, and this path is bound by the host links")
    ap.add_argument("--gpus-per-beam", type=int, default=1, metavar="K",
                    help="give every beam K GPUs of the -c list: its stage spreads the beam's chunks over them as "
                         "channel groups (paf_baseband2power -d g0,g1,...), so one beam uses K host links")
    ap.add_argument("--dry-run", action="store_true", help="print the commands, run nothing")
    ap.add_argument("--timeout", type=float, default=None)
    args = ap.parse_args(argv)

    conf = read_conf(args.cfname)
    if args.ndf:
        conf.diskdb_ndf = args.ndf
    if args.nblk:
        conf.diskdb_nbuf = args.nblk
    if args.visiblegpu not in ("", "all"):
        os.environ["CUDA_VISIBLE_DEVICES"] = args.visiblegpu
    extra = ["-s", "1"] if args.average else []
    plans = plan(conf, args.directory, args.dfname, [] if args.spread else args.gpu, bool(args.memcheck),
                 extra_stage_args=extra, ngpus_box=args.spread, gpus_per_beam=args.gpus_per_beam)
    if args.dry_run:
        for bp in plans:
            for cmd in bp.create + bp.stages + bp.destroy:
                print(" ".join(cmd))
        return 0
    write_key_files(conf, args.directory, plans)
    return run(plans, args.timeout)


if __name__ == "__main__":
    sys.exit(main())
