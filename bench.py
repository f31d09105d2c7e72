#!/usr/bin/env python
"""bench.py — input GB/s and real-time factor of the baseband->power hot path on B200.

One step = one beam-integration per beam stream: one input ring block of
8192 data frames (2 818 572 288 B, paf-baseband2power.conf:9 /
paf-baseband2power.py:67) unpacked, detected and integrated into 336 float32
(1344 B).  Default workload = BASELINE.json configs[1]: a single beam, a
continuous stream of integrations.

  value     kernel-only: blocks already resident in HBM (4 rotating blocks, each
            22x larger than L2), K steps on one stream, ONE kernel launch per step
            (b2p_integrate_device, PDL-chained), CUDA events, max over ranks.
  e2e       the same through the C ABI with HOST buffers: b2p_integrate_host
            (pinned ring block -> chunked H2D overlapped with the kernels -> D2H of
            the spectrum) inside the timed region.  N > 1 also measures the
            channel-group mode: every GPU takes a link-weighted share of the chunks
            of every beam (strided H2D), so unequal host links finish together.
  roofline  fused kernel only, per-launch duration from CUDA events recorded on
            the launching stream inside the timed region, against the measured
            HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU oracle (a port of the specification; the reference has no
            kernel to time, kernel.cu:1-7) on all host cores, bounded sample —
            the very routine `--impl reference` runs.
  ring_e2e / live_replay  BASELINE.json configs[3]/[4]: the executables through
            SysV rings (and loopback UDP), one pipeline per GPU on every rank.

`--impl reference` times that CPU port alone (rank 0 only), same metric/config.
Multi-GPU: one process per GPU (torchrun), beams shard by rank, no collective on
the data path; only the spectra are gathered.  Ranks are spread over the box's
GPUs (rank i -> GPU floor(i*D/N)) so that a partial run does not crowd one host
bridge; the map is printed in `placement`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "baseband_input_throughput"
UNIT = "GB/s"
BLOCK_NDF = 8192
CPU_NOTE = ("reference kernel absent (kernel.cu:1-7): C port of the specification, gcc -O3 -march=native "
            "-fopenmp, frames handed to threads dynamically")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def _ncu_traffic():
    """dram bytes per fused launch from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report that, do not invent clocks
            self.err = repr(e)

    _NAMES = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
        0x100: "display_clock_setting", 0x10: "sync_boost",
    }

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": round(max(self.power), 1) if self.power else None}


# --------------------------------------------------------------------------- CPU port

def _cpu_port(native=True):
    """The CPU oracle built for this host (the only place bench.py executes oracle/)."""
    import oracle
    L = None
    if native:
        try:
            L = oracle.lib(oracle.build(native=True))
        except Exception:
            L = None
    return oracle, (L or oracle.lib())


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_measure(ndf: int, min_seconds: float = 4.0, min_passes: int = 50, max_seconds: float = 25.0):
    """Time the CPU port on one beam-integration held in host RAM.  ONE routine for both the
    `--impl reference` arm and the in-line `cpu_baseline`, so the two cannot drift apart: same
    allocation (anonymous mmap, transparent huge pages requested, first touched by all
    threads), same thread count (all the process may use — explicit, torchrun exports
    OMP_NUM_THREADS=1), same policy: 3 untimed passes, then whole passes until >= min_seconds
    AND >= min_passes (capped at max_seconds); median and best are both reported."""
    import mmap
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    oracle, L = _cpu_port()
    g = oracle.Geometry()
    nbytes = ndf * g.frame_bytes
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except Exception:
        pass
    block = np.frombuffer(mm, dtype=np.uint8)
    # a short synthetic piece (Gaussian, sigma 512) tiled over the block: the arithmetic has no
    # data-dependent branch, so the content does not change the rate; the piece is 64 frames
    piece = oracle.synth_fill(min(64, ndf), seed=1, mode=1)
    threads = _host_threads()
    n = piece.nbytes

    def fill(i):
        m = min(n, nbytes - i * n)
        block[i * n:i * n + m] = piece[:m]

    with ThreadPoolExecutor(max_workers=threads) as ex:   # first touch from many threads
        list(ex.map(fill, range((nbytes + n - 1) // n)))
    sums = np.zeros(g.nchan, dtype=np.uint64)
    for _ in range(3):
        oracle.accumulate_omp(block, ndf, g, sums=sums, nthreads=threads, L=L)
    times = []
    t_begin = time.perf_counter()
    while True:
        sums[:] = 0
        t0 = time.perf_counter()
        oracle.accumulate_omp(block, ndf, g, sums=sums, nthreads=threads, L=L)
        oracle.finish(sums, 1.0)
        times.append(time.perf_counter() - t0)
        el = time.perf_counter() - t_begin
        if (el >= min_seconds and len(times) >= min_passes) or el >= max_seconds:
            break
    n1 = min(ndf, 1024)                           # one single-thread pass over 1/8 block, for the per-core figure
    t0 = time.perf_counter()
    oracle.accumulate(block[: n1 * g.frame_bytes], n1, g, L=L)
    t_single = (time.perf_counter() - t0) * (ndf / n1)
    med, best = statistics.median(times), min(times)
    t_int = ndf * 128 * 27.0 / 32.0 * 1e-6
    del block
    try:
        mm.close()
    except BufferError:
        pass
    return {
        "value": round(nbytes / med / 1e9, 3), "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"{len(times)} passes over 1 beam-integration of {ndf} frames ({nbytes} B) in host RAM "
                   f"(THP-advised anonymous memory) in {round(sum(times), 1)} s; median; "
                   f"best {round(nbytes / best / 1e9, 3)} GB/s"),
        "best": round(nbytes / best / 1e9, 3), "passes": len(times),
        "ms_per_pass_median": round(med * 1e3, 3),
        "realtime_factor": round(t_int / med, 3),
        "single_thread_GBps": round(nbytes / t_single / 1e9, 3),
        "note": CPU_NOTE,
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference
    never wrote one (kernel.cu:1-7), so this is the oracle port on all host threads."""
    if rank != 0:
        return
    # each "step" is one pass; the pass count comes from the time policy, not from --steps,
    # so that the figure equals the b200 arm's in-line cpu_baseline on the same box
    cpu = cpu_port_measure(args.ndf)
    emit({
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": cpu["passes"], "warmup": 3,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": cpu["ms_per_pass_median"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": _config(args, 1),
        "realtime_factor": cpu["realtime_factor"],
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def _config(args, nbeam):
    """Identical in both arms (the driver compares the dicts); run-specific facts — resolved
    kernel, split count, GPU placement — live in top-level keys of the b200 line."""
    return {
        "workload": ("BASELINE.json configs[1]: single beam, continuous stream of integrations, "
                     "1 input ring block (8192 frames x 48 chunks x 7168 B = 2818572288 B) -> 336 x float32 per step"
                     if nbeam == 1 else
                     f"BASELINE.json configs[2]: {nbeam} beams batched per step, each 1 ring block of 2818572288 B"),
        "nbeam_per_gpu": nbeam, "ndf": args.ndf, "nchunk": 48, "nch_per_chunk": 7, "nsamp_df": 128,
        "mode": "exact-uint64",
        "l2": "inputs larger than L2: up to 4 rotating 2.8 GB blocks per beam (value), 2 rotating pinned blocks (e2e)",
        "parallelism": f"beams sharded by rank, {args.gpus} x independent, no collective on the data path",
    }


_REAL_STDOUT = None


def _guard_stdout():
    """Only the one JSON line may reach stdout: libraries (NCCL prints its version there)
    are pointed at stderr for the run; emit() writes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def _mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- SysV shared blocks

class SysVBlock:
    """A host block in a SysV shared-memory segment, so that every rank (one process per GPU)
    can read every beam's ring block — what a PSRDADA ring is (shmget/shmat), minus the ring."""
    IPC_CREAT, IPC_RMID = 0o1000, 0

    def __init__(self, key: int, nbytes: int, create: bool):
        libc = ctypes.CDLL(None, use_errno=True)
        libc.shmget.restype = ctypes.c_int
        libc.shmget.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        libc.shmat.restype = ctypes.c_void_p
        libc.shmat.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        libc.shmdt.argtypes = [ctypes.c_void_p]
        libc.shmctl.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self.libc, self.nbytes, self.owner = libc, nbytes, create
        self.id = libc.shmget(key, nbytes, (self.IPC_CREAT | 0o666) if create else 0o666)
        if self.id < 0:
            raise OSError(ctypes.get_errno(), f"shmget key {key:#x} ({nbytes} B)")
        self.ptr = libc.shmat(self.id, None, 0)
        if self.ptr in (None, ctypes.c_void_p(-1).value):
            raise OSError(ctypes.get_errno(), "shmat")

    def close(self):
        if self.ptr:
            self.libc.shmdt(ctypes.c_void_p(self.ptr))
            self.ptr = None
        if self.owner and self.id >= 0:
            self.libc.shmctl(self.id, self.IPC_RMID, None)
            self.id = -1


class Run:
    """What the legs share: rank layout, collectives, the loaded library."""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nbeam", type=int, default=1, help="beam streams per GPU per step")
    ap.add_argument("--ndf", type=int, default=BLOCK_NDF, help="data frames per ring block")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ldg", "tma"])
    ap.add_argument("--nsplit", type=int, default=0, help="time splits per chunk (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 32)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ring", action="store_true", help="skip the executables-through-the-rings leg")
    ap.add_argument("--no-live", action="store_true", help="skip the UDP live-replay leg")
    ap.add_argument("--ring-blocks", type=int, default=64, help="integrations streamed through the rings (configs[1]: 64)")
    ap.add_argument("--live-blocks", type=int, default=2, help="ring blocks of live UDP replay per beam")
    ap.add_argument("--beamset", type=int, default=36, help="extra kernel-only point: full beam set on one GPU (0 = skip)")
    ap.add_argument("--placement", default="spread", choices=["spread", "identity"],
                    help="rank -> GPU map when the box has more GPUs than ranks")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _guard_stdout()

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from paf_baseband2power_b200 import BMF, Baseband2Power, PinnedBuffer, device_info
    from paf_baseband2power_b200 import _lib
    from paf_baseband2power_b200 import api as b2p_api
    from paf_baseband2power_b200.sharding import beams_for_rank, gather_spectra, gpu_for_rank

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    ngpu_box = torch.cuda.device_count()
    env_map = os.environ.get("B2P_BENCH_GPUS")
    if env_map:
        gpu_map = [int(x) for x in env_map.split(",")][:world]
    else:
        gpu_map = [gpu_for_rank(r, world, ngpu_box, args.placement) for r in range(world)]
    gpu = gpu_map[local] if local < len(gpu_map) else local
    torch.cuda.set_device(gpu)
    host_pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", gpu))
        host_pg = dist.new_group(backend="gloo")   # spectra are gathered host-side, 1344 B per beam

    R = Run()
    R.args, R.rank, R.world, R.gpu, R.host_pg = args, rank, world, gpu, host_pg
    R.gpu_map = gpu_map
    R.vcpus = _host_threads()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_pg)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_gather_floats(x: float) -> list:
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def gather_objects(obj):
        if world == 1:
            return [obj]
        out = [None] * world if rank == 0 else None
        dist.gather_object(obj, out, dst=0, group=host_pg)
        return out

    R.barrier, R.host_barrier, R.max_over_ranks, R.gather_objects = barrier, host_barrier, max_over_ranks, gather_objects

    nbeam, ndf = args.nbeam, args.ndf
    g = BMF
    blk = ndf * g.frame_bytes
    R.g, R.ndf, R.blk = g, ndf, blk
    # rotate between distinct blocks so that no step can find its input in L2 (126 MB); one
    # set is already 22x L2 per beam, so fewer sets are used when many beams fill the HBM
    nrot = max(1, min(4, int(0.45 * torch.cuda.get_device_properties(gpu).total_memory // (nbeam * blk))))
    # ---- device-resident inputs: nrot distinct blocks per beam ----
    dev_in = torch.empty(nrot * nbeam * blk, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    R.lib, R.dev_in = lib, dev_in
    wpb = blk // 8
    for r in range(nrot):
        for b in range(nbeam):
            beam_id = rank * nbeam + b
            rc = lib.b2p_synth_fill_device(gpu, dev_in.data_ptr() + (r * nbeam + b) * blk, ndf, 48, 7, 128, 1,
                                           1000 + beam_id, r * wpb, 1, None)
            assert rc == 0
    out_dev = torch.empty(nbeam * g.nchan, dtype=torch.float32, device="cuda")
    st = Baseband2Power(device_id=gpu, nbeam=nbeam, kernel=args.kernel, nsplit=args.nsplit)
    # The kernels run on the context's own stream (stream=None through the C ABI): only
    # there may a fused kernel start under the tail of its predecessor (PDL).  torch wraps
    # that same stream so the torch.cuda.Event pair below is recorded on it.
    xstream = torch.cuda.ExternalStream(st.stream, device=torch.device("cuda", gpu))
    ptrs = [[dev_in.data_ptr() + (r * nbeam + b) * blk for b in range(nbeam)] for r in range(nrot)]
    torch.cuda.synchronize()

    def step(i):
        st.integrate_device(ptrs[i % nrot], ndf, out_dev, None)   # ONE launch per integration

    sampler = ClockSampler(gpu)
    for i in range(args.warmup):
        step(i)
    barrier()
    l0 = st.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record(xstream)
    for i in range(args.steps):
        step(i)
    e1.record(xstream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = st.launch_count - l0
    ms_step = ms_total / args.steps
    value = world * nbeam * blk / (ms_step * 1e-3) / 1e9

    # the fused kernel alone: one CUDA-event pair per launch on the launching stream
    # (the events serialise neighbouring launches, so this is the isolated duration)
    st.set_timing(True)
    for i in range(args.steps):
        step(i)
    barrier()
    fused_ms, fused_n = st.fused_time_ms()
    st.set_timing(False)

    # parity spot check of what the timed loop last produced (beam 0, last block)
    last_out = out_dev[: g.nchan].cpu().numpy()

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    e2e_last = None
    pinned = None
    ke = args.e2e_steps or min(args.steps, 32)
    R.ke = ke
    if not args.no_e2e:
        hrot = 2
        pinned = [[PinnedBuffer(blk) for _ in range(nbeam)] for _ in range(hrot)]
        for r in range(hrot):
            for b in range(nbeam):
                rc = lib.b2p_memcpy_d2h(gpu, pinned[r][b].ptr, dev_in.data_ptr() + (r * nbeam + b) * blk, blk)
                assert rc == 0
        my_beams = beams_for_rank(world * nbeam, rank, world)
        ste = Baseband2Power(device_id=gpu, nbeam=nbeam, kernel=args.kernel)

        def e2e_step(i):
            sp = ste.integrate_host(pinned[i % hrot], ndf)   # H2D pieces + kernels + D2H of the spectra
            if world > 1:                           # the only cross-rank traffic of the path
                allsp = gather_spectra(sp, my_beams, world * nbeam, group=host_pg)
                return sp, allsp
            return sp, sp

        for i in range(3):
            spec, allspec = e2e_step(i)
        # plain pinned H2D with every rank copying at once: the link ceiling the e2e number sits under
        host_barrier()
        h2d_link = b2p_api.probe_h2d([gpu], nbytes=1 << 30, reps=3)[0]
        links = all_gather_floats(h2d_link)
        links_equal_bytes = None
        if world > 1:
            # Equal byte counts let the fast links finish first and the slow ones then run alone,
            # which reads too high.  Second pass: bytes in proportion to the first pass's rates, so
            # every link stays loaded to the end — what the links deliver TOGETHER.
            links_equal_bytes = links
            nb2 = max(1 << 20, int((1 << 30) * h2d_link / max(links)) & ~255)
            host_barrier()
            h2d_link = b2p_api.probe_h2d([gpu], nbytes=nb2, reps=3)[0]
            links = all_gather_floats(h2d_link)
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            spec, allspec = e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        e2e_ms = max_over_ranks(dt * 1e3) / ke
        by_beam = round(world * nbeam * blk / (e2e_ms * 1e-3) / 1e9, 3)
        e2e = {"value": by_beam, "unit": UNIT,
               "h2d_bytes_per_step": nbeam * blk, "d2h_bytes_per_step": nbeam * g.out_bytes,
               "ms_per_step": round(e2e_ms, 3), "steps": ke, "host_threads_per_gpu": 1,
               "mode": "by_beam",
               "realtime_factor": round(world * nbeam * g.t_integration_s * ndf / BLOCK_NDF / (e2e_ms * 1e-3), 2),
               "h2d_link_GBps_per_gpu": round(h2d_link, 2),
               "h2d_link_GBps_all_ranks": [round(x, 2) for x in links],
               "h2d_links_sum_GBps": round(sum(links), 2),
               "h2d_link_probe": ("all ranks copy at once, bytes per rank in proportion to a first pass's rates so that "
                                  "every link stays loaded to the end" if world > 1 else "3 x 1 GiB pinned cudaMemcpyAsync"),
               "h2d_link_GBps_equal_bytes_pass": [round(x, 2) for x in links_equal_bytes] if links_equal_bytes else None,
               "frac_of_h2d_link": round(nbeam * blk / (e2e_ms * 1e-3) / 1e9 / h2d_link, 4),
               "frac_of_h2d_links_sum": round(by_beam / sum(links), 4),
               "path": "b2p_integrate_host (pinned ring block, 256-frame pieces, 3 staging buffers, one launch per piece, spectrum D2H)"
                       + (" + gloo gather of spectra to rank 0" if world > 1 else "")}
        e2e_last = spec[0].copy()
        ste.close()

        # ---- N > 1: channel-group mode, link-weighted (unequal host links finish together) ----
        if world > 1 and nbeam == 1:
            e2e["by_beam"] = {"value": by_beam, "ms_per_step": round(e2e_ms, 3)}
            try:
                cg = _channel_group_e2e(R, links, allspec, (ke - 1) % hrot)
            except SystemExit:
                raise
            except Exception as e:   # the by-beam figure stands
                cg = {"error": repr(e)[:300]}
            e2e["by_channel_group"] = cg
            best = cg.get("all_beams") or {}
            if best.get("value", 0) > e2e["value"]:
                e2e.update({"value": best["value"], "ms_per_step": best["ms_per_step"], "steps": best["steps"],
                            "mode": "by_channel_group",
                            "realtime_factor": round(world * g.t_integration_s * ndf / BLOCK_NDF / (best["ms_per_step"] * 1e-3), 2),
                            "frac_of_h2d_links_sum": round(best["value"] / sum(links), 4),
                            "path": best["path"]})
    clocks = sampler.stop()

    # ---- CPU baseline + parity check against the oracle (rank 0, N=1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle, L = _cpu_port()
        og = oracle.Geometry()
        if args.no_e2e:
            host = np.empty(blk, dtype=np.uint8)
            lib.b2p_memcpy_d2h(gpu, host.ctypes.data, dev_in.data_ptr() + ((args.steps - 1) % nrot) * nbeam * blk, blk)
            check_against = last_out
        else:
            host = pinned[(ke - 1) % 2][0].array
            check_against = e2e_last
        sums = oracle.accumulate_omp(host, ndf, og, nthreads=R.vcpus, L=L)
        want = oracle.finish(sums, 1.0)
        parity = bool(np.array_equal(want.view(np.uint32), check_against.view(np.uint32)))
        if not parity:
            raise SystemExit("bench.py: GPU spectrum differs from the CPU oracle — number withheld")
        del host
    if pinned is not None:
        for row in pinned:
            for p in row:
                p.free()
        pinned = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_port_measure(ndf)               # the routine `--impl reference` runs

    # ---- configs[1]/[3]: the same stream through the real process surface, on EVERY rank ----
    # (SysV rings + executables, one pipeline and one ring pair per GPU: paf-baseband2power.py:88-95,114-115)
    ring = None
    if not args.no_e2e and not args.no_ring and nbeam == 1:
        ring = _ring_leg(R)

    # ---- configs[4]: UDP live-rate replay, one beam per GPU on every rank ----
    live = None
    if not args.no_e2e and not args.no_live and nbeam == 1:
        live = _live_leg(R)

    # ---- extra: full beam set on one GPU, kernel-only (configs[2]) ----
    beamset = None
    if args.beamset and nbeam == 1 and args.beamset > 1 and world == 1:
        try:
            nb = args.beamset
            del dev_in
            R.dev_in = None
            torch.cuda.empty_cache()
            big = torch.empty(nb * blk, dtype=torch.uint8, device="cuda")
            for b in range(nb):
                lib.b2p_synth_fill_device(gpu, big.data_ptr() + b * blk, ndf, 48, 7, 128, 1, 2000 + rank * nb + b, 0, 1, None)
            bout = torch.empty(nb * g.nchan, dtype=torch.float32, device="cuda")
            sb = Baseband2Power(device_id=gpu, nbeam=nb, kernel=args.kernel)
            bp = [big.data_ptr() + b * blk for b in range(nb)]
            bx = torch.cuda.ExternalStream(sb.stream, device=torch.device("cuda", gpu))
            for _ in range(3):
                sb.integrate_device(bp, ndf, bout, None)
            barrier()
            reps = 5
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(bx)
            for _ in range(reps):
                sb.integrate_device(bp, ndf, bout, None)
            b1.record(bx)
            barrier()
            bms = max_over_ranks(b0.elapsed_time(b1)) / reps
            beamset = {"nbeam_per_gpu": nb, "ms_per_step": round(bms, 3),
                       "value": round(world * nb * blk / (bms * 1e-3) / 1e9, 1), "unit": UNIT,
                       "hbm_resident_bytes": nb * blk,
                       "realtime_factor_per_beamset": round(g.t_integration_s / (bms * 1e-3), 1)}
            sb.close()
            del big
        except Exception as e:  # informational leg only
            beamset = {"error": repr(e)[:200]}

    peak, peak_kind = _peaks()
    per_launch_ms = fused_ms / max(fused_n, 1)
    alg_bytes = nbeam * (blk + g.out_bytes)
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    tr = _ncu_traffic()
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs: copy, read+write)",
                "traffic": (tr or {}).get("dram_bytes_per_launch"),
                "traffic_source": ("ncu, not this run: " + (tr or {}).get("source", "profiles/roofline_traffic.json")) if tr else None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel": {"ldg": "b2p_fused_ldg256_bmf", "tma": "b2p_fused_tma_bmf"}.get(st.kernel, st.kernel),
                "launch_ms": round(per_launch_ms, 5), "launches_timed": fused_n,
                "timing": "isolated per-launch CUDA events on the launching stream, second pass of K steps "
                          "(the launch includes the cross-CTA reduce and the float32 finish)",
                "achieved_in_stream": round(alg_bytes / (ms_step * 1e-3) / 1e9, 1),
                "frac_of_nominal_8000": round(achieved / 8000.0, 4)}
    if rank == 0:
        emit({
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": _config(args, nbeam),
            "kernel_info": {"kernel": st.kernel, "nsplit": st.nsplit,
                            "launches_per_integration": launches / args.steps},
            "placement": {"policy": "env B2P_BENCH_GPUS" if env_map else args.placement, "gpus_on_box": ngpu_box,
                          "rank_to_gpu": gpu_map, "host_vcpus": R.vcpus,
                          "host_mem_available_GB": round((_mem_available_bytes() or 0) / 1e9, 1)},
            "realtime_factor": round(world * nbeam * g.t_integration_s * ndf / BLOCK_NDF / (ms_step * 1e-3), 1),
            "samples_per_s": round(world * nbeam * ndf * 128 * g.nchan / (ms_step * 1e-3), 1),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "parity_vs_oracle": parity,
            "ring_e2e": ring, "live_replay": live, "beamset": beamset, "device": device_info(gpu)["name"],
        })
    st.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- legs (N-rank aware)

def _channel_group_e2e(R, links, by_beam_spec, src_rot):
    """Channel-group sharding across the ranks: the (beam, chunk) units of a step — laid out
    beam-major — are cut into one run per GPU in proportion to what its host link delivers, so
    unequal links finish together; a GPU reads its chunk columns straight out of the full-frame
    blocks (strided H2D).  The blocks live in SysV shared memory (as ring blocks do) and are
    page-locked by every rank that reads them.  The shares start from the link probe and are
    refined from the measured per-GPU completion times of a few untimed steps (links that share
    a host bridge slow each other down, which no solo probe sees).  Returns timings for one
    beam over N GPUs and for N beams over N GPUs."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from paf_baseband2power_b200 import Baseband2Power
    from paf_baseband2power_b200.sharding import gather_unit_ranges, plan_units
    rank, world, gpu, lib, blk, ndf, g, args = R.rank, R.world, R.gpu, R.lib, R.blk, R.ndf, R.g, R.args

    base = [0]
    if rank == 0:
        base[0] = 0x5B200000 | ((os.getpid() & 0xFFF) << 8)
    dist.broadcast_object_list(base, src=0, group=R.host_pg)
    mine = SysVBlock(base[0] + rank, blk, create=True)
    # the block the by-beam loop finished on, so that the two modes can be compared bit for bit
    rc = lib.b2p_memcpy_d2h(gpu, mine.ptr, R.dev_in.data_ptr() + src_rot * blk, blk)
    assert rc == 0
    R.host_barrier()
    blocks = [mine if r == rank else SysVBlock(base[0] + r, blk, create=False) for r in range(world)]
    registered = set()
    t_reg = [0.0]

    def need(beam):
        if beam not in registered:
            t0 = time.perf_counter()
            rc = lib.b2p_host_register(blocks[beam].ptr, blk)
            assert rc == 0, "cudaHostRegister of a shared beam block failed"
            t_reg[0] += time.perf_counter() - t0
            registered.add(beam)

    def all_gather(x: float) -> list:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        outl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [float(o.item()) for o in outl]

    out = {"link_GBps_probe": [round(x, 2) for x in links],
           "host_blocks": "SysV shared memory, one 2.8 GB block per beam, page-locked by the ranks that read it"}

    def run_mode(nb, steps, refine):
        shares = list(links)
        history = []
        ctxs, plan = [], None

        def build(shares):
            nonlocal ctxs, plan
            for *_, c in ctxs:
                c.close()
            plan = plan_units(shares, nb, 48)
            ctxs = []
            for beam, first, n in plan[rank]:
                need(beam)
                ctxs.append((beam, first, n, Baseband2Power(device_id=gpu, nbeam=1, kernel=args.kernel, nchunk=n,
                                                            first_chunk=first, nchunk_total=48)))

        def one():
            """-> (spectra on rank 0, this rank's seconds from first issue to last spectrum)"""
            t0 = time.perf_counter()
            for beam, first, n, c in ctxs:                 # queue everything, then wait
                c.accumulate_host_async([blocks[beam].ptr], ndf, finish=True)
            part = np.zeros((nb, g.nchan), dtype=np.float32)
            for beam, first, n, c in ctxs:
                part[beam, 7 * first:7 * (first + n)] = c.wait_output()[0]
            busy = time.perf_counter() - t0
            return gather_unit_ranges(part, plan, 7, group=R.host_pg), busy   # 1344 B per beam per rank

        build(shares)
        for it in range(refine + 1):
            R.barrier()
            one()                                           # warm (staging buffers, first touch)
            R.barrier()
            _, busy = one()
            times = all_gather(busy)
            units = [sum(n for _, _, n in plan[r]) for r in range(world)]
            history.append({"units_per_gpu": units, "busy_ms_per_gpu": [round(t * 1e3, 2) for t in times]})
            live = [t for t, u in zip(times, units) if u > 0]
            if it == refine or max(live) / min(live) < 1.04:
                break
            # what each link delivered under the real contention of this step
            rate = [u / t if u > 0 else 0.0 for u, t in zip(units, times)]
            # a GPU that got no unit keeps its probe rate scaled like the others
            scale = sum(r_ for r_ in rate if r_ > 0) / max(1e-9, sum(l for l, r_ in zip(links, rate) if r_ > 0))
            new = [r_ if r_ > 0 else l * scale for r_, l in zip(rate, links)]
            shares = [0.5 * (a / sum(shares)) + 0.5 * (b_ / sum(new)) for a, b_ in zip(shares, new)]   # damped
            build(shares)
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            full, _ = one()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        R.barrier()
        ms = R.max_over_ranks(dt * 1e3) / steps
        ok = None
        if rank == 0 and by_beam_spec is not None:
            ok = bool(np.array_equal(full.view(np.uint32), np.asarray(by_beam_spec)[:nb].view(np.uint32)))
        units = [sum(n for _, _, n in plan[r]) for r in range(world)]
        for *_, c in ctxs:
            c.close()
        ctxs = []
        return {"beams_per_step": nb, "value": round(nb * blk / (ms * 1e-3) / 1e9, 3), "unit": UNIT,
                "ms_per_step": round(ms, 3), "steps": steps,
                "realtime_factor": round(nb * g.t_integration_s * ndf / BLOCK_NDF / (ms * 1e-3), 2),
                "units_per_gpu": units, "units_total": nb * 48, "unit_is": "one chunk column of one beam (58.7 MB)",
                "plan": [[list(x) for x in plan[r]] for r in range(world)],
                "refinement": history,
                "bit_identical_to_by_beam_spectra": ok,
                "path": "b2p_accumulate_host_async + b2p_wait_output on chunk-group shard contexts (strided H2D, pitch 344064 B); "
                        "(beam, chunk) units per GPU from the link probe, refined from measured completion times; "
                        "gloo gather of the channel ranges"}

    try:
        out["single_beam"] = run_mode(1, max(4, R.ke), 3)       # one beam's block over N host links
        out["all_beams"] = run_mode(world, max(4, R.ke // 2), 4)
        out["register_s_per_rank"] = round(t_reg[0], 2)
        for k in ("single_beam", "all_beams"):
            if rank == 0 and out[k]["bit_identical_to_by_beam_spectra"] is False:
                raise SystemExit("bench.py: channel-group spectra differ from the by-beam spectra — number withheld")
    finally:
        for beam in registered:
            lib.b2p_host_unregister(blocks[beam].ptr)
        R.host_barrier()
        for b in blocks:
            b.close()
    return out


def _load_tool(name):
    import importlib.util
    spec_ = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    return mod


def _ring_plan(world, want_bufs, blk_ndf):
    """Ring geometry that fits the host: the reference block (8192 frames, 2.8 GB) when
    memory allows, else shorter blocks with the integration spanning several (-n 8192)."""
    avail = _mem_available_bytes()
    ndf_blk, nbufs = blk_ndf, want_bufs
    frame = 48 * 7168
    while avail is not None and world * nbufs * ndf_blk * frame > 0.6 * avail and ndf_blk > 512:
        ndf_blk //= 2
    return ndf_blk, nbufs, avail


def _ring_leg(R):
    from paf_baseband2power_b200.sharding import ring_keys_for_beam
    rank, world, gpu, args, ndf = R.rank, R.world, R.gpu, R.args, R.ndf
    synced = [False]

    def start_together():       # every rank passes this exactly once, also when its pipeline fails
        if not synced[0]:
            synced[0] = True
            R.host_barrier()

    try:
        mod = _load_tool("run_ring_e2e")
        ndf_blk, nbufs, _ = _ring_plan(world, 3, ndf)
        per_int = ndf // ndf_blk
        salt = (os.getppid() & 0x3F) * 0x10000
        keys = ring_keys_for_beam(rank, 0x1B200 + salt, 0x1B2A0 + salt, 0x100)
        res = mod.run(ndf=ndf_blk, nbufs=nbufs, nblocks=args.ring_blocks * per_int, gpu=gpu, kernel=args.kernel,
                      keys=keys, ndf_integration=ndf if per_int > 1 else 0,
                      producer_threads=max(1, R.vcpus // world), seed=1 + rank, start_barrier=start_together)
        res.pop("_spectra", None)
        res["rank"] = rank
    except Exception as e:
        res = {"rank": rank, "error": repr(e)[:300]}
    start_together()
    allres = R.gather_objects(res)
    if rank != 0:
        R.host_barrier()        # rank 0 alone runs the one-stage-all-GPUs pipeline below
        return None
    good = [r for r in allres if "error" not in r]
    out = {"path": "paf_memdb -> ring -> paf_baseband2power -> ring -> paf_dbdisk: one pipeline and one ring pair per GPU, all running at once",
           "integrations_per_beam": args.ring_blocks, "pipelines": len(allres)}
    if good:
        busy = max(r["stage_busy_s"] for r in good)
        total = sum(r["blocks"] * r["ndf_per_block"] * 48 * 7168 for r in good)
        t_data = sum(r["blocks"] * r["ndf_per_block"] * 128 * 27.0 / 32.0 * 1e-6 for r in good)
        out.update({"aggregate_GBps": round(total / busy / 1e9, 3),
                    "aggregate_realtime_factor": round(t_data / busy, 2),
                    "slowest_stage_busy_s": busy,
                    "sum_of_per_rank_GBps": round(sum(r["stage_GBps"] for r in good), 3),
                    "per_rank_GBps": [r.get("stage_GBps") for r in allres],
                    "note": "aggregate = all bytes / the slowest pipeline's busy time; the pipelines are independent, "
                            "so while all run the box ingests the sum of the per-rank rates"})
        if world == 1:                              # same keys as in round 1's line
            out.update({k: v for k, v in good[0].items() if k not in ("rank",)})
    out["per_rank"] = allres
    if world > 1:
        # one beam, ONE stage process, its channel groups over all GPUs of this run (-d list):
        # the single-process form of channel-group sharding, through the same rings
        try:
            keys = ring_keys_for_beam(world, 0x1B200 + salt, 0x1B2A0 + salt, 0x100)
            res = mod.run(ndf=ndf_blk, nbufs=nbufs, nblocks=args.ring_blocks * per_int,
                          gpu=",".join(str(x) for x in R.gpu_map), kernel=args.kernel, keys=keys,
                          ndf_integration=ndf if per_int > 1 else 0, producer_threads=R.vcpus, seed=99)
            res.pop("_spectra", None)
            out["single_beam_one_stage_all_gpus"] = res
        except Exception as e:
            out["single_beam_one_stage_all_gpus"] = {"error": repr(e)[:300]}
    R.host_barrier()
    return out


def _live_leg(R):
    from paf_baseband2power_b200.sharding import ring_keys_for_beam
    rank, world, gpu, args, ndf = R.rank, R.world, R.gpu, R.args, R.ndf
    try:
        mod = _load_tool("run_live_replay")
    except Exception as e:
        return {"error": repr(e)[:300]} if rank == 0 else None
    ndf_blk, nbufs, _ = _ring_plan(world, 4, ndf)
    per_int = ndf // ndf_blk
    # two sender threads per beam: with UDP segmentation offload two keep the line rate, and more
    # threads only fight the capture threads for cores (measured: 4 per beam fell behind at N=4)
    threads = 2 if R.vcpus >= 4 * world else 1
    trials = []
    for ti, rate in enumerate((1.0, 0.5, 0.25)):
        salt = (os.getppid() & 0x3F) * 0x10000 + ti * 0x1000
        keys = ring_keys_for_beam(rank, 0x2C200 + salt, 0x2C2A0 + salt, 0x100)
        synced = [False]

        def start_together(synced=synced):
            if not synced[0]:
                synced[0] = True
                R.host_barrier()

        try:
            res = mod.run(ndf=ndf_blk, nblocks=args.live_blocks * per_int, rate_frac=rate, threads=threads, gpu=gpu,
                          keys=keys, port=21000 + 64 * rank + 8 * ti, nbufs=nbufs,
                          ndf_integration=ndf if per_int > 1 else 0, start_barrier=start_together)
            res["rank"] = rank
        except Exception as e:
            res = {"rank": rank, "error": repr(e)[:300]}
        start_together()
        allres = R.gather_objects(res)
        done = [True]
        if rank == 0:
            good = [r for r in allres if "error" not in r and "packets_expected" in r]
            exp = sum(r["packets_expected"] for r in good)
            got = sum(r["packets_received"] for r in good)
            zf = sum(r["packets_zero_filled"] for r in good)
            ok = bool(exp) and len(good) == len(allres)
            trials.append({"rate_x_line": rate, "beams": len(allres), "beams_ok": len(good),
                           "packets_expected": exp, "packets_received": got, "packets_zero_filled": zf,
                           "received_frac": round(got / exp, 5) if exp else None,
                           "sustained_realtime_factor": rate if (ok and zf == 0) else None,
                           "aggregate_wire_GBps": round(sum(r.get("replay_GBps", 0.0) for r in good), 2),
                           "gpu_stage_headroom_x_realtime": [r.get("stage_headroom_x_realtime") for r in allres],
                           "per_rank": allres})
            done[0] = ok and got >= 0.999 * exp
        if world > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(done, src=0, group=R.host_pg)
        if done[0]:
            break
    if rank != 0:
        return None
    return {"path": "bmf_replay -> UDP loopback -> paf_capture -> ring -> paf_baseband2power -> ring -> paf_dbdisk: one beam per GPU, all running at once",
            "host_vcpus": R.vcpus, "sender_threads_per_beam": threads, "blocks_per_beam": args.live_blocks,
            "line_rate": "9259 frames/s x 48 packets of 7232 B = 3.21 GB/s per beam (capture.h:27-32)",
            "note": "loopback UDP is two kernel copies per packet on the host CPUs: this leg measures the host, the GPU stage idles",
            "trials": trials}


if __name__ == "__main__":
    main()
