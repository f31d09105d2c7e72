#!/usr/bin/env python
"""bench.py — input GB/s and real-time factor of the baseband->power hot path on B200.

One step = one beam-integration per beam stream: one input ring block of
8192 data frames (2 818 572 288 B, paf-baseband2power.conf:9 /
paf-baseband2power.py:67) unpacked, detected and integrated into 336 float32
(1344 B).  Default workload = BASELINE.json configs[1]: a single beam, a
continuous stream of integrations.

  value     kernel-only: blocks already resident in HBM (4 rotating blocks, each
            22x larger than L2), K steps on one stream (fused kernel + reduce/finish
            kernel per step, PDL-chained), CUDA events, max over ranks.
  e2e       the same through the C ABI with HOST buffers: b2p_accumulate_host
            (pinned ring block -> chunked H2D overlapped with the kernel) +
            b2p_finish (D2H of the spectrum) inside the timed region.
  roofline  fused kernel only, per-launch duration from CUDA events recorded on
            the launching stream inside the timed region, against the measured
            HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU oracle (a port of the specification; the reference has no
            kernel to time, kernel.cu:1-7) on all host cores, bounded sample.

`--impl reference` times that CPU port alone (rank 0 only), same metric/config.
Multi-GPU: one process per GPU (torchrun), beams shard by rank, no collective on
the data path; only the spectra are gathered.
"""
from __future__ import annotations

import argparse
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


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference
    never wrote one (kernel.cu:1-7), so this is the oracle port on all host threads."""
    if rank != 0:
        return
    import numpy as np
    oracle, L = _cpu_port()
    g = oracle.Geometry()
    ndf = args.ndf
    block = np.empty(ndf * g.frame_bytes, dtype=np.uint8)
    gc = g.c()
    import ctypes
    L.b2p_oracle_synth_fill(block.ctypes.data, ndf, ctypes.byref(gc), 1, 0, 1)
    threads = _host_threads()   # explicit: torchrun exports OMP_NUM_THREADS=1
    sums = np.zeros(g.nchan, dtype=np.uint64)
    for _ in range(args.warmup):
        oracle.accumulate_omp(block, ndf, g, sums=sums, nthreads=threads, L=L)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sums[:] = 0
        oracle.accumulate_omp(block, ndf, g, sums=sums, nthreads=threads, L=L)
        oracle.finish(sums, 1.0)
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    gbs = block.nbytes / (ms * 1e-3) / 1e9
    t_int = ndf * 128 * 27.0 / 32.0 * 1e-6
    sample = f"{args.steps} steps x 1 beam-integration of {ndf} frames ({block.nbytes} B) on host RAM"
    emit(({
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": _config(args, 1),
        "realtime_factor": round(t_int / (ms * 1e-3), 3),
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "reference kernel absent (kernel.cu:1-7): C port of the specification, gcc -O3 -march=native -fopenmp"},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _config(args, nbeam):
    return {
        "workload": ("BASELINE.json configs[1]: single beam, continuous stream of integrations, "
                     "1 input ring block (8192 frames x 48 chunks x 7168 B = 2818572288 B) -> 336 x float32 per step"
                     if nbeam == 1 else
                     f"BASELINE.json configs[2]: {nbeam} beams batched per step, each 1 ring block of 2818572288 B"),
        "nbeam_per_gpu": nbeam, "ndf": args.ndf, "nchunk": 48, "nch_per_chunk": 7, "nsamp_df": 128,
        "mode": "exact-uint64", "kernel": args.kernel,
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


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nbeam", type=int, default=1, help="beam streams per GPU per step")
    ap.add_argument("--ndf", type=int, default=8192, help="data frames per ring block")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ldg", "tma"])
    ap.add_argument("--nsplit", type=int, default=0, help="time splits per chunk (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 32)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ring", action="store_true", help="skip the executables-through-the-rings leg")
    ap.add_argument("--beamset", type=int, default=36, help="extra kernel-only point: full beam set on one GPU (0 = skip)")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    host_pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_pg = dist.new_group(backend="gloo")   # spectra are gathered host-side, 1344 B per beam

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nbeam, ndf = args.nbeam, args.ndf
    g = BMF
    blk = ndf * g.frame_bytes
    # rotate between distinct blocks so that no step can find its input in L2 (126 MB); one
    # set is already 22x L2 per beam, so fewer sets are used when many beams fill the HBM
    nrot = max(1, min(4, int(0.45 * torch.cuda.get_device_properties(local).total_memory // (nbeam * blk))))
    # ---- device-resident inputs: nrot distinct blocks per beam ----
    dev_in = torch.empty(nrot * nbeam * blk, dtype=torch.uint8, device="cuda")
    from paf_baseband2power_b200 import _lib
    lib = _lib.load()
    wpb = blk // 8
    for r in range(nrot):
        for b in range(nbeam):
            beam_id = rank * nbeam + b
            rc = lib.b2p_synth_fill_device(local, dev_in.data_ptr() + (r * nbeam + b) * blk, ndf, 48, 7, 128, 1,
                                           1000 + beam_id, r * wpb, 1, None)
            assert rc == 0
    out_dev = torch.empty(nbeam * g.nchan, dtype=torch.float32, device="cuda")
    st = Baseband2Power(device_id=local, nbeam=nbeam, kernel=args.kernel, nsplit=args.nsplit)
    # The kernels run on the context's own stream (stream=None through the C ABI): only
    # there may a fused kernel start under the tail of its predecessor (PDL).  torch wraps
    # that same stream so the torch.cuda.Event pair below is recorded on it.
    xstream = torch.cuda.ExternalStream(st.stream, device=torch.device("cuda", local))
    ptrs = [[dev_in.data_ptr() + (r * nbeam + b) * blk for b in range(nbeam)] for r in range(nrot)]
    torch.cuda.synchronize()

    def step(i):
        st.integrate_device(ptrs[i % nrot], ndf, out_dev, None)   # ONE launch per integration

    sampler = ClockSampler(local)
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
    if not args.no_e2e:
        ke = args.e2e_steps or min(args.steps, 32)
        hrot = 2
        pinned = [[PinnedBuffer(blk) for _ in range(nbeam)] for _ in range(hrot)]
        for r in range(hrot):
            for b in range(nbeam):
                rc = lib.b2p_memcpy_d2h(local, pinned[r][b].ptr, dev_in.data_ptr() + (r * nbeam + b) * blk, blk)
                assert rc == 0
        from paf_baseband2power_b200.sharding import beams_for_rank, gather_spectra
        my_beams = beams_for_rank(world * nbeam, rank, world)
        ste = Baseband2Power(device_id=local, nbeam=nbeam, kernel=args.kernel)

        def e2e_step(i):
            sp = ste.integrate_host(pinned[i % hrot], ndf)   # H2D pieces + kernels + D2H of the spectra
            if world > 1:                           # the only cross-rank traffic of the path
                allsp = gather_spectra(sp, my_beams, world * nbeam, group=host_pg)
                return sp, allsp
            return sp, sp

        for i in range(3):
            spec, _ = e2e_step(i)
        # plain pinned H2D of one block: the link ceiling the e2e number sits under
        link0, link1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_t = torch.empty(blk, dtype=torch.uint8, device="cuda")
        hsrc = torch.from_numpy(pinned[0][0].array)
        stage_t.copy_(hsrc, non_blocking=True)
        torch.cuda.synchronize()
        link0.record()
        for _ in range(3):
            stage_t.copy_(hsrc, non_blocking=True)
        link1.record()
        torch.cuda.synchronize()
        h2d_link = 3 * blk / (link0.elapsed_time(link1) * 1e-3) / 1e9
        del stage_t
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            spec, _ = e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        e2e_ms = max_over_ranks(dt * 1e3) / ke
        e2e = {"value": round(world * nbeam * blk / (e2e_ms * 1e-3) / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": nbeam * blk, "d2h_bytes_per_step": nbeam * g.out_bytes,
               "ms_per_step": round(e2e_ms, 3), "steps": ke, "host_threads_per_gpu": 1,
               "realtime_factor": round(world * nbeam * g.t_integration_s * ndf / 8192 / (e2e_ms * 1e-3), 2),
               "h2d_link_GBps_per_gpu": round(h2d_link, 2),
               "frac_of_h2d_link": round(nbeam * blk / (e2e_ms * 1e-3) / 1e9 / h2d_link, 4),
               "path": "b2p_accumulate_host (pinned ring block, 256-frame pieces, 3 staging buffers) + b2p_finish"
                       + (" + gloo gather of spectra to rank 0" if world > 1 else "")}
        e2e_last = spec[0].copy()
        ste.close()
    clocks = sampler.stop()

    # ---- CPU baseline + parity check against the oracle (rank 0, N=1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle, L = _cpu_port()
        og = oracle.Geometry()
        if args.no_e2e:
            host = np.empty(blk, dtype=np.uint8)
            lib.b2p_memcpy_d2h(local, host.ctypes.data, dev_in.data_ptr() + ((args.steps - 1) % nrot) * nbeam * blk, blk)
            check_against = last_out
        else:
            host = pinned[(ke - 1) % hrot][0].array
            check_against = e2e_last
        threads = _host_threads()
        times = []
        t_begin = time.perf_counter()
        while True:
            sums = np.zeros(og.nchan, dtype=np.uint64)
            t0 = time.perf_counter()
            oracle.accumulate_omp(host, ndf, og, sums=sums, nthreads=threads, L=L)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > 10.0 or len(times) >= 200:
                break
        t0 = time.perf_counter()                      # one single-thread pass, for the per-core figure
        oracle.accumulate(host, ndf, og, L=L)
        t_single = time.perf_counter() - t0
        want = oracle.finish(sums, 1.0)
        parity = bool(np.array_equal(want.view(np.uint32), check_against.view(np.uint32)))
        med = statistics.median(times)
        cpu = {"value": round(blk / med / 1e9, 3), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{len(times)} passes over 1 beam-integration ({blk} B) in host RAM, median; best {round(blk / min(times) / 1e9, 3)} GB/s",
               "realtime_factor": round(g.t_integration_s * ndf / 8192 / med, 3),
               "single_thread_GBps": round(blk / t_single / 1e9, 3),
               "note": "reference kernel absent (kernel.cu:1-7): C port of the specification, gcc -O3 -march=native -fopenmp"}
        if not parity:
            raise SystemExit("bench.py: GPU spectrum differs from the CPU oracle — number withheld")

    # ---- extra: the same stream through the real process surface (SysV rings + executables) ----
    ring = None
    if rank == 0 and world == 1 and not args.no_e2e and not args.no_ring and nbeam == 1:
        try:
            import importlib.util
            spec_ = importlib.util.spec_from_file_location("run_ring_e2e", os.path.join(ROOT, "tools", "run_ring_e2e.py"))
            mod = importlib.util.module_from_spec(spec_)
            spec_.loader.exec_module(mod)
            ring = mod.run(ndf=ndf, nbufs=4, nblocks=16, gpu=local, kernel=args.kernel)
        except Exception as e:  # informational leg only
            ring = {"error": repr(e)[:300]}

    # ---- extra: full beam set on one GPU, kernel-only (configs[2]) ----
    beamset = None
    if args.beamset and nbeam == 1 and args.beamset > 1:
        try:
            nb = args.beamset
            del dev_in
            torch.cuda.empty_cache()
            big = torch.empty(nb * blk, dtype=torch.uint8, device="cuda")
            for b in range(nb):
                lib.b2p_synth_fill_device(local, big.data_ptr() + b * blk, ndf, 48, 7, 128, 1, 2000 + rank * nb + b, 0, 1, None)
            bout = torch.empty(nb * g.nchan, dtype=torch.float32, device="cuda")
            sb = Baseband2Power(device_id=local, nbeam=nb, kernel=args.kernel)
            bp = [big.data_ptr() + b * blk for b in range(nb)]
            bx = torch.cuda.ExternalStream(sb.stream, device=torch.device("cuda", local))
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
                "algorithmic_bytes_per_launch": alg_bytes, "kernel": f"b2p_fused_{st.kernel}_bmf",
                "launch_ms": round(per_launch_ms, 5), "launches_timed": fused_n,
                "timing": "isolated per-launch CUDA events on the launching stream, second pass of K steps",
                "achieved_in_stream": round(alg_bytes / (ms_step * 1e-3) / 1e9, 1),
                "frac_of_nominal_8000": round(achieved / 8000.0, 4)}
    if rank == 0:
        emit(({
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": _config(args, nbeam) | {"kernel": st.kernel, "nsplit": st.nsplit},
            "realtime_factor": round(world * nbeam * g.t_integration_s * ndf / 8192 / (ms_step * 1e-3), 1),
            "samples_per_s": round(world * nbeam * ndf * 128 * g.nchan / (ms_step * 1e-3), 1),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "parity_vs_oracle": parity,
            "ring_e2e": ring, "beamset": beamset, "device": device_info(local)["name"],
        }))
    st.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
