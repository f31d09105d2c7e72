"""Diagnostic (multi-GPU box): aggregate pinned H2D bandwidth with and without NUMA-local
host buffers.  Run under torchrun; prints per-rank and aggregate GB/s."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paf_baseband2power_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
libc = ctypes.CDLL(None, use_errno=True)


def gpu_numa_node(dev):
    bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
    except Exception:
        pass
    if not bus:
        return -1, None
    bus = bus.lower()
    if len(bus.split(":")[0]) == 8:
        bus = bus[4:]
    p = f"/sys/bus/pci/devices/{bus}/numa_node"
    try:
        return int(open(p).read()), bus
    except Exception:
        return -1, bus


def node_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:
        return set()
    out = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out |= set(range(int(a), int(b or a) + 1))
    return out


def set_mempolicy(mode, node):
    mask = ctypes.c_ulong(1 << node) if node >= 0 else ctypes.c_ulong(0)
    SYS_set_mempolicy = 238
    r = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)
    return r, ctypes.get_errno()


def measure(tag, nbytes=2 << 30):
    p = ctypes.c_void_p()
    assert lib.b2p_host_alloc(ctypes.byref(p), nbytes) == 0
    ctypes.memset(p.value, 1, nbytes)
    host = torch.frombuffer((ctypes.c_uint8 * nbytes).from_address(p.value), dtype=torch.uint8)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rate = 5 * nbytes / dt / 1e9
    t = torch.tensor([rate], device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        rs = [round(float(x.item()), 1) for x in allr]
        print(tag, "per-rank GB/s", rs, "aggregate", round(sum(rs), 1), flush=True)
    del dev, host
    lib.b2p_host_free(p)
    dist.barrier()


node, bus = gpu_numa_node(local)
allowed = os.sched_getaffinity(0)
print(f"rank {rank} gpu {local} bus {bus} numa {node} allowed_cpus {len(allowed)} node_cpus&allowed {len(node_cpus(node) & allowed) if node >= 0 else None}", flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m | head -14; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; ls /sys/devices/system/node/ | grep node; cat /proc/self/status | grep -i mems_allowed_list")
dist.barrier()
measure("default")
if node >= 0:
    r = set_mempolicy(2, node)      # MPOL_BIND
    cpus = node_cpus(node) & allowed
    if cpus:
        os.sched_setaffinity(0, cpus)
    print(f"rank {rank} set_mempolicy rc {r} cpus {len(cpus)}", flush=True)
measure("numa-local")
dist.destroy_process_group()
