"""Beam -> GPU sharding and the host-side gather of integrated spectra.

The hot path shards by beam with no collective: each beam is an independent
stream with its own ring key and its own stage process in the reference's design
(-a/-b keys and -d GPU index per process, paf_baseband2power.cu:23-26; the
launcher starts one stage per GPU, paf-baseband2power.py:26,88-92).  The only
cross-rank traffic is the result: NCHAN float32 (1344 B) per beam per
integration, gathered to the host of rank 0.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def beams_for_rank(nbeam_total: int, rank: int, world: int) -> List[int]:
    """Round-robin: beam b lives on rank b % world."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, nbeam_total, world))


def rank_of_beam(beam: int, world: int) -> int:
    return beam % world


def ring_keys_for_beam(beam: int, base_in: int = 0xDADA, base_out: int = 0xADAD, stride: int = 0x10):
    """One input and one output ring key per beam (each key uses key and key+1)."""
    return base_in + beam * stride, base_out + beam * stride


def gather_spectra(local: np.ndarray, local_beams: Sequence[int], nbeam_total: int, group=None, dst: int = 0):
    """Gather per-rank spectra [nlocal, nchan] to rank `dst`; returns [nbeam_total, nchan]
    (beam-ordered) on `dst`, None elsewhere.  Host tensors: use a gloo group."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.float32)
    nchan = local.shape[1]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = np.zeros((nbeam_total, nchan), dtype=np.float32)
        out[list(local_beams)] = local
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = (nbeam_total + world - 1) // world
    pad = np.zeros((nmax, nchan), dtype=np.float32)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros((nbeam_total, nchan), dtype=np.float32)
    for r in range(world):
        ids = beams_for_rank(nbeam_total, r, world)
        out[ids] = bufs[r].numpy()[: len(ids)]
    return out
