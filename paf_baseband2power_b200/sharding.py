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


# ---------------------------------------------------------------- placement and channel groups

def gpu_for_rank(rank: int, world: int, ngpus: int, policy: str = "spread") -> int:
    """GPU a rank (or a beam's pipeline) runs on when the box has `ngpus` GPUs.

    `spread`: rank i -> GPU floor(i * ngpus / world), i.e. 4 ranks on an 8-GPU box take GPUs
    0, 2, 4, 6 instead of 0..3 — neighbouring GPUs usually hang off the same host bridge and
    share its bandwidth, and the end-to-end rate of this path is the host link.  `identity`:
    rank i -> GPU i (the reference launcher's -c index, paf-baseband2power.py:26)."""
    if ngpus <= 0:
        return rank
    if policy == "identity" or world >= ngpus:
        return rank % ngpus
    return (rank * ngpus) // world


def chunk_ranges(counts: Sequence[int]) -> List[tuple]:
    """[(first_chunk, nchunk)] per shard for consecutive chunk counts (zeros allowed)."""
    out, first = [], 0
    for n in counts:
        if n < 0:
            raise ValueError("negative chunk count")
        out.append((first, int(n)))
        first += int(n)
    return out


def split_chunks(weights: Sequence[float] | None, n: int, nchunk: int = 48) -> List[int]:
    """nchunk chunks over n shards in proportion to weights (largest remainder, ties to the
    lower index) — the host-side twin of b2p_split_chunks, for planning without a GPU."""
    w = [1.0] * n if weights is None else [float(x) for x in weights]
    if len(w) != n or any(x < 0 for x in w) or sum(w) <= 0:
        raise ValueError("weights must be n non-negative numbers, not all zero")
    share = [x / sum(w) * nchunk for x in w]
    counts = [int(s) for s in share]
    frac = [s - c for s, c in zip(share, counts)]
    while sum(counts) < nchunk:
        best = max(range(n), key=lambda i: (frac[i], -i))
        counts[best] += 1
        frac[best] = -1.0
    return counts


def gather_channel_groups(local: np.ndarray, counts: Sequence[int], nch_per_chunk: int = 7, group=None,
                          dst: int = 0):
    """Channel-group sharding: rank r holds spectra [nbeam, counts[r]*nch] for chunks
    chunk_ranges(counts)[r]; returns [nbeam, sum(counts)*nch] on `dst` (None elsewhere).
    The ranges are disjoint, so the gather is a placement, not a reduction."""
    import torch
    import torch.distributed as dist

    ranges = chunk_ranges(counts)
    nchan = sum(counts) * nch_per_chunk
    local = np.ascontiguousarray(local, dtype=np.float32)
    nbeam = local.shape[0]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[1] != nchan:
            raise ValueError("single rank must hold every channel")
        return local.copy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if len(counts) != world:
        raise ValueError("one chunk count per rank")
    first, n = ranges[rank]
    if local.shape[1] != n * nch_per_chunk:
        raise ValueError(f"rank {rank} holds {local.shape[1]} channels, its range has {n * nch_per_chunk}")
    pad = np.zeros((nbeam, nchan), dtype=np.float32)          # fixed-size message: 1344 B per beam
    pad[:, first * nch_per_chunk:(first + n) * nch_per_chunk] = local
    t = torch.from_numpy(pad)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros((nbeam, nchan), dtype=np.float32)
    for r, (f, m) in enumerate(ranges):
        sl = slice(f * nch_per_chunk, (f + m) * nch_per_chunk)
        out[:, sl] = bufs[r].numpy()[:, sl]
    return out


def plan_units(shares: Sequence[float], nbeams: int, nchunk: int = 48) -> List[List[tuple]]:
    """Finer than whole chunk columns: the nbeams*nchunk (beam, chunk) units, laid out
    beam-major, are cut into len(shares) consecutive runs in proportion to `shares`; shard r
    gets, per beam it touches, one interval (beam, first_chunk, nchunk).  With 8 beams a unit
    is 1/384 of the step instead of the 1/48 of a chunk column, so unequal host links can be
    matched to ~1 %.  Returns one interval list per shard (possibly empty)."""
    w = [float(x) for x in shares]
    if not w or any(x < 0 for x in w) or sum(w) <= 0:
        raise ValueError("shares must be non-negative and not all zero")
    total = nbeams * nchunk
    bounds, acc = [0], 0.0
    for x in w[:-1]:
        acc += x / sum(w)
        bounds.append(min(total, max(bounds[-1], int(round(acc * total)))))
    bounds.append(total)
    plan = []
    for r in range(len(w)):
        lo, hi = bounds[r], bounds[r + 1]
        items = []
        while lo < hi:
            beam, first = divmod(lo, nchunk)
            n = min(hi - lo, nchunk - first)
            items.append((beam, first, n))
            lo += n
        plan.append(items)
    return plan


def gather_unit_ranges(local: np.ndarray, plan: Sequence[Sequence[tuple]], nch_per_chunk: int = 7, group=None,
                       dst: int = 0):
    """`local` is this rank's [nbeams, nchan] array with its own (beam, chunk-range) intervals of
    `plan` filled in; returns the complete [nbeams, nchan] on `dst` (None elsewhere).  Intervals
    are disjoint, so this is a placement of 1344-byte rows, not a reduction."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.float32)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.copy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if len(plan) != world:
        raise ValueError("one interval list per rank")
    t = torch.from_numpy(local)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros_like(local)
    for r, items in enumerate(plan):
        src = bufs[r].numpy()
        for beam, first, n in items:
            sl = slice(first * nch_per_chunk, (first + n) * nch_per_chunk)
            out[beam, sl] = src[beam, sl]
    return out
