"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): beams shard round-robin, every rank
integrates only its beams, rank 0 gathers the spectra.  The per-beam arithmetic here is the
oracle standing in for the GPU stage (this file runs without a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from paf_baseband2power_b200.sharding import beams_for_rank, gather_spectra, rank_of_beam, ring_keys_for_beam

NBEAM, NDF = 7, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _beam_spectrum(beam):
    import oracle
    blk = oracle.synth_fill(NDF, seed=900 + beam, mode=1)
    return oracle.finish(oracle.accumulate(blk), 1.0)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = beams_for_rank(NBEAM, rank, world)
    local = np.stack([_beam_spectrum(b) for b in mine]) if mine else np.zeros((0, 336), np.float32)
    out = gather_spectra(local, mine, NBEAM)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([_beam_spectrum(b) for b in range(NBEAM)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_round_robin_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            ids = beams_for_rank(36, r, world)
            assert all(rank_of_beam(b, world) == r for b in ids)
            seen += ids
        assert sorted(seen) == list(range(36))
    sizes = [len(beams_for_rank(36, r, 8)) for r in range(8)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        beams_for_rank(4, 4, 4)


def test_ring_keys_do_not_collide():
    keys = set()
    for b in range(36):
        kin, kout = ring_keys_for_beam(b)
        for k in (kin, kin + 1, kout, kout + 1):   # data ring + header ring at key+1
            assert k not in keys
            keys.add(k)


def test_single_process_gather_is_identity():
    local = np.arange(2 * 336, dtype=np.float32).reshape(2, 336)
    out = gather_spectra(local, [0, 1], 2)
    assert np.array_equal(out, local)


# ------------------------------------------------ channel-group sharding (round 2)

def _cg_worker(rank, world, port, counts, q):
    import oracle
    from paf_baseband2power_b200.sharding import chunk_ranges, gather_channel_groups
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = chunk_ranges(counts)[rank]
    nbeam = 2
    local = np.zeros((nbeam, 7 * n), np.float32)
    for b in range(nbeam):
        # the oracle on this rank's chunk columns only stands in for the GPU shard
        blk = oracle.synth_fill(NDF, seed=700 + b, mode=1).reshape(NDF, 48, 7168)
        sub = np.ascontiguousarray(blk[:, first:first + n]).reshape(-1)
        if n:
            gsub = oracle.Geometry(nchunk=n)
            local[b] = oracle.finish(oracle.accumulate(sub, NDF, gsub), 1.0)
    out = gather_channel_groups(local, counts)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[24, 24], [20, 28], [48, 0], [10, 17, 21]])
def test_channel_group_gather_over_gloo(counts):
    """Every rank integrates its chunk range of the same beams; rank 0 places the disjoint
    channel ranges side by side: equal to the oracle's spectrum of the whole block."""
    import oracle
    world = len(counts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, counts, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([oracle.finish(oracle.accumulate(oracle.synth_fill(NDF, seed=700 + b, mode=1)), 1.0) for b in range(2)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_split_chunks_host_twin():
    from paf_baseband2power_b200.sharding import chunk_ranges, split_chunks
    assert split_chunks(None, 8) == [6] * 8
    assert split_chunks([23, 23, 23, 23, 35, 35, 35, 35], 8) == [5, 5, 5, 5, 7, 7, 7, 7]
    assert split_chunks([55, 55], 2) == [24, 24]
    c = split_chunks([1.0, 2.5, 0.0, 3.1], 4)
    assert sum(c) == 48 and c[2] == 0
    assert chunk_ranges([5, 0, 7]) == [(0, 5), (5, 0), (5, 7)]
    for n in range(1, 17):
        for nchunk in (1, 7, 48):
            assert sum(split_chunks([1 + (i * 7) % 5 for i in range(n)], n, nchunk)) == nchunk
    with pytest.raises(ValueError):
        split_chunks([0, 0], 2)


def test_gpu_placement_spreads_ranks():
    from paf_baseband2power_b200.sharding import gpu_for_rank
    assert [gpu_for_rank(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]     # two per half of an 8-GPU box
    assert [gpu_for_rank(r, 2, 8) for r in range(2)] == [0, 4]
    assert [gpu_for_rank(r, 8, 8) for r in range(8)] == list(range(8))
    assert [gpu_for_rank(r, 1, 8) for r in range(1)] == [0]
    assert [gpu_for_rank(r, 4, 8, "identity") for r in range(4)] == [0, 1, 2, 3]
    assert [gpu_for_rank(r, 4, 2) for r in range(4)] == [0, 1, 0, 1]        # more beams than GPUs
    for world in range(1, 9):
        for ngpu in range(world, 17):
            gpus = [gpu_for_rank(r, world, ngpu) for r in range(world)]
            assert len(set(gpus)) == world and max(gpus) < ngpu


def _unit_worker(rank, world, port, shares, nbeams, q):
    import oracle
    from paf_baseband2power_b200.sharding import gather_unit_ranges, plan_units
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = plan_units(shares, nbeams)
    part = np.zeros((nbeams, 336), np.float32)
    for beam, first, n in plan[rank]:
        blk = oracle.synth_fill(NDF, seed=800 + beam, mode=1).reshape(NDF, 48, 7168)
        sub = np.ascontiguousarray(blk[:, first:first + n]).reshape(-1)
        part[beam, 7 * first:7 * (first + n)] = oracle.finish(oracle.accumulate(sub, NDF, oracle.Geometry(nchunk=n)), 1.0)
    out = gather_unit_ranges(part, plan)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shares,nbeams", [([1, 1], 2), ([26, 43, 35], 3), ([1, 0, 2], 2), ([5, 1], 1)])
def test_unit_plan_gather_over_gloo(shares, nbeams):
    """(beam, chunk) units cut into one run per rank: every rank integrates its intervals, rank 0
    places them; equal to the oracle's spectra of the whole blocks."""
    import oracle
    from paf_baseband2power_b200.sharding import plan_units
    world = len(shares)
    plan = plan_units(shares, nbeams)
    covered = sorted((b, c) for items in plan for b, f, n in items for c in range(f, f + n))
    assert covered == [(b, c) for b in range(nbeams) for c in range(48)]      # a partition, nothing twice
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_unit_worker, args=(r, world, port, shares, nbeams, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([oracle.finish(oracle.accumulate(oracle.synth_fill(NDF, seed=800 + b, mode=1)), 1.0) for b in range(nbeams)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
