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
