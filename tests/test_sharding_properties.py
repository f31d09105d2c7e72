"""Property tests of the host-side planning helpers of channel-group sharding (no GPU, no
process group): whatever the link weights, every (beam, chunk) unit is owned by exactly one
shard, ranges are consecutive, and the split follows the weights to within one unit."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from paf_baseband2power_b200.sharding import (beams_for_rank, chunk_ranges, gpu_for_rank, plan_units,
                                               rank_of_beam, ring_keys_for_beam, split_chunks)

weights = st.lists(st.floats(min_value=0.0, max_value=100.0, allow_nan=False), min_size=1, max_size=16).filter(
    lambda w: sum(w) > 1e-6)


@settings(max_examples=300, deadline=None)
@given(weights, st.integers(min_value=1, max_value=96))
def test_split_chunks_partitions_and_follows_the_weights(w, nchunk):
    counts = split_chunks(w, len(w), nchunk)
    assert sum(counts) == nchunk and all(c >= 0 for c in counts)
    total = sum(w)
    for c, x in zip(counts, w):
        assert abs(c - x / total * nchunk) < 1.0 + 1e-9          # largest remainder: off by < 1 chunk
        if x == 0.0:
            assert c == 0                                          # a dead link gets nothing
    ranges = chunk_ranges(counts)
    assert ranges[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert ranges[-1][0] + ranges[-1][1] == nchunk


@settings(max_examples=300, deadline=None)
@given(weights, st.integers(min_value=1, max_value=12), st.sampled_from([6, 48]))
def test_plan_units_covers_every_unit_exactly_once(w, nbeams, nchunk):
    plan = plan_units(w, nbeams, nchunk)
    assert len(plan) == len(w)
    owner = -np.ones((nbeams, nchunk), dtype=int)
    flat = []
    for r, items in enumerate(plan):
        for beam, first, n in items:
            assert n >= 1 and 0 <= first and first + n <= nchunk and 0 <= beam < nbeams
            assert (owner[beam, first:first + n] == -1).all()     # disjoint
            owner[beam, first:first + n] = r
            flat.append((beam * nchunk + first, n, r))
    assert (owner >= 0).all()                                      # complete
    flat.sort()
    assert all(a[0] + a[1] == b[0] for a, b in zip(flat, flat[1:]))   # beam-major, consecutive runs
    assert [x[2] for x in flat] == sorted(x[2] for x in flat)          # shard r's run precedes shard r+1's
    total, units = sum(w), nbeams * nchunk
    for r, items in enumerate(plan):
        got = sum(n for _, _, n in items)
        assert abs(got - w[r] / total * units) <= 1.0 + 1e-9        # rounded cumulative bounds


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=64), st.integers(min_value=1, max_value=16))
def test_beams_round_robin_is_a_partition(nbeam, world):
    seen = []
    for r in range(world):
        mine = beams_for_rank(nbeam, r, world)
        assert all(rank_of_beam(b, world) == r for b in mine)
        seen += mine
    assert sorted(seen) == list(range(nbeam))
    keys = [k for b in range(nbeam) for kk in ring_keys_for_beam(b) for k in (kk, kk + 1)]
    assert len(set(keys)) == len(keys) or nbeam > (0xDADA - 0xADAD) // 0x10   # in/out key ranges do not meet


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=16), st.integers(min_value=1, max_value=16))
def test_spread_placement_uses_distinct_gpus_when_it_can(world, ngpus):
    gpus = [gpu_for_rank(r, world, ngpus, "spread") for r in range(world)]
    assert all(0 <= g < ngpus for g in gpus)
    if world <= ngpus:
        assert len(set(gpus)) == world and gpus == sorted(gpus)
        assert gpus[0] == 0
    ident = [gpu_for_rank(r, world, ngpus, "identity") for r in range(world)]
    assert ident == [r % ngpus for r in range(world)]
