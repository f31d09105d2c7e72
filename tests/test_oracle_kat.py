"""The oracle against the hand-computable KATs, its numpy twin, and the order-independence claim.

PARITY UNPINNED: the reference holds no implementation or vectors for this path
(kernel.cu:1-7) — these closed forms stand in for them.
"""
import numpy as np
import pytest

from oracle import b2p_oracle_np as onp
from tests import kat


@pytest.mark.parametrize("big_endian", [True, False])
def test_kats_c_and_numpy(oracle_mod, big_endian):
    g = oracle_mod.Geometry(big_endian=big_endian)
    for name, (block, want) in kat.all_kats(ndf=3, big_endian=big_endian).items():
        got_c = oracle_mod.accumulate(block, g=g)
        got_np = onp.channel_sums(block, big_endian=big_endian)
        assert np.array_equal(got_c, want), name
        assert np.array_equal(got_np, want), name
        assert np.array_equal(oracle_mod.accumulate_omp(block, g=g, nthreads=3), want), name


def test_wrong_endianness_is_detected(oracle_mod):
    block, want = kat.kat_ones(ndf=2)
    wrong = oracle_mod.accumulate(block, g=oracle_mod.Geometry(big_endian=False))
    assert np.array_equal(wrong, want * np.uint64(256 * 256))


def test_full_scale_is_2_pow_52(oracle_mod):
    """One full integration (2^20 samples) of -32768 is exactly 2^52 per channel; checked
    arithmetically from a short block (linearity), then the float32 conversion."""
    block, want = kat.kat_min(ndf=2)
    got = oracle_mod.accumulate(block)
    per_frame = got // np.uint64(2)
    full = per_frame * np.uint64(8192)
    assert np.all(full == np.uint64(1 << 52))
    assert np.all(oracle_mod.finish(full, 1.0) == np.float32(2.0 ** 52))
    assert np.all(oracle_mod.finish(full, 2.0 ** -20) == np.float32(2.0 ** 32))


@pytest.mark.parametrize("geom", [(2, 3, 4), (5, 7, 16), (48, 7, 128), (3, 1, 2), (4, 32, 6)])
def test_random_geometries_c_vs_numpy(oracle_mod, geom):
    nchunk, nch, nsamp = geom
    g = oracle_mod.Geometry(nchunk=nchunk, nch_per_chunk=nch, nsamp_df=nsamp)
    rng = np.random.default_rng(1234 + nchunk)
    ndf = 5
    x = rng.integers(-32768, 32768, size=ndf * g.frame_bytes // 2, dtype=np.int64).astype(">i2")
    block = x.view(np.uint8)
    assert np.array_equal(oracle_mod.accumulate(block, g=g),
                          onp.channel_sums(block, nchunk=nchunk, nch=nch, nsamp=nsamp))


@pytest.mark.parametrize("mode", [0, 1])
def test_synth_stream_c_vs_numpy(oracle_mod, mode):
    a = oracle_mod.synth_fill(3, seed=99, first_word=12345, mode=mode)
    b = onp.synth_block(3, 99, 12345, mode)
    assert np.array_equal(a, b)
    # a block is a pure function of (seed, absolute word index): splitting it changes nothing
    g = oracle_mod.Geometry()
    wpf = g.frame_bytes // 8
    tail = oracle_mod.synth_fill(2, seed=99, first_word=12345 + wpf, mode=mode)
    assert np.array_equal(a[g.frame_bytes:], tail)


def test_order_independence_and_float_modes(oracle_mod):
    block = oracle_mod.synth_fill(8, seed=5, mode=1)
    exact = oracle_mod.accumulate(block)
    f64 = oracle_mod.accumulate_f64(block)
    assert np.array_equal(f64.astype(np.uint64), exact)  # double accumulation is exact here
    # accumulate in two calls == one call (integration spans calls)
    g = oracle_mod.Geometry()
    part = oracle_mod.accumulate(block[: 3 * g.frame_bytes])
    part = oracle_mod.accumulate(block[3 * g.frame_bytes:], sums=part)
    assert np.array_equal(part, exact)
    # a careless fp32 running sum drifts; the spec's float-mode bound is 1e-6
    naive = oracle_mod.accumulate_f32_naive(block)
    rel = np.abs(naive.astype(np.float64) - exact.astype(np.float64)) / exact.astype(np.float64)
    assert rel.max() < 1e-3


def test_finish_scale(oracle_mod):
    sums = np.array([0, 1, (1 << 24) + 1, (1 << 52), 123456789012345], dtype=np.uint64)
    out = oracle_mod.finish(sums, 1.0)
    assert np.array_equal(out, sums.astype(np.float32))
    mean = oracle_mod.finish(sums, 2.0 ** -20)
    assert np.array_equal(mean, (sums.astype(np.float32) * np.float32(2.0 ** -20)))
    assert np.array_equal(onp.finish(sums, 2.0 ** -20), mean)
