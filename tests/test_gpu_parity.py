"""Parity of the CUDA path (through the C ABI) against the CPU oracle — bit-exact.

Integer sums are compared as uint64, the float32 output as raw bit patterns.
Float mode: <= 1e-6 relative to the exact value (BASELINE.json north_star).
"""
import numpy as np
import pytest

from tests import kat

pytestmark = pytest.mark.gpu

KERNELS = ["ldg", "tma"]


def _run_device(b2p, block, ndf, kernel="ldg", nbeam=1, **kw):
    st = b2p.Baseband2Power(kernel=kernel, nbeam=nbeam, **kw)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    per = block.nbytes // nbeam
    st.accumulate_device([dev.ptr + b * per for b in range(nbeam)], ndf)
    sums = st.read_sums() if kw.get("mode", "exact") == "exact" else None
    out = st.finish()
    dev.free()
    st.close()
    return sums, out


def test_unpack_all_65536_patterns(b2p):
    v = np.arange(65536, dtype=np.uint32)
    be = ((v & 0xFF) << 8 | (v >> 8)).astype(np.uint16).view(np.int16).astype(np.int32)
    le = v.astype(np.uint16).view(np.int16).astype(np.int32)
    assert np.array_equal(b2p.selftest_unpack(0, True), be)
    assert np.array_equal(b2p.selftest_unpack(0, False), le)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("big_endian", [True, False])
def test_kats(b2p, oracle_mod, kernel, big_endian):
    for name, (block, want) in kat.all_kats(ndf=5, big_endian=big_endian).items():
        sums, out = _run_device(b2p, block, 5, kernel, big_endian=big_endian)
        assert np.array_equal(sums[0], want), (kernel, name)
        assert np.array_equal(out[0].view(np.uint32), oracle_mod.finish(want).view(np.uint32)), name


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ndf", [1, 2, 7, 8, 9, 37, 64, 300, 1000])
@pytest.mark.parametrize("mode", [0, 1])
def test_synth_blocks_bit_exact(b2p, oracle_mod, kernel, ndf, mode):
    block = oracle_mod.synth_fill(ndf, seed=20240517 + ndf, mode=mode)
    want = oracle_mod.accumulate_omp(block)
    sums, out = _run_device(b2p, block, ndf, kernel)
    assert np.array_equal(sums[0], want)
    assert np.array_equal(out[0].view(np.uint32), oracle_mod.finish(want).view(np.uint32))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("nsplit", [1, 3, 37, 200])
def test_nsplit_does_not_change_the_answer(b2p, oracle_mod, kernel, nsplit):
    block = oracle_mod.synth_fill(50, seed=3, mode=0)
    want = oracle_mod.accumulate_omp(block)
    sums, _ = _run_device(b2p, block, 50, kernel, nsplit=nsplit)
    assert np.array_equal(sums[0], want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_scale_mean(b2p, oracle_mod, kernel):
    block = oracle_mod.synth_fill(16, seed=8, mode=1)
    want = oracle_mod.finish(oracle_mod.accumulate(block), 2.0 ** -20)
    _, out = _run_device(b2p, block, 16, kernel, scale=2.0 ** -20)
    assert np.array_equal(out[0].view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("kernel", KERNELS)
def test_multibeam_batched(b2p, oracle_mod, kernel):
    nbeam, ndf = 5, 24
    g = oracle_mod.Geometry()
    blocks = [oracle_mod.synth_fill(ndf, seed=100 + b, mode=b % 2) for b in range(nbeam)]
    sums, out = _run_device(b2p, np.concatenate(blocks), ndf, kernel, nbeam=nbeam)
    for b in range(nbeam):
        want = oracle_mod.accumulate_omp(blocks[b], g=g)
        assert np.array_equal(sums[b], want), b
        assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(want).view(np.uint32))


@pytest.mark.parametrize("kernel", KERNELS)
def test_integration_spans_calls_and_resets(b2p, oracle_mod, kernel):
    g = oracle_mod.Geometry()
    block = oracle_mod.synth_fill(40, seed=77, mode=1)
    want = oracle_mod.accumulate_omp(block)
    st = b2p.Baseband2Power(kernel=kernel)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    for f0, n in [(0, 13), (13, 1), (14, 26)]:
        st.accumulate_device([dev.ptr + f0 * g.frame_bytes], n)
    assert np.array_equal(st.read_sums()[0], want)
    out = st.finish()[0]
    assert np.array_equal(out.view(np.uint32), oracle_mod.finish(want).view(np.uint32))
    assert not st.read_sums().any()  # finish resets the integration
    st.accumulate_device([dev], 40)
    assert np.array_equal(st.finish()[0].view(np.uint32), out.view(np.uint32))  # idempotent
    assert st.launch_count > 0
    dev.free()
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("pinned", [True, False])
def test_host_path_staged(b2p, oracle_mod, kernel, pinned):
    ndf = 70
    block = oracle_mod.synth_fill(ndf, seed=31, mode=1)
    want = oracle_mod.accumulate_omp(block)
    st = b2p.Baseband2Power(kernel=kernel, stage_ndf=16, nstage_bufs=2)
    if pinned:
        pb = b2p.PinnedBuffer(block.nbytes)
        pb.array[:] = block
        src = pb
    else:
        src = block
    st.accumulate_host([src], ndf)
    assert np.array_equal(st.read_sums()[0], want)
    st.accumulate_host([src], ndf)   # second block of the same integration
    out = st.finish()[0]
    assert np.array_equal(out.view(np.uint32), oracle_mod.finish(want * np.uint64(2)).view(np.uint32))
    if pinned:
        st.accumulate_host_mapped([src], ndf)
        assert np.array_equal(st.read_sums()[0], want)
        pb.free()
    st.close()


@pytest.mark.parametrize("geom", [(2, 3, 4), (5, 7, 16), (3, 1, 2), (4, 32, 6), (6, 8, 128), (48, 7, 64)])
def test_other_geometries(b2p, oracle_mod, geom):
    nchunk, nch, nsamp = geom
    g = oracle_mod.Geometry(nchunk=nchunk, nch_per_chunk=nch, nsamp_df=nsamp)
    ndf = 11
    block = oracle_mod.synth_fill(ndf, seed=5, mode=0, g=g)
    want = oracle_mod.accumulate(block, g=g)
    sums, _ = _run_device(b2p, block, ndf, "ldg", nchunk=nchunk, nch_per_chunk=nch, nsamp_df=nsamp)
    assert np.array_equal(sums[0], want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_float_mode_within_1e6(b2p, oracle_mod, kernel):
    block = oracle_mod.synth_fill(256, seed=9, mode=1)
    exact = oracle_mod.accumulate_omp(block).astype(np.float64)
    _, out = _run_device(b2p, block, 256, kernel, mode="float")
    rel = np.abs(out[0].astype(np.float64) - exact) / exact
    assert rel.max() <= 1e-6, rel.max()   # tolerance stated by BASELINE.json north_star


def test_device_synth_matches_host_generator(b2p, oracle_mod):
    g = oracle_mod.Geometry()
    for mode in (0, 1):
        host = oracle_mod.synth_fill(3, seed=42, first_word=987654321, mode=mode)
        dev = b2p.DeviceBuffer(host.nbytes)
        dev.synth_fill(3, seed=42, first_word=987654321, mode=mode)
        assert np.array_equal(dev.download(), host)
        dev.free()


def test_bad_arguments_fail_loudly(b2p):
    with pytest.raises(b2p.B2pError):
        b2p.Baseband2Power(nbeam=0)
    with pytest.raises(b2p.B2pError):
        b2p.Baseband2Power(nch_per_chunk=3, nsamp_df=1)  # packet not a multiple of 16 B
    st = b2p.Baseband2Power()
    with pytest.raises(b2p.B2pError):
        st.accumulate_device([0], 4)       # NULL pointer
    with pytest.raises(b2p.B2pError):
        st.accumulate_device([8], 4)       # misaligned
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_chained_integrations_do_not_race(b2p, oracle_mod, kernel):
    """Back-to-back integrations on the context's own stream: kernels are PDL-chained and a
    fused kernel starts under the tail of its predecessor; every spectrum must still be exact."""
    g = oracle_mod.Geometry()
    nblk, ndf, nrun = 5, 300, 30
    blocks = [oracle_mod.synth_fill(ndf, seed=500 + i, mode=i % 2) for i in range(nblk)]
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    dev = [b2p.DeviceBuffer(b.nbytes) for b in blocks]
    for d, b in zip(dev, blocks):
        d.upload(b)
    outs = b2p.DeviceBuffer(nrun * g.nchan * 4)
    st = b2p.Baseband2Power(kernel=kernel)
    for i in range(nrun):
        st.accumulate_device([dev[i % nblk]], ndf)
        st.finish_device(outs.ptr + i * g.nchan * 4)
    b2p.device_sync(0)   # the context stream is non-blocking: a legacy-stream copy does not wait for it
    got = outs.download().view(np.float32).reshape(nrun, g.nchan)
    for i in range(nrun):
        assert np.array_equal(got[i].view(np.uint32), want[i % nblk].view(np.uint32)), i
    # two accumulates per integration as well (fold kernel between them)
    for i in range(nrun):
        st.accumulate_device([dev[i % nblk]], ndf)
        st.accumulate_device([dev[(i + 1) % nblk]], ndf)
        st.finish_device(outs.ptr + i * g.nchan * 4)
    b2p.device_sync(0)
    got = outs.download().view(np.float32).reshape(nrun, g.nchan)
    for i in range(nrun):
        s = oracle_mod.accumulate_omp(blocks[i % nblk]) + oracle_mod.accumulate_omp(blocks[(i + 1) % nblk])
        assert np.array_equal(got[i].view(np.uint32), oracle_mod.finish(s).view(np.uint32)), i
    st.close()
    for d in dev:
        d.free()
    outs.free()


@pytest.mark.parametrize("kernel", KERNELS)
def test_caller_stream(b2p, oracle_mod, kernel):
    """A caller-owned (torch) stream: producer kernel -> our kernels -> consumer on one stream."""
    torch = pytest.importorskip("torch")
    g = oracle_mod.Geometry()
    ndf = 64
    block = oracle_mod.synth_fill(ndf, seed=77, mode=1)
    want = oracle_mod.finish(oracle_mod.accumulate_omp(block))
    s = torch.cuda.Stream()
    src = torch.from_numpy(block).cuda()
    with torch.cuda.stream(s):
        for _ in range(5):
            buf = src.clone()                         # produced on the same stream right before
            out = torch.empty(g.nchan, dtype=torch.float32, device="cuda")
            st = b2p.Baseband2Power(kernel=kernel)
            st.accumulate_device([buf], ndf, s.cuda_stream)
            st.finish_device(out, s.cuda_stream)
            res = out.cpu().numpy()
            assert np.array_equal(res.view(np.uint32), want.view(np.uint32))
            st.close()


def test_empty_inputs(b2p, oracle_mod):
    st = b2p.Baseband2Power()
    assert not st.finish().any()                      # finish with nothing accumulated -> zeros
    dev = b2p.DeviceBuffer(48 * 7168)
    st.accumulate_device([dev], 0)                    # zero frames is a no-op, not an error
    assert not st.read_sums().any()
    assert st.launch_count == 1                       # only the finish above launched anything
    st.reset()
    dev.free()
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_block_bit_exact(b2p, oracle_mod, kernel):
    """One full ring block (8192 frames, 2 818 572 288 B) against the oracle, device and host paths."""
    ndf = 8192
    block = oracle_mod.synth_fill(ndf, seed=4242, mode=1)
    want = oracle_mod.accumulate_omp(block, nthreads=len(__import__("os").sched_getaffinity(0)))
    st = b2p.Baseband2Power(kernel=kernel)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.synth_fill(ndf, seed=4242, mode=1)            # device generator, same stream
    st.accumulate_device([dev], ndf)
    assert np.array_equal(st.read_sums()[0], want)
    out = st.finish()[0]
    assert np.array_equal(out.view(np.uint32), oracle_mod.finish(want).view(np.uint32))
    st.accumulate_host([block], ndf)                  # pageable host memory, staged path
    assert np.array_equal(st.finish()[0].view(np.uint32), out.view(np.uint32))
    dev.free()
    st.close()


def test_full_scale_full_block_is_2_pow_52(b2p):
    """Every component -32768 over a full integration: each channel is exactly 2^52 (the largest
    value the spec can produce) — no overflow anywhere in the chain; mean scaling gives 2^32."""
    ndf = 8192
    dev = b2p.DeviceBuffer(ndf * 48 * 7168)
    piece = np.tile(np.array([0x80, 0x00], dtype=np.uint8), 64 * 48 * 7168 // 2)   # 64 frames of 0x8000
    for i in range(ndf // 64):
        dev.upload(piece, offset=i * piece.nbytes)
    for kernel in KERNELS:
        st = b2p.Baseband2Power(kernel=kernel)
        st.accumulate_device([dev], ndf)
        assert np.all(st.read_sums() == np.uint64(1 << 52))
        assert np.all(st.finish() == np.float32(2.0 ** 52))
        st.close()
        st = b2p.Baseband2Power(kernel=kernel, scale=2.0 ** -20)
        st.accumulate_device([dev], ndf)
        assert np.all(st.finish() == np.float32(2.0 ** 32))
        st.close()
    dev.free()


def test_maximum_beam_count(b2p, oracle_mod):
    nbeam, ndf = 64, 4
    blocks = [oracle_mod.synth_fill(ndf, seed=b, mode=0) for b in range(nbeam)]
    sums, _ = _run_device(b2p, np.concatenate(blocks), ndf, "ldg", nbeam=nbeam)
    for b in (0, 1, 31, 63):
        assert np.array_equal(sums[b], oracle_mod.accumulate(blocks[b])), b
    with pytest.raises(b2p.B2pError):
        b2p.Baseband2Power(nbeam=65)


@pytest.mark.parametrize("kernel", KERNELS)
def test_mixed_host_and_device_calls_multibeam(b2p, oracle_mod, kernel):
    """One integration fed by host blocks, device blocks and mapped host blocks for 3 beams."""
    nbeam, ndf = 3, 40
    g = oracle_mod.Geometry()
    hb = [oracle_mod.synth_fill(ndf, seed=600 + b, mode=1) for b in range(nbeam)]
    db = [oracle_mod.synth_fill(ndf, seed=700 + b, mode=0) for b in range(nbeam)]
    st = b2p.Baseband2Power(kernel=kernel, nbeam=nbeam, stage_ndf=16)
    dev = b2p.DeviceBuffer(nbeam * db[0].nbytes)
    dev.upload(np.concatenate(db))
    pins = [b2p.PinnedBuffer(hb[b].nbytes) for b in range(nbeam)]
    for b in range(nbeam):
        pins[b].array[:] = hb[b]
    st.accumulate_host(hb, ndf)
    st.accumulate_device([dev.ptr + b * db[0].nbytes for b in range(nbeam)], ndf)
    st.accumulate_host_mapped(pins, ndf)
    sums = st.read_sums()
    out = st.finish()
    for b in range(nbeam):
        want = oracle_mod.accumulate(hb[b]) * np.uint64(2) + oracle_mod.accumulate(db[b])
        assert np.array_equal(sums[b], want), b
        assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(want).view(np.uint32))
    for p in pins:
        p.free()
    dev.free()
    st.close()


def test_float_mode_spans_calls(b2p, oracle_mod):
    block = oracle_mod.synth_fill(96, seed=12, mode=1)
    exact = oracle_mod.accumulate_omp(block).astype(np.float64)
    g = oracle_mod.Geometry()
    st = b2p.Baseband2Power(mode="float")
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    for f0, n in [(0, 32), (32, 1), (33, 63)]:
        st.accumulate_device([dev.ptr + f0 * g.frame_bytes], n)
    out = st.finish()[0].astype(np.float64)
    assert (np.abs(out - exact) / exact).max() <= 1e-6
    with pytest.raises(b2p.B2pError):
        st.read_sums()                      # exact-mode accessor
    dev.free()
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_cuda_graph_capture_and_replay(b2p, oracle_mod, kernel):
    """The hot path captured once into a CUDA graph (fused + finish kernel, PDL edge included) and
    replayed on new data: no host-side state is needed between replays (the TMA kernel's work
    counter is put back by the reduce kernel)."""
    torch = pytest.importorskip("torch")
    g = oracle_mod.Geometry()
    ndf = 96
    blocks = [oracle_mod.synth_fill(ndf, seed=800 + i, mode=i % 2) for i in range(3)]
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    s = torch.cuda.Stream()
    buf = torch.from_numpy(blocks[0]).cuda()
    out = torch.zeros(g.nchan, dtype=torch.float32, device="cuda")
    st = b2p.Baseband2Power(kernel=kernel)
    st.accumulate_device([buf], ndf, s.cuda_stream)          # warm-up outside the capture
    st.finish_device(out, s.cuda_stream)
    s.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        st.accumulate_device([buf], ndf, torch.cuda.current_stream().cuda_stream)
        st.finish_device(out, torch.cuda.current_stream().cuda_stream)
    for rep in range(6):
        i = rep % 3
        buf.copy_(torch.from_numpy(blocks[i]).cuda())
        out.zero_()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want[i].view(np.uint32)), (kernel, rep)
    st.close()


# ------------------------------------------------------------------ round 2
# one launch per integration, channel-group shards, asynchronous host path


@pytest.mark.parametrize("kernel", KERNELS)
def test_integrate_device_is_one_launch(b2p, oracle_mod, kernel):
    """b2p_integrate_device: accumulate + cross-CTA reduce + finish inside a single kernel."""
    g = oracle_mod.Geometry()
    ndf = 300
    block = oracle_mod.synth_fill(ndf, seed=901, mode=1)
    want_s = oracle_mod.accumulate_omp(block)
    st = b2p.Baseband2Power(kernel=kernel)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    out = b2p.DeviceBuffer(g.nchan * 4)
    n0 = st.launch_count
    st.integrate_device([dev], ndf, out)
    assert st.launch_count == n0 + 1
    b2p.device_sync(0)
    got = out.download().view(np.float32)
    assert np.array_equal(got.view(np.uint32), oracle_mod.finish(want_s).view(np.uint32))
    assert not st.read_sums().any()                   # the integration was closed and cleared
    # earlier accumulate calls are folded into the integration the launch closes
    st.accumulate_device([dev], 100)
    st.accumulate_device([dev.ptr + 100 * g.frame_bytes], 50)
    st.integrate_device([dev.ptr + 150 * g.frame_bytes], 150, out)
    b2p.device_sync(0)
    got = out.download().view(np.float32)
    assert np.array_equal(got.view(np.uint32), oracle_mod.finish(want_s).view(np.uint32))
    dev.free()
    out.free()
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_chained_single_launch_integrations(b2p, oracle_mod, kernel):
    """60 integrations back to back, one PDL-chained launch each, results kept on the device."""
    g = oracle_mod.Geometry()
    nblk, ndf, nrun = 4, 260, 60
    blocks = [oracle_mod.synth_fill(ndf, seed=930 + i, mode=i % 2) for i in range(nblk)]
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    dev = [b2p.DeviceBuffer(b.nbytes) for b in blocks]
    for d, b in zip(dev, blocks):
        d.upload(b)
    outs = b2p.DeviceBuffer(nrun * g.nchan * 4)
    st = b2p.Baseband2Power(kernel=kernel)
    n0 = st.launch_count
    for i in range(nrun):
        st.integrate_device([dev[i % nblk]], ndf, outs.ptr + i * g.nchan * 4)
    assert st.launch_count == n0 + nrun
    b2p.device_sync(0)
    got = outs.download().view(np.float32).reshape(nrun, g.nchan)
    for i in range(nrun):
        assert np.array_equal(got[i].view(np.uint32), want[i % nblk].view(np.uint32)), i
    st.close()
    for d in dev:
        d.free()
    outs.free()


@pytest.mark.parametrize("kernel", KERNELS)
def test_multibeam_integrate_with_float_and_scale(b2p, oracle_mod, kernel):
    nbeam, ndf = 3, 40
    g = oracle_mod.Geometry()
    blocks = [oracle_mod.synth_fill(ndf, seed=940 + b, mode=1) for b in range(nbeam)]
    dev = b2p.DeviceBuffer(nbeam * blocks[0].nbytes)
    dev.upload(np.concatenate(blocks))
    out = b2p.DeviceBuffer(nbeam * g.nchan * 4)
    for mode in ("exact", "float"):
        st = b2p.Baseband2Power(kernel=kernel, nbeam=nbeam, mode=mode, scale=2.0 ** -12)
        st.integrate_device([dev.ptr + b * blocks[0].nbytes for b in range(nbeam)], ndf, out)
        b2p.device_sync(0)
        got = out.download().view(np.float32).reshape(nbeam, g.nchan)
        for b in range(nbeam):
            s = oracle_mod.accumulate(blocks[b])
            if mode == "exact":
                assert np.array_equal(got[b].view(np.uint32), oracle_mod.finish(s, 2.0 ** -12).view(np.uint32))
            else:
                ref = s.astype(np.float64) * 2.0 ** -12
                assert (np.abs(got[b] - ref) / ref).max() <= 1e-6
        st.close()
    dev.free()
    out.free()


def _shard_ranges(counts):
    first = 0
    for n in counts:
        if n:
            yield first, n
        first += n


SPLITS = [[6] * 8, [5, 5, 5, 5, 7, 7, 7, 7], [48], [1, 47], [13, 0, 35], [3, 9, 11, 25]]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("counts", SPLITS)
def test_chunk_group_shards_concatenate_bit_exact(b2p, oracle_mod, kernel, counts):
    """Every shard reads its chunk columns out of the full-frame block (device-strided, staged
    2-D H2D, and zero-copy mapped); the channel ranges put side by side equal the oracle's
    spectrum of the whole block."""
    g = oracle_mod.Geometry()
    ndf = 70
    block = oracle_mod.synth_fill(ndf, seed=950, mode=1)
    want = oracle_mod.accumulate_omp(block)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    pin = b2p.PinnedBuffer(block.nbytes)
    pin.array[:] = block
    got_dev = np.zeros(g.nchan, dtype=np.uint64)
    got_host = np.zeros(g.nchan, dtype=np.float32)
    got_map = np.zeros(g.nchan, dtype=np.uint64)
    for first, n in _shard_ranges(counts):
        st = b2p.Baseband2Power(kernel=kernel, nchunk=n, first_chunk=first, nchunk_total=48, stage_ndf=16)
        assert st.nchan == 7 * n and st.source_frame_bytes == g.frame_bytes
        sl = slice(7 * first, 7 * (first + n))
        st.accumulate_device([dev], ndf)              # pointer = frame 0 of the full stream
        got_dev[sl] = st.read_sums()[0]
        st.reset()
        got_host[sl] = st.integrate_host([pin], ndf)[0]
        st.accumulate_host_mapped([pin], ndf)
        got_map[sl] = st.read_sums()[0]
        st.close()
    assert np.array_equal(got_dev, want)
    assert np.array_equal(got_map, want)
    assert np.array_equal(got_host.view(np.uint32), oracle_mod.finish(want).view(np.uint32))
    pin.free()
    dev.free()


def test_shard_rejects_bad_ranges(b2p):
    with pytest.raises(b2p.B2pError):
        b2p.Baseband2Power(nchunk=6, first_chunk=44, nchunk_total=48)
    with pytest.raises(b2p.B2pError):
        b2p.Baseband2Power(nchunk=6, first_chunk=-1, nchunk_total=48)
    with pytest.raises(b2p.B2pError):
        b2p.ShardGroup([0, 0], [6, 6])                # does not add up to 48


@pytest.mark.parametrize("counts", [[6] * 8, [5, 5, 5, 5, 7, 7, 7, 7], [48, 0], [20, 28]])
def test_shard_group_on_one_gpu(b2p, oracle_mod, counts):
    """b2p_group_*: issue on every shard, wait afterwards; here all shards share GPU 0."""
    nbeam, ndf = 2, 90
    blocks = [oracle_mod.synth_fill(ndf, seed=960 + b, mode=b % 2) for b in range(nbeam)]
    pins = [b2p.PinnedBuffer(b.nbytes) for b in blocks]
    for p, b in zip(pins, blocks):
        p.array[:] = b
    want = [oracle_mod.accumulate_omp(b) for b in blocks]
    grp = b2p.ShardGroup([0] * len(counts), counts, nbeam=nbeam)
    assert [(f, n) for _, f, n in grp.shards] == list(_shard_ranges(counts))
    out = grp.integrate_host(pins, ndf)
    for b in range(nbeam):
        assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(want[b]).view(np.uint32)), b
    # an integration over two blocks: accumulate, then integrate closes it
    grp.accumulate_host(pins, ndf)
    out = grp.integrate_host(pins[::-1], ndf)
    for b in range(nbeam):
        s = want[b] + want[nbeam - 1 - b]
        assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(s).view(np.uint32)), b
    grp.accumulate_host(pins, ndf)
    with pytest.raises(b2p.B2pError):
        grp.rebalance()                               # only between integrations
    out = grp.finish()
    for b in range(nbeam):
        assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(want[b]).view(np.uint32)), b
    # measured rebalancing moves chunks between shards; whatever split it lands on, the
    # spectrum stays the oracle's
    for _ in range(4):
        grp.rebalance()
        assert sum(n for _, _, n in grp.shards) == 48
        out = grp.integrate_host(pins, ndf)
        for b in range(nbeam):
            assert np.array_equal(out[b].view(np.uint32), oracle_mod.finish(want[b]).view(np.uint32)), b
    grp.close()
    for p in pins:
        p.free()


def test_split_chunks_and_probe(b2p):
    assert b2p.split_chunks(None, 8) == [6] * 8
    assert b2p.split_chunks([23, 23, 23, 23, 35, 35, 35, 35], 8) == [5, 5, 5, 5, 7, 7, 7, 7]
    assert sum(b2p.split_chunks([1.0, 2.5, 0.0, 3.1], 4)) == 48
    assert b2p.split_chunks([1.0, 2.5, 0.0, 3.1], 4)[2] == 0
    with pytest.raises(b2p.B2pError):
        b2p.split_chunks([0.0, 0.0], 2)
    rates = b2p.probe_h2d([0], nbytes=64 << 20, reps=2)
    assert len(rates) == 1 and rates[0] > 1.0         # GB/s; any real link is far above this


@pytest.mark.parametrize("kernel", KERNELS)
def test_async_host_path(b2p, oracle_mod, kernel):
    """accumulate_host_async / wait_input / wait_output on two contexts driven by one thread."""
    ndf = 64
    blocks = [oracle_mod.synth_fill(ndf, seed=970 + i, mode=1) for i in range(2)]
    pins = [b2p.PinnedBuffer(b.nbytes) for b in blocks]
    for p, b in zip(pins, blocks):
        p.array[:] = b
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    sts = [b2p.Baseband2Power(kernel=kernel, stage_ndf=8, nstage_bufs=2) for _ in range(2)]
    for rep in range(3):
        for st, p in zip(sts, pins):
            st.accumulate_host_async([p], ndf, finish=True)
        for st in sts:
            st.wait_input()
        for i, st in enumerate(sts):
            assert np.array_equal(st.wait_output()[0].view(np.uint32), want[i].view(np.uint32)), (rep, i)
    with pytest.raises(b2p.B2pError):
        sts[0].wait_output()                          # nothing queued any more
    for st in sts:
        st.close()
    for p in pins:
        p.free()


def test_reset_mid_integration_then_many_tma_launches(b2p, oracle_mod):
    """a reset that drops work must leave every device-side counter clean — run past
    the ticket ring (4096 launches) afterwards and compare with the oracle."""
    g = oracle_mod.Geometry()
    ndf = 8
    block = oracle_mod.synth_fill(ndf, seed=980, mode=1)
    want = oracle_mod.finish(oracle_mod.accumulate(block))
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    out = b2p.DeviceBuffer(g.nchan * 4)
    st = b2p.Baseband2Power(kernel="tma")
    st.accumulate_device([dev], ndf)
    st.reset()
    for i in range(4200):
        st.integrate_device([dev], ndf, out)
        if i in (0, 4095, 4096, 4199):
            b2p.device_sync(0)
            got = out.download().view(np.float32)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), i
    st.close()
    dev.free()
    out.free()


@pytest.mark.parametrize("kernel", KERNELS)
def test_launches_on_different_streams_are_ordered(b2p, oracle_mod, kernel):
    """accumulate on one stream, finish on another, accumulate again on the context's
    own stream — the library orders each launch behind the previous one."""
    torch = pytest.importorskip("torch")
    g = oracle_mod.Geometry()
    ndf = 500
    block = oracle_mod.synth_fill(ndf, seed=990, mode=1)
    want = oracle_mod.finish(oracle_mod.accumulate_omp(block))
    buf = torch.from_numpy(block).cuda()
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = torch.zeros(20, g.nchan, dtype=torch.float32, device="cuda")
    st = b2p.Baseband2Power(kernel=kernel)
    for i in range(20):
        st.accumulate_device([buf], ndf, sa.cuda_stream)
        st.finish_device(outs[i], sb.cuda_stream)
        st.accumulate_device([buf], ndf, None)
        st.finish_device(outs[i], sa.cuda_stream)
    torch.cuda.synchronize()
    b2p.device_sync(0)
    res = outs.cpu().numpy()
    for i in range(20):
        assert np.array_equal(res[i].view(np.uint32), want.view(np.uint32)), i
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ndf", [16, 255, 256, 257, 1024, 4100])
def test_piece_shapes(b2p, oracle_mod, kernel, ndf):
    """Launch shapes of the host path (whole-wave split counts for short pieces)."""
    block = oracle_mod.synth_fill(ndf, seed=ndf, mode=0)
    want = oracle_mod.accumulate_omp(block)
    st = b2p.Baseband2Power(kernel=kernel)
    dev = b2p.DeviceBuffer(block.nbytes)
    dev.upload(block)
    st.accumulate_device([dev], ndf)
    assert np.array_equal(st.read_sums()[0], want)
    dev.free()
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_cuda_graph_of_single_launch_integration(b2p, oracle_mod, kernel):
    torch = pytest.importorskip("torch")
    g = oracle_mod.Geometry()
    ndf = 96
    blocks = [oracle_mod.synth_fill(ndf, seed=860 + i, mode=i % 2) for i in range(3)]
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    s = torch.cuda.Stream()
    buf = torch.from_numpy(blocks[0]).cuda()
    out = torch.zeros(2, g.nchan, dtype=torch.float32, device="cuda")
    st = b2p.Baseband2Power(kernel=kernel)
    st.integrate_device([buf], ndf, out[0], s.cuda_stream)
    s.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        cs = torch.cuda.current_stream().cuda_stream
        st.integrate_device([buf], ndf, out[0], cs)   # two integrations per replay
        st.integrate_device([buf], ndf, out[1], cs)
    for rep in range(6):
        i = rep % 3
        buf.copy_(torch.from_numpy(blocks[i]).cuda())
        out.zero_()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        res = out.cpu().numpy()
        for k in range(2):
            assert np.array_equal(res[k].view(np.uint32), want[i].view(np.uint32)), (kernel, rep, k)
    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_output_queue_runs_integrations_ahead(b2p, oracle_mod, kernel):
    """Up to 4 finished integrations may wait to be collected: a stage queues the next ring block
    before it collects the previous spectrum.  Spectra come back oldest first."""
    ndf = 48
    blocks = [oracle_mod.synth_fill(ndf, seed=1100 + i, mode=1) for i in range(4)]
    pins = [b2p.PinnedBuffer(b.nbytes) for b in blocks]
    for p, b in zip(pins, blocks):
        p.array[:] = b
    want = [oracle_mod.finish(oracle_mod.accumulate_omp(b)) for b in blocks]
    st = b2p.Baseband2Power(kernel=kernel, stage_ndf=16)
    for p in pins:
        st.accumulate_host_async([p], ndf, finish=True)
    with pytest.raises(b2p.B2pError):
        st.accumulate_host_async([pins[0]], ndf, finish=True)     # a fifth: collect first
    with pytest.raises(b2p.B2pError):
        st.finish()                                               # not while spectra are queued
    st.reset()                                                    # drops the frames of the refused call and the queue
    for rep in range(2):                                          # two ahead, steady state
        st.accumulate_host_async([pins[0]], ndf, finish=True)
        for i in range(1, 4):
            st.accumulate_host_async([pins[i]], ndf, finish=True)
            st.wait_input()
            got = st.wait_output()[0]                             # spectrum of block i-1
            assert np.array_equal(got.view(np.uint32), want[i - 1].view(np.uint32)), (rep, i)
        assert np.array_equal(st.wait_output()[0].view(np.uint32), want[3].view(np.uint32))
    st.close()
    for p in pins:
        p.free()


@pytest.mark.parametrize("kernel", KERNELS)
def test_resizable_shard_moves_its_chunk_range_in_place(b2p, oracle_mod, kernel):
    """b2p_set_chunk_range: what b2p_group_rebalance does to follow the measured link rates."""
    g = oracle_mod.Geometry()
    ndf = 60
    block = oracle_mod.synth_fill(ndf, seed=1200, mode=1)
    want = oracle_mod.finish(oracle_mod.accumulate_omp(block))
    pin = b2p.PinnedBuffer(block.nbytes)
    pin.array[:] = block
    st = b2p.Baseband2Power(kernel=kernel, nchunk=6, first_chunk=0, nchunk_total=48, resizable=True)
    for first, n in [(0, 6), (6, 1), (7, 41), (0, 48), (47, 1), (10, 13)]:
        st.set_chunk_range(first, n)
        assert st.nchan == 7 * n
        got = st.integrate_host([pin], ndf)[0]
        assert np.array_equal(got.view(np.uint32), want[7 * first:7 * (first + n)].view(np.uint32)), (first, n)
        st.accumulate_host_mapped([pin], ndf)          # and the zero-copy path on the new range
        sums = st.read_sums()[0]
        assert np.array_equal(oracle_mod.finish(sums).view(np.uint32), want[7 * first:7 * (first + n)].view(np.uint32))
        st.reset()
    with pytest.raises(b2p.B2pError):
        st.set_chunk_range(40, 9)                      # past the end of the frame
    st.close()
    fixed = b2p.Baseband2Power(kernel=kernel, nchunk=6, first_chunk=0, nchunk_total=48)
    with pytest.raises(b2p.B2pError):
        fixed.set_chunk_range(6, 6)                    # not created resizable
    fixed.close()
    pin.free()
