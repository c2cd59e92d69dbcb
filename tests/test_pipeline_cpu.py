"""Host-side logic of the block executor on CPU (system space): gulping with
mismatched sizes, input overlap, rare commits, header views -- the behaviours
the hot-path blocks rely on (reference: test/test_pipeline_cpu.py patterns)."""
from copy import deepcopy

import numpy as np

import bifrost_b200 as bf
from bifrost_b200.pipeline import TransformBlock, Pipeline
from bifrost_b200.blocks import array_source, callback_sink, copy


def header(shape, dtype='f32', labels=None):
    nd = len(shape)
    return {'_tensor': {'dtype': dtype, 'shape': list(shape),
                        'labels': labels or ['time'] + ['d%i' % i for i in range(1, nd)],
                        'scales': [[0, 1.0] for _ in range(nd)],
                        'units': ['s'] + [None] * (nd - 1)},
            'name': 'test', 'gulp_nframe': 1}


class Collect(object):
    def __init__(self):
        self.chunks, self.headers = [], []

    def seq(self, iseq):
        self.headers.append(deepcopy(iseq.header))

    def data(self, ispan):
        self.chunks.append(np.array(ispan.data.copy('system')))


def test_copy_chain_with_mixed_gulp_sizes():
    data = np.arange(50 * 3, dtype=np.float32).reshape(50, 3)
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, header([-1, 3]), gulp_nframe=7)
        b = copy(src, gulp_nframe=4)
        b = copy(b, gulp_nframe=9)
        callback_sink(b, out.seq, out.data, gulp_nframe=5)
        p.run()
    got = np.concatenate(out.chunks, axis=0)
    np.testing.assert_array_equal(got, data)
    assert out.headers[0]['_tensor']['shape'] == [-1, 3]


def test_small_gulps_after_a_large_span_copy_linearly(monkeypatch):
    """A gulp-1 consumer behind a 600-frame producer: spans are views of the
    ring, so the only frames ever copied inside the executor are the ghost
    copies at the wrap of each ring -- O(frames), not O(n^2)."""
    from bifrost_b200 import ring as rg
    copied = [0]
    real = rg.copy_array

    def counting(dst, src):
        copied[0] += int(np.prod(src.shape))
        return real(dst, src)

    monkeypatch.setattr(rg, 'copy_array', counting)
    data = np.arange(600 * 2, dtype=np.float32).reshape(600, 2)
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, header([-1, 2]), gulp_nframe=600)
        b = copy(src, gulp_nframe=1)
        callback_sink(b, out.seq, out.data, gulp_nframe=7)
        p.run()
    np.testing.assert_array_equal(np.concatenate(out.chunks, axis=0), data)
    assert copied[0] < 3 * data.size


def test_rings_wrap_through_their_ghost_region():
    """Many more frames than the ring holds, gulps that do not divide its
    capacity, an overlapping reader and a second reader on the same ring: every
    span must come back intact across the wrap (ghost copies in both
    directions), and the storage is allocated once per sequence."""
    data = np.arange(1000 * 3, dtype=np.float32).reshape(1000, 3)
    a, b = Collect(), Collect()
    with Pipeline() as p:
        with bf.block_scope(buffer_factor=2):
            src = array_source(data, header([-1, 3]), gulp_nframe=7)
            c = copy(src, gulp_nframe=5)
            m = MovingSum(c, gulp_nframe=11)
        callback_sink(m, a.seq, a.data, gulp_nframe=4)
        callback_sink(c, b.seq, b.data, gulp_nframe=13)
        p.run()
    np.testing.assert_array_equal(np.concatenate(b.chunks, axis=0), data)
    got = np.concatenate(a.chunks, axis=0)
    want = np.array([data[k:k + 4].sum(axis=0) for k in range(1000 - 3)])
    np.testing.assert_array_equal(got[:len(want)], want)
    for blk in (src, c, m):
        st = blk.orings[0].stats
        assert st['allocations'] == 1 and st['ghost_copies'] > 10


class MovingSum(TransformBlock):
    """Needs `overlap` frames of history, like FdmtBlock (blocks/fdmt.py:112-124)."""
    overlap = 3

    def on_sequence(self, iseq):
        return deepcopy(iseq.header)

    def define_input_overlap_nframe(self, iseq):
        return self.overlap

    def on_data(self, ispan, ospan):
        if ispan.nframe <= self.overlap:
            return 0
        x = np.asarray(ispan.data)
        y = np.asarray(ospan.data)
        n = ispan.nframe - self.overlap
        for k in range(n):
            y[k] = x[k:k + self.overlap + 1].sum(axis=0)


def test_input_overlap_is_re_presented():
    data = np.arange(40, dtype=np.float32).reshape(40, 1)
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, header([-1, 1]), gulp_nframe=8)
        b = MovingSum(src, gulp_nframe=8)
        callback_sink(b, out.seq, out.data)
        p.run()
    got = np.concatenate(out.chunks, axis=0)[:, 0]
    want = np.array([data[k:k + 4, 0].sum() for k in range(40 - 3)])
    np.testing.assert_array_equal(got[:len(want)], want)


class EveryN(TransformBlock):
    """Commits one frame every n inputs, like AccumulateBlock."""

    def __init__(self, iring, n):
        super(EveryN, self).__init__(iring, gulp_nframe=1)
        self.n = n

    def on_sequence(self, iseq):
        self.count = 0
        return deepcopy(iseq.header)

    def on_data(self, ispan, ospan):
        x, y = np.asarray(ispan.data), np.asarray(ospan.data)
        if self.count == 0:
            y[...] = x
        else:
            y[...] = y + x
        self.count += 1
        if self.count == self.n:
            self.count = 0
            return 1
        return 0


def test_rare_commits_accumulate_into_the_same_span():
    data = np.ones((24, 2), np.float32) * np.arange(24)[:, None]
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, header([-1, 2]), gulp_nframe=5)
        b = EveryN(src, 8)
        callback_sink(b, out.seq, out.data)
        p.run()
    got = np.concatenate(out.chunks, axis=0)
    want = data.reshape(3, 8, 2).sum(axis=1)
    np.testing.assert_array_equal(got, want)


def test_views_rewrite_headers_without_moving_data():
    data = np.arange(6 * 4 * 8, dtype=np.float32).reshape(6, 4, 8)
    hdr = header([-1, 4, 8], labels=['time', 'freq', 'fine'])
    hdr['_tensor']['scales'] = [[0, 1.0], [100.0, 8.0], [0.0, 1.0]]
    hdr['_tensor']['units'] = ['s', 'MHz', 'MHz']
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, hdr, gulp_nframe=2)
        v = bf.views.merge_axes(src, 'freq', 'fine', label='freq')
        v = bf.views.rename_axis(v, 'freq', 'chan')
        callback_sink(v, out.seq, out.data)
        p.run()
    t = out.headers[0]['_tensor']
    assert t['shape'] == [-1, 32] and t['labels'] == ['time', 'chan'] and t['scales'][1] == [100.0, 1.0]
    np.testing.assert_array_equal(np.concatenate(out.chunks, 0), data.reshape(6, 32))
    out2 = Collect()
    with Pipeline() as p:
        src = array_source(data, hdr, gulp_nframe=3)
        v = bf.views.split_axis(src, 'fine', 2, label='half')
        callback_sink(v, out2.seq, out2.data)
        p.run()
    assert out2.headers[0]['_tensor']['shape'] == [-1, 4, 4, 2]
    np.testing.assert_array_equal(np.concatenate(out2.chunks, 0), data.reshape(6, 4, 4, 2))


def test_remaining_views_add_delete_reverse_astype_reinterpret():
    """views/basic_views.py:51-145 of the reference: header-only transforms."""
    data = np.arange(4 * 1 * 6, dtype=np.float32).reshape(4, 1, 6)
    hdr = header([-1, 1, 6], labels=['time', 'pol', 'freq'])
    hdr['_tensor']['scales'] = [[0, 1.0], [0, 1], [1500.0, -0.5]]
    hdr['_tensor']['units'] = ['s', None, 'MHz']
    out = Collect()
    with Pipeline() as p:
        src = array_source(data, hdr, gulp_nframe=2)
        v = bf.views.delete_axis(src, 'pol')
        v = bf.views.reverse_scale(v, 'freq')
        v = bf.views.add_axis(v, 1, label='beam', scale=[0, 1], units=None)
        v = bf.views.reinterpret_axis(v, 'freq', label='chan', scale=[7.0, 2.0], units='kHz')
        callback_sink(v, out.seq, out.data)
        p.run()
    t = out.headers[0]['_tensor']
    assert t['shape'] == [-1, 1, 6] and t['labels'] == ['time', 'beam', 'chan']
    assert t['scales'][2] == [7.0, 2.0] and t['units'][2] == 'kHz'
    np.testing.assert_array_equal(np.concatenate(out.chunks, 0), data)
    out2 = Collect()
    with Pipeline() as p:
        src = array_source(data, hdr, gulp_nframe=4)
        v = bf.views.reverse_scale(src, 2)
        callback_sink(v, out2.seq, out2.data)
        p.run()
    assert out2.headers[0]['_tensor']['scales'][2] == [1500.0, 0.5]


def test_block_scope_and_space_validation():
    data = np.zeros((4, 2), np.float32)
    with Pipeline() as p:
        with bf.block_scope(gulp_nframe=2, fuse=True):
            src = array_source(data, header([-1, 2]), gulp_nframe=2)
            b = bf.blocks.reduce(src, 'd1', 2)            # needs cuda space
            assert b.gulp_nframe == 2 and b.fuse
        try:
            p.run()
        except ValueError as e:
            assert 'space' in str(e)
        else:
            raise AssertionError("system-space input to a cuda-only block must be rejected")


def test_correlate_block_header_and_gulp_contract():
    """blocks/correlate.py:51-74 of the reference: six-axis output tensor, time
    step scaled by the integration length, gulp capped at one integration and
    required to divide it."""
    import pytest
    from bifrost_b200 import blocks

    class Seq(object):
        def __init__(self, gulp):
            self.header = {'_tensor': {'dtype': 'ci8', 'shape': [-1, 4, 16, 2],
                                       'labels': ['time', 'freq', 'station', 'pol'],
                                       'scales': [[0, 1e-3], [100.0, 0.1], None, None],
                                       'units': ['s', 'MHz', None, None]},
                           'name': 'corr', 'gulp_nframe': gulp}

    with Pipeline():
        src = array_source(np.zeros((4, 4, 16, 2), dtype=bf.DataType('ci8').as_numpy_dtype()), Seq(32).header, gulp_nframe=32)
        blk = blocks.correlate(src, nframe_per_integration=96)
    ohdr = blk.on_sequence(Seq(32))
    t = ohdr['_tensor']
    assert t['dtype'] == 'cf32' and t['shape'] == [-1, 4, 16, 2, 16, 2]
    assert t['labels'] == ['time', 'freq', 'station_i', 'pol_i', 'station_j', 'pol_j']
    assert t['scales'][0] == [0, 1e-3 * 96] and t['units'] == ['s', 'MHz', None, None, None, None]
    assert ohdr['matrix_fill_mode'] == 'lower' and ohdr['gulp_nframe'] == 32
    assert blk.on_sequence(Seq(200))['gulp_nframe'] == 96        # capped at one integration
    with pytest.raises(ValueError):
        blk.on_sequence(Seq(40))                                  # 40 does not divide 96
    blk.gulp_nframe = 48                                          # the user's own gulp wins
    assert blk.on_sequence(Seq(40))['gulp_nframe'] == 40
    blk.gulp_nframe = 36
    with pytest.raises(ValueError):
        blk.on_sequence(Seq(32))


def test_fuse_scope_collapses_the_guppi_chain():
    """block_scope(fuse=True): transpose -> fft -> detect -> merge_axes -> reduce ->
    accumulate becomes one SpectrometerBlock between the same rings; a chain that
    is not all inside the scope is left alone (structure only: nothing runs)."""
    from bifrost_b200 import blocks, views
    from bifrost_b200.blocks.spectrometer import SpectrometerBlock
    hdr = {'_tensor': {'dtype': 'ci8', 'shape': [-1, 4, 4096, 2], 'labels': ['time', 'freq', 'fine_time', 'pol'],
                       'scales': [[0, 1.], [1200.0, 100.0], [0, 1e-8], None], 'units': ['s', 'MHz', 's', None]},
           'name': 'guppi', 'gulp_nframe': 1}
    x = np.zeros((2, 4, 4096, 2), dtype=bf.DataType('ci8').as_numpy_dtype())

    def build(fuse_all):
        seen = []
        p = Pipeline()
        with p:
            src = array_source(x, hdr, gulp_nframe=1)
            with bf.block_scope(fuse=True):
                c = blocks.transpose(src, ['time', 'pol', 'freq', 'fine_time'])
                c = blocks.fft(c, axes='fine_time', axis_labels='fine_freq', apply_fftshift=True)
                c = blocks.detect(c, mode='stokes')
                c = views.merge_axes(c, 'freq', 'fine_freq')
                c = blocks.reduce(c, 'freq', 4)
                if fuse_all:
                    c = blocks.accumulate(c, 8)
            if not fuse_all:
                c = blocks.accumulate(c, 8)
            sink = callback_sink(c, None, lambda s: seen.append(s))
        return p, src, sink

    p, src, sink = build(True)
    p._fuse_chains()
    kinds = [type(b).__name__ for b in p.blocks]
    assert kinds.count('SpectrometerBlock') == 1
    assert not {'TransposeBlock', 'FftBlock', 'DetectBlock', 'ReduceBlock', 'AccumulateBlock'} & set(kinds)
    spec = [b for b in p.blocks if isinstance(b, SpectrometerBlock)][0]
    assert spec.f_avg == 4 and spec.n_int == 8
    assert src.orings[0].consumers == [spec] and spec.orings[0].consumers == [sink]
    p, src, sink = build(False)
    p._fuse_chains()
    kinds = [type(b).__name__ for b in p.blocks]
    assert 'SpectrometerBlock' not in kinds and 'FftBlock' in kinds


def test_blocks_keep_the_reference_status_logs():
    """Every block writes <name>/bind, /in, /out, /sequence0 and /perf through
    bfProcLog* (what the reference's pipeline.py:346-364,445-451,486,649 writes and
    its like_top / like_bmon tools read); core= binds the block's thread."""
    import os
    from bifrost_b200 import proclog
    data = np.arange(40 * 2, dtype=np.float32).reshape(40, 2)
    out = Collect()
    cores = sorted(os.sched_getaffinity(0))
    seen = {}

    def data_cb(ispan):
        seen['core'] = bf.affinity.get_core()
        seen['logs'] = proclog.load_by_pid(os.getpid(), include_rings=True)
        out.data(ispan)

    with Pipeline() as p:
        src = array_source(data, header([-1, 2]), gulp_nframe=8, name='logsrc')
        b = copy(src, gulp_nframe=8, name='logcopy')
        callback_sink(b, out.seq, data_cb, gulp_nframe=8, name='logsink', core=cores[-1])
        p.run()
    np.testing.assert_array_equal(np.concatenate(out.chunks, axis=0), data)
    assert seen['core'] == cores[-1]
    logs = seen['logs']
    assert logs['logsink']['bind'] == {'ncore': 1, 'core0': cores[-1]}
    assert logs['logcopy']['in']['nring'] == 1 and logs['logcopy']['out']['nring'] == 1
    assert logs['logcopy']['in']['ring0'] == logs['logsrc']['out']['ring0']
    assert set(logs['logcopy']['perf']) == {'acquire_time', 'reserve_time', 'process_time'}
    assert logs['logsrc']['perf']['acquire_time'] == -1
    assert logs['logcopy']['sequence0']['name'] == 'test'


def test_views_that_split_or_merge_the_frame_axis():
    """views.split_axis / merge_axes on the frame axis (basic_views.py:154-158,
    191-193 of the reference): no data moves, the reader of the view counts in
    the new frames and gulp_nframe is rescaled."""
    data = np.arange(40 * 6, dtype=np.float32).reshape(40, 6)
    out = Collect()
    hdr = header([-1, 6])
    hdr['gulp_nframe'] = 8
    with Pipeline() as p:
        src = array_source(data, hdr, gulp_nframe=8)
        v = bf.views.split_axis(src, 'time', 4, label='fine')
        callback_sink(v, out.seq, out.data)                    # gulp from the header: 8 -> 2 new frames
        p.run()
    t = out.headers[0]['_tensor']
    assert t['shape'] == [-1, 4, 6] and t['labels'][:2] == ['time', 'fine']
    assert t['scales'][0][1] == 4.0 and t['scales'][1] == [0, 1.0]
    assert out.headers[0]['gulp_nframe'] == 2
    assert all(c.shape == (2, 4, 6) for c in out.chunks)
    np.testing.assert_array_equal(np.concatenate(out.chunks, axis=0), data.reshape(10, 4, 6))

    # and back: [-1, 4, 6] frames of 4 fine samples -> 40 frames of one
    hdr = header([-1, 4, 6], labels=['time', 'fine', 'chan'])
    hdr['_tensor']['units'] = ['s', 's', None]
    hdr['_tensor']['scales'] = [[0, 4.0], [0, 1.0], [0, 1.0]]
    hdr['gulp_nframe'] = 3
    out = Collect()
    with Pipeline() as p:
        src = array_source(data.reshape(10, 4, 6), hdr, gulp_nframe=3)
        v = bf.views.merge_axes(src, 'time', 'fine')
        callback_sink(v, out.seq, out.data)                    # gulp 3 -> 12 new frames
        p.run()
    t = out.headers[0]['_tensor']
    assert t['shape'] == [-1, 6] and t['labels'] == ['time', 'chan'] and t['scales'][0] == [0, 1.0]
    assert out.headers[0]['gulp_nframe'] == 12
    assert [c.shape[0] for c in out.chunks] == [12, 12, 12, 4]
    np.testing.assert_array_equal(np.concatenate(out.chunks, axis=0), data)

    # a gulp of the view that is not a whole number of ring frames is refused
    with Pipeline() as p:
        src = array_source(data.reshape(10, 4, 6), hdr, gulp_nframe=3)
        v = bf.views.merge_axes(src, 'time', 'fine')
        callback_sink(v, out.seq, out.data, gulp_nframe=6)
        try:
            p.run()
            raise AssertionError('expected a ValueError')
        except ValueError as e:
            assert 'whole number of ring frames' in str(e)


def test_a_signal_shuts_a_running_pipeline_down():
    """Pipeline.shutdown_on_signals (pipeline.py:271-286 of the reference): an
    endless source and its consumers stop at their next ring call and run()
    returns."""
    import os
    import signal
    import threading
    import time
    import warnings
    from bifrost_b200.blocks.testing import NumpySourceBlock

    class Endless(NumpySourceBlock):
        def on_data(self, reader, ospans):
            reader.pos = 0                                   # never runs out
            return super(Endless, self).on_data(reader, ospans)

    seen = [0]

    def count(ispan):
        seen[0] += ispan.nframe

    data = np.arange(8 * 2, dtype=np.float32).reshape(8, 2)
    old = signal.getsignal(signal.SIGUSR1)
    try:
        with Pipeline() as p:
            src = Endless(data, header([-1, 2]), 4)
            b = copy(src, gulp_nframe=4)
            callback_sink(b, None, count, gulp_nframe=4)
            p.shutdown_on_signals([signal.SIGUSR1])
            threading.Timer(0.3, os.kill, (os.getpid(), signal.SIGUSR1)).start()
            t0 = time.time()
            with warnings.catch_warnings(record=True) as caught:
                warnings.simplefilter('always')
                p.run()
        assert time.time() - t0 < 20
        assert seen[0] > 0
        assert any('shutting down pipeline' in str(w.message) for w in caught)
    finally:
        signal.signal(signal.SIGUSR1, old)


def test_unguaranteed_readers_are_lapped_instead_of_holding_the_writer_back():
    """guarantee=False (pipeline.py:518-537,590-643 of the reference): a slow
    sink loses whole gulps, never sees a torn one, and does not slow the source
    or the guaranteed sink next to it down; a transform behind an unguaranteed
    reader keeps its output cadence with on_skip's zeros for what was lost."""
    import time
    nframe, gulp = 4000, 8
    data = (np.arange(nframe * 3, dtype=np.float32).reshape(nframe, 3) + 1)       # no zeros in the data
    fast, slow, behind = Collect(), Collect(), Collect()
    slow.offsets, slow.skipped = [], []
    t_fast = [0.]

    def fast_data(ispan):
        fast.data(ispan)
        t_fast[0] = time.time()

    def slow_data(ispan):
        chunk = np.array(ispan.data.copy('system'))
        time.sleep(0.002)
        if ispan.nframe_overwritten:           # the writer came by meanwhile: the copy may be torn
            slow.skipped.append((ispan.frame_offset, ispan.frame_offset + ispan.nframe))
        else:
            slow.offsets.append(ispan.frame_offset)
            slow.chunks.append(chunk)

    class SlowSink(bf.blocks.testing.CallbackSinkBlock):
        def on_skip(self, islice):
            slow.skipped.append((islice.start, islice.stop))

    class SlowCopy(TransformBlock):
        def on_sequence(self, iseq):
            return deepcopy(iseq.header)

        def on_data(self, ispan, ospan):
            time.sleep(0.002)
            bf.copy_array(ospan.data, ispan.data)

    with Pipeline() as p:
        with bf.block_scope(buffer_nframe=4 * gulp):
            src = array_source(data, header([-1, 3]), gulp_nframe=gulp)
        callback_sink(src, None, fast_data, gulp_nframe=gulp)
        SlowSink(src, None, slow_data, gulp_nframe=gulp, guarantee=False)
        callback_sink(SlowCopy(src, gulp_nframe=gulp, guarantee=False), None, behind.data, gulp_nframe=gulp)
        t0 = time.time()
        p.run()
        t_all = time.time() - t0
    # the guaranteed sink got everything, and nobody waited for the slow ones: had they
    # held the source back the run would have taken 500 gulps x 2 ms
    np.testing.assert_array_equal(np.concatenate(fast.chunks, axis=0), data)
    assert t_all < 0.8 * (nframe // gulp) * 0.002
    # the slow sink: whole gulps, each intact, the rest reported lost, nothing twice
    assert slow.skipped and len(slow.chunks) < nframe // gulp
    for off, chunk in zip(slow.offsets, slow.chunks):
        np.testing.assert_array_equal(chunk, data[off:off + gulp])
    seen = sorted([(o, o + gulp) for o in slow.offsets] + slow.skipped)
    assert seen[0][0] == 0 and all(a[1] == b[0] for a, b in zip(seen[:-1], seen[1:]))
    # the transform: every frame accounted for, lost gulps are zeros, the others the data
    out = np.concatenate(behind.chunks, axis=0)
    assert out.shape == data.shape
    blocks = out.reshape(-1, gulp, 3)
    zero = ~blocks.reshape(len(blocks), -1).any(axis=1)
    assert zero.any() and not zero.all()
    np.testing.assert_array_equal(blocks[~zero], data.reshape(-1, gulp, 3)[~zero])
