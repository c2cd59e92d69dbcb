"""The hot-path blocks driven through the executor on the GPU: the GUPPI
spectrometer chain of the reference README / testbench/gpuspec_simple.py
(unfused blocks vs the fused spectrometer block vs the fp64 oracle), the FDMT
block with its max_delay input overlap, and transpose + reduce (BASELINE
config 1 run in CUDA space)."""
from copy import deepcopy

import numpy as np
import pytest

import bifrost_b200 as bf
from bifrost_b200 import blocks, views
from bifrost_b200.pipeline import Pipeline
from oracle import fdmt as ofdmt
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_spectrometer import make as make_voltages, oracle_chain, close  # noqa: E402

pytestmark = pytest.mark.gpu


class Collect(object):
    def __init__(self):
        self.chunks, self.headers = [], []

    def seq(self, iseq):
        self.headers.append(deepcopy(iseq.header))

    def data(self, ispan):
        self.chunks.append(np.array(ispan.data.copy('system')))


def guppi_header(nchan, nfft):
    return {'_tensor': {'dtype': 'ci8', 'shape': [-1, nchan, nfft, 2],
                        'labels': ['time', 'freq', 'fine_time', 'pol'],
                        'scales': [[0, nfft * nchan / 400e6], [1200.0, 400.0 / nchan],
                                   [0, nchan / 400e6], None],
                        'units': ['s', 'MHz', 's', None]},
            'name': 'guppi', 'gulp_nframe': 1}


def test_guppi_chain_unfused_vs_fused_vs_oracle():
    nframe, nchan, nfft, f_avg, n_int = 8, 4, 4096, 4, 4
    x = make_voltages(nframe, nchan, nfft, 42)
    unfused, fused = Collect(), Collect()
    with Pipeline() as p:
        src = blocks.array_source(x, guppi_header(nchan, nfft), gulp_nframe=2)
        b = blocks.copy(src, space='cuda')
        c = blocks.transpose(b, ['time', 'pol', 'freq', 'fine_time'])
        c = blocks.fft(c, axes='fine_time', axis_labels='fine_freq', apply_fftshift=True)
        c = blocks.detect(c, mode='stokes')
        c = views.merge_axes(c, 'freq', 'fine_freq')
        c = blocks.reduce(c, 'freq', f_avg)
        c = blocks.accumulate(c, n_int)
        c = blocks.copy(c, space='cuda_host')
        blocks.callback_sink(c, unfused.seq, unfused.data)
        with bf.block_scope(fuse=True):
            d = blocks.spectrometer(b, f_avg=f_avg, n_int=n_int, gulp_nframe=2)
        d = blocks.copy(d, space='cuda_host')
        blocks.callback_sink(d, fused.seq, fused.data)
        p.run()
    a = np.concatenate(unfused.chunks, 0)
    b_ = np.concatenate(fused.chunks, 0)
    assert a.shape == (nframe // n_int, 4, nchan * nfft // f_avg) == b_.shape
    want = np.stack([oracle_chain(x[i * n_int:(i + 1) * n_int], f_avg) for i in range(nframe // n_int)])
    for i in range(want.shape[0]):
        close(a[i], want[i])
        close(b_[i], want[i])
    t = unfused.headers[0]['_tensor']
    assert t['labels'] == ['time', 'pol', 'freq'] and t['shape'] == [-1, 4, nchan * nfft // f_avg]
    assert fused.headers[0]['_tensor']['shape'] == t['shape']


def test_fuse_scope_runs_the_chain_as_one_kernel():
    """The same chain written block by block under block_scope(fuse=True): the
    executor swaps in the fused kernel (one launch per gulp) and the result is
    the oracle's."""
    nframe, nchan, nfft, f_avg, n_int = 8, 4, 4096, 4, 4
    x = make_voltages(nframe, nchan, nfft, 43)
    out = Collect()
    with Pipeline() as p:
        src = blocks.array_source(x, guppi_header(nchan, nfft), gulp_nframe=2)
        b = blocks.copy(src, space='cuda')
        with bf.block_scope(fuse=True):
            c = blocks.transpose(b, ['time', 'pol', 'freq', 'fine_time'])
            c = blocks.fft(c, axes='fine_time', axis_labels='fine_freq', apply_fftshift=True)
            c = blocks.detect(c, mode='stokes')
            c = views.merge_axes(c, 'freq', 'fine_freq')
            c = blocks.reduce(c, 'freq', f_avg)
            c = blocks.accumulate(c, n_int)
        c = blocks.copy(c, space='cuda_host')
        blocks.callback_sink(c, out.seq, out.data)
        before = bf.launch_count()
        p.run()
        launches = bf.launch_count() - before
    assert 'SpectrometerBlock' in [type(b).__name__ for b in p.blocks]
    # one fused launch per 2-frame gulp (+ one strided device copy per committed frame); the
    # unfused chain needs five launches per gulp and one accumulate per frame
    assert launches <= 2 * (nframe // 2)
    got = np.concatenate(out.chunks, 0)
    want = np.stack([oracle_chain(x[i * n_int:(i + 1) * n_int], f_avg) for i in range(nframe // n_int)])
    for i in range(want.shape[0]):
        close(got[i], want[i])
    assert out.headers[0]['_tensor']['labels'] == ['time', 'pol', 'freq']


def test_fdmt_block_with_overlap_equals_one_big_transform():
    nchan, ntime, gulp = 64, 4000, 1000
    f0, df, dt = 1000.0, 400.0 / 64, 1e-3
    rng = np.random.default_rng(9)
    x = rng.integers(-100, 100, size=(nchan, ntime)).astype(np.int8)
    hdr = {'_tensor': {'dtype': 'i8', 'shape': [nchan, -1], 'labels': ['freq', 'time'],
                       'scales': [[f0, df], [0, dt]], 'units': ['MHz', 's']},
           'name': 'fil', 'gulp_nframe': gulp}
    out = Collect()
    with Pipeline() as p:
        src = blocks.array_source(x, hdr, gulp_nframe=gulp, frame_axis=1)
        b = blocks.copy(src, space='cuda')
        b = blocks.fdmt(b, max_delay=50)
        b = blocks.copy(b, space='system')
        blocks.callback_sink(b, out.seq, out.data)
        p.run()
    got = np.concatenate(out.chunks, axis=1)
    want = ofdmt.fdmt(x, 50, f0, df)
    n = got.shape[1]
    assert n >= ntime - 50 - gulp
    np.testing.assert_array_equal(got, want[:, :n])
    assert out.headers[0]['_tensor']['labels'] == ['dispersion', 'time']


def test_config1_transpose_reduce_blocks():
    rng = np.random.default_rng(10)
    x = rng.normal(size=(4096, 256)).astype(np.float32)
    hdr = {'_tensor': {'dtype': 'f32', 'shape': [-1, 256, 1], 'labels': ['time', 'freq', 'pol'],
                       'scales': [[0, 1e-3], [100.0, 0.1], None], 'units': ['s', 'MHz', None]},
           'name': 'c1', 'gulp_nframe': 512}
    out = Collect()
    with Pipeline() as p:
        src = blocks.array_source(x.reshape(4096, 256, 1), hdr, gulp_nframe=512)
        b = blocks.copy(src, space='cuda')
        b = blocks.transpose(b, ['time', 'pol', 'freq'])
        b = blocks.reduce(b, 'freq', 4)
        b = blocks.reduce(b, 'time', 8)
        b = blocks.copy(b, space='system')
        blocks.callback_sink(b, out.seq, out.data)
        p.run()
    got = np.concatenate(out.chunks, 0)
    want = x.reshape(512, 8, 64, 4).sum(3).sum(1).reshape(512, 1, 64)
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_unpack_block_ci4_to_ci8():
    """blocks/unpack.py of the reference: ci4 voltages -> ci8 (gunpack path, bit-exact)."""
    from oracle import unpack as ounpack
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, size=(64, 8, 32), dtype=np.uint8)
    x = bf.ndarray(raw.view(bf.DataType('ci4').as_numpy_dtype()), dtype='ci4')
    hdr = {'_tensor': {'dtype': 'ci4', 'shape': [-1, 8, 32], 'labels': ['time', 'freq', 'station'],
                       'scales': [[0, 1e-3], [100.0, 0.1], None], 'units': ['s', 'MHz', None]},
           'name': 'u', 'gulp_nframe': 16}
    out = Collect()
    with Pipeline() as p:
        src = blocks.array_source(x, hdr, gulp_nframe=16)
        b = blocks.copy(src, space='cuda')
        b = blocks.unpack(b, 'ci8')
        b = blocks.copy(b, space='system')
        blocks.callback_sink(b, out.seq, out.data)
        p.run()
    assert out.headers[0]['_tensor']['dtype'] == 'ci8'
    got = np.concatenate(out.chunks, 0)
    got = np.stack([got['re'], got['im']], -1)
    want = ounpack.unpack(raw, 4, True, gpu=True).reshape(64, 8, 32, 2)
    np.testing.assert_array_equal(got, want)


def test_file_to_file_guppi_spectrometer_sigproc(tmp_path):
    """GUPPI RAW file -> device -> fused spectrometer -> host -> sigproc filterbank
    (SURVEY 8f.4): the file body equals what a callback sink sees, the header
    carries the channelisation."""
    from bifrost_b200 import guppi_raw, sigproc
    nblock, nchan, nfft = 8, 4, 4096
    rng = np.random.default_rng(12)
    data = rng.integers(-100, 100, size=(nblock, nchan, nfft, 2, 2), dtype=np.int8)
    path = str(tmp_path / 'volt.raw')
    with open(path, 'wb') as f:
        for b in range(nblock):
            guppi_raw.write_header(dict(BACKEND='GUPPI', TELESCOP='GBT', SRC_NAME='FAKE', OBSFREQ=1500.0,
                                        OBSBW=100.0, OBSNCHAN=nchan, NPOL=4, NBITS=8, BLOCSIZE=nchan * nfft * 4,
                                        PKTIDX=b, PKTSIZE=nchan * nfft * 4, STT_IMJD=58849, STT_SMJD=0,
                                        DIRECTIO=1), f)
            data[b].tofile(f)
    out = Collect()
    with Pipeline() as p:
        src = blocks.read_guppi_raw([path], gulp_nframe=2)
        b = blocks.copy(src, space='cuda')
        b = blocks.spectrometer(b, f_avg=4, n_int=4, gulp_nframe=2)
        b = blocks.copy(b, space='system')
        blocks.callback_sink(b, out.seq, out.data)
        blocks.write_sigproc(b, path=str(tmp_path))
        p.run()
    want = np.concatenate(out.chunks, 0)
    assert want.shape == (nblock // 4, 4, nchan * nfft // 4)
    with open(str(tmp_path / 'volt.raw.fil'), 'rb') as f:
        h = sigproc.read_header(f)
        body = np.fromfile(f, dtype=np.float32)
    assert h['nifs'] == 4 and h['nchans'] == nchan * nfft // 4 and h['nbits'] == 32 and h['data_type'] == 1
    assert h['foff'] == pytest.approx(100.0 / nchan / nfft * 4) and h['telescope_id'] == 6
    np.testing.assert_array_equal(body.reshape(want.shape), want)
    # and the values are the spectrometer's (oracle chain), not just self-consistent
    x = data.view(bf.DataType('ci8').as_numpy_dtype()).reshape(nblock, nchan, nfft, 2)
    for i in range(nblock // 4):
        ref = oracle_chain(x[4 * i:4 * i + 4], 4)
        close(want[i], ref)
