"""Tolerance (north_star: 1e-5 relative): every output within 1e-5 of its own
magnitude, with the rms of Stokes I as the floor for the near-empty bins --
``close()`` below.  (The inputs carry a strong tone per channel, so the largest
absolute errors sit on bins that are ~20x the rms.)

The fused spectrometer kernel against (1) the oracle chain
fft(x/128, fftshift) -> stokes -> sum of f_avg bins -> sum over frames in fp64
and (2) the unfused C-ABI ops run one after the other on the GPU."""
import numpy as np
import pytest

import bifrost_b200 as bf
from oracle import fft as offt

pytestmark = pytest.mark.gpu


def make(nframe, nchan, nfft, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((nframe, nchan, nfft, 2), dtype=bf.DataType('ci8').as_numpy_dtype())
    x['re'] = rng.integers(-127, 128, size=x.shape)
    x['im'] = rng.integers(-127, 128, size=x.shape)
    # one tone per coarse channel (bin-position check)
    t = np.arange(nfft)
    for c in range(nchan):
        k = (37 * c + 11) % nfft
        tone = 40 * np.exp(2j * np.pi * k * t / nfft)
        x['re'][:, c, :, 0] = np.clip(x['re'][:, c, :, 0] // 4 + np.rint(tone.real), -127, 127)
        x['im'][:, c, :, 0] = np.clip(x['im'][:, c, :, 0] // 4 + np.rint(tone.imag), -127, 127)
    return x


def close(got, want, tol=1e-5):
    scale = np.sqrt(np.mean(np.asarray(want)[0] ** 2))
    err = np.abs(got - want)
    lim = tol * np.maximum(np.abs(want), scale)
    assert (err <= lim).all(), (float((err / lim).max()), float(err.max()), float(scale))


def oracle_chain(x, f_avg):
    v = (x['re'].astype(np.float64) + 1j * x['im'].astype(np.float64)) / 128.0
    spec = np.fft.fftshift(np.fft.fft(v, axis=2), axes=2)        # [frame, chan, fine, pol]
    X, Y = spec[..., 0], spec[..., 1]
    xy = X * np.conj(Y)
    st = np.stack([abs(X)**2 + abs(Y)**2, abs(X)**2 - abs(Y)**2, 2 * xy.real, -2 * xy.imag])
    nframe, nchan, nfft = X.shape
    st = st.reshape(4, nframe, nchan * nfft // f_avg, f_avg).sum(-1)
    return st.sum(1)


@pytest.mark.parametrize("f_avg", [1, 2, 4, 8, 16, 32])
def test_fused_matches_oracle(f_avg):
    nframe, nchan, nfft = 3, 6, 4096
    x = make(nframe, nchan, nfft, 100 + f_avg)
    d_x = bf.asarray(x, space='cuda')
    d_o = bf.asarray(np.full((4, nchan * nfft // f_avg), np.nan, np.float32), space='cuda')
    bf.spectrometer(d_x, d_o, nfft, f_avg, beta=0.0)
    got = np.asarray(d_o.copy('system'))
    want = oracle_chain(x, f_avg)
    close(got, want)
    # the tone lands in the right (shifted) bin of channel 0
    k = (11 + nfft // 2) % nfft
    assert np.argmax(got[0, :nfft // f_avg]) == k // f_avg


def test_beta_accumulates_across_calls():
    nframe, nchan, nfft, f_avg = 2, 4, 4096, 4
    x1, x2 = make(nframe, nchan, nfft, 1), make(nframe, nchan, nfft, 2)
    d_o = bf.zeros((4, nchan * nfft // f_avg), 'f32', 'cuda')
    bf.spectrometer(bf.asarray(x1, space='cuda'), d_o, nfft, f_avg, beta=0.0)
    bf.spectrometer(bf.asarray(x2, space='cuda'), d_o, nfft, f_avg, beta=1.0)
    got = np.asarray(d_o.copy('system'))
    want = oracle_chain(x1, f_avg) + oracle_chain(x2, f_avg)
    close(got, want)


def test_fused_matches_unfused_ops():
    """Same gulp through bfTranspose, bfFft, bfDetect, bfReduce, bfAccumulate."""
    nframe, nchan, nfft, f_avg = 4, 8, 4096, 4
    x = make(nframe, nchan, nfft, 7)
    d_x = bf.asarray(x, space='cuda')
    d_t = bf.empty((nframe, 2, nchan, nfft), 'ci8', 'cuda')
    bf.transpose(d_t, d_x, (0, 3, 1, 2))
    d_f = bf.empty((nframe, 2, nchan, nfft), 'cf32', 'cuda')
    plan = bf.fft.Fft()
    plan.init(d_t, d_f, axes=[3], apply_fftshift=True)
    plan.execute(d_t, d_f)
    d_d = bf.empty((nframe, 4, nchan, nfft), 'f32', 'cuda')
    bf.detect(d_f, d_d, 'stokes', 1)
    d_r = bf.empty((nframe, 4, nchan * nfft // f_avg), 'f32', 'cuda')
    bf.reduce(d_d.reshape(nframe, 4, nchan * nfft), d_r, 'sum')
    d_a = bf.empty((1, 4, nchan * nfft // f_avg), 'f32', 'cuda')
    for k in range(nframe):
        bf.accumulate(d_r[k:k + 1], d_a, 0.0 if k == 0 else 1.0)
    unfused = np.asarray(d_a.copy('system'))[0]
    d_o = bf.empty((4, nchan * nfft // f_avg), 'f32', 'cuda')
    bf.spectrometer(d_x, d_o, nfft, f_avg)
    fused = np.asarray(d_o.copy('system'))
    close(fused, unfused)


def test_unsupported_shapes_are_reported():
    from bifrost_b200.libbifrost import _bf
    x = bf.empty((2, 4, 1024, 2), 'ci8', 'cuda')
    o = bf.empty((4, 4 * 1024 // 4), 'f32', 'cuda')
    assert _bf.bfSpectrometerFused(x.as_BFarray(), o.as_BFarray(), 1024, 4, 0.0) == \
        _bf.BF_STATUS_UNSUPPORTED_SHAPE


def test_baseline_config3_full_size_against_fp64():
    """BASELINE config 3 at the size bench.py times -- 32 frames x 4096 coarse
    channels x 4096 fine samples x 2 pol ci8 (2 GiB), f_avg = 4 -- value-checked on
    64 coarse channels (incl. both ends) against the fp64 chain."""
    import torch
    nframe, nchan, nfft, f_avg = 32, 4096, 4096, 4
    g = torch.Generator(device='cuda').manual_seed(3)
    raw = torch.randint(-127, 128, (nframe, nchan, nfft, 2, 2), dtype=torch.int8, device='cuda', generator=g)
    d_x = bf.ndarray(space='cuda', shape=(nframe, nchan, nfft, 2), dtype='ci8', buffer=raw.data_ptr())   # zero-copy view
    d_o = bf.zeros((4, nchan * nfft // f_avg), 'f32', 'cuda')
    bf.spectrometer(d_x, d_o, nfft, f_avg, beta=0.0)
    got = np.asarray(d_o.copy('system')).reshape(4, nchan, nfft // f_avg)
    rng = np.random.default_rng(0)
    chans = np.unique(np.concatenate([[0, 1, nchan - 2, nchan - 1], rng.integers(0, nchan, 60)]))
    sub = raw[:, torch.as_tensor(chans, device='cuda')].cpu().numpy()        # [frame, 64, fine, pol, re/im]
    x = np.zeros(sub.shape[:4], dtype=bf.DataType('ci8').as_numpy_dtype())
    x['re'], x['im'] = sub[..., 0], sub[..., 1]
    want = oracle_chain(x, f_avg).reshape(4, len(chans), nfft // f_avg)
    close(got[:, chans], want)
