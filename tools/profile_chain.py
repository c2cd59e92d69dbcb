"""One launch each of the HBM-bound chain ops (transpose, fft, detect, reduce,
accumulate, unpack) on a 4-frame GUPPI gulp, for ncu captures (not a benchmark)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bifrost_b200 as bf

rng = np.random.default_rng(3)
nframe, nchan, nfft, npol = 4, 4096, 4096, 2
raw = rng.integers(-127, 128, size=(nframe, nchan, nfft, npol, 2), dtype=np.int8)
x = raw.view(bf.DataType('ci8').as_numpy_dtype()).reshape(nframe, nchan, nfft, npol)
d_x = bf.asarray(x, space='cuda')
d_t = bf.empty((nframe, npol, nchan, nfft), 'ci8', 'cuda')
d_f = bf.empty((nframe, npol, nchan, nfft), 'cf32', 'cuda')
d_d = bf.empty((nframe, 4, nchan, nfft), 'f32', 'cuda')
d_r = bf.empty((nframe, 4, nchan * nfft // 4), 'f32', 'cuda')
d_acc = bf.zeros((1, 4, nchan * nfft // 4), 'f32', 'cuda')
plan = bf.fft.Fft()
plan.init(d_t, d_f, axes=[3], apply_fftshift=True)
for _ in range(2):
    bf.transpose(d_t, d_x, (0, 3, 1, 2))
    plan.execute(d_t, d_f)
    bf.detect(d_f, d_d, 'stokes', 1)
    bf.reduce(d_d.reshape(nframe, 4, nchan * nfft), d_r, 'sum')
    bf.accumulate(d_r[0:1], d_acc, 1.0)
# other FFT lengths: 1024 (fast kernel) and 131072 (register pass A + fast pass B)
for nfft_ in (1024, 131072):
    nb = (1 << 25) // nfft_
    xi = raw.reshape(-1)[:nb * nfft_ * 2].reshape(nb, nfft_, 2).view(bf.DataType('ci8').as_numpy_dtype()).reshape(nb, nfft_)
    d_i = bf.asarray(xi, space='cuda')
    d_o = bf.empty((nb, nfft_), 'cf32', 'cuda')
    pl = bf.fft.Fft()
    pl.init(d_i, d_o, axes=[1])
    for _ in range(2):
        pl.execute(d_i, d_o)
    del d_i, d_o, pl
# ci4 -> ci8 unpack of a 256 MB packed buffer
nbyte = 1 << 28
p8 = bf.asarray(rng.integers(0, 256, size=(nbyte,), dtype=np.uint8), space='cuda')
try:
    src = p8.view('ci4')
    dst = bf.empty((nbyte,), 'ci8', 'cuda')
    for _ in range(2):
        bf.unpack(src, dst)
except Exception as e:          # view API differences must not lose the other captures
    print('unpack skipped:', e)
bf.device.stream_synchronize()
print('done')
