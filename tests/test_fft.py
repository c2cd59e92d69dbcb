"""bfFft parity against fp64 numpy.fft (the reference's own gold,
test/test_fft.py:36-51).  Tolerance: max |err| <= 1e-5 * rms(gold), whatever the
length (north_star: 1e-5 relative; the reference asserts rtol 1e-1 /
atol 1e-6*mean)."""
import numpy as np
import pytest

import bifrost_b200 as bf
from bifrost_b200.fft import Fft
from oracle import fft as offt

pytestmark = pytest.mark.gpu


def compare(result, gold, tol=1e-5):
    gold = np.asarray(gold)
    rms = np.sqrt(np.mean(np.abs(gold) ** 2)) + 1e-30
    err = np.abs(np.asarray(result) - gold).max()
    assert err <= tol * rms, (err, rms)


def run(x, oshape, odtype, axes, inverse=False, fftshift=False):
    d_in = bf.asarray(x, space='cuda')
    d_out = bf.empty(oshape, dtype=odtype, space='cuda')
    plan = Fft()
    plan.init(d_in, d_out, axes=axes, apply_fftshift=fftshift)
    plan.execute(d_in, d_out, inverse)
    return np.asarray(d_out.copy('system'))


def c2c_case(shape, axes, rng):
    x = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    for inverse in (False, True):
        for shift in (False, True):
            got = run(x, shape, 'cf32', axes, inverse, shift)
            want = offt.fft(x.astype(np.complex128), axes, inverse=inverse, fftshift=shift)
            compare(got, want)


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_c2c_1d_pow2_lengths(n):
    c2c_case((max(2, 65536 // n), n), [1], np.random.default_rng(n))


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 12, 33, 65, 100, 255, 1000])
def test_c2c_1d_other_lengths(n):
    c2c_case((7, n), [1], np.random.default_rng(n))


def test_c2c_nd_reference_shapes():
    """test/test_fft.py:100-163 (reduced sizes where the originals are huge)."""
    rng = np.random.default_rng(1234)
    c2c_case((256, 256), [0], rng)
    c2c_case((256, 256), [1], rng)
    c2c_case((256, 256), [0, 1], rng)
    for axes in ([0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]):
        c2c_case((32, 32, 32), axes, rng)
    for axes in ([0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3], [0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]):
        c2c_case((16, 16, 16, 16), axes, rng)
    c2c_case((33, 31, 65, 16), [0, 2], rng)


@pytest.mark.parametrize("shape,axes_list", [
    ((16777216,), [[0]]),                                                     # shape1D
    ((2048, 2048), [[0], [1], [0, 1]]),                                       # shape2D
    ((128, 128, 128), [[0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]]),    # shape3D
    ((32, 32, 32, 32), [[0, 1], [0, 3], [1, 2], [2, 3], [0, 1, 2], [0, 2, 3], [1, 2, 3]]),   # shape4D
])
def test_c2c_at_the_reference_sizes(shape, axes_list):
    """The shapes test/test_fft.py:57-66 itself runs (its 1-D case is 2^24
    points), forward and inverse + fftshift."""
    rng = np.random.default_rng(len(shape))
    x = rng.normal(size=shape + (2,)).astype(np.float32).view(np.complex64)[..., 0]
    xd = x.astype(np.complex128)
    for axes in axes_list:
        for inverse, shift in ((False, False), (True, True)):
            got = run(x, shape, 'cf32', axes, inverse, shift)
            compare(got, offt.fft(xd, axes, inverse=inverse, fftshift=shift))


def test_r2c_and_c2r_at_the_reference_sizes():
    """test/test_fft.py:193-204: real transforms of shape2D and shape3D."""
    rng = np.random.default_rng(6)
    for shape, axes in (((2048, 2048), [0, 1]), ((128, 128, 128), [0, 1, 2])):
        x = rng.normal(size=shape).astype(np.float32)
        oshape = list(shape)
        oshape[-1] = shape[-1] // 2 + 1
        gold = np.fft.rfftn(x.astype(np.float64), axes=axes)
        compare(run(x, oshape, 'cf32', axes), gold)
        spec = gold.astype(np.complex64)
        back = run(spec, shape, 'f32', axes)
        compare(back, np.fft.irfftn(spec.astype(np.complex128), s=list(shape), axes=axes) * np.prod(shape))


@pytest.mark.parametrize("n", [1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 20])
def test_c2c_large_four_step(n):
    """Lengths above the single-pass limit (8192): n = n1 * n2 with n2 = min(4096, n/16);
    131072 is the README-exact GUPPI fine_time length (SURVEY 8d config 3 stretch)."""
    rng = np.random.default_rng(5)
    x = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(np.complex64)
    for inverse, shift in [(False, False), (True, False), (False, True), (True, True)]:
        got = run(x, (2, n), 'cf32', [1], inverse, shift)
        want = offt.fft(x.astype(np.complex128), [1], inverse=inverse, fftshift=shift)
        compare(got, want)


@pytest.mark.parametrize("shape,axes", [((2, 10000), [1]), ((12345,), [0]), ((3, 20000), [1]), ((100000,), [0]),
                                        ((9000, 6), [0]), ((4, 9000), [0, 1])])
def test_c2c_long_lengths_that_are_not_powers_of_two(shape, axes):
    """Bluestein on the power-of-two kernels (the reference takes any length
    through cuFFT): forward / inverse, with and without fftshift, odd lengths,
    a strided axis, and a 2-D transform with one such axis."""
    c2c_case(shape, axes, np.random.default_rng(shape[-1]))


@pytest.mark.parametrize("dtype,scale", [('ci8', 127), ('ci16', 32767), ('ci4', 7)])
def test_integer_complex_inputs(dtype, scale):
    """GUPPI-chain input: ci8 [time, pol, freq, fine_time], FFT over fine_time
    with fftshift (blocks/fft.py:118-137 -> callback_load_ci8)."""
    rng = np.random.default_rng(3)
    shape = (2, 2, 8, 1024)
    re = rng.integers(-scale, scale + 1, size=shape)
    im = rng.integers(-scale, scale + 1, size=shape)
    if dtype == 'ci4':
        x = np.zeros(shape, dtype=bf.DataType('ci4').as_numpy_dtype())
        x['re_im'] = ((re & 0xF) << 4 | (im & 0xF)).astype(np.uint8)
        want_in = ((re << 4) + 1j * (im << 4)) * offt.SCALE['ci4']
    else:
        x = np.zeros(shape, dtype=bf.DataType(dtype).as_numpy_dtype())
        x['re'], x['im'] = re, im
        want_in = (re + 1j * im) * offt.SCALE[dtype]
    got = run(x, shape, 'cf32', [3], False, True)
    compare(got, offt.fft(want_in.astype(np.complex128), [3], fftshift=True))


def test_r2c_and_c2r():
    rng = np.random.default_rng(4)
    for shape, axes in [((64, 256), [1]), ((64, 256), [0, 1]), ((16, 32, 64), [0, 1, 2]),
                        ((16, 32, 64), [0, 2]), ((8, 12, 30), [1, 2]), ((33, 31, 65, 16), [0, 2])]:
        x = rng.normal(size=shape).astype(np.float32)
        oshape = list(shape)
        oshape[axes[-1]] = shape[axes[-1]] // 2 + 1
        got = run(x, oshape, 'cf32', axes)
        compare(got, np.fft.rfftn(x.astype(np.float64), axes=axes))
        # c2r back (unnormalised): even lengths only, like the reference's tests
        if shape[axes[-1]] % 2 == 0:
            spec = np.fft.rfftn(x.astype(np.float64), axes=axes).astype(np.complex64)
            back = run(spec, shape, 'f32', axes)
            norm = np.prod([shape[a] for a in axes])
            compare(back, np.fft.irfftn(spec.astype(np.complex128), s=[shape[a] for a in axes], axes=axes) * norm)


@pytest.mark.parametrize("n,batch", [(1 << 14, 3), (1 << 15, 1), (1 << 20, 2), (1 << 24, 1)])
def test_r2c_and_c2r_longer_than_one_pass(n, batch):
    """test/test_fft.py:57,194-201 runs r2c / c2r at 2^24 points: the half-length
    complex transform + fix-up path (run_axis_real_long), f32 and integer inputs."""
    rng = np.random.default_rng(n)
    x = rng.normal(size=(batch, n)).astype(np.float32)
    spec = np.fft.rfft(x.astype(np.float64), axis=1)
    compare(run(x, (batch, n // 2 + 1), 'cf32', [1]), spec)
    back = run(spec.astype(np.complex64), (batch, n), 'f32', [1])
    compare(back, np.fft.irfft(spec.astype(np.complex64).astype(np.complex128), n=n, axis=1) * n)
    if n <= (1 << 20):
        for dtype, scale in [(np.int16, 32767), (np.int8, 127)]:
            xi = (rng.normal(size=(batch, n)) * scale / 4).astype(dtype)
            compare(run(xi, (batch, n // 2 + 1), 'cf32', [1]),
                    np.fft.rfft(xi.astype(np.float64) / (scale + 1), axis=1))


def test_r2c_integer_and_misaligned():
    """test/test_fft.py:83-99: i8/i16 inputs at odd byte offsets."""
    rng = np.random.default_rng(6)
    for dtype, scale in [(np.int16, 32767), (np.int8, 127)]:
        for mis in range(4):
            n = 512
            base = (rng.normal(size=(6, n + mis)) * scale / 4).astype(dtype)
            d_base = bf.asarray(base, space='cuda')
            view = d_base[:, mis:]
            d_out = bf.empty((6, n // 2 + 1), dtype='cf32', space='cuda')
            plan = Fft()
            plan.init(view, d_out, axes=[1])
            plan.execute(view, d_out)
            got = np.asarray(d_out.copy('system'))
            want = np.fft.rfft(base[:, mis:].astype(np.float64) / (scale + 1), axis=1)
            compare(got, want)


def test_double_precision():
    rng = np.random.default_rng(7)
    x = rng.normal(size=(5, 512)) + 1j * rng.normal(size=(5, 512))
    got = run(x.astype(np.complex128), (5, 512), 'cf64', [1])
    compare(got, np.fft.fft(x, axis=1), tol=1e-13)


def test_status_codes():
    from bifrost_b200.libbifrost import _bf
    import ctypes
    a = bf.empty((4, 64), dtype='f32', space='cuda')
    b = bf.empty((4, 64), dtype='f32', space='cuda')
    plan = Fft()
    axes = (ctypes.c_int * 1)(1)
    size = ctypes.c_size_t()
    assert _bf.bfFftInit(plan.obj, a.as_BFarray(), b.as_BFarray(), 1, axes, 0, ctypes.byref(size)) == \
        _bf.BF_STATUS_INVALID_DTYPE                      # real -> real
    c = bf.empty((4, 40), dtype='cf32', space='cuda')
    assert _bf.bfFftInit(plan.obj, a.as_BFarray(), c.as_BFarray(), 1, axes, 0, ctypes.byref(size)) == \
        _bf.BF_STATUS_INVALID_SHAPE                      # r2c needs n/2+1
    d = bf.empty((4, 33), dtype='cf32', space='cuda')
    assert _bf.bfFftInit(plan.obj, a.as_BFarray(), d.as_BFarray(), 1, axes, 1, ctypes.byref(size)) == \
        _bf.BF_STATUS_UNSUPPORTED                        # fftshift on a real transform
