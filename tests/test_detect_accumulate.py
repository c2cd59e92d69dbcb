"""detect / accumulate parity: the fixed kernels behind bfDetect, bfAccumulate
and the bfMap strings of blocks/detect.py:87-136 and blocks/accumulate.py:67
against the oracle's restatement of those formulae (the reference has no test
for detect: parity unpinned by the reference, the formulae are the spec)."""
import numpy as np
import pytest

import bifrost_b200 as bf
from oracle import ops as oracle

pytestmark = pytest.mark.gpu
NPOL_OUT = {'stokes': 4, 'coherence': 4, 'jones': 2, 'stokes_i': 1}


def make_cf32(rng, shape):
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)


@pytest.mark.parametrize("mode", ['stokes', 'coherence', 'jones', 'stokes_i'])
@pytest.mark.parametrize("shape,axis", [((3, 2, 40, 64), 1), ((2, 5, 33), 0), ((6, 17, 2), 2),
                                        ((1, 2, 4096), 1)])
def test_detect_modes(mode, shape, axis):
    rng = np.random.default_rng(11)
    x = make_cf32(rng, shape)
    oshape = list(shape)
    oshape[axis] = NPOL_OUT[mode]
    d_x = bf.asarray(x, space='cuda')
    d_y = bf.empty(oshape, dtype='cf32' if mode == 'jones' else 'f32', space='cuda')
    bf.detect(d_x, d_y, mode, axis)
    got = np.asarray(d_y.copy('system'))
    want = oracle.detect(x, mode, axis)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    # independent fp64 statement of the physics
    x64 = np.moveaxis(x.astype(np.complex128), axis, 0)
    if mode == 'stokes':
        ref = np.stack([abs(x64[0])**2 + abs(x64[1])**2, abs(x64[0])**2 - abs(x64[1])**2,
                        2 * (x64[0] * x64[1].conj()).real, -2 * (x64[0] * x64[1].conj()).imag])
        np.testing.assert_allclose(np.moveaxis(got, axis, 0), ref, rtol=1e-5, atol=1e-5)


def test_detect_scalar_and_integer_inputs():
    rng = np.random.default_rng(12)
    x = make_cf32(rng, (7, 100, 3))
    d_y = bf.empty(x.shape, dtype='f32', space='cuda')
    bf.detect(bf.asarray(x, space='cuda'), d_y, 'scalar')
    np.testing.assert_allclose(np.asarray(d_y.copy('system')), oracle.detect(x, 'scalar'), rtol=1e-6)
    xi = np.empty((4, 2, 300), dtype=bf.DataType('ci8').as_numpy_dtype())
    xi['re'] = rng.integers(-127, 128, size=xi.shape)
    xi['im'] = rng.integers(-127, 128, size=xi.shape)
    d_s = bf.empty((4, 4, 300), dtype='f32', space='cuda')
    bf.detect(bf.asarray(xi, space='cuda'), d_s, 'stokes', 1)
    np.testing.assert_array_equal(np.asarray(d_s.copy('system')), oracle.detect(xi, 'stokes', 1))


def test_detect_through_bfmap_strings():
    """The exact strings DetectBlock.on_data builds (blocks/detect.py:96-136)."""
    rng = np.random.default_rng(13)
    x = make_cf32(rng, (5, 2, 64))
    axis = 1
    inds = ['i%i' % i for i in range(x.ndim)]
    inds[axis] = '%i'
    inds_pol = ','.join(inds)
    inds_ = [inds_pol % i for i in range(4)]
    names = inds[:axis] + inds[axis + 1:]
    func = """
                Complex<b_type> x = a(%s);
                Complex<b_type> y = a(%s);
                auto xx = x.mag2();
                auto yy = y.mag2();
                auto xy = x*y.conj();
                b(%s) = xx + yy;
                b(%s) = xx - yy;
                b(%s) =  2*xy.real;
                b(%s) = -2*xy.imag;
                """ % (inds_[0], inds_[1], inds_[0], inds_[1], inds_[2], inds_[3])
    d_x = bf.asarray(x, space='cuda')
    d_y = bf.empty((5, 4, 64), dtype='f32', space='cuda')
    bf.map(func, shape=(5, 64), axis_names=names, data={'a': d_x, 'b': d_y})
    np.testing.assert_allclose(np.asarray(d_y.copy('system')), oracle.detect(x, 'stokes', 1), rtol=1e-6, atol=1e-6)
    d_z = bf.empty(x.shape, dtype='f32', space='cuda')
    bf.map("b = Complex<b_type>(a).mag2()", {'a': d_x, 'b': d_z})
    np.testing.assert_allclose(np.asarray(d_z.copy('system')), oracle.detect(x, 'scalar'), rtol=1e-6)
    with pytest.raises(RuntimeError):
        bf.map("b = a + 1", {'a': d_x, 'b': d_z})


def test_accumulate():
    rng = np.random.default_rng(14)
    frames = rng.normal(size=(8, 4, 1000)).astype(np.float32)
    d_b = bf.asarray(np.full((4, 1000), np.nan, np.float32), space='cuda')   # beta=0 must ignore it
    want = None
    for k in range(8):
        beta = 0. if k == 0 else 1.
        d_a = bf.asarray(frames[k], space='cuda')
        if k % 2:
            bf.accumulate(d_a, d_b, beta)
        else:
            bf.map("b = beta * b + (b_type)a", {'a': d_a, 'b': d_b, 'beta': beta})
        want = oracle.accumulate(frames[k], want if want is not None else frames[k], beta)
    np.testing.assert_array_equal(np.asarray(d_b.copy('system')), want)
    # complex and integer inputs, odd length (scalar path)
    c = (rng.normal(size=(3, 77)) + 1j * rng.normal(size=(3, 77))).astype(np.complex64)
    d_c = bf.asarray(c, space='cuda')
    d_acc = bf.asarray(c, space='cuda')
    bf.accumulate(d_c, d_acc, 1.0)
    np.testing.assert_array_equal(np.asarray(d_acc.copy('system')), c + c)
    i8 = rng.integers(-100, 100, size=(5, 33)).astype(np.int8)
    d_f = bf.zeros((5, 33), dtype='f32', space='cuda')
    bf.accumulate(bf.asarray(i8, space='cuda'), d_f, 1.0)
    np.testing.assert_array_equal(np.asarray(d_f.copy('system')), i8.astype(np.float32))
