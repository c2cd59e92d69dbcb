"""bfTranspose parity (bit-exact) -- the sweep of test/test_transpose.py:46-74
(all permutations x shapes x element sizes 1..16) plus 4-D, odd element sizes,
strided views and the GUPPI-chain permutation."""
import itertools

import numpy as np
import pytest

import bifrost_b200 as bf

pytestmark = pytest.mark.gpu


def run(a, axes, dtype=None):
    d_a = bf.asarray(a, space='cuda')
    oshape = [a.shape[i] for i in axes]
    d_b = bf.empty(oshape, dtype=d_a.bf.dtype, space='cuda')
    bf.memset_array(d_b, 0)
    bf.transpose(d_b, d_a, axes)
    return np.asarray(d_b.copy('system'))


def elem_dtype(size):
    return {1: np.uint8, 2: np.uint16, 4: np.float32, 8: np.complex64, 16: np.complex128}[size]


@pytest.mark.parametrize("shape", [(7, 5, 3), (32, 32, 32), (127, 65, 33), (256, 64, 16), (3, 300, 129)])
@pytest.mark.parametrize("esize", [1, 2, 4, 8, 16])
def test_all_3d_permutations(shape, esize):
    rng = np.random.default_rng(1234)
    raw = rng.integers(0, 256, size=shape + (esize,), dtype=np.uint8)
    a = raw.view(elem_dtype(esize)).reshape(shape)
    for axes in itertools.permutations(range(3)):
        got = run(a, axes)
        want = np.transpose(a, axes)
        assert got.tobytes() == np.ascontiguousarray(want).tobytes(), (shape, esize, axes)


@pytest.mark.parametrize("esize", [1, 2, 4, 8])
def test_short_inner_dims_interleave_paths(esize):
    """2- and 4-long dims innermost on either side ((de)interleave kernels),
    aligned and misaligned."""
    rng = np.random.default_rng(8)
    for shape in [(6, 64, 2), (3, 5, 128, 4), (4, 2, 96), (2, 4, 7, 48), (5, 50, 2), (3, 31, 4)]:
        raw = rng.integers(0, 256, size=shape + (esize,), dtype=np.uint8)
        a = raw.view(elem_dtype(esize)).reshape(shape)
        for axes in itertools.permutations(range(len(shape))):
            got = run(a, axes)
            assert got.tobytes() == np.ascontiguousarray(np.transpose(a, axes)).tobytes(), (shape, esize, axes)
    base = rng.integers(0, 255, size=(4, 130, 2), dtype=np.uint8)
    d = bf.asarray(base, space='cuda')
    view = d[:, 1:129, :]                       # misaligned rows
    out = bf.empty((4, 2, 128), 'u8', 'cuda')
    bf.transpose(out, view, (0, 2, 1))
    np.testing.assert_array_equal(np.asarray(out.copy('system')), np.transpose(base[:, 1:129, :], (0, 2, 1)))


def test_4d_and_5d():
    rng = np.random.default_rng(2)
    a = rng.integers(-128, 127, size=(3, 17, 4, 50), dtype=np.int8)
    for axes in itertools.permutations(range(4)):
        np.testing.assert_array_equal(run(a, axes), np.transpose(a, axes))
    b = rng.normal(size=(2, 3, 5, 7, 11)).astype(np.float32)
    for axes in [(4, 3, 2, 1, 0), (0, 2, 4, 1, 3), (1, 0, 3, 2, 4), (3, 4, 0, 1, 2)]:
        np.testing.assert_array_equal(run(b, axes), np.transpose(b, axes))


def test_guppi_chain_permutation():
    """[time, freq, fine_time, pol] ci8 -> [time, pol, freq, fine_time]
    (testbench/gpuspec_simple.py:47; the reference's vector_read special case)."""
    rng = np.random.default_rng(3)
    a = np.empty((2, 64, 1024, 2), dtype=bf.DataType('ci8').as_numpy_dtype())
    a['re'] = rng.integers(-127, 128, size=a.shape)
    a['im'] = rng.integers(-127, 128, size=a.shape)
    got = run(a, (0, 3, 1, 2))
    want = np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))
    assert got.tobytes() == want.tobytes()


def test_negative_axes_and_config1():
    rng = np.random.default_rng(4)
    a = rng.normal(size=(4096, 256)).astype(np.float32)        # BASELINE config 1
    np.testing.assert_array_equal(run(a, (-1, -2)), a.T)
    np.testing.assert_array_equal(run(a, (1, 0)), a.T)


def test_strided_device_views_and_misaligned_bases():
    rng = np.random.default_rng(5)
    base = rng.integers(0, 255, size=(40, 50, 70), dtype=np.uint8)
    d = bf.asarray(base, space='cuda')
    for sl in [np.s_[1:33, 3:35, 5:69], np.s_[::2, 1:, ::3], np.s_[3:, :, 1:2]]:
        view = d[sl]
        hview = base[sl]
        for axes in [(2, 1, 0), (1, 0, 2), (0, 2, 1), (2, 0, 1)]:
            out = bf.empty([hview.shape[i] for i in axes], dtype='u8', space='cuda')
            bf.transpose(out, view, axes)
            np.testing.assert_array_equal(np.asarray(out.copy('system')), np.transpose(hview, axes))


def test_odd_element_sizes():
    """Elements of 3/6/12 bytes (vector dtypes) move as opaque words."""
    from bifrost_b200.libbifrost import _bf, _check
    import ctypes
    rng = np.random.default_rng(6)
    for esize in [3, 6, 12]:
        shape = (9, 14, 5)
        raw = rng.integers(0, 256, size=shape + (esize,), dtype=np.uint8)
        d_in = bf.asarray(raw, space='cuda')
        d_out = bf.empty((5, 9, 14, esize), dtype='u8', space='cuda')
        a, b = d_in.as_BFarray(), d_out.as_BFarray()
        vec = (8 | 0x100) | ((esize - 1) << 12)          # u8 vector of length esize
        for arr, shp in ((a, shape), (b, (5, 9, 14))):
            arr.ndim = 3
            arr.dtype = vec
            strides = [shp[1] * shp[2] * esize, shp[2] * esize, esize]
            for i in range(3):
                arr.shape[i] = shp[i]
                arr.strides[i] = strides[i]
        axes = (ctypes.c_int * 3)(2, 0, 1)
        _check(_bf.bfTranspose(a, b, axes))
        got = np.asarray(d_out.copy('system'))
        np.testing.assert_array_equal(got, np.transpose(raw, (2, 0, 1, 3)))


def test_noncontiguous_copy_uses_device_path():
    rng = np.random.default_rng(7)
    base = rng.normal(size=(6, 20, 30)).astype(np.float32)
    d = bf.asarray(base, space='cuda')
    got = np.asarray(d[:, 2:18:3, 1::2].copy('system'))
    np.testing.assert_array_equal(got, base[:, 2:18:3, 1::2])


def test_errors():
    from bifrost_b200.libbifrost import _bf
    import ctypes
    a = bf.empty((4, 6), dtype='f32', space='cuda')
    b = bf.empty((6, 5), dtype='f32', space='cuda')
    axes = (ctypes.c_int * 2)(1, 0)
    assert _bf.bfTranspose(a.as_BFarray(), b.as_BFarray(), axes) == _bf.BF_STATUS_INVALID_SHAPE
    c = bf.empty((6, 4), dtype='i32', space='cuda')
    assert _bf.bfTranspose(a.as_BFarray(), c.as_BFarray(), axes) == _bf.BF_STATUS_INVALID_DTYPE
    e = bf.empty((5, 5), dtype='f32', space='cuda')
    d = bf.empty((5, 5), dtype='f32', space='cuda')
    bad = (ctypes.c_int * 2)(0, 0)
    assert _bf.bfTranspose(e.as_BFarray(), d.as_BFarray(), bad) == _bf.BF_STATUS_INVALID_ARGUMENT
