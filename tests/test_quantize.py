"""bfQuantize: the reference's known answers (test/test_quantize.py:33-50) and
random data against the oracle, bit-exact."""
import numpy as np
import pytest

import bifrost_b200 as bf
from oracle import quantize as oquant

KNOWN_IN = np.array([[0.4 + 0.5j, 1.4 + 1.5j], [2.4 + 2.5j, 3.4 + 3.5j], [4.4 + 4.5j, 5.4 + 5.5j]], np.complex64)
KNOWN_OUT = np.array([[(0, 0), (1, 2)], [(2, 2), (3, 4)], [(4, 4), (5, 6)]])


@pytest.mark.parametrize("kind", ['i8', 'i16', 'i32'])
def test_oracle_reproduces_reference_known_answers(kind):
    np.testing.assert_array_equal(oquant.quantize(KNOWN_IN, kind), KNOWN_OUT)


@pytest.mark.gpu
@pytest.mark.parametrize("odtype", ['ci8', 'ci16', 'ci32'])
def test_gpu_known_answers(odtype):
    d_in = bf.asarray(KNOWN_IN, space='cuda')
    d_out = bf.empty((3, 2), dtype=odtype, space='cuda')
    bf.quantize(d_in, d_out)
    out = np.asarray(d_out.copy('system'))
    np.testing.assert_array_equal(np.stack([out['re'], out['im']], -1), KNOWN_OUT)


@pytest.mark.gpu
@pytest.mark.parametrize("odtype", ['i8', 'i16', 'i32', 'u8', 'u16', 'u32', 'ci8', 'ci16'])
@pytest.mark.parametrize("scale", [1.0, 0.37, 1e6])
def test_gpu_matches_oracle(odtype, scale):
    rng = np.random.default_rng(5)
    n = 100003
    cplx = odtype.startswith('c')
    x = (rng.normal(size=n) * 300).astype(np.float32)
    x[:8] = [0.5, 1.5, 2.5, -0.5, -1.5, 1e20, -1e20, 127.5]
    if odtype.endswith('32'):
        x[5:7] = [1e3, -1e3]          # saturation of 32-bit outputs is undefined in the reference
    if cplx:
        x = (x + 1j * (rng.normal(size=n) * 300).astype(np.float32)).astype(np.complex64)
    d_out = bf.empty((n,), dtype=odtype, space='cuda')
    bf.quantize(bf.asarray(x, space='cuda'), d_out, scale)
    out = np.asarray(d_out.copy('system'))
    got = np.stack([out['re'], out['im']], -1) if cplx else out
    np.testing.assert_array_equal(got, oquant.quantize(x, odtype.lstrip('c'), scale))


@pytest.mark.gpu
def test_status_codes():
    from bifrost_b200.libbifrost import _bf
    a = bf.empty((4,), 'f32', 'cuda')
    b = bf.empty((5,), 'i8', 'cuda')
    c = bf.empty((4,), 'ci8', 'cuda')
    assert _bf.bfQuantize(a.as_BFarray(), b.as_BFarray(), 1.0) == _bf.BF_STATUS_INVALID_SHAPE
    assert _bf.bfQuantize(a.as_BFarray(), c.as_BFarray(), 1.0) == _bf.BF_STATUS_INVALID_DTYPE
