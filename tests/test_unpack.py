"""bfUnpack parity.  CPU: the oracle against every known-answer vector of the
reference (test/test_unpack.py:33-97, test/test_gunpack.py:37-103).  GPU: the
kernel against the same vectors and against the oracle on random data for all
supported dtype / endianness / align_msb / conjugate combinations."""
import itertools

import numpy as np
import pytest

import bifrost_b200 as bf
from oracle import unpack as ounpack

# (input bytes, big_endian, conjugated) -> always the same answer
KNOWN = [
    ([[0x10, 0x32], [0x54, 0x76], [0x98, 0xBA]], False, False),
    ([[0x01, 0x23], [0x45, 0x67], [0x89, 0xAB]], True, False),
    ([[0xF0, 0xD2], [0xB4, 0x96], [0x78, 0x5A]], False, True),
    ([[0x0F, 0x2D], [0x4B, 0x69], [0x87, 0xA5]], True, True),
]
ANSWER = np.array([[(0, 1), (2, 3)], [(4, 5), (6, 7)], [(-8, -7), (-6, -5)]], dtype=np.int8)


@pytest.mark.parametrize("vec,big_endian,conj", KNOWN)
def test_oracle_reproduces_reference_known_answers(vec, big_endian, conj):
    got = ounpack.unpack(np.array(vec, np.uint8), 4, True, byte_reverse=big_endian, conjugate=conj)
    np.testing.assert_array_equal(got.reshape(3, 2, 2), ANSWER)


@pytest.mark.gpu
@pytest.mark.parametrize("vec,big_endian,conj", KNOWN)
@pytest.mark.parametrize("odtype", ['ci8', 'cf32'])
def test_gpu_known_answers(vec, big_endian, conj, odtype):
    iarray = bf.ndarray(np.array(vec, np.uint8).reshape(3, 2).view(bf.DataType('ci4').as_numpy_dtype()),
                        dtype='ci4')
    if big_endian:
        iarray = iarray.byteswap()
    if conj:
        iarray = iarray.conj()
    d_in = bf.asarray(iarray, space='cuda')
    d_out = bf.empty((3, 2), dtype=odtype, space='cuda')
    bf.unpack(d_in, d_out)
    out = np.asarray(d_out.copy('system'))
    if odtype == 'ci8':
        got = np.stack([out['re'], out['im']], -1)
    else:
        got = np.stack([out.real, out.imag], -1)
    np.testing.assert_array_equal(got, ANSWER)


@pytest.mark.gpu
def test_gpu_matches_oracle_all_modes():
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, size=(37, 61), dtype=np.uint8)   # 141 uint4 vectors + 1 tail byte
    for (idt, nbit, signed, cplx) in [('i4', 4, True, False), ('ci4', 4, True, True), ('i2', 2, True, False),
                                      ('ci2', 2, True, True), ('i1', 1, True, False), ('ci1', 1, True, True),
                                      ('u4', 4, False, False), ('u2', 2, False, False)]:
        per = 8 // nbit // (2 if cplx else 1)
        shape = (37, 61 * per)
        odts = (['ci8', 'cf32', 'cf64'] if cplx else ['i8', 'f32', 'f64']) if signed else ['u8']
        for odt, big, msb, conj in itertools.product(odts, [False, True], [False, True],
                                                    [False, True] if cplx else [False]):
            d_raw = bf.asarray(raw, space='cuda')            # packed bytes on the device
            d_in = bf.ndarray(space='cuda', buffer=d_raw.ctypes.data, shape=shape, dtype=idt,
                              native=not big, conjugated=conj)
            d_out = bf.empty(shape, dtype=odt, space='cuda')
            bf.unpack(d_in, d_out, align_msb=msb)
            out = np.asarray(d_out.copy('system'))
            want = ounpack.unpack(raw, nbit, signed, byte_reverse=big, align_msb=msb,
                                  conjugate=conj, gpu=True)
            if odt == 'ci8':
                got = np.stack([out['re'], out['im']], -1).reshape(37, -1)
            elif odt in ('cf32', 'cf64'):
                got = np.stack([out.real, out.imag], -1).reshape(37, -1)
            else:
                got = out
            np.testing.assert_array_equal(got, want.astype(got.dtype), err_msg=str((idt, odt, big, msb, conj)))


@pytest.mark.gpu
def test_status_codes():
    from bifrost_b200.libbifrost import _bf
    a = bf.empty((4, 8), 'ci4', 'cuda')
    b = bf.empty((4, 8), 'i8', 'cuda')
    assert _bf.bfUnpack(a.as_BFarray(), b.as_BFarray(), 0) == _bf.BF_STATUS_INVALID_DTYPE
    h = bf.empty((4, 8), 'ci4', 'system')
    c = bf.empty((4, 8), 'ci8', 'cuda')
    assert _bf.bfUnpack(h.as_BFarray(), c.as_BFarray(), 0) == _bf.BF_STATUS_UNSUPPORTED_SPACE
    d = bf.empty((4, 8), 'ci8', 'cuda')
    e = bf.empty((4, 8), 'ci16', 'cuda')
    assert _bf.bfUnpack(d.as_BFarray(), e.as_BFarray(), 0) == _bf.BF_STATUS_UNSUPPORTED_DTYPE


# ---- system-space arrays: the host path of the ABI (src/unpack.cpp:48-197) ----
@pytest.mark.parametrize("vec,big_endian,conj", KNOWN)
@pytest.mark.parametrize("odtype", ['ci8', 'cf32'])
def test_host_known_answers(vec, big_endian, conj, odtype):
    """The reference's own known-answer vectors (test/test_unpack.py:33-97) through
    bfUnpack on system-space arrays -- runs without a GPU."""
    iarray = bf.ndarray(np.array(vec, np.uint8).reshape(3, 2).view(bf.DataType('ci4').as_numpy_dtype()),
                        dtype='ci4')
    if big_endian:
        iarray = iarray.byteswap()
    if conj:
        iarray = iarray.conj()
    out = bf.empty((3, 2), dtype=odtype, space='system')
    bf.unpack(iarray, out)
    out = np.asarray(out)
    got = np.stack([out['re'], out['im']], -1) if odtype == 'ci8' else np.stack([out.real, out.imag], -1)
    np.testing.assert_array_equal(got, ANSWER)


def test_host_matches_oracle_all_modes():
    """Every dtype / endianness / align_msb / conjugate combination on the host,
    against the oracle's CPU convention (signed 1-bit: bit 1 -> -1, bit 0 -> 0)."""
    rng = np.random.default_rng(2)
    raw = rng.integers(0, 256, size=(5, 23), dtype=np.uint8)
    for (idt, nbit, signed, cplx) in [('i4', 4, True, False), ('ci4', 4, True, True), ('i2', 2, True, False),
                                      ('ci2', 2, True, True), ('i1', 1, True, False), ('ci1', 1, True, True),
                                      ('u4', 4, False, False), ('u2', 2, False, False)]:
        per = 8 // nbit // (2 if cplx else 1)
        shape = (5, 23 * per)
        odts = (['ci8', 'cf32'] if cplx else ['i8', 'f64']) if signed else ['u8']
        for odt, big, msb, conj in itertools.product(odts, [False, True], [False, True],
                                                     [False, True] if cplx else [False]):
            a = bf.ndarray(space='system', shape=shape, dtype=idt, buffer=raw.ctypes.data)
            a.bf.native = not big
            a.bf.conjugated = conj
            out = bf.empty(shape, dtype=odt, space='system')
            bf.unpack(a, out, align_msb=msb)
            o = np.asarray(out)
            if cplx:
                got = (np.stack([o['re'], o['im']], -1) if odt == 'ci8' else np.stack([o.real, o.imag], -1))
            else:
                got = o
            want = ounpack.unpack(raw, nbit, signed, byte_reverse=big, align_msb=msb, conjugate=conj, gpu=False)
            np.testing.assert_array_equal(got.reshape(want.shape).astype(np.float64), want.astype(np.float64),
                                          err_msg=str((idt, odt, big, msb, conj)))
