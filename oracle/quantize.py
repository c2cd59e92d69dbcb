"""Quantize oracle (numpy).  TEST INFRASTRUCTURE -- see oracle/__init__.py.
Follows src/quantize.cpp:45-90: out = IntType(rint(clip(in * scale))), clip to
[-max, +max] (signed) or [0, max] (unsigned), rint = round half to even.
Pinned by the known answers of test/test_quantize.py:33-50."""
import numpy as np

_RANGE = {'i8': (-127, 127), 'i16': (-32767, 32767), 'i32': (-2147483647, 2147483647),
          'u8': (0, 255), 'u16': (0, 65535), 'u32': (0, 4294967295)}
_NP = {'i8': np.int8, 'i16': np.int16, 'i32': np.int32, 'u8': np.uint8, 'u16': np.uint16, 'u32': np.uint32}


def quantize(x, kind, scale=1.0):
    """x: float32 or complex64 array; kind: 'i8', 'u16', ... (complex input gives a
    trailing axis of length 2: re, im)."""
    x = np.asarray(x)
    if np.iscomplexobj(x):
        x = np.stack([x.real, x.imag], -1)
    v = (x.astype(np.float32) * np.float32(scale)).astype(np.float32)
    lo, hi = _RANGE[kind]
    # The reference clips in the float type: F(-max) and F(+max) (quantize.cpp:60-62).
    # For 32-bit outputs those are -2^31 / +2^31 (or 2^32): the low one converts
    # exactly (INT_MIN), the high one is out of range -- undefined in the
    # reference, saturating on the GPU, which the second clip reproduces.
    v = np.clip(v, np.float32(lo), np.float32(hi))
    r = np.rint(v.astype(np.float64))
    return np.minimum(r, hi).astype(_NP[kind])


def quantize_packed(x, nbit, scale=1.0):
    """Sub-byte outputs (src/guantize.cu:146-348): values clip to [-7, 7] (4-bit)
    or [-1, 1] (2-bit) and round half to even; 1-bit is (x * scale >= 0).  The
    first value of every byte goes to the most significant bits.  Returns uint8
    [..., n * nbit / 8].  NOTE: the reference's 1-bit kernel masks its first
    three values away (0x08 / 0x04 / 0x02 on bits 7..5); `reference_1bit_mask`
    is the part of the byte on which it agrees with this definition."""
    x = np.asarray(x)
    if np.iscomplexobj(x):
        x = np.stack([x.real, x.imag], -1).reshape(x.shape[:-1] + (2 * x.shape[-1],))
    v = (x.astype(np.float32) * np.float32(scale)).astype(np.float32)
    if nbit == 1:
        q = (v >= 0).astype(np.uint8)
    else:
        lim = 7 if nbit == 4 else 1
        q = np.rint(np.clip(v, -lim, lim)).astype(np.int8).view(np.uint8) & np.uint8((1 << nbit) - 1)
    per = 8 // nbit
    q = q.reshape(q.shape[:-1] + (q.shape[-1] // per, per))
    out = np.zeros(q.shape[:-1], np.uint8)
    for k in range(per):
        out |= (q[..., k] << np.uint8(8 - nbit * (k + 1))).astype(np.uint8)
    return out


reference_1bit_mask = 0x1F
