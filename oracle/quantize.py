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
