"""bf.quantize (mirrors python/bifrost/quantize.py:35-41 -> bfQuantize)."""
from bifrost_b200.libbifrost import _bf, _check
from bifrost_b200.ndarray import asarray


def quantize(src, dst, scale=1.):
    """dst = integer(rint(clip(src * scale))); f32/cf32 -> 8/16/32-bit (complex) integers."""
    _check(_bf.bfQuantize(asarray(src).as_BFarray(), asarray(dst).as_BFarray(), float(scale)))
    return dst
