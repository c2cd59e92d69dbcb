"""bf.unpack (mirrors python/bifrost/unpack.py:36-40 -> bfUnpack)."""
from bifrost_b200.libbifrost import _bf, _check
from bifrost_b200.ndarray import asarray


def unpack(src, dst, align_msb=False):
    _check(_bf.bfUnpack(asarray(src).as_BFarray(), asarray(dst).as_BFarray(), align_msb))
    return dst
