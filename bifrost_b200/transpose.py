"""bf.transpose (mirrors python/bifrost/transpose.py:42-51 -> bfTranspose)."""
import ctypes
from bifrost_b200.libbifrost import _bf, _check
from bifrost_b200.ndarray import asarray


def transpose(dst, src, axes=None):
    """dst = src.transpose(axes) on the device; axes defaults to reversal."""
    if axes is None:
        axes = list(reversed(range(len(dst.shape))))
    axes = list(axes)
    dst_bf = asarray(dst).as_BFarray()
    src_bf = asarray(src).as_BFarray()
    axes_array = (ctypes.c_int * len(axes))(*axes)
    _check(_bf.bfTranspose(src_bf, dst_bf, axes_array))
    return dst
