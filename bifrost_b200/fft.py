"""bf.fft.Fft (mirrors python/bifrost/fft.py:38-75 -> bfFft*)."""
import ctypes
from bifrost_b200.libbifrost import _bf, _check, _get, BifrostObject
from bifrost_b200.ndarray import asarray


class Fft(BifrostObject):
    def __init__(self):
        BifrostObject.__init__(self, _bf.bfFftCreate, _bf.bfFftDestroy)

    def init(self, iarray, oarray, axes=None, apply_fftshift=False):
        if isinstance(axes, int):
            axes = [axes]
        if axes is None:
            ndim, axes_c = 1, None
        else:
            ndim = len(axes)
            axes_c = (ctypes.c_int * ndim)(*axes)
        self.workspace_size = _get(_bf.bfFftInit, self.obj,
                                   asarray(iarray).as_BFarray(),
                                   asarray(oarray).as_BFarray(),
                                   ndim, axes_c, apply_fftshift)

    def execute(self, iarray, oarray, inverse=False):
        return self.execute_workspace(iarray, oarray, None, 0, inverse)

    def execute_workspace(self, iarray, oarray, workspace_ptr, workspace_size, inverse=False):
        _check(_bf.bfFftExecute(self.obj, asarray(iarray).as_BFarray(),
                                asarray(oarray).as_BFarray(), inverse,
                                workspace_ptr, workspace_size))
        return oarray
