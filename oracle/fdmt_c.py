"""ctypes wrapper of oracle/fdmt_c.c (OpenMP C restatement of the reference
FDMT execution).  TEST / BENCHMARK INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from oracle.fdmt import FdmtPlan

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'libfdmt_oracle.so')
_lib = None
_KIND = {np.dtype(np.int8): 0, np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.uint16): 3,
         np.dtype(np.int32): 4, np.dtype(np.uint32): 5, np.dtype(np.float32): 6}


def available():
    global _lib
    if _lib is None and os.path.exists(_PATH):
        try:
            _lib = ctypes.CDLL(_PATH)
            _lib.fdmt_c_max_threads.restype = ctypes.c_int
        except OSError:
            _lib = None
    return _lib is not None


def max_threads():
    return _lib.fdmt_c_max_threads() if available() else 1


class Plan(object):
    def __init__(self, nchan, max_delay, f0, df, exponent=-2.0):
        assert available()
        self.p = FdmtPlan(nchan, max_delay, f0, df, exponent)
        self.offsets = np.ascontiguousarray(self.p.row_offsets[0], dtype=np.int64)
        self.nrow = np.ascontiguousarray(self.p.nrow, dtype=np.int64)
        self.src = [np.zeros((1, 2), np.int64)] + [np.ascontiguousarray(s, dtype=np.int64) for s in self.p.srcrows[1:]]
        self.dly = [np.zeros(1, np.int64)] + [np.ascontiguousarray(d, dtype=np.int64) for d in self.p.delays[1:]]
        PL = ctypes.POINTER(ctypes.c_long)
        self._src_ptrs = (PL * self.p.nstep)(*[s.ctypes.data_as(PL) for s in self.src])
        self._dly_ptrs = (PL * self.p.nstep)(*[d.ctypes.data_as(PL) for d in self.dly])
        self._bufs = None

    def execute(self, x, out, threads=0):
        """x: [nchan, ntime] contiguous; out: [max_delay, ntime] float32 (cells the
        reference does not write are left as they are)."""
        x = np.ascontiguousarray(x)
        nchan, ntime = x.shape
        assert out.dtype == np.float32 and out.flags['C_CONTIGUOUS']
        need = int(self.p.nrow_max) * ntime
        if self._bufs is None or self._bufs[0].size < need:
            self._bufs = (np.empty(need, np.float32), np.empty(need, np.float32))
        PL = ctypes.POINTER(ctypes.c_long)
        rc = _lib.fdmt_c_execute(
            ctypes.c_void_p(x.ctypes.data), ctypes.c_int(_KIND[x.dtype]), ctypes.c_long(nchan),
            ctypes.c_long(ntime), ctypes.c_int(1 if self.p.reverse_band else 0),
            self.offsets.ctypes.data_as(PL), ctypes.c_int(self.p.nstep), self.nrow.ctypes.data_as(PL),
            self._src_ptrs, self._dly_ptrs, ctypes.c_void_p(out.ctypes.data), ctypes.c_long(out.shape[-1]),
            ctypes.c_void_p(self._bufs[0].ctypes.data), ctypes.c_void_p(self._bufs[1].ctypes.data),
            ctypes.c_int(int(threads)))
        assert rc == 0
        return out
