"""bf.map and the fixed-kernel entry points behind it.

The reference's ``bf.map`` JIT-compiles arbitrary expressions with NVRTC
(python/bifrost/map.py -> bfMap, src/map.cpp).  On the hot path it is used for
exactly two things: the detect expressions (blocks/detect.py:87-136) and the
accumulate expression (blocks/accumulate.py:67).  This build ships those as
fixed sm_100a kernels; ``bf.map`` recognises those expressions and returns
BF_STATUS_UNSUPPORTED (RuntimeError) for anything else.
"""
import ctypes

import numpy as np

from bifrost_b200.libbifrost import _bf, _check, _array, BFarray
from bifrost_b200.ndarray import asarray, ndarray

DETECT_MODES = {'scalar': 0, 'jones': 1, 'stokes': 2, 'stokes_i': 3, 'coherence': 4}


def detect(idata, odata, mode='stokes', axis=None):
    """odata = detect(idata): |x|^2 ('scalar') or polarisation products along
    `axis` (length-2 pol axis of idata).  Semantics of blocks/detect.py:86-138."""
    mode = DETECT_MODES[mode] if isinstance(mode, str) else int(mode)
    if axis is None:
        axis = 0
    _check(_bf.bfDetect(asarray(idata).as_BFarray(), asarray(odata).as_BFarray(),
                        mode, int(axis)))
    return odata


def accumulate(idata, odata, beta=1.0):
    """odata = beta*odata + idata  (blocks/accumulate.py:63-74)."""
    _check(_bf.bfAccumulate(asarray(idata).as_BFarray(), asarray(odata).as_BFarray(),
                            float(beta)))
    return odata


def _is_scalar(x):
    return isinstance(x, (int, float, complex, np.number))


def map(func_string, data, axis_names=None, shape=None, func_name=None,
        extra_code=None, block_shape=None, block_axes=None):
    """Apply `func_string` to the arrays in `data` (dict name -> array/scalar).
    Signature of python/bifrost/map.py:map; only the hot-path expressions are
    compiled in (see module docstring)."""
    narg = len(data)
    ndim = len(shape) if shape is not None else 0
    arg_arrays, arg_names, keepalive = [], [], []
    for key, arg in data.items():
        if _is_scalar(arg):
            arr = np.array(arg)
            if isinstance(arg, int):
                arr = arr.astype(np.int64)
            elif isinstance(arg, float):
                arr = arr.astype(np.float64)
            arr = arr.reshape(1).view(ndarray)
            arr.flags['WRITEABLE'] = False
            arg = arr
        arg = asarray(arg)
        keepalive.append(arg)
        arg_arrays.append(arg.as_BFarray())
        arg_names.append(key)
    if block_axes is not None and axis_names is not None:
        block_axes = [axis_names.index(a) if isinstance(a, str) else a for a in block_axes]
    _check(_bf.bfMap(ndim, _array(shape, dtype=ctypes.c_long), _array(axis_names),
                     narg, _array(arg_arrays), _array(arg_names),
                     func_name.encode() if isinstance(func_name, str) else func_name,
                     func_string.encode() if isinstance(func_string, str) else func_string,
                     extra_code.encode() if isinstance(extra_code, str) else extra_code,
                     _array(block_shape), _array(block_axes)))


def clear_map_cache():
    _check(_bf.bfMapClearCache())
