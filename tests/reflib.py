"""Loader for the reference CUDA library built by oracle/ref_build.sh (GPU
oracle).  Returns None when oracle/_ref/libbifrost_ref.so did not travel."""
import ctypes
import os

from bifrost_b200.libbifrost import _PROTOTYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, 'oracle', '_ref', 'libbifrost_ref.so')
_cache = {}


class _Ref(object):
    pass


def load():
    if 'ref' in _cache:
        return _cache['ref']
    ref = None
    if os.path.exists(_PATH):
        try:
            lib = ctypes.CDLL(_PATH, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)
            ref = _Ref()
            for name, (res, args) in _PROTOTYPES.items():
                fn = getattr(lib, name, None)
                if fn is None:
                    continue
                fn.restype, fn.argtypes = res, args
                setattr(ref, name, fn)
        except OSError:
            ref = None
    _cache['ref'] = ref
    return ref
