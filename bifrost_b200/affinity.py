"""Thread placement (ref: python/bifrost/affinity.py; C side csrc/ring.cpp)."""
from bifrost_b200.libbifrost import _bf, _check, _get, _array

import ctypes


def get_core():
    """The core the calling thread is bound to, -1 if it may run on several."""
    return _get(_bf.bfAffinityGetCore)


def set_core(core):
    """Binds the calling thread to `core`; -1 unbinds."""
    _check(_bf.bfAffinitySetCore(core))


def set_openmp_cores(cores):
    _check(_bf.bfAffinitySetOpenMPCores(len(cores), _array(cores, ctypes.c_int)))
