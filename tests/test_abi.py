"""CPU checks of the drop-in boundary: the shared library loads, exports every
entry point include/bifrost_b200.h declares, and the ABI constants/struct
layout match the reference's headers (src/bifrost/*.h)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'bifrost_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bf[A-Z]\w*)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, 'bifrost_b200', 'lib', 'libbifrost_b200.so'))
    names = declared_functions()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    from bifrost_b200.libbifrost import EXPORTED_SYMBOLS
    assert set(declared_functions()) == set(EXPORTED_SYMBOLS)


def test_forwarding_headers_exist():
    for name in ['common', 'memory', 'array', 'cuda', 'transpose', 'reduce', 'fdmt',
                 'fft', 'linalg', 'unpack', 'map']:
        assert os.path.exists(os.path.join(ROOT, 'include', 'bifrost', name + '.h'))


def test_bfarray_layout_and_constants():
    from bifrost_b200.libbifrost import _bf, _th, BFarray
    assert ctypes.sizeof(BFarray) == 168                      # array.h:206-222
    assert BFarray.shape.offset == 24 and BFarray.strides.offset == 88
    assert _bf.BF_DTYPE_CI8 == (8 | 0x100000)                 # array.h:163
    assert _bf.BF_DTYPE_CF32 == (32 | 0x200 | 0x100000)
    assert _bf.BF_DTYPE_U8 == (8 | 0x100)
    assert _bf.BF_SPACE_CUDA == 2 and _bf.BF_SPACE_CUDA_HOST == 3   # memory.h:43-49
    assert _bf.BF_STATUS_UNSUPPORTED == 48 and _bf.BF_STATUS_INTERNAL_ERROR == 99
    assert int(_th.BFreduce_enum.pwrsum) == 5                  # reduce.h:44-55
    assert int(_th.BFspace_enum.cuda) == 2
    assert _bf.bfGetStatusString(13) == b'BF_STATUS_INVALID_SHAPE'
    assert _bf.bfGetSpaceString(3) == b'cuda_host'
    assert _bf.bfGetAlignment() == 4096
    assert _bf.bfGetCudaEnabled() == 1


def test_host_only_runtime_paths():
    """bfMalloc/bfMemcpy/bfArrayCopy in system space work without a device."""
    import bifrost_b200 as bf
    a = bf.asarray(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    b = bf.empty_like(a)
    bf.copy_array(b, a)
    np.testing.assert_array_equal(np.asarray(b), np.asarray(a))
    c = bf.zeros((5, 7), dtype='ci8')
    assert np.asarray(c)['re'].sum() == 0
    assert c.bf.dtype == 'ci8' and c.as_BFarray().strides[1] == 2
    # strided host copy through the C ABI
    src = bf.asarray(np.arange(60, dtype=np.int16).reshape(3, 4, 5))
    view = src[:, 1:3, ::2]
    dst = bf.empty(view.shape, dtype='i16')
    from bifrost_b200.libbifrost import _bf, _check
    _check(_bf.bfArrayCopy(dst.as_BFarray(), view.as_BFarray()))
    np.testing.assert_array_equal(np.asarray(dst), np.asarray(src)[:, 1:3, ::2])


def test_invalid_arguments_return_status_not_crash():
    from bifrost_b200.libbifrost import _bf
    assert _bf.bfTranspose(None, None, None) == _bf.BF_STATUS_INVALID_POINTER
    assert _bf.bfReduce(None, None, 0) == _bf.BF_STATUS_INVALID_POINTER
    assert _bf.bfFdmtInit(None, 16, 4, 1., 1., -2., 2, None, None) == _bf.BF_STATUS_INVALID_HANDLE
    n = ctypes.c_int()
    assert _bf.bfFdmtPlanQuery(1, 4, 1., 1., -2., -1, ctypes.byref(n), None) == \
        _bf.BF_STATUS_INVALID_ARGUMENT
    with pytest.raises(RuntimeError):
        from bifrost_b200.libbifrost import _check
        _check(_bf.BF_STATUS_INVALID_SHAPE)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under bifrost_b200/ (Python or
    C++/CUDA) may import, include or load it, and the binding must fail loudly
    when the CUDA library is missing (no CPU fallback)."""
    import re
    pkg = os.path.join(ROOT, 'bifrost_b200')
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.hpp', '.h')):
                text = open(os.path.join(dirpath, f), errors='replace').read()
                if re.search(r'^\s*(from|import)\s+oracle\b|#include\s+"[^"]*oracle|libfdmt_oracle|oracle/_ref', text, re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == []
    src = open(os.path.join(pkg, 'libbifrost.py')).read()
    assert 'raise' in src and 'libbifrost_b200' in src
