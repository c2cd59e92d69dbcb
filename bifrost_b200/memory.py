"""Raw space-aware allocation helpers (mirrors python/bifrost/memory.py)."""
import ctypes
from bifrost_b200.libbifrost import _bf, _check, _get, _string2space


def space_accessible(space, from_spaces):
    """True if memory in `space` can be dereferenced from any of `from_spaces`."""
    space = str(space)
    if from_spaces == 'any':
        return True
    from_spaces = set(str(s) for s in from_spaces)
    if space in from_spaces:
        return True
    if space in ('cuda_host', 'cuda_managed'):
        return 'cuda' in from_spaces or 'system' in from_spaces
    return False


def raw_malloc(size, space):
    ptr = ctypes.c_void_p()
    _check(_bf.bfMalloc(ctypes.byref(ptr), int(size), _string2space(str(space))))
    return ptr.value


def raw_free(ptr, space='auto'):
    _check(_bf.bfFree(ptr, _string2space(str(space))))


def raw_get_space(ptr):
    return _get(_bf.bfGetSpace, ptr)


def alignment():
    return _bf.bfGetAlignment()
