"""``bf.ndarray``: a numpy subclass whose buffer may live in any Bifrost memory
space, carrying the Bifrost dtype/space metadata needed to build a ``BFarray``.

Same user-facing behaviour as the reference's python/bifrost/ndarray.py
(asarray/empty/zeros/*_like/copy_array/memset_array, ``arr.bf.space``,
``arr.bf.dtype``, ``arr.copy(space=...)``, ``arr.as_BFarray()``, slicing of
device arrays without touching the data, packed sub-byte dtypes whose numpy
shape has the last dim divided by 8/nbit -- ref ndarray.py:235-245,329-332).

Ownership differs from the reference: device/host allocations are owned by an
``_Allocation`` object that sits at the bottom of numpy's ``.base`` chain, so
views keep their memory alive and nothing is freed twice.
"""

import ctypes

import numpy as np

from bifrost_b200 import device
from bifrost_b200.DataType import DataType
from bifrost_b200.Space import Space
from bifrost_b200.libbifrost import _bf, _check, _space2string, BFarray
from bifrost_b200.memory import raw_malloc, raw_free, raw_get_space, space_accessible


class _Allocation(object):
    """Owns `nbyte` bytes at `ptr` in `space`; exposes them to numpy through the
    array interface without numpy ever dereferencing device memory."""

    def __init__(self, nbyte, space, ptr=None):
        self.nbyte = max(int(nbyte), 1)
        self.space = str(space)
        self.owned = ptr is None
        self.ptr = raw_malloc(self.nbyte, self.space) if ptr is None else int(ptr)
        self.__array_interface__ = {'data': (self.ptr, False), 'shape': (self.nbyte,),
                                    'typestr': '|u1', 'version': 3}

    def __del__(self):
        if getattr(self, 'owned', False) and self.ptr:
            try:
                raw_free(self.ptr, self.space)
            except Exception:
                pass
            self.ptr = 0


class BFArrayInfo(object):
    def __init__(self, space, dtype, native=True, conjugated=False):
        self.space = str(space)
        self.dtype = DataType(dtype)
        self.native = native
        self.conjugated = conjugated


def _packed_shape(shape, dtype):
    """Numpy-side shape/itemsize for a logical shape (folds sub-byte packing)."""
    bits = dtype.itemsize_bits
    shape = list(shape)
    if bits < 8:
        per_byte = 8 // bits
        if not shape or shape[-1] % per_byte:
            raise ValueError("Array cannot be packed")
        shape[-1] //= per_byte
        return shape, 1
    return shape, bits // 8


class ndarray(np.ndarray):
    def __new__(cls, base=None, space=None, shape=None, dtype=None, buffer=None,
                offset=0, strides=None, native=None, conjugated=None):
        if isinstance(shape, (int, np.integer)):
            shape = [int(shape)]
        if base is not None:
            if (shape is not None or buffer is not None or offset != 0 or
                    strides is not None or native is not None):
                raise ValueError('Invalid combination of arguments when base is specified')
            if isinstance(base, BFarray):
                nd = base.ndim
                return ndarray.__new__(cls, space=_space2string(base.space),
                                       buffer=int(base.data or 0),
                                       shape=list(base.shape)[:nd],
                                       dtype=DataType(int(base.dtype)),
                                       strides=list(base.strides)[:nd])
            if hasattr(base, '__cuda_array_interface__') and not isinstance(base, np.ndarray):
                cai = base.__cuda_array_interface__
                np_dtype = np.dtype(cai['typestr'])
                cstrides = cai.get('strides')
                return ndarray.__new__(cls, space='cuda', buffer=int(cai['data'][0]),
                                       shape=list(cai['shape']), dtype=np_dtype,
                                       strides=list(cstrides) if cstrides else None)
            if dtype is not None:
                dtype = DataType(dtype)
            if space is None and dtype is None:
                if not isinstance(base, np.ndarray):
                    base = np.asarray(base)
                obj = base.view(cls)
                if conjugated is not None:
                    obj.bf.conjugated = conjugated
                return obj
            # Copy `base` into a new array in `space` (converting dtype on host)
            if not isinstance(base, np.ndarray):
                base = np.array(base, dtype=dtype.as_numpy_dtype() if dtype else None)
            if not isinstance(base, ndarray):
                if dtype is not None and DataType(base.dtype) != dtype:
                    base = base.astype(dtype.as_numpy_dtype())
                base = base.view(ndarray)
            if dtype is not None and base.bf.dtype != dtype:
                raise TypeError(f"Unable to convert type {base.bf.dtype} to {dtype} "
                                "during array construction")
            if space is None:
                space = base.bf.space
            if conjugated is None:
                conjugated = base.bf.conjugated
            logical_shape = list(base.shape)
            if base.bf.dtype.itemsize_bits < 8 and logical_shape:
                logical_shape[-1] *= 8 // base.bf.dtype.itemsize_bits
            obj = ndarray.__new__(cls, space=space, shape=logical_shape,
                                  dtype=base.bf.dtype, native=base.bf.native,
                                  conjugated=conjugated)
            copy_array(obj, base)
            return obj
        # ---- new array (optionally over an existing buffer address)
        dtype = DataType('f32' if dtype is None else dtype)
        native = True if native is None else native
        conjugated = False if conjugated is None else conjugated
        if shape is None:
            raise ValueError('Either base or shape must be specified')
        np_shape, itemsize = _packed_shape(shape, dtype)
        if strides is None:
            strides = []
            acc = itemsize
            for n in reversed(np_shape):
                strides.insert(0, acc)
                acc *= n
        strides = tuple(int(s) for s in strides)
        if len(np_shape):
            extent = 1 + sum((n - 1) * abs(s) for n, s in zip(np_shape, strides) if n > 0)
            nbyte = 0 if 0 in np_shape else extent - 1 + itemsize
        else:
            nbyte = itemsize
        if buffer is None:
            space = 'system' if space is None else str(Space(space))
            alloc = _Allocation(nbyte + offset, space)
        else:
            if space is None:
                space = _space2string(raw_get_space(buffer))
            space = str(Space(space))
            alloc = _Allocation(nbyte + offset, space, ptr=buffer)
        np_dtype = np.dtype(dtype.as_numpy_dtype())
        if not native:
            np_dtype = np_dtype.newbyteorder()
        raw = np.asarray(alloc)   # uint8 view; keeps `alloc` alive via .base
        obj = np.ndarray.__new__(cls, tuple(np_shape), np_dtype, raw, offset, strides)
        obj.bf = BFArrayInfo(space, dtype, native, conjugated)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        if isinstance(obj, ndarray) and hasattr(obj, 'bf'):
            self.bf = BFArrayInfo(obj.bf.space, obj.bf.dtype, obj.bf.native, obj.bf.conjugated)
            if self.dtype != obj.dtype:
                # numpy-level view()/field access changed the element type
                try:
                    self.bf.dtype = DataType(self.dtype)
                except TypeError:
                    pass
        else:
            try:
                dt = DataType(self.dtype)
            except TypeError:
                dt = DataType('u8')
            self.bf = BFArrayInfo('system', dt, self.dtype.isnative, False)

    # ------------------------------------------------------------ C ABI ----
    def as_BFarray(self):
        a = BFarray()
        a.data = self.ctypes.data
        a.space = Space(self.bf.space).as_BFspace()
        a.dtype = self.bf.dtype.as_BFdtype()
        a.immutable = not self.flags['WRITEABLE']
        nd = len(self.shape)
        a.ndim = nd
        if nd == 0:   # the backend has no 0-d arrays (ref ndarray.py:322-326)
            a.ndim = 1
            a.shape[0] = 1
            a.strides[0] = self.bf.dtype.itemsize
        for d in range(nd):
            a.shape[d] = self.shape[d]
            a.strides[d] = self.strides[d]
        bits = self.bf.dtype.itemsize_bits
        if bits < 8 and nd:
            a.shape[nd - 1] *= 8 // bits
        a.big_endian = not self.bf.native
        a.conjugated = self.bf.conjugated
        return a

    # ------------------------------------------------------------- views ----
    def conj(self):
        v = self.view(ndarray)
        v.bf.conjugated = not self.bf.conjugated
        return v

    def byteswap(self, inplace=False):
        """Swap the byte order of the data and flip the endianness flag."""
        if inplace:
            self.bf.native = not self.bf.native
            return np.ndarray.byteswap(self, True)
        out = ndarray(self, space=self.bf.space if not space_accessible(self.bf.space, ['system']) else 'system')
        out.bf.native = not self.bf.native
        if space_accessible(out.bf.space, ['system']):
            np.ndarray.byteswap(out, True)
        else:
            raise NotImplementedError("byteswap of device arrays: byteswap on the host first")
        return out

    def view(self, dtype=None, type_=None):
        if type_ is not None and dtype is None:
            dtype = type_
        if dtype is None:
            return super(ndarray, self).view()
        if isinstance(dtype, type) and issubclass(dtype, np.ndarray):
            return super(ndarray, self).view(dtype)
        dtype_bf = DataType(dtype)
        v = super(ndarray, self).view(np.dtype(dtype_bf.as_numpy_dtype()))
        v.bf.dtype = dtype_bf
        return v

    def _host(self):
        return self if space_accessible(self.bf.space, ['system']) else self.copy(space='system')

    def __repr__(self):
        return np.asarray(self._host()).__repr__()

    def __str__(self):
        return np.asarray(self._host()).__str__()

    def tofile(self, fid, sep="", format="%s"):
        return np.asarray(self._host()).tofile(fid, sep, format)

    def astype(self, dtype):
        dtype_bf = DataType(dtype)
        host = self._host()
        if self.bf.space == 'cuda_managed':
            device.stream_synchronize()
        src = np.asarray(host)
        if self.bf.dtype.is_complex and self.bf.dtype.is_integer:
            src = src['re'].astype(np.float64) + 1j * src['im'].astype(np.float64)
        if dtype_bf.is_complex and dtype_bf.is_integer:
            out = np.empty(src.shape, dtype=dtype_bf.as_numpy_dtype())
            out['re'] = np.real(src).astype(dtype_bf.as_real().as_numpy_dtype())
            out['im'] = np.imag(src).astype(dtype_bf.as_real().as_numpy_dtype())
        else:
            if not dtype_bf.is_complex and np.iscomplexobj(src):
                src = src.real
            out = src.astype(dtype_bf.as_numpy_dtype())
        out = out.view(ndarray)
        out.bf.dtype = dtype_bf
        if not space_accessible(self.bf.space, ['system']):
            out = ndarray(out, space=self.bf.space)
        return out

    def copy(self, space=None, order='C'):
        if order != 'C':
            raise NotImplementedError('Only order="C" is supported')
        if space is None:
            space = self.bf.space
        return ndarray(self, space=space)

    def _key_is_scalar(self, key):
        if isinstance(key, tuple):
            return (len(key) == self.ndim and
                    all(isinstance(k, (int, np.integer)) for k in key))
        return self.ndim == 1 and isinstance(key, (int, np.integer))

    def __getitem__(self, key):
        if self._key_is_scalar(key) and not space_accessible(self.bf.space, ['system']):
            return np.asarray(self._host())[key]
        return super(ndarray, self).__getitem__(key)

    def __setitem__(self, key, val):
        if space_accessible(self.bf.space, ['system']):
            if self.bf.space == 'cuda_managed':
                device.stream_synchronize()
            super(ndarray, self).__setitem__(key, val)
            return
        if self._key_is_scalar(key):
            key = tuple(slice(k, k + 1) for k in (key if isinstance(key, tuple) else (key,)))
        dst = super(ndarray, self).__getitem__(key)
        src = np.broadcast_to(np.asarray(val, dtype=dst.dtype), dst.shape) \
            if not isinstance(val, ndarray) else val
        copy_array(dst, src)

    def __reduce__(self):
        raise TypeError("bf.ndarray cannot be pickled")


# ------------------------------------------------------------------ factory ----
def asarray(arr, space=None):
    if isinstance(arr, ndarray) and (space is None or str(space) == arr.bf.space):
        return arr
    return ndarray(arr, space=space)


def empty(shape, dtype='f32', space=None, **kwargs):
    return ndarray(shape=shape, dtype=dtype, space=space, **kwargs)


def zeros(shape, dtype='f32', space=None, **kwargs):
    ret = empty(shape, dtype, space, **kwargs)
    memset_array(ret, 0)
    return ret


def empty_like(arr, space=None):
    arr = asarray(arr)
    shape = list(arr.shape)
    if arr.bf.dtype.itemsize_bits < 8 and shape:
        shape[-1] *= 8 // arr.bf.dtype.itemsize_bits
    return ndarray(shape=shape, dtype=arr.bf.dtype,
                   space=arr.bf.space if space is None else space,
                   native=arr.bf.native, conjugated=arr.bf.conjugated)


def zeros_like(arr, space=None):
    ret = empty_like(arr, space)
    memset_array(ret, 0)
    return ret


def copy_array(dst, src):
    """Copy `src` into `dst` across spaces (ref: ndarray.py:96-112).  Host<->host
    goes through numpy; anything touching the device goes through bfArrayCopy
    and is synchronised when the spaces differ."""
    dst_bf = asarray(dst)
    src_bf = asarray(src)
    if (space_accessible(dst_bf.bf.space, ['system']) and
            space_accessible(src_bf.bf.space, ['system'])):
        if 'cuda_managed' in (src_bf.bf.space, dst_bf.bf.space):
            device.stream_synchronize()
        np.copyto(np.asarray(dst_bf), np.asarray(src_bf))
    else:
        src_host = space_accessible(src_bf.bf.space, ['system'])
        dst_host = space_accessible(dst_bf.bf.space, ['system'])
        if dst_host and not src_host and not (src_bf.flags['C_CONTIGUOUS'] and
                                              dst_bf.flags['C_CONTIGUOUS']):
            # strided device -> host: pack on the device first, then one memcpy
            packed = empty_like(src_bf)
            _check(_bf.bfArrayCopy(packed.as_BFarray(), src_bf.as_BFarray()))
            tmp = empty_like(src_bf, space='system')
            _check(_bf.bfArrayCopy(tmp.as_BFarray(), packed.as_BFarray()))
            device.stream_synchronize()
            np.copyto(np.asarray(dst_bf), np.asarray(tmp))
            return dst
        if src_host and not src_bf.flags['C_CONTIGUOUS'] \
                and dst_bf.flags['C_CONTIGUOUS']:
            src_bf = np.ascontiguousarray(np.asarray(src_bf)).view(ndarray)
            src_bf.bf = BFArrayInfo('system', dst_bf.bf.dtype, dst_bf.bf.native,
                                    dst_bf.bf.conjugated)
        _check(_bf.bfArrayCopy(dst_bf.as_BFarray(), src_bf.as_BFarray()))
        if dst_bf.bf.space != src_bf.bf.space:
            device.stream_synchronize()
    return dst


def memset_array(dst, value):
    dst_bf = asarray(dst)
    _check(_bf.bfArrayMemset(dst_bf.as_BFarray(), value))
    return dst
