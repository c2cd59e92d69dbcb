"""Bifrost dtype strings <-> BFdtype <-> numpy dtypes.

Same vocabulary as the reference (python/bifrost/DataType.py:28-39):
``i`` signed int, ``u`` unsigned int, ``f`` float, ``ci`` complex signed int,
``cf`` complex float, followed by the number of bits per real component,
e.g. ``ci8`` = 8+8-bit complex integer, ``cf32`` = numpy complex64.
Complex integers map to numpy structured dtypes with fields ``re``/``im``
(ref: DataType.py:55-60); ``ci4`` is one packed byte, high nibble = real.
"""

import numpy as np

from bifrost_b200.libbifrost import _bf

ci4 = np.dtype([('re_im', np.uint8)])
ci8 = np.dtype([('re', np.int8), ('im', np.int8)])
ci16 = np.dtype([('re', np.int16), ('im', np.int16)])
ci32 = np.dtype([('re', np.int32), ('im', np.int32)])
ci64 = np.dtype([('re', np.int64), ('im', np.int64)])
cf16 = np.dtype([('re', np.float16), ('im', np.float16)])

_KIND_BITS = {'i': _bf.BF_DTYPE_INT_TYPE, 'u': _bf.BF_DTYPE_UINT_TYPE,
              'f': _bf.BF_DTYPE_FLOAT_TYPE}
_BITS_KIND = {v: k for k, v in _KIND_BITS.items()}
_VALID_NBIT = {'i': (1, 2, 4, 8, 16, 32, 64), 'u': (1, 2, 4, 8, 16, 32, 64),
               'f': (16, 32, 64), 'ci': (1, 2, 4, 8, 16, 32, 64), 'cf': (16, 32, 64)}
_NUMPY = {
    'i': {8: np.int8, 16: np.int16, 32: np.int32, 64: np.int64},
    'u': {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64},
    'f': {16: np.float16, 32: np.float32, 64: np.float64},
    # sub-byte kinds are stored packed in bytes
    'ci': {1: np.int8, 2: np.int8, 4: ci4, 8: ci8, 16: ci16, 32: ci32, 64: ci64},
    'cf': {16: cf16, 32: np.complex64, 64: np.complex128},
}
_STRUCT_COMPLEX = {ci4: ('ci', 4), ci8: ('ci', 8), ci16: ('ci', 16), ci32: ('ci', 32),
                   ci64: ('ci', 64), cf16: ('cf', 16)}


class DataType(object):
    def __init__(self, t=None):
        self._veclen = 1
        if isinstance(t, DataType):
            self._kind, self._nbit, self._veclen = t._kind, t._nbit, t._veclen
        elif isinstance(t, str):
            i = 0
            while i < len(t) and not t[i].isdigit():
                i += 1
            self._kind, self._nbit = t[:i], int(t[i:])
        elif isinstance(t, tuple):
            self._kind, self._nbit, self._veclen = t
        elif isinstance(t, (int, np.integer)) and not isinstance(t, bool):
            t = int(t)
            self._nbit = t & _bf.BF_DTYPE_NBIT_BITS
            self._kind = _BITS_KIND[t & _bf.BF_DTYPE_TYPE_BITS]
            if t & _bf.BF_DTYPE_COMPLEX_BIT:
                self._kind = 'c' + self._kind
            self._veclen = 1 + ((t & _bf.BF_DTYPE_VECTOR_BITS) >> _bf.BF_DTYPE_VECTOR_BIT0)
        else:
            t = np.dtype(t)   # TypeError if invalid
            if t in _STRUCT_COMPLEX:
                self._kind, self._nbit = _STRUCT_COMPLEX[t]
            elif t.kind == 'c':
                self._kind, self._nbit = 'cf', t.itemsize * 4
            elif t.kind in 'iuf':
                self._kind, self._nbit = t.kind, t.itemsize * 8
            elif t.kind == 'b':
                self._kind, self._nbit = 'u', 8
            else:
                raise TypeError(f"Unsupported data type: {t}")
        if self._kind not in _VALID_NBIT or self._nbit not in _VALID_NBIT[self._kind]:
            raise TypeError(f"Unsupported data type: {self._kind}{self._nbit}")

    def __eq__(self, other):
        other = DataType(other)
        return (self._kind, self._nbit, self._veclen) == (other._kind, other._nbit, other._veclen)

    def __ne__(self, other):
        return not (self == other)

    def __hash__(self):
        return hash((self._kind, self._nbit, self._veclen))

    def __str__(self):
        s = f"{self._kind}{self._nbit}"
        return s if self._veclen == 1 else f"{s}[{self._veclen}]"

    __repr__ = __str__

    def as_BFdtype(self):
        val = self._nbit | _KIND_BITS[self._kind[-1]]
        if self.is_complex:
            val |= _bf.BF_DTYPE_COMPLEX_BIT
        return val | ((self._veclen - 1) << _bf.BF_DTYPE_VECTOR_BIT0)

    def as_numpy_dtype(self):
        kind, nbit = self._kind, self._nbit
        if kind in ('i', 'u') and nbit < 8:
            base = np.dtype(np.int8 if kind == 'i' else np.uint8)   # packed
        else:
            base = np.dtype(_NUMPY[kind][nbit])
        if self._veclen == 1:
            return base
        return np.dtype(','.join((str(base),) * self._veclen))

    @property
    def is_complex(self):
        return self._kind[0] == 'c'

    @property
    def is_real(self):
        return not self.is_complex

    @property
    def is_signed(self):
        return 'i' in self._kind or 'f' in self._kind

    @property
    def is_floating_point(self):
        return 'f' in self._kind

    @property
    def is_integer(self):
        return 'i' in self._kind or 'u' in self._kind

    @property
    def nbit(self):
        return self._nbit

    @property
    def kind(self):
        return self._kind

    def as_floating_point(self):
        if self.is_floating_point:
            return self
        return DataType(('cf' if self.is_complex else 'f', 32 if self._nbit <= 24 else 64,
                         self._veclen))

    def as_integer(self, nbit=None):
        kind = self._kind.replace('f', 'i') if self.is_floating_point else self._kind
        return DataType((kind, self._nbit if nbit is None else nbit, self._veclen))

    def as_real(self):
        return DataType((self._kind[1:], self._nbit, self._veclen)) if self.is_complex else self

    def as_complex(self):
        return self if self.is_complex else DataType(('c' + self._kind, self._nbit, self._veclen))

    def as_nbit(self, nbit):
        return DataType((self._kind, nbit, self._veclen))

    def as_vector(self, veclen):
        return DataType((self._kind, self._nbit, veclen))

    @property
    def itemsize_bits(self):
        return self._nbit * (2 if self.is_complex else 1) * self._veclen

    @property
    def itemsize(self):
        bits = self.itemsize_bits
        if bits < 8:
            raise ValueError('itemsize is undefined when nbit < 8')
        return bits // 8
