"""Unpack oracle (numpy).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates the reference's bit manipulation literally:
  CPU  src/unpack.cpp:48-197   (the path its known-answer tests pin)
  GPU  src/gunpack.cu:41-183   (identical except signed 1-bit, :168-183)
using the same shift/mask sequences on numpy integers."""
import numpy as np


def _spread(ival, nbit, byte_reverse):
    """Place the 8/nbit fields of each byte in the top bits of 8/nbit bytes.
    Returns uint8 array [..., 8/nbit] in output memory order."""
    ival = ival.astype(np.uint64)
    if nbit == 4:
        if byte_reverse:
            o = ival
            o = (o | (o << np.uint64(12))) & np.uint64(0xF0F0)
        else:
            o = ival << np.uint64(4)
            o = (o | (o << np.uint64(4))) & np.uint64(0xF0F0)
        nb = 2
    elif nbit == 2:
        o = ival << np.uint64(6)
        o = (o | (o << np.uint64(12))) & np.uint64(0x03C003C0)
        o = (o | (o << np.uint64(6))) & np.uint64(0xC0C0C0C0)
        nb = 4
    else:
        o = ival << np.uint64(7)
        o = (o | (o << np.uint64(28))) & np.uint64(0x0000078000000780)
        o = (o | (o << np.uint64(14))) & np.uint64(0x0180018001800180)
        o = (o | (o << np.uint64(7))) & np.uint64(0x8080808080808080)
        nb = 8
    out = np.stack([((o >> np.uint64(8 * j)) & np.uint64(0xFF)).astype(np.uint8)
                    for j in range(nb)], axis=-1)        # little-endian memory order
    if byte_reverse and nbit != 4:
        out = out[..., ::-1]                             # byteswap()
    return out


def unpack(packed_bytes, nbit, signed, byte_reverse=False, align_msb=False,
           conjugate=False, gpu=False):
    """packed_bytes: uint8 array.  Returns int8/uint8 array with last dim
    multiplied by 8/nbit.  gpu=True selects the GPU's signed 1-bit mapping."""
    b = np.asarray(packed_bytes, dtype=np.uint8)
    if signed and nbit == 1 and gpu:
        sp = _spread(~b, 1, False) | np.uint8(0x40)      # gunpack.cu:172-177
        if byte_reverse:
            sp = sp[..., ::-1]
        vals = sp.view(np.int8)
        if not align_msb:
            vals = vals >> 6                              # rshift_subwords<6>
    else:
        sp = _spread(b, nbit, byte_reverse)
        if signed:
            vals = sp.view(np.int8)
            if not align_msb:
                vals = vals >> (8 - nbit)                 # arithmetic
        else:
            vals = sp
            if not align_msb:
                vals = vals >> (8 - nbit)
    vals = vals.copy()
    if conjugate and signed:
        vals[..., 1::2] = (-vals[..., 1::2].astype(np.int16)).astype(np.int8)
    return vals.reshape(b.shape[:-1] + (b.shape[-1] * (8 // nbit),)) if b.ndim else vals
