"""unpack block (mirrors python/bifrost/blocks/unpack.py:41-90): 1/2/4-bit
integer or complex-integer samples to 8 bits through bfUnpack (csrc/unpack.cu).
Device arrays only in this build (the reference also accepts system space)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.unpack import unpack as bf_unpack


class UnpackBlock(TransformBlock):
    def __init__(self, iring, dtype, align_msb=False, *args, **kwargs):
        super(UnpackBlock, self).__init__(iring, *args, **kwargs)
        self.dtype = dtype
        self.align_msb = align_msb

    def define_valid_input_spaces(self):
        return ('cuda',)

    def on_sequence(self, iseq):
        ihdr = iseq.header
        ohdr = deepcopy(ihdr)
        itype = DataType(ihdr['_tensor']['dtype'])
        self.itype = itype
        # the user may pass nbit instead of an explicit dtype
        if isinstance(self.dtype, int):
            otype = itype.as_nbit(self.dtype)
        else:
            otype = DataType(self.dtype)
        ohdr['_tensor']['dtype'] = str(otype)
        return ohdr

    def on_data(self, ispan, ospan):
        bf_unpack(ispan.data, ospan.data, self.align_msb)


def unpack(iring, dtype, *args, **kwargs):
    """Unpack i/u2, i/u4, ci2, ci4 (also 1-bit) data to i8 / ci8 (or f32 / cf32)."""
    return UnpackBlock(iring, dtype, *args, **kwargs)
