"""quantize block (mirrors python/bifrost/blocks/quantize.py:41-89): requantise
f32 / cf32 to an integer type (dtype or number of bits) after scaling.  Device
arrays only in this build."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.quantize import quantize as bf_quantize


class QuantizeBlock(TransformBlock):
    def __init__(self, iring, dtype, scale=1., *args, **kwargs):
        super(QuantizeBlock, self).__init__(iring, *args, **kwargs)
        self.dtype = dtype
        self.scale = scale

    def define_valid_input_spaces(self):
        return ('cuda',)

    def on_sequence(self, iseq):
        ihdr = iseq.header
        ohdr = deepcopy(ihdr)
        itype = DataType(ihdr['_tensor']['dtype'])
        self.itype = itype
        if isinstance(self.dtype, int):            # number of bits instead of a dtype
            otype = itype.as_integer(self.dtype)
        else:
            otype = DataType(self.dtype)
        ohdr['_tensor']['dtype'] = str(otype)
        return ohdr

    def on_data(self, ispan, ospan):
        bf_quantize(ispan.data, ospan.data, self.scale)


def quantize(iring, dtype, scale=1., *args, **kwargs):
    """Requantise [c]f32 data to a (complex) integer type of 8, 16 or 32 bits."""
    return QuantizeBlock(iring, dtype, scale, *args, **kwargs)
