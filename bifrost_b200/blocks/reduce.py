"""reduce block (mirrors python/bifrost/blocks/reduce.py:38-118 -> bfReduce)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.reduce import reduce as bf_reduce


class ReduceBlock(TransformBlock):
    def __init__(self, iring, axis, factor=None, op='sum', *args, **kwargs):
        super(ReduceBlock, self).__init__(iring, *args, **kwargs)
        self.specified_axis = axis
        self.specified_factor = factor
        self.op = op

    def define_valid_input_spaces(self):
        return ('cuda',)

    def define_output_nframes(self, input_nframe):
        return input_nframe // self.frame_factor

    def on_sequence(self, iseq):
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        ohdr = deepcopy(ihdr)
        otensor = ohdr['_tensor']
        otensor['dtype'] = 'f32'
        if itensor['dtype'] in ['cf32', 'ci8', 'ci16'] and not self.op.startswith('pwr'):
            otensor['dtype'] = 'cf32'
        self.axis = (itensor['labels'].index(self.specified_axis)
                     if isinstance(self.specified_axis, str) else self.specified_axis)
        frame_axis = itensor['shape'].index(-1)
        self.frame_factor = 1
        self.factor = self.specified_factor
        if self.axis == frame_axis:
            if self.factor is None:
                raise ValueError("Cannot reduce all of the frame axis")
            self.frame_factor = self.factor
            ohdr['gulp_nframe'] = max((ihdr.get('gulp_nframe') or self.factor) // self.factor, 1)
        else:
            if self.factor is None:
                self.factor = otensor['shape'][self.axis]
            if otensor['shape'][self.axis] % self.factor:
                raise ValueError("Reduce factor does not divide axis length")
            otensor['shape'][self.axis] //= self.factor
        if 'scales' in otensor:
            s = otensor['scales'][self.axis]
            otensor['scales'][self.axis] = [s[0], s[1] * self.factor]
        return ohdr

    def on_data(self, ispan, ospan):
        bf_reduce(ispan.data, ospan.data, self.op)


def reduce(iring, axis, factor=None, op='sum', *args, **kwargs):
    """Reduce `axis` by `factor` (None = all) with op sum/mean/min/max/stderr/pwr*."""
    return ReduceBlock(iring, axis, factor, op, *args, **kwargs)
