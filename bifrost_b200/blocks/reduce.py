"""reduce block (same contract as python/bifrost/blocks/reduce.py:38-118; the
data path is bfReduce)."""
from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.reduce import reduce as bf_reduce
from bifrost_b200.blocks import _header as H

_COMPLEX_IN = ('cf32', 'ci8', 'ci16')


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
        ohdr, otensor = H.derive(ihdr)
        # power ops give real output; plain ops keep complex input complex
        keeps_complex = ihdr['_tensor']['dtype'] in _COMPLEX_IN and not self.op.startswith('pwr')
        otensor['dtype'] = 'cf32' if keeps_complex else 'f32'
        self.axis = H.axis_index(otensor, self.specified_axis)
        self.factor, self.frame_factor = self.specified_factor, 1
        if self.axis == H.frame_axis(otensor):
            # along time: fewer frames come out, the shape entry stays -1
            if self.factor is None:
                raise ValueError("Cannot reduce all of the frame axis")
            self.frame_factor = self.factor
            ohdr['gulp_nframe'] = max((ihdr.get('gulp_nframe') or self.factor) // self.factor, 1)
        else:
            length = otensor['shape'][self.axis]
            if self.factor is None:
                self.factor = length
            if length % self.factor:
                raise ValueError("Reduce factor does not divide axis length")
            otensor['shape'][self.axis] = length // self.factor
        H.scale_step(otensor, self.axis, self.factor)
        return ohdr

    def on_data(self, ispan, ospan):
        bf_reduce(ispan.data, ospan.data, self.op)


def reduce(iring, axis, factor=None, op='sum', *args, **kwargs):
    """Reduce `axis` by `factor` (None = all) with op sum/mean/min/max/stderr/pwr*."""
    return ReduceBlock(iring, axis, factor, op, *args, **kwargs)
