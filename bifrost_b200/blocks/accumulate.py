"""accumulate block (mirrors python/bifrost/blocks/accumulate.py:38-96)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.map import accumulate as bf_accumulate


class AccumulateBlock(TransformBlock):
    def __init__(self, iring, nframe, dtype=None, gulp_nframe=1, *args, **kwargs):
        assert gulp_nframe == 1
        super(AccumulateBlock, self).__init__(iring, gulp_nframe=1, *args, **kwargs)
        self.nframe = nframe
        self.dtype = dtype

    def define_valid_input_spaces(self):
        return ('cuda',)

    def on_sequence(self, iseq):
        ohdr = deepcopy(iseq.header)
        otensor = ohdr['_tensor']
        if 'scales' in otensor:
            fax = otensor['shape'].index(-1)
            s = otensor['scales'][fax]
            otensor['scales'][fax] = [s[0], s[1] * self.nframe]
        if self.dtype is not None:
            otensor['dtype'] = self.dtype
        ohdr['gulp_nframe'] = 1
        self.frame_count = 0
        return ohdr

    def on_data(self, ispan, ospan):
        beta = 0. if self.frame_count == 0 else 1.
        bf_accumulate(ispan.data, ospan.data, beta)
        self.frame_count += 1
        if self.frame_count == self.nframe:
            self.frame_count = 0
            return 1
        return 0


def accumulate(iring, nframe, dtype=None, *args, **kwargs):
    """Sum `nframe` frames, one at a time, before emitting one frame."""
    return AccumulateBlock(iring, nframe, dtype, *args, **kwargs)
