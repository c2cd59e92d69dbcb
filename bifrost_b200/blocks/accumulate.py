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
        self.frame_count = 0                     # frames summed into the open output frame
        hdr_out = deepcopy(iseq.header)
        hdr_out['gulp_nframe'] = 1
        t = hdr_out['_tensor']
        if self.dtype is not None:
            t['dtype'] = self.dtype
        if 'scales' in t:                        # one output frame spans nframe input frames
            frame_axis = t['shape'].index(-1)
            start, step = t['scales'][frame_axis]
            t['scales'][frame_axis] = [start, step * self.nframe]
        return hdr_out

    def on_data(self, ispan, ospan):
        # the first frame of an integration overwrites, the others add (b = beta*b + a)
        bf_accumulate(ispan.data, ospan.data, 0. if self.frame_count == 0 else 1.)
        self.frame_count = (self.frame_count + 1) % self.nframe
        return 1 if self.frame_count == 0 else 0


def accumulate(iring, nframe, dtype=None, *args, **kwargs):
    """Sum `nframe` frames, one at a time, before emitting one frame."""
    return AccumulateBlock(iring, nframe, dtype, *args, **kwargs)
