"""Source / sink helpers for feeding synthetic data through block chains --
the pattern of the reference's own tests (test/test_pipeline.py:51-73)."""
import numpy as np

from bifrost_b200.pipeline import SourceBlock, SinkBlock
from bifrost_b200.ndarray import copy_array


class _Reader(object):
    def __init__(self, arr):
        self.arr, self.pos = arr, 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class NumpySourceBlock(SourceBlock):
    """Emits `array` (frame axis = axis 0 unless `frame_axis`) with `header`."""

    def __init__(self, array, header, gulp_nframe, frame_axis=0, space='system', *args, **kwargs):
        super(NumpySourceBlock, self).__init__(['array'], gulp_nframe, space=space, *args, **kwargs)
        self.array, self.header, self.frame_axis = array, header, frame_axis

    def create_reader(self, sourcename):
        return _Reader(self.array)

    def on_sequence(self, reader, sourcename):
        return [dict(self.header)]

    def on_data(self, reader, ospans):
        ospan = ospans[0]
        fax = self.frame_axis
        n = min(ospan.nframe, reader.arr.shape[fax] - reader.pos)
        if n > 0:
            sl = [slice(None)] * reader.arr.ndim
            sl[fax] = slice(reader.pos, reader.pos + n)
            osl = [slice(None)] * reader.arr.ndim
            osl[fax] = slice(0, n)
            copy_array(ospan.data[tuple(osl)], np.ascontiguousarray(reader.arr[tuple(sl)]))
            reader.pos += n
        return [n]


class CallbackSinkBlock(SinkBlock):
    def __init__(self, iring, seq_callback=None, data_callback=None, *args, **kwargs):
        super(CallbackSinkBlock, self).__init__(iring, *args, **kwargs)
        self.seq_callback, self.data_callback = seq_callback, data_callback

    def on_sequence(self, iseq):
        if self.seq_callback is not None:
            self.seq_callback(iseq)

    def on_data(self, ispan):
        if self.data_callback is not None:
            self.data_callback(ispan)


def array_source(array, header, gulp_nframe, **kwargs):
    return NumpySourceBlock(array, header, gulp_nframe, **kwargs)


def callback_sink(iring, seq_callback=None, data_callback=None, **kwargs):
    return CallbackSinkBlock(iring, seq_callback, data_callback, **kwargs)
