"""transpose block (same contract as python/bifrost/blocks/transpose.py:38-92;
the data path is bfTranspose)."""
from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.transpose import transpose as bf_transpose
from bifrost_b200.blocks import _header as H


class TransposeBlock(TransformBlock):
    def __init__(self, iring, axes, *args, **kwargs):
        super(TransposeBlock, self).__init__(iring, *args, **kwargs)
        self.specified_axes = axes

    def define_valid_input_spaces(self):
        return ('cuda',)          # the reference's numpy fallback for 'system' is not part of the hot path

    def on_sequence(self, iseq):
        ohdr, otensor = H.derive(iseq.header)
        itensor = iseq.header['_tensor']
        self.axes = [H.axis_index(itensor, ax) for ax in self.specified_axes]
        H.remap_axes(itensor, otensor, self.axes)
        return ohdr

    def on_data(self, ispan, ospan):
        bf_transpose(ospan.data, ispan.data, self.axes)


def transpose(iring, axes, *args, **kwargs):
    """Permute the axes of the data stream (axes: indices or labels)."""
    return TransposeBlock(iring, axes, *args, **kwargs)
