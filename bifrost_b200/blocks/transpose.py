"""transpose block (mirrors python/bifrost/blocks/transpose.py:38-92 -> bfTranspose)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.transpose import transpose as bf_transpose


class TransposeBlock(TransformBlock):
    def __init__(self, iring, axes, *args, **kwargs):
        super(TransposeBlock, self).__init__(iring, *args, **kwargs)
        self.specified_axes = axes

    def define_valid_input_spaces(self):
        return ('cuda',)          # the reference's numpy fallback for 'system' is not part of the hot path

    def on_sequence(self, iseq):
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        self.axes = [itensor['labels'].index(ax) if isinstance(ax, str) else ax
                     for ax in self.specified_axes]
        ohdr = deepcopy(ihdr)
        otensor = ohdr['_tensor']
        for item in ('shape', 'labels', 'scales', 'units'):
            if item in itensor:
                otensor[item] = [itensor[item][ax] for ax in self.axes]
        return ohdr

    def on_data(self, ispan, ospan):
        bf_transpose(ospan.data, ispan.data, self.axes)


def transpose(iring, axes, *args, **kwargs):
    """Permute the axes of the data stream (axes: indices or labels)."""
    return TransposeBlock(iring, axes, *args, **kwargs)
