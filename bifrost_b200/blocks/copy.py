"""copy block (mirrors python/bifrost/blocks/copy.py:37-77 -> bfArrayCopy)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.ndarray import copy_array


class CopyBlock(TransformBlock):
    def __init__(self, iring, space=None, *args, **kwargs):
        super(CopyBlock, self).__init__(iring, *args, **kwargs)
        if space is not None:
            self.orings[0].space = space

    def on_sequence(self, iseq):
        return deepcopy(iseq.header)

    def on_data(self, ispan, ospan):
        copy_array(ospan.data, ispan.data)


def copy(iring, space=None, *args, **kwargs):
    """Copy data, possibly to another space ('system', 'cuda', 'cuda_host')."""
    return CopyBlock(iring, space, *args, **kwargs)
