"""detect block (mirrors python/bifrost/blocks/detect.py:38-160).  The
reference builds bfMap strings; this build calls the fixed kernels directly
(bfDetect) -- the strings are still accepted by bf.map (csrc/map.cu)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.map import detect as bf_detect


class DetectBlock(TransformBlock):
    def __init__(self, iring, mode, axis=None, *args, **kwargs):
        super(DetectBlock, self).__init__(iring, *args, **kwargs)
        self.specified_axis = axis
        self.mode = mode.lower()

    def define_valid_input_spaces(self):
        return ('cuda',)

    def on_sequence(self, iseq):
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        itype = DataType(itensor['dtype'])
        if not itype.is_complex:
            raise TypeError("Input data must be complex")
        self.axis = self.specified_axis
        if 'labels' not in itensor and self.axis is None:
            raise TypeError("Polarization (pol) index must be labelled, or axis must be set manually")
        elif self.axis is None and self.mode != 'scalar' and 'pol' in itensor['labels']:
            self.axis = itensor['labels'].index('pol')
        elif isinstance(self.axis, str):
            self.axis = itensor['labels'].index(self.axis)
        ohdr = deepcopy(ihdr)
        otensor = ohdr['_tensor']
        if self.axis is not None:
            self.npol = otensor['shape'][self.axis]
            if self.npol not in [1, 2]:
                raise ValueError("Axis must have length 1 or 2")
            if self.mode in ('stokes', 'coherence') and self.npol == 2:
                otensor['shape'][self.axis] = 4
            if self.mode == 'stokes_i' and self.npol == 2:
                otensor['shape'][self.axis] = 1
            if 'labels' in otensor:
                otensor['labels'][self.axis] = 'pol'
        else:
            self.npol = 1
        otype = itype if (self.mode == 'jones' and self.npol == 2) else itype.as_real()
        otensor['dtype'] = str(otype.as_floating_point())
        return ohdr

    def on_data(self, ispan, ospan):
        if self.npol == 1:
            bf_detect(ispan.data, ospan.data, 'scalar')
        else:
            bf_detect(ispan.data, ospan.data, self.mode, self.axis)


def detect(iring, mode, axis=None, *args, **kwargs):
    """Square-law detection: 'scalar', 'jones', 'stokes', 'stokes_i', 'coherence'."""
    return DetectBlock(iring, mode, axis, *args, **kwargs)
