"""detect block (same contract as python/bifrost/blocks/detect.py:38-160).  The
reference builds bfMap strings; this build calls the fixed kernels directly
(bfDetect) -- the strings are still accepted by bf.map (csrc/map.cu)."""
from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.map import detect as bf_detect
from bifrost_b200.blocks import _header as H

# output length of a 2-long pol axis per mode (absent: unchanged)
_POL_OUT = {'stokes': 4, 'coherence': 4, 'stokes_i': 1}


class DetectBlock(TransformBlock):
    def __init__(self, iring, mode, axis=None, *args, **kwargs):
        super(DetectBlock, self).__init__(iring, *args, **kwargs)
        self.specified_axis = axis
        self.mode = mode.lower()

    def define_valid_input_spaces(self):
        return ('cuda',)

    def _pol_axis(self, tensor):
        """The axis the Jones / Stokes products run over, or None (scalar)."""
        axis = self.specified_axis
        if axis is None:
            if 'labels' not in tensor:
                raise TypeError("Polarization (pol) index must be labelled, or axis must be set manually")
            if self.mode != 'scalar' and 'pol' in tensor['labels']:
                return tensor['labels'].index('pol')
            return None
        return H.axis_index(tensor, axis)

    def on_sequence(self, iseq):
        ohdr, otensor = H.derive(iseq.header)
        itype = DataType(otensor['dtype'])
        if not itype.is_complex:
            raise TypeError("Input data must be complex")
        self.axis = self._pol_axis(otensor)
        self.npol = 1
        if self.axis is not None:
            self.npol = otensor['shape'][self.axis]
            if self.npol not in (1, 2):
                raise ValueError("Axis must have length 1 or 2")
            if self.npol == 2:
                otensor['shape'][self.axis] = _POL_OUT.get(self.mode, 2)
            if 'labels' in otensor:
                otensor['labels'][self.axis] = 'pol'
        # only a 2-pol Jones product stays complex
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
