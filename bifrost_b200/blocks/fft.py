"""fft block (mirrors python/bifrost/blocks/fft.py:38-137 -> bfFftInit/bfFftExecute)."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.DataType import DataType
from bifrost_b200.fft import Fft
from bifrost_b200.units import transform_units


class FftBlock(TransformBlock):
    def __init__(self, iring, axes, inverse=False, real_output=False, axis_labels=None,
                 apply_fftshift=False, *args, **kwargs):
        super(FftBlock, self).__init__(iring, *args, **kwargs)
        if not isinstance(axes, (list, tuple)):
            axes = [axes]
        if not isinstance(axis_labels, (list, tuple)):
            axis_labels = [axis_labels]
        self.specified_axes = list(axes)
        self.real_output = real_output
        self.inverse = inverse
        self.axis_labels = list(axis_labels)
        self.apply_fftshift = apply_fftshift
        self.fft = Fft()
        self.plan_key = None

    def define_valid_input_spaces(self):
        return ('cuda',)

    def on_sequence(self, iseq):
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        itype = DataType(itensor['dtype']).as_floating_point()
        self.axes = [itensor['labels'].index(ax) if isinstance(ax, str) else ax
                     for ax in self.specified_axes]
        axes = self.axes
        shape = [itensor['shape'][ax] for ax in axes]
        otype = itype.as_real() if self.real_output else itype.as_complex()
        ohdr = deepcopy(ihdr)
        otensor = ohdr['_tensor']
        otensor['dtype'] = str(otype)
        if itype.is_real and otype.is_complex:
            self.mode = 'r2c'
        elif itype.is_complex and otype.is_real:
            self.mode = 'c2r'
        else:
            self.mode = 'c2c'
        if itensor['shape'].index(-1) in axes:
            raise KeyError("Cannot transform frame axis; reshape the data stream first")
        if self.mode == 'r2c':
            otensor['shape'][axes[-1]] = otensor['shape'][axes[-1]] // 2 + 1
        elif self.mode == 'c2r':
            otensor['shape'][axes[-1]] = (otensor['shape'][axes[-1]] - 1) * 2
            shape[-1] = (shape[-1] - 1) * 2
        for i, (ax, length) in enumerate(zip(axes, shape)):
            if 'units' in otensor:
                otensor['units'][ax] = transform_units(otensor['units'][ax], -1)
            if 'scales' in otensor:
                scale = otensor['scales'][ax][1]
                otensor['scales'][ax] = [0, 1. / (scale * length)]
            if 'labels' in otensor and i < len(self.axis_labels) and self.axis_labels[i] is not None:
                otensor['labels'][ax] = self.axis_labels[i]
        return ohdr

    def on_data(self, ispan, ospan):
        idata, odata = ispan.data, ospan.data
        key = (idata.shape, odata.shape, idata.strides, odata.strides)
        if key != self.plan_key:
            self.fft.init(idata, odata, axes=self.axes, apply_fftshift=self.apply_fftshift)
            self.plan_key = key
        self.fft.execute(idata, odata, inverse=self.inverse)


def fft(iring, axes, inverse=False, real_output=False, axis_labels=None,
        apply_fftshift=False, *args, **kwargs):
    """FFT over the given axes (indices or labels) of the data stream."""
    return FftBlock(iring, axes, inverse, real_output, axis_labels, apply_fftshift, *args, **kwargs)
