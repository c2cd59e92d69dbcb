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
        hdr_out = deepcopy(iseq.header)
        t = hdr_out['_tensor']
        labels = t.get('labels')
        self.axes = axes = [labels.index(a) if isinstance(a, str) else a for a in self.specified_axes]
        if t['shape'].index(-1) in axes:
            raise KeyError("Cannot transform frame axis; reshape the data stream first")
        t_in = DataType(t['dtype']).as_floating_point()
        t_res = t_in.as_real() if self.real_output else t_in.as_complex()
        self.mode = {(True, False): 'r2c', (False, True): 'c2r'}.get((t_in.is_real, t_res.is_real), 'c2c')
        t['dtype'] = str(t_res)
        lengths = [t['shape'][a] for a in axes]             # transform lengths (full size for c2r)
        last = axes[-1]
        if self.mode == 'r2c':
            t['shape'][last] = t['shape'][last] // 2 + 1    # Hermitian half
        elif self.mode == 'c2r':
            lengths[-1] = t['shape'][last] = (t['shape'][last] - 1) * 2
        for k, (a, n) in enumerate(zip(axes, lengths)):
            if 'units' in t:
                t['units'][a] = transform_units(t['units'][a], -1)
            if 'scales' in t:
                t['scales'][a] = [0, 1. / (t['scales'][a][1] * n)]   # bin width = 1 / (step * length)
            new_label = self.axis_labels[k] if k < len(self.axis_labels) else None
            if labels is not None and new_label is not None:
                t['labels'][a] = new_label
        return hdr_out

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
