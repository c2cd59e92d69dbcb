"""correlate block (mirrors python/bifrost/blocks/correlate.py:37-138 ->
bfLinAlgMatMul(a=NULL, b=x))."""
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.linalg import LinAlg


class CorrelateBlock(TransformBlock):
    def __init__(self, iring, nframe_per_integration, *args, **kwargs):
        super(CorrelateBlock, self).__init__(iring, *args, **kwargs)
        self.nframe_per_integration = nframe_per_integration
        self.linalg = LinAlg()

    def define_valid_input_spaces(self):
        return ('cuda',)

    def define_output_nframes(self, input_nframe):
        return 1

    def on_sequence(self, iseq):
        self.nframe_integrated = 0
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        if itensor['labels'] != ['time', 'freq', 'station', 'pol']:
            raise ValueError("Expected axes ['time', 'freq', 'station', 'pol'], got %s" % itensor['labels'])
        ohdr = deepcopy(ihdr)
        ot = ohdr['_tensor']
        ot['dtype'] = 'cf32'
        for key in ('shape', 'labels', 'scales', 'units'):
            if key in ot:
                time_, freq, stand, pol = itensor[key]
                ot[key] = [time_, freq, stand, pol, deepcopy(stand), deepcopy(pol)]
        ot['labels'] = ['time', 'freq', 'station_i', 'pol_i', 'station_j', 'pol_j']
        if 'scales' in ot:
            s = ot['scales'][0]
            ot['scales'][0] = [s[0], s[1] * self.nframe_per_integration]
        ohdr['matrix_fill_mode'] = 'lower'
        ohdr['gulp_nframe'] = 1
        return ohdr

    def on_data(self, ispan, ospan):
        idata, odata = ispan.data, ospan.data
        ntime, nchan, nstand, npol = idata.shape
        # [time, freq, stand*pol] -> [freq, time, stand*pol] view; out [freq, n, n]
        x = idata.reshape(ntime, nchan, nstand * npol).transpose(1, 0, 2)
        c = odata.reshape(nchan, nstand * npol, nstand * npol)
        beta = 0. if self.nframe_integrated == 0 else 1.
        self.linalg.matmul(1., None, x, beta, c)
        self.nframe_integrated += ispan.nframe
        if self.nframe_integrated >= self.nframe_per_integration:
            self.nframe_integrated = 0
            return 1
        return 0


def correlate(iring, nframe_per_integration, *args, **kwargs):
    """Cross-multiply ['time','freq','station','pol'] voltages into lower-triangular visibilities."""
    return CorrelateBlock(iring, nframe_per_integration, *args, **kwargs)
