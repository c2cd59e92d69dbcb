"""correlate block (same contract as python/bifrost/blocks/correlate.py:37-138;
the data path is bfLinAlgMatMul(a=NULL, b=x): the int8 tensor-core correlator)."""
from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.linalg import LinAlg
from bifrost_b200.blocks import _header as H

_IN_AXES = ['time', 'freq', 'station', 'pol']


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
        if ihdr['_tensor']['labels'] != _IN_AXES:
            raise ValueError("Expected axes %s, got %s" % (_IN_AXES, ihdr['_tensor']['labels']))
        ohdr, otensor = H.derive(ihdr)
        otensor['dtype'] = 'cf32'
        # [time, freq, station, pol] -> [time, freq, station_i, pol_i, station_j, pol_j]
        H.remap_axes(ihdr['_tensor'], otensor, [0, 1, 2, 3, 2, 3])
        otensor['labels'] = ['time', 'freq', 'station_i', 'pol_i', 'station_j', 'pol_j']
        H.scale_step(otensor, 0, self.nframe_per_integration)
        ohdr['matrix_fill_mode'] = 'lower'
        # an integration is made of whole input gulps: the gulp the block reads
        # (its own setting, else the upstream one, capped at one integration) has
        # to divide nframe_per_integration (blocks/correlate.py:67-74)
        ohdr['gulp_nframe'] = min(ihdr.get('gulp_nframe') or 1, self.nframe_per_integration)
        gulp = self.gulp_nframe or ohdr['gulp_nframe']
        if self.nframe_per_integration % gulp:
            raise ValueError("gulp_nframe (%i) does not divide nframe_per_integration (%i)"
                             % (gulp, self.nframe_per_integration))
        return ohdr

    def _gulp(self, iseq):
        # the input is read in gulps that tile an integration
        return self.gulp_nframe or min(iseq.header.get('gulp_nframe') or 1, self.nframe_per_integration)

    def on_data(self, ispan, ospan):
        idata, odata = ispan.data, ospan.data
        ntime, nchan, nstand, npol = idata.shape
        # [time, freq, stand*pol] -> [freq, time, stand*pol] view; out [freq, n, n]
        x = idata.reshape(ntime, nchan, nstand * npol).transpose(1, 0, 2)
        c = odata.reshape(nchan, nstand * npol, nstand * npol)
        self.linalg.matmul(1., None, x, 1. if self.nframe_integrated else 0., c)
        self.nframe_integrated += ispan.nframe
        if self.nframe_integrated > self.nframe_per_integration:
            raise RuntimeError("integration overran: gulp does not tile nframe_per_integration")
        if self.nframe_integrated == self.nframe_per_integration:
            self.nframe_integrated = 0
            return 1
        return 0


def correlate(iring, nframe_per_integration, *args, **kwargs):
    """Cross-multiply ['time','freq','station','pol'] voltages into lower-triangular visibilities."""
    return CorrelateBlock(iring, nframe_per_integration, *args, **kwargs)
