"""fdmt block (mirrors python/bifrost/blocks/fdmt.py:38-168 -> bfFdmt*)."""
import math
from copy import deepcopy

from bifrost_b200.pipeline import TransformBlock
from bifrost_b200.fdmt import Fdmt
from bifrost_b200.units import convert_units


class FdmtBlock(TransformBlock):
    def __init__(self, iring, max_dm=None, max_delay=None, max_diagonal=None,
                 exponent=-2.0, negative_delays=False, *args, **kwargs):
        super(FdmtBlock, self).__init__(iring, *args, **kwargs)
        if sum(m is not None for m in [max_dm, max_delay, max_diagonal]) != 1:
            raise ValueError("Must specify exactly one of: max_dm, max_delay, max_diagonal")
        self.space = self.orings[0].space
        self.max_value = max_dm or max_delay or max_diagonal or 0.
        self.max_mode = ('dm' if max_dm is not None else
                         'delay' if max_delay is not None else 'diagonal')
        self.kdm = 4.148741601e3        # MHz**2 cm**3 s / pc
        self.dm_units = 'pc cm^-3'
        self.exponent = exponent
        self.negative_delays = negative_delays
        self.fdmt = Fdmt()

    def define_valid_input_spaces(self):
        return ('cuda',)

    def _dispersion_span(self, nchan, first_MHz, chan_MHz, tsamp_s):
        """(max_delay in samples, max_dm) from whichever limit the user gave
        (blocks/fdmt.py:70-93 of the reference: the DM that sweeps max_delay
        samples between the band edges under the nu^-2 law)."""
        sweep = first_MHz ** -2 - (first_MHz + nchan * chan_MHz) ** -2
        limit, mode = self.max_value, self.max_mode
        if mode == 'diagonal':                       # a slope of `limit` samples per channel
            self.max_mode, mode = 'delay', 'delay'
            self.max_value = limit = int(math.ceil(nchan * limit))
        if mode == 'dm':
            ndelay = int(math.ceil(abs(self.kdm / tsamp_s * limit * sweep)))
            dm = limit
        else:
            ndelay = int(limit)
            dm = ndelay * tsamp_s / (self.kdm * abs(sweep))
        return ndelay, (-dm if self.negative_delays else dm)

    def on_sequence(self, iseq):
        hdr_in = iseq.header
        t_in = hdr_in['_tensor']
        if list(t_in['labels'][-2:]) != ['freq', 'time']:
            raise KeyError("Expected axes [..., 'freq', 'time'], got %s" % t_in['labels'])
        nchan = t_in['shape'][-2]
        funit, tunit = t_in['units'][-2], t_in['units'][-1]
        first_raw, chan_raw = t_in['scales'][-2]
        first_MHz = convert_units(first_raw, funit, 'MHz')
        chan_MHz = convert_units(chan_raw, funit, 'MHz')
        tsamp_s = convert_units(t_in['scales'][-1][1], tunit, 's')
        self.max_delay, max_dm = self._dispersion_span(nchan, first_MHz, chan_MHz, tsamp_s)
        self.dm_step = max_dm / self.max_delay
        self.fdmt.init(nchan, self.max_delay, first_MHz, chan_MHz, self.exponent, 'cuda')
        dm0 = 0.
        if 'refdm' in hdr_in:
            dm0 = convert_units(hdr_in['refdm'], hdr_in['refdm_units'], self.dm_units)
        hdr_out = deepcopy(hdr_in)
        t_out = hdr_out['_tensor']
        t_out['dtype'] = 'f32'
        for field, value in (('shape', self.max_delay), ('labels', 'dispersion'),
                             ('scales', (dm0, self.dm_step)), ('units', self.dm_units)):
            t_out[field][-2] = value                 # the freq axis becomes the trial axis
        hdr_out.update(max_dm=max_dm, max_dm_units=self.dm_units,
                       cfreq=first_raw + 0.5 * (nchan - 1) * chan_raw, cfreq_units=funit,
                       bw=nchan * chan_raw, bw_units=funit)
        return hdr_out

    def define_input_overlap_nframe(self, iseq):
        return self.max_delay

    def on_data(self, ispan, ospan):
        if ispan.nframe <= self.max_delay:
            return 0
        self.fdmt.execute(ispan.data, ospan.data, negative_delays=self.negative_delays)


def fdmt(iring, max_dm=None, max_delay=None, max_diagonal=None, exponent=-2.0,
         negative_delays=False, *args, **kwargs):
    """Fast Dispersion Measure Transform of a [..., 'freq', 'time'] stream."""
    return FdmtBlock(iring, max_dm, max_delay, max_diagonal, exponent, negative_delays,
                     *args, **kwargs)
