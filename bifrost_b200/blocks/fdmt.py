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

    def on_sequence(self, iseq):
        ihdr = iseq.header
        itensor = ihdr['_tensor']
        labels = itensor['labels']
        if labels[-1] != 'time' or labels[-2] != 'freq':
            raise KeyError("Expected axes [..., 'freq', 'time'], got %s" % labels)
        nchan = itensor['shape'][-2]
        f0_, df_ = itensor['scales'][-2]
        t0_, dt_ = itensor['scales'][-1]
        f0 = convert_units(f0_, itensor['units'][-2], 'MHz')
        df = convert_units(df_, itensor['units'][-2], 'MHz')
        dt = convert_units(dt_, itensor['units'][-1], 's')
        if self.max_mode == 'diagonal':
            self.max_mode = 'delay'
            self.max_value = int(math.ceil(nchan * self.max_value))
        fac = f0 ** -2 - (f0 + nchan * df) ** -2
        if self.max_mode == 'dm':
            max_dm = self.max_value
            self.max_delay = int(math.ceil(abs(self.kdm / dt * max_dm * fac)))
        else:
            self.max_delay = int(self.max_value)
            max_dm = self.max_delay * dt / (self.kdm * abs(fac))
        if self.negative_delays:
            max_dm = -max_dm
        self.dm_step = max_dm / self.max_delay
        self.fdmt.init(nchan, self.max_delay, f0, df, self.exponent, 'cuda')
        ohdr = deepcopy(ihdr)
        refdm = convert_units(ihdr['refdm'], ihdr['refdm_units'], self.dm_units) if 'refdm' in ihdr else 0.
        ot = ohdr['_tensor']
        ot['dtype'] = 'f32'
        ot['shape'][-2] = self.max_delay
        ot['labels'][-2] = 'dispersion'
        ot['scales'][-2] = (refdm, self.dm_step)
        ot['units'][-2] = self.dm_units
        ohdr['max_dm'] = max_dm
        ohdr['max_dm_units'] = self.dm_units
        ohdr['cfreq'] = f0_ + 0.5 * (nchan - 1) * df_
        ohdr['cfreq_units'] = itensor['units'][-2]
        ohdr['bw'] = nchan * df_
        ohdr['bw_units'] = itensor['units'][-2]
        return ohdr

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
