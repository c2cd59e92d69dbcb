"""GUPPI RAW source block (mirrors python/bifrost/blocks/guppi_raw.py:40-131):
one frame per block, tensor [time, freq, fine_time, pol] of ci<NBITS> in system
space -- the input of the spectrometer chain (testbench/gpuspec_simple.py)."""
from bifrost_b200.pipeline import SourceBlock
from bifrost_b200 import guppi_raw


def _mjd2unix(mjd):
    return (mjd - 40587) * 86400


class GuppiRawSourceBlock(SourceBlock):
    def __init__(self, sourcenames, gulp_nframe=1, *args, **kwargs):
        super(GuppiRawSourceBlock, self).__init__(sourcenames, gulp_nframe=gulp_nframe, *args, **kwargs)

    def create_reader(self, sourcename):
        return open(sourcename, 'rb')

    # GUPPI card -> (output header key, conversion); cards that are absent give None
    _PASSTHROUGH = (('AZ', 'az_start', None), ('ZA', 'za_start', None),
                    ('RA', 'raj', lambda deg: deg * (24. / 360.)),        # degrees -> hours
                    ('DEC', 'dej', None), ('SRC_NAME', 'source_name', None),
                    ('CHAN_DM', 'refdm', None), ('TELESCOP', 'telescope', None),
                    ('BACKEND', 'machine', None))

    @staticmethod
    def _time_axis(card):
        """Unix time of the block's first sample and the sample interval (s)."""
        chan_bw_MHz = card['OBSBW'] / card['OBSNCHAN']
        tsamp = 1. / chan_bw_MHz / 1e6                   # negative for high -> low channel order
        bytes_per_sample = card['BLOCSIZE'] / card['NTIME']
        elapsed = card['PKTIDX'] * card['PKTSIZE'] / (bytes_per_sample / tsamp)
        mjd = card['STT_IMJD'] + (card['STT_SMJD'] + elapsed) / 86400.
        return _mjd2unix(mjd), tsamp

    def on_sequence(self, reader, sourcename):
        card = guppi_raw.read_header(reader, 0)
        self.stream_pos = card.nbyte
        nbit, nchan, ntime = card['NBITS'], card['OBSNCHAN'], card['NTIME']
        if nbit not in (4, 8, 16, 32, 64):
            raise ValueError("Unsupported NBITS: %r" % nbit)
        chan_bw_MHz = card['OBSBW'] / nchan
        first_chan_MHz = card['OBSFREQ'] - 0.5 * (nchan - 1) * chan_bw_MHz
        t0, tsamp = self._time_axis(card)
        self.blocsize = card['BLOCSIZE']
        # one block = one frame of the tensor below: an explicit NTIME card that
        # disagrees with BLOCSIZE would silently misalign every frame
        if self.blocsize * 8 != nchan * ntime * card['NPOL'] * 2 * nbit:
            raise ValueError("BLOCSIZE (%d) does not match OBSNCHAN x NTIME x NPOL x 2 x NBITS / 8 (%d)"
                             % (self.blocsize, nchan * ntime * card['NPOL'] * 2 * nbit // 8))
        tensor = dict(dtype='ci%d' % nbit,
                      shape=[-1, nchan, ntime, card['NPOL']],
                      labels=['time', 'freq', 'fine_time', 'pol'],       # 'time' counts blocks
                      scales=[(t0, abs(tsamp) * ntime), (first_chan_MHz, chan_bw_MHz), (0, tsamp), None],
                      units=['s', 'MHz', 's', None],
                      gulp_nframe=1)
        ohdr = {'_tensor': tensor, 'refdm_units': 'pc cm^-3', 'rawdatafile': sourcename,
                'coord_frame': 'topocentric', 'name': sourcename,
                'time_tag': int(round(t0 * 2 ** 32))}                   # 32.32 fixed point
        for key, okey, conv in self._PASSTHROUGH:
            val = card.get(key)
            if key == 'RA' and val is None:
                val = 0.
            ohdr[okey] = conv(val) if (conv is not None and val is not None) else val
        self.already_read_header = True
        return [ohdr]

    def on_data(self, reader, ospans):
        """One block per frame: every block after the first starts with its own
        header, which is parsed and skipped (any gulp_nframe works; the reference
        assumes gulp_nframe = 1)."""
        ospan = ospans[0]
        import numpy as np
        assert ospan.data.flags['C_CONTIGUOUS'], "GUPPI frames are read into a contiguous span"
        buf = np.asarray(ospan.data).reshape(-1).view(np.uint8)
        nframe = 0
        for i in range(ospan.nframe):
            if not self.already_read_header:
                try:
                    card = guppi_raw.read_header(reader, self.stream_pos)
                except guppi_raw.EndOfFile:
                    break                                   # clean EOF between blocks
                # (a file that stops inside a header raises: it is corrupt, not finished)
                self.stream_pos += card.nbyte
            self.already_read_header = False
            view = buf[i * self.blocsize:(i + 1) * self.blocsize]
            nbyte = reader.readinto(view)
            if nbyte == 0:
                break
            if nbyte < self.blocsize:
                raise IOError("Block data is truncated")
            self.stream_pos += nbyte
            nframe += 1
        return [nframe]


def read_guppi_raw(filenames, gulp_nframe=1, *args, **kwargs):
    """Read GUPPI RAW files.  Output: [time, freq, fine_time, pol], ci*, system space."""
    return GuppiRawSourceBlock(filenames, gulp_nframe, *args, **kwargs)
