"""GUPPI RAW source block (mirrors python/bifrost/blocks/guppi_raw.py:40-131):
one frame per block, tensor [time, freq, fine_time, pol] of ci<NBITS> in system
space -- the input of the spectrometer chain (testbench/gpuspec_simple.py)."""
from bifrost_b200.pipeline import SourceBlock
from bifrost_b200 import guppi_raw


def _mjd2unix(mjd):
    return (mjd - 40587) * 86400


def _get_with_default(obj, key, default=None):
    return obj[key] if key in obj else default


class GuppiRawSourceBlock(SourceBlock):
    def __init__(self, sourcenames, gulp_nframe=1, *args, **kwargs):
        super(GuppiRawSourceBlock, self).__init__(sourcenames, gulp_nframe=gulp_nframe, *args, **kwargs)

    def create_reader(self, sourcename):
        return open(sourcename, 'rb')

    def on_sequence(self, reader, sourcename):
        ihdr = guppi_raw.read_header(reader)
        nbit = ihdr['NBITS']
        assert nbit in (4, 8, 16, 32, 64)
        nchan = ihdr['OBSNCHAN']
        bw_MHz = ihdr['OBSBW']
        cfreq_MHz = ihdr['OBSFREQ']
        df_MHz = bw_MHz / nchan
        f0_MHz = cfreq_MHz - 0.5 * (nchan - 1) * df_MHz
        dt_s = 1. / df_MHz / 1e6                 # negative when OBSBW is (high -> low channel order)
        byte_offset = ihdr['PKTIDX'] * ihdr['PKTSIZE']
        frame_nbyte = ihdr['BLOCSIZE'] / ihdr['NTIME']
        offset_secs = byte_offset / (frame_nbyte / dt_s)
        tstart_mjd = ihdr['STT_IMJD'] + (ihdr['STT_SMJD'] + offset_secs) / 86400.
        tstart_unix = _mjd2unix(tstart_mjd)
        self.blocsize = ihdr['BLOCSIZE']
        ohdr = {
            '_tensor': {
                'dtype': 'ci' + str(nbit),
                'shape': [-1, nchan, ihdr['NTIME'], ihdr['NPOL']],
                'labels': ['time', 'freq', 'fine_time', 'pol'],
                'scales': [(tstart_unix, abs(dt_s) * ihdr['NTIME']), (f0_MHz, df_MHz), (0, dt_s), None],
                'units': ['s', 'MHz', 's', None],
                'gulp_nframe': 1,
            },
            'az_start': _get_with_default(ihdr, 'AZ'),
            'za_start': _get_with_default(ihdr, 'ZA'),
            'raj': _get_with_default(ihdr, 'RA', 0.) * (24. / 360.),
            'dej': _get_with_default(ihdr, 'DEC'),
            'source_name': _get_with_default(ihdr, 'SRC_NAME'),
            'refdm': _get_with_default(ihdr, 'CHAN_DM'),
            'refdm_units': 'pc cm^-3',
            'telescope': _get_with_default(ihdr, 'TELESCOP'),
            'machine': _get_with_default(ihdr, 'BACKEND'),
            'rawdatafile': sourcename,
            'coord_frame': 'topocentric',
        }
        ohdr['time_tag'] = int(round(tstart_unix * 2 ** 32))
        ohdr['name'] = sourcename
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
                    guppi_raw.read_header(reader)
                except IOError:
                    break                                   # clean EOF between blocks
            self.already_read_header = False
            view = buf[i * self.blocsize:(i + 1) * self.blocsize]
            nbyte = reader.readinto(view)
            if nbyte == 0:
                break
            if nbyte < self.blocsize:
                raise IOError("Block data is truncated")
            nframe += 1
        return [nframe]


def read_guppi_raw(filenames, gulp_nframe=1, *args, **kwargs):
    """Read GUPPI RAW files.  Output: [time, freq, fine_time, pol], ci*, system space."""
    return GuppiRawSourceBlock(filenames, gulp_nframe, *args, **kwargs)
