"""GUPPI RAW block headers (mirrors python/bifrost/guppi_raw.py:62-99).

A file is a sequence of blocks; each block is a header of 80-character cards
("KEY     = value", strings in single quotes, closed by an "END" card, padded to
a 512-byte boundary when DIRECTIO is non-zero) followed by BLOCSIZE bytes of
samples laid out [chan][time][pol][re,im] with NBITS bits per component.
NTIME = BLOCSIZE * 8 / (2 * NPOL * OBSNCHAN * NBITS) is derived when absent, and
NPOL = 4 (which counts the complex components) is read as 2, as the reference
does."""

RECORD_LEN = 80
DIRECTIO_ALIGN_NBYTE = 512


def _parse_value(text):
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    if not text or text[0] not in "'\"":
        raise ValueError("Invalid header value: %r" % text)
    return text[1:-1].rstrip()


class Header(dict):
    """A parsed block header; `nbyte` = bytes consumed from the stream (cards + padding)."""
    nbyte = 0


class EndOfFile(IOError):
    """Clean end of the file: no byte of another header follows."""


def read_header(f, offset=None):
    """Reads one block header from the binary file object `f` (positioned at its
    first card) and leaves `f` at the first data byte.  Works on pipes: padding
    is consumed with read(), never seek(); a caller reading a pipe passes the
    absolute byte `offset` of the header so that DIRECTIO padding is computed
    from the position in the stream.  Raises EndOfFile when nothing follows,
    IOError when the file stops inside a header."""
    hdr = Header()
    nread = 0
    while True:
        record = f.read(RECORD_LEN)
        if len(record) == 0 and nread == 0:
            raise EndOfFile("no further block")
        if len(record) < RECORD_LEN:
            raise IOError("EOF reached in middle of header")
        nread += RECORD_LEN
        record = record.decode()
        if record.startswith('END'):
            break
        key, val = record.split('=', 1)
        key = key.strip()
        if key in hdr:
            raise KeyError("Duplicate header key: %s" % key)
        hdr[key] = _parse_value(val.strip())
    if 'DIRECTIO' in hdr:
        # the reference pads whenever the key is present (guppi_raw.py:88-91)
        try:
            pos = f.tell()
        except (IOError, OSError):
            pos = (offset or 0) + nread
        npad = DIRECTIO_ALIGN_NBYTE - pos % DIRECTIO_ALIGN_NBYTE
        f.read(npad)
        nread += npad
    hdr.nbyte = nread
    if 'NPOL' in hdr:
        hdr['NPOL'] = 1 if hdr['NPOL'] == 1 else 2
    if 'NTIME' not in hdr:
        hdr['NTIME'] = hdr['BLOCSIZE'] * 8 // (hdr['OBSNCHAN'] * hdr['NPOL'] * 2 * hdr['NBITS'])
    return hdr


def write_header(hdr, f):
    """Writes `hdr` as a block header (the inverse of read_header; used to make
    synthetic files -- the reference has no writer)."""
    nbyte = 0
    for key, val in hdr.items():
        if isinstance(val, str):
            text = "'%-8s'" % val
        elif isinstance(val, bool):
            text = str(int(val))
        else:
            text = repr(val) if isinstance(val, float) else str(val)
        card = ('%-8s= %s' % (key[:8], text)).ljust(RECORD_LEN)[:RECORD_LEN]
        f.write(card.encode())
        nbyte += RECORD_LEN
    f.write('END'.ljust(RECORD_LEN).encode())
    nbyte += RECORD_LEN
    if 'DIRECTIO' in hdr:
        f.write(b' ' * (DIRECTIO_ALIGN_NBYTE - f.tell() % DIRECTIO_ALIGN_NBYTE))
