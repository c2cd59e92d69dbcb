"""FFT oracle (numpy, fp64).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

numpy.fft in double precision is the reference's own gold standard
(test/test_fft.py:36-51); this module adds the reference's load-callback
semantics on top of it (src/fft_kernels.cu:96-197, src/fft.cu:294-296):
integer scaling, unnormalised inverse, and where the fftshift is applied."""
import numpy as np

SCALE = {'ci4': 1. / 128, 'ci8': 1. / 128, 'ci16': 1. / 32768, 'i8': 1. / 128,
         'i16': 1. / 32768, 'u8': 1. / 256, 'u16': 1. / 65536}


def to_complex128(x, bf_dtype):
    """Decode an array as the reference's load callbacks would (before scaling)."""
    if bf_dtype == 'ci4':
        b = x['re_im'].astype(np.int8) if x.dtype.names else x.astype(np.int8)
        re = (b & np.int8(-16)).astype(np.int8).astype(np.float64)        # value << 4
        im = ((b.astype(np.int16) << 4) & 0xFF).astype(np.uint8).view(np.int8).astype(np.float64)
        return re + 1j * im
    if x.dtype.names:
        return x['re'].astype(np.float64) + 1j * x['im'].astype(np.float64)
    return x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)


def fft(x, axes, bf_dtype=None, inverse=False, fftshift=False, real_out_len=None):
    """Reference semantics of bfFftExecute for an input already decoded to
    float64/complex128 (apply SCALE[bf_dtype] first for integer types).
    c2c: [i]fftn unnormalised; forward shift = fftshift(output), inverse shift =
    ifftshift(input).  real input -> rfftn.  real_out_len -> irfftn * N."""
    axes = list(axes)
    if real_out_len is not None:
        shape = [x.shape[a] for a in axes]
        shape[-1] = real_out_len
        norm = np.prod(shape)
        return np.fft.irfftn(x, s=shape, axes=axes) * norm
    if not np.iscomplexobj(x):
        return np.fft.rfftn(x, axes=axes)
    if inverse:
        if fftshift:
            x = np.fft.ifftshift(x, axes=axes)
        norm = np.prod([x.shape[a] for a in axes])
        return np.fft.ifftn(x, axes=axes) * norm
    y = np.fft.fftn(x, axes=axes)
    if fftshift:
        y = np.fft.fftshift(y, axes=axes)
    return y
