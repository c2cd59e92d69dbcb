"""Element-wise / layout oracles (numpy).  TEST INFRASTRUCTURE -- see
oracle/__init__.py.  fp32 arithmetic in the reference's operation order."""
import numpy as np

f32 = np.float32


def transpose(x, axes):
    """numpy.transpose is the reference's own CPU path (blocks/transpose.py:80)."""
    return np.ascontiguousarray(np.transpose(x, axes))


def _as_complex_pair(x):
    """Returns (re, im) float32 for complex64 or structured ci* input."""
    if x.dtype.names:
        return x['re'].astype(f32), x['im'].astype(f32)
    return x.real.astype(f32), x.imag.astype(f32)


def reduce(x, factor, axis, op='sum'):
    """src/reduce.cu:62-121,156-212,366-484,610-730: strictly left-to-right
    fp32 combination of `factor` consecutive elements along `axis`.
    x: real array, complex64, or structured ci8/ci16."""
    is_complex = bool(x.dtype.names) or np.iscomplexobj(x)
    n = x.shape[axis]
    if factor is None:
        factor = n
    assert n % factor == 0
    power = op.startswith('pwr')
    base = op[3:] if power else op
    x = np.moveaxis(x, axis, -1)
    lead = x.shape[:-1]
    x = x.reshape(lead + (n // factor, factor))

    def fold(vals_first, vals_rest):
        acc = vals_first
        for k in range(1, factor):
            v = vals_rest(k)
            if base in ('sum', 'mean', 'stderr'):
                acc = (acc + v).astype(f32)
            elif base == 'min':
                acc = np.minimum(acc, v)
            elif base == 'max':
                acc = np.maximum(acc, v)
        return acc

    if not is_complex:
        xf = x.astype(f32)
        if power:
            # first element squared in the input type (exact for ints), later
            # ones as fma(v, v, acc)
            first = (x[..., 0].astype(np.float64) ** 2).astype(f32)
            acc = first
            for k in range(1, factor):
                v = xf[..., k]
                sq = v.astype(np.float64) * v.astype(np.float64)
                if base in ('sum', 'mean', 'stderr'):
                    acc = (acc.astype(np.float64) + sq).astype(f32)     # fused multiply-add
                elif base == 'min':
                    acc = np.minimum(acc, sq.astype(f32))
                else:
                    acc = np.maximum(acc, sq.astype(f32))
        else:
            acc = fold(xf[..., 0], lambda k: xf[..., k])
        if base == 'mean':
            acc = (acc.astype(np.float64) * (1. / factor)).astype(f32)
        elif base == 'stderr':
            acc = (acc.astype(np.float64) * (1. / np.float64(np.sqrt(f32(factor))))).astype(f32)
        out = acc
    else:
        re, im = _as_complex_pair(x)
        if power:
            def mag2(k):
                r, i = re[..., k].astype(np.float64), im[..., k].astype(np.float64)
                return ((r * r).astype(f32).astype(np.float64) + i * i).astype(f32)   # fma(y,y,x*x)
            acc = fold(mag2(0), mag2)
            if base == 'mean':
                acc = (acc.astype(np.float64) * (1. / factor)).astype(f32)
            elif base == 'stderr':
                acc = (acc.astype(np.float64) * (1. / np.float64(np.sqrt(f32(factor))))).astype(f32)
            out = acc
        else:
            assert base in ('sum', 'mean', 'stderr')
            ar = fold(re[..., 0], lambda k: re[..., k])
            ai = fold(im[..., 0], lambda k: im[..., k])
            if base == 'mean':
                s = f32(1. / factor)
                ar, ai = ar * s, ai * s
            elif base == 'stderr':
                s = f32(1. / np.float64(np.sqrt(f32(factor))))
                ar, ai = ar * s, ai * s
            out = (ar + 1j * ai).astype(np.complex64)
    return np.ascontiguousarray(np.moveaxis(out, -1, axis))


def _fma(a, b, c):
    """float32 fused multiply-add (exact product in float64, one rounding)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def detect(x, mode, axis=None):
    """python/bifrost/blocks/detect.py:86-136 with src/Complex.hpp arithmetic
    (mag2 = fma(y,y,x*x); x*conj(y) via operator*=, Complex.hpp:194-200,217).
    Integer-complex input is converted unscaled."""
    re, im = _as_complex_pair(x)
    if mode == 'scalar':
        return _fma(im, im, (re * re).astype(f32))
    re = np.moveaxis(re, axis, 0)
    im = np.moveaxis(im, axis, 0)
    assert re.shape[0] == 2
    xr, xi, yr, yi = re[0], im[0], re[1], im[1]
    xx = _fma(xi, xi, (xr * xr).astype(f32))
    yy = _fma(yi, yi, (yr * yr).astype(f32))
    if mode == 'stokes_i':
        out = np.stack([(xx + yy).astype(f32)])
    elif mode in ('stokes', 'jones'):
        # xy = x * conj(y):  re = xr*yr - xi*(-yi);  im = xi*yr + xr*(-yi)
        xy_re = _fma(-xi, -yi, (xr * yr).astype(f32))
        xy_im = _fma(xr, -yi, (xi * yr).astype(f32))
        if mode == 'stokes':
            out = np.stack([(xx + yy).astype(f32), (xx - yy).astype(f32),
                            (f32(2) * xy_re).astype(f32), (f32(-2) * xy_im).astype(f32)])
        else:
            out = np.stack([xx + 1j * yy, xy_re + 1j * xy_im]).astype(np.complex64)
    elif mode == 'coherence':
        # conj(x) * y:  re = xr*yr - (-xi)*yi;  im = (-xi)*yr + xr*yi
        c_re = _fma(xi, yi, (xr * yr).astype(f32))
        c_im = _fma(xr, yi, ((-xi) * yr).astype(f32))
        out = np.stack([xx, yy, c_re, c_im])
    else:
        raise ValueError(mode)
    return np.ascontiguousarray(np.moveaxis(out, 0, axis))


def accumulate(a, b, beta):
    """blocks/accumulate.py:67: b = beta*b + (b_type)a."""
    a = a.astype(b.dtype)
    if beta == 0:
        return a.copy()
    return (b.dtype.type(beta) * b + a).astype(b.dtype)
