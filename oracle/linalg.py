"""Correlator oracle (numpy).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

c[..., i, j] = alpha * sum_t conj(x[..., t, i]) * x[..., t, j] + beta * c   for i >= j
(everything above the diagonal untouched) -- the b^H.b form of bfLinAlgMatMul
as used by CorrelateBlock (python/bifrost/blocks/correlate.py:79-103); formula
of test/test_linalg.py:168-185 and test/test_pipeline.py:258-298.  Integer
inputs are summed exactly before the float32 conversion: the products are
formed in float64 (BLAS), which holds every partial sum of 16-bit inputs over
up to 2^21 samples without rounding (|sum| < 2^53)."""
import numpy as np


def _to_int_pair(x):
    if x.dtype.names:
        return x['re'].astype(np.float64), x['im'].astype(np.float64)
    return None


def correlate(x, c=None, alpha=1.0, beta=0.0):
    """x: [..., ntime, n] structured ci8/ci16 or complex; returns cf32 [..., n, n]
    with the lower triangle updated and the rest of `c` (zeros if None) kept."""
    pair = _to_int_pair(x)
    if pair is not None:
        a, b = pair
        at, bt = np.swapaxes(a, -1, -2), np.swapaxes(b, -1, -2)
        re = at @ a + bt @ b                      # sum a_i a_j + b_i b_j
        im = at @ b - bt @ a                      # sum a_i b_j - b_i a_j
        full = re.astype(np.float32) + 1j * im.astype(np.float32)
    else:
        x128 = x.astype(np.complex128)
        full = np.swapaxes(x128.conj(), -1, -2) @ x128
    n = full.shape[-1]
    out = np.zeros(full.shape, np.complex64) if c is None else np.array(c, dtype=np.complex64)
    il = np.tril_indices(n)
    new = (np.float32(alpha) * full.astype(np.complex64))
    if beta != 0:
        new = new + np.float32(beta) * out
    out[..., il[0], il[1]] = new[..., il[0], il[1]]
    return out
