"""FDMT oracle (numpy).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows the reference line by line where arithmetic matters:
  plan   src/fdmt.cu:298-530   (BFfdmt_impl::init and helpers)
  init   src/fdmt.cu:52-92     (fdmt_init_kernel)
  step   src/fdmt.cu:95-155    (fdmt_exec_kernel)
  exec   src/fdmt.cu:629-718   (BFfdmt_impl::execute; diagonal final store)
All data arithmetic is float32 in the reference's operation order, so the
result is expected to be bit-identical to the reference GPU output.
"""
import math

import numpy as np


def _cpow(x, g):
    # std::pow(std::complex<double>, double): real pow for positive reals
    # (libstdc++ <complex>), principal branch otherwise.
    if x > 0:
        return complex(math.pow(x, g), 0.0)
    return complex(x, 0.0) ** g


class FdmtPlan(object):
    def __init__(self, nchan, max_delay, f0, df, exponent=-2.0):
        self.reverse_band = df < 0            # fdmt.cu:344-350
        if self.reverse_band:
            f0 += (nchan - 1) * df
            df = -df
        self.nchan, self.max_delay = int(nchan), int(max_delay)
        self.f0, self.df, self.exponent = float(f0), float(df), float(exponent)
        self._build()

    def cfreq(self, chan):                     # fdmt.cu:298-300
        return self.f0 + self.df * chan

    def rel_delay(self, flo, fhi, fmin=None, fmax=None):   # fdmt.cu:301-326
        if fmin is None:
            fmin, fmax = self.cfreq(0), self.cfreq(self.nchan - 1)
        g = self.exponent
        numer = _cpow(flo, g) - _cpow(fhi, g)
        denom = _cpow(fmin, g) - _cpow(fmax, g)
        eps = np.finfo(np.float64).eps
        if abs(denom) ** 2 < eps * eps:
            return 0.0
        return (numer / denom).real

    def subband_ndelay(self, f0, df):          # fdmt.cu:327-332
        fracdelay = self.rel_delay(f0, f0 + df)
        return int(math.ceil(fracdelay * (self.max_delay - 1))) + 1

    def _build(self):
        nchan = self.nchan
        # -- merge tree (fdmt.cu:364-387)
        parents = [None]
        nsub = nchan
        while nsub > 1:
            step = len(parents)
            cur = []
            for sb in range(0, nsub, 2):
                p0, p1 = sb, sb + 1
                if nsub % 2:
                    if (step - 1) % 2:
                        p0 -= 1
                        p1 -= 1
                    elif p1 == nsub:
                        p1 = -1
                cur.append((p0, p1))
            parents.append(cur)
            nsub = len(cur)
        nstep = len(parents)
        # -- channels per sub-band (fdmt.cu:391-405)
        nchans = [[1] * nchan]
        for step in range(1, nstep):
            prev = nchans[step - 1]
            nchans.append([(prev[p0] if p0 != -1 else 0) + (prev[p1] if p1 != -1 else 0)
                           for (p0, p1) in parents[step]])
        # -- channel / row offsets (fdmt.cu:407-436)
        chan_off, row_off = [], []
        for step in range(nstep):
            co, ro = [], []
            chan0 = row = 0
            for n in nchans[step]:
                f0 = self.cfreq(chan0) - (0.5 * self.df if step == 0 else 0.0)
                df = self.df * (1 if step == 0 else n - 1)
                co.append(chan0)
                ro.append(row)
                chan0 += n
                row += self.subband_ndelay(f0, df)
            co.append(chan0)
            ro.append(row)
            chan_off.append(co)
            row_off.append(ro)
        self.row_offsets = row_off
        self.nrow = [ro[-1] for ro in row_off]
        self.nrow_max = max(self.nrow)
        # -- per-row source tables (fdmt.cu:444-526)
        self.srcrows = [None]
        self.delays = [None]
        for step in range(1, nstep):
            src = np.full((self.nrow[step], 2), -1, dtype=np.int64)
            dly = np.zeros(self.nrow[step], dtype=np.int64)
            for sb, (p0, p1) in enumerate(parents[step]):
                p0_nchan = nchans[step - 1][p0] if p0 != -1 else 1
                p1_nchan = nchans[step - 1][p1] if p1 != -1 else 1
                p0_chan0 = chan_off[step - 1][p0 if p0 != -1 else p1]
                p1_chan0 = chan_off[step - 1][p1 if p1 != -1 else p0]
                if p1 == -1:
                    p1_chan0 += p0_nchan - 1
                flo = self.cfreq(p0_chan0)
                fmidlo = self.cfreq(p0_chan0 + (p0_nchan - 1))
                fmidhi = self.cfreq(p1_chan0)
                fhi = self.cfreq(p1_chan0 + (p1_nchan - 1))
                cmidlo = self.rel_delay(flo, fmidlo, flo, fhi)
                cmidhi = self.rel_delay(flo, fmidhi, flo, fhi)
                beg, end = row_off[step][sb], row_off[step][sb + 1]
                for delay in range(end - beg):
                    dmidlo = int(_c_round(delay * cmidlo))
                    dmidhi = int(_c_round(delay * cmidhi))
                    drest = delay - dmidhi
                    prev_mid1 = row_off[step - 1][p1] if p1 != -1 else -1
                    prev_end = row_off[step - 1][p1 + 1] if p1 != -1 else -1
                    if p1 != -1 and drest >= prev_end - prev_mid1:   # fdmt.cu:500-503
                        drest -= 1
                    r = beg + delay
                    src[r, 0] = row_off[step - 1][p0] + dmidlo if p0 != -1 else -1
                    src[r, 1] = row_off[step - 1][p1] + drest if p1 != -1 else -1
                    dly[r] = dmidhi
            self.srcrows.append(src)
            self.delays.append(dly)
        self.nstep = nstep


def _c_round(x):
    """C ::round -- half away from zero."""
    return math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5)


def fdmt_init(plan, x, reverse_time=False):
    """Step 0 (fdmt.cu:52-92): x [nchan, ntime] -> state [nrow0, ntime] float32."""
    nchan, ntime = x.shape
    offs = plan.row_offsets[0]
    state = np.full((offs[-1], ntime), np.nan, dtype=np.float32)
    t = np.arange(ntime)
    for c in range(nchan):
        c_ = nchan - 1 - c if plan.reverse_band else c
        row = x[c_].astype(np.float32)
        acc = np.zeros(ntime, dtype=np.float32)
        for d in range(offs[c + 1] - offs[c]):
            valid = t >= d
            t_ = (ntime - 1 - t) if reverse_time else t
            idx = t_ - d
            ok = valid & (idx >= 0)       # reference reads out of bounds where idx < 0
            acc[ok] = acc[ok] + row[idx[ok]]
            scale = np.float32(1.0) / np.float32(d + 1)
            state[offs[c] + d, valid] = acc[valid] * scale
    return state


def fdmt_step(plan, step, state):
    """One merge step (fdmt.cu:95-155), not final: returns [nrow, ntime] float32."""
    ntime = state.shape[1]
    src, dly = plan.srcrows[step], plan.delays[step]
    out = np.zeros((plan.nrow[step], ntime), dtype=np.float32)
    for r in range(plan.nrow[step]):
        s0, s1, d = int(src[r, 0]), int(src[r, 1]), int(dly[r])
        if s0 != -1:
            out[r] = state[s0]
        if s1 != -1 and d < ntime:
            out[r, d:] = out[r, d:] + state[s1, :ntime - d]
    return out


def fdmt(x, max_delay, f0, df, exponent=-2.0, negative_delays=False, out=None,
         plan=None):
    """Full transform of x [..., nchan, ntime] -> [..., max_delay, ntime] float32.
    Cells the reference never writes keep the value they have in `out`
    (default NaN-free sentinel: they are left as given, or 0 if out is None)."""
    x = np.asarray(x)
    if x.ndim > 2:
        lead = x.shape[:-2]
        xo = x.reshape((-1,) + x.shape[-2:])
        if out is None:
            out = np.zeros(lead + (max_delay, x.shape[-1]), dtype=np.float32)
        oo = out.reshape((-1,) + out.shape[-2:])
        for b in range(xo.shape[0]):
            fdmt(xo[b], max_delay, f0, df, exponent, negative_delays, oo[b], plan)
        return out
    nchan, ntime = x.shape
    if plan is None:
        plan = FdmtPlan(nchan, max_delay, f0, df, exponent)
    if out is None:
        out = np.zeros((max_delay, ntime), dtype=np.float32)
    state = fdmt_init(plan, x, negative_delays)
    for step in range(1, plan.nstep):
        state = fdmt_step(plan, step, state)
    # final diagonal store (fdmt.cu:120-124,150-151,702-708)
    for r in range(plan.nrow[-1]):
        if r >= ntime:
            continue
        vals = state[r, r:]
        if negative_delays:
            # t_ = ntime-1-t, stride+1: column ntime-1-t+r for t = r..ntime-1
            out[r, r:][::-1] = vals
        else:
            out[r, :ntime - r] = vals
    return out
