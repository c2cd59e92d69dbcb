"""CPU check of the FDMT tile-pass planner (csrc/fdmt_tiles.cuh).

bfFdmtTileQuery returns the work-item tables bfFdmtExecute uploads for a pass;
this test interprets them with numpy exactly as fdmt_tile_kernel does (stage,
then one phase per step, float32 adds, -0.0f before t = 0) and compares every
row the pass produces with the oracle's state of that step, bit for bit.  It
pins the host logic (windows, offsets, alignment rules, flags) without a GPU;
the kernels themselves are covered by tests/test_fdmt.py -m gpu."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bifrost_b200.libbifrost import _bf  # noqa: E402
from oracle import fdmt as ofdmt  # noqa: E402

NO_A, NO_B, SLOW = 1, 2, 4
F32 = np.float32


def query(nchan, md, f0, df, s0, s1, D, nwarp, raw):
    hdr = (ctypes.c_int * 8)()
    st = _bf.bfFdmtTileQuery(nchan, md, f0, df, -2.0, s0, s1, D, nwarp, int(raw), hdr, None, None)
    if st != 0:
        return None
    T, nprog, nphase, slots, smem_floats, raw_bytes, edge, nitem = list(hdr)
    items = (ctypes.c_int * (4 * nitem))()
    aux = (ctypes.c_int * (4 * nprog * nwarp * slots))()
    assert _bf.bfFdmtTileQuery(nchan, md, f0, df, -2.0, s0, s1, D, nwarp, int(raw), hdr, items, aux) == 0
    items = np.frombuffer(items, np.int32).reshape(nprog, nphase, nwarp, slots, 4).copy()
    aux = np.frombuffer(aux, np.int32).reshape(nprog, nwarp, slots, 4).copy()
    return dict(T=T, nprog=nprog, nphase=nphase, slots=slots, smem=smem_floats, raw_bytes=raw_bytes,
                edge=edge, items=items, aux=aux)


def oracle_states(x, md, f0, df):
    plan = ofdmt.FdmtPlan(x.shape[0], md, f0, df)
    states = [ofdmt.fdmt_init(plan, x)]
    for s in range(1, plan.nstep):
        states.append(ofdmt.fdmt_step(plan, s, states[-1]))
    return plan, states


def window(row, t_start, n, fill):
    """row[t_start : t_start + n] with `fill` outside the row."""
    out = np.full(n, fill, row.dtype)
    lo, hi = max(0, t_start), min(len(row), t_start + n)
    if hi > lo:
        out[lo - t_start:hi - t_start] = row[lo:hi]
    return out


def state0_fast(raw, off, n, dd, signed):
    """fdmt_tiles.cuh tile_state0: newest-to-oldest float32 sum, one multiply."""
    cast = np.int8 if signed else np.uint8
    acc = raw[off:off + n].view(cast).astype(F32)
    for k in range(1, dd + 1):
        acc = acc + raw[off - k:off - k + n].view(cast).astype(F32)
    return acc * (F32(1.0) / F32(dd + 1)) if dd else acc


def state0_exact(raw, off, n, dd, t, signed):
    """tile_state0_exact: -0.0 before t = 0, NaN for t < dd."""
    cast = np.int8 if signed else np.uint8
    acc = np.zeros(n, F32)
    for k in range(dd + 1):
        acc = acc + raw[off - k:off - k + n].view(cast).astype(F32)
    val = acc * (F32(1.0) / F32(dd + 1))
    val = np.where(t < dd, F32(np.nan), val)
    return np.where(t < 0, F32(-0.0), val).astype(F32)


def run_pass(tp, src_state, x, ntime, tiles, signed=True):
    """Interprets the item tables for the given tile indices; returns {row: {t: value}} as
    a dense array [nrow_out_max][ntime] with NaN-pattern sentinel where nothing was written."""
    raw = tp['raw_bytes'] > 0
    written = {}
    for prog in range(tp['nprog']):
        for tile in tiles:
            t0 = tile * tp['T']
            smem = np.zeros(tp['smem'] + 64, F32)
            rbuf = np.zeros(tp['raw_bytes'] + 64, np.uint8)
            # ---- stage
            for it in tp['items'][prog, 0].reshape(-1, 4):
                if it[3] == 0:
                    continue
                if raw:
                    rbuf[it[2]:it[2] + it[3]] = window(x[it[0]].view(np.uint8), t0 + it[1], it[3], 0)
                else:
                    vals = window(src_state[it[0]], t0 + it[1], 4 * it[3], F32(0))
                    tt = t0 + it[1] + np.arange(4 * it[3])
                    vals[tt < 0] = F32(-0.0)
                    smem[it[2]:it[2] + 4 * it[3]] = vals
            # ---- merge phases
            for phase in range(1, tp['nphase']):
                last = phase == tp['nphase'] - 1
                new = smem.copy()
                for (wi, si), it in np.ndenumerate(np.zeros(tp['items'].shape[2:4])):
                    x_, y_, z_, w_ = tp['items'][prog, phase, wi, si]
                    if raw and phase == 1:
                        nvec, flags = w_ & 0xFF, w_ >> 16
                    else:
                        nvec, flags = w_ & 0xFFFF, w_ >> 16
                    if nvec == 0:
                        continue
                    n = 4 * nvec
                    if raw and phase == 1:
                        smax_r, delay, dd0x, dd1x = tp['aux'][prog, wi, si]
                        if flags == 0 and t0 >= tp['edge']:
                            dd0, dd1 = (w_ >> 8) & 0xF, (w_ >> 12) & 0xF
                            val = state0_fast(rbuf, y_, n, dd0, signed) + state0_fast(rbuf, z_, n, dd1, signed)
                        else:
                            t = t0 - smax_r + np.arange(n)
                            va = np.zeros(n, F32) if flags & NO_A else state0_exact(rbuf, y_, n, dd0x, t, signed)
                            val = va
                            if not flags & NO_B:
                                val = va + state0_exact(rbuf, z_, n, dd1x, t - delay, signed)
                            if flags & NO_A:
                                val = np.where(t < 0, F32(-0.0), val).astype(F32)
                        tw = t0 - smax_r if not (flags == 0 and t0 >= tp['edge']) else None
                    else:
                        a = smem[y_:y_ + n]
                        b = smem[z_:z_ + n]
                        if flags & NO_A:
                            val = F32(0) + b
                        elif flags & NO_B:
                            val = a.copy()
                        else:
                            val = a + b
                    if not last:
                        new[x_:x_ + n] = val
                    else:
                        # last phase rows have smin = smax = 0: sample w is time t0 + w
                        row = written.setdefault(int(x_), {})
                        for w in range(n):
                            t = t0 + w
                            if 0 <= t < ntime and w < tp['T']:
                                row[t] = val[w]
                smem = new
    return written


def check(tp, want_state, written, ntime, tiles):
    assert written, "the pass wrote nothing"
    nrow_seen = 0
    for row, cells in written.items():
        ts = np.array(sorted(cells))
        got = np.array([cells[t] for t in ts], F32)
        want = want_state[row][ts]
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32),
                                      err_msg='row %d' % row)
        nrow_seen += 1
    return nrow_seen


CASES = [
    # nchan, max_delay, f0, bw, ntime
    (64, 50, 1000., 400., 1500),
    (100, 70, 1200., 300., 1300),      # odd band counts: absent parents (flags)
    (256, 120, 1000., 400., 1200),
]


@pytest.mark.parametrize("nchan,md,f0,bw,ntime", CASES)
@pytest.mark.parametrize("split", [(2, 3), (3, 5), (4, 99)])
def test_float_tile_pass_reproduces_the_oracle_states(nchan, md, f0, bw, ntime, split):
    rng = np.random.default_rng(nchan + md)
    x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    df = bw / nchan
    plan, states = oracle_states(x, md, f0, df)
    s0, s1 = split[0], min(split[1], plan.nstep - 1)
    if s0 > s1:
        pytest.skip('plan has too few steps')
    tp = query(nchan, md, f0, df, s0, s1, 8, 4, raw=False)
    assert tp is not None
    ntile = -(-ntime // tp['T'])
    tiles = sorted({0, 1, ntile // 2, ntile - 1})
    written = run_pass(tp, states[s0 - 1], None, ntime, tiles)
    nrow = check(tp, states[s1], written, ntime, tiles)
    assert nrow == plan.nrow[s1]            # every row of the pass's last step is produced


@pytest.mark.parametrize("nchan,md,f0,bw,ntime", CASES + [(48, 400, 60., 30., 1400)])   # last: step-0 delays > 3
@pytest.mark.parametrize("signed", [True, False])
def test_raw_tile_pass_reproduces_the_oracle_states(nchan, md, f0, bw, ntime, signed):
    rng = np.random.default_rng(nchan * 3 + md)
    x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    if not signed:
        x = x.view(np.uint8)
    df = bw / nchan
    plan, states = oracle_states(x, md, f0, df)
    s1 = min(3, plan.nstep - 1)
    tp = query(nchan, md, f0, df, 1, s1, 32, 4, raw=True)
    if tp is None:
        pytest.skip('raw pass not tileable for this plan (falls back to the head kernel)')
    ntile = -(-ntime // tp['T'])
    tiles = sorted({0, 1, ntile - 1})
    written = run_pass(tp, None, x, ntime, tiles, signed=signed)
    nrow = check(tp, states[s1], written, ntime, tiles)
    assert nrow == plan.nrow[s1]


def test_item_tables_respect_the_alignment_rules():
    """dst and a offsets are multiples of 4 floats, windows are whole float4s and
    at most 96 vectors long, stage time offsets are multiples of 4."""
    tp = query(4096, 794, 1000., 400. / 4096, 6, 9, 24, 8, raw=False)
    assert tp is not None and tp['T'] % 4 == 0 and tp['T'] >= 64
    it = tp['items']
    stage = it[:, 0].reshape(-1, 4)
    stage = stage[stage[:, 3] > 0]
    assert (stage[:, 1] % 4 == 0).all() and (stage[:, 1] <= 0).all() and (stage[:, 2] % 4 == 0).all()
    assert (stage[:, 3] <= 96).all()
    for phase in range(1, tp['nphase']):
        m = it[:, phase].reshape(-1, 4)
        m = m[(m[:, 3] & 0xFFFF) > 0]
        assert ((m[:, 3] & 0xFFFF) <= 96).all()
        assert (m[:, 1] % 4 == 0).all()
        if phase < tp['nphase'] - 1:
            assert (m[:, 0] % 4 == 0).all()
            assert (m[:, 0] + 4 * (m[:, 3] & 0xFFFF) <= tp['smem']).all()
        assert (m[:, 2] + 4 * (m[:, 3] & 0xFFFF) + 4 <= tp['smem'] + 16).all()


def test_random_plans_and_passes_reproduce_the_oracle():
    """Seeded fuzz over plan geometry (odd channel counts, reversed bands, low
    frequencies with long step-0 windows), pass ranges, delay-block sizes, warp
    counts, raw and float passes; the first, last and a random tile of each."""
    rng = np.random.default_rng(20240922)
    ncase = 0
    while ncase < 120:
        nchan = int(rng.integers(2, 160))
        md = int(rng.integers(1, 120))
        f0 = float(rng.uniform(30, 3000))
        bw = float(rng.uniform(5, f0 * 0.6)) * (1 if rng.random() < 0.8 else -1)
        df = bw / nchan
        ntime = int(rng.integers(md + 5, md + 700))
        n = ctypes.c_int(0)
        if _bf.bfFdmtPlanQuery(nchan, md, f0, df, -2.0, -1, ctypes.byref(n), None) != 0 or n.value < 2:
            continue
        nstep = n.value
        x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
        plan, states = oracle_states(x, md, f0, df)
        s0 = int(rng.integers(1, nstep))
        s1 = int(rng.integers(s0, nstep))
        raw = s0 == 1 and rng.random() < 0.7
        if s0 == 1 and not raw:
            s0 = 2
            if s0 > s1:
                continue
        tp = query(nchan, md, f0, df, s0, s1, int(rng.choice([4, 8, 16, 32])), int(rng.choice([1, 2, 4, 8])), raw)
        if tp is None:
            continue
        ntile = -(-ntime // tp['T'])
        tiles = sorted({0, ntile - 1, int(rng.integers(0, ntile))})
        written = run_pass(tp, None if raw else states[s0 - 1], x if raw else None, ntime, tiles)
        assert check(tp, states[s1], written, ntime, tiles) == plan.nrow[s1]
        ncase += 1
