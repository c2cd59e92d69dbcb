"""CPU check of the packed-integer FDMT schedule (csrc/fdmt_packed.cuh).

bfFdmtPackedQuery returns the very tables bfFdmtExecute uploads for 1-byte
inputs; this file interprets them with numpy -- staging, per-warp register
rows, shared-memory rows, workspaces, bias removal, diagonal store -- and
compares the result with the oracle bit for bit.  Shared memory and the
workspaces start out poisoned, so an op that reads a sample no earlier op (or
staging) wrote shows up as a wrong output.  No GPU needed.
"""
import ctypes

import numpy as np
import pytest

from bifrost_b200.libbifrost import _bf
from oracle import fdmt as ofdmt

NO_A, NO_B, STORE_G, BYTES, GROUP4 = 1, 2, 4, 8, 16
WO_SHIFT, H_SHIFT, NVEC_SHIFT = 8, 10, 16
POISON = -(1 << 40)


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def query(nchan, md, f0, df):
    hdr = np.zeros(24, np.int32)
    assert _bf.bfFdmtPackedQuery(nchan, md, f0, df, -2.0, -1, _ip(hdr), None, None, None) == 0
    passes = []
    for k in range(int(hdr[0])):
        h = np.zeros(24, np.int32)
        assert _bf.bfFdmtPackedQuery(nchan, md, f0, df, -2.0, k, _ip(h), None, None, None) == 0
        keys = ('s0 s1 nlev esize src_kind dst_kind T nprog nwarp slots src_slots data_bytes '
                'lookback nrow_out smem_bytes nops lv fused prefetch early').split()
        p = dict(zip(keys, (int(v) for v in h)))
        p['fused'] = bool(p['fused'])
        ops = np.zeros((p['nprog'], p['nlev'], p['nwarp'], p['slots'], 4), np.int32)
        src = np.zeros((p['nprog'], p['src_slots'], 4), np.int32)
        ph = np.zeros((p['nprog'], 4), np.int32)
        assert _bf.bfFdmtPackedQuery(nchan, md, f0, df, -2.0, k, _ip(h), _ip(ops), _ip(src), _ip(ph)) == 0
        p.update(ops=ops, src=src, hdr=ph)
        passes.append(p)
    return passes


def geometry(passes, ntime):
    n = len(passes)
    g = [dict() for _ in range(n)]
    g[-1]['tb'] = 0
    g[-1]['nt'] = -(-ntime // passes[-1]['T'])
    g[-1]['te'] = g[-1]['nt'] * passes[-1]['T']
    for k in range(n - 2, -1, -1):
        g[k]['tb'] = -((passes[k + 1]['lookback'] - g[k + 1]['tb'] + 7) // 8 * 8)
        g[k]['nt'] = -(-(g[k + 1]['te'] - g[k]['tb']) // passes[k]['T'])
        g[k]['te'] = g[k]['tb'] + g[k]['nt'] * passes[k]['T']
    return g


class Machine(object):
    """State of one pass: interprets a (program, tile) exactly as the kernel
    does -- shared memory is a byte-addressed array of elements (one int64 /
    float64 cell per element, POISON / NaN where nothing was written)."""

    def __init__(self, p, g, x, xi, signed, ws_prev, tb_prev, ring_prev, ws, ring, out):
        self.p, self.g, self.x, self.xi, self.signed = p, g, x, xi, signed
        self.ws_prev, self.tb_prev, self.ring_prev = ws_prev, tb_prev, ring_prev
        self.ws, self.ring, self.out = ws, ring, out
        self.ntime = x.shape[1]

    def run(self, prog, tile):
        p, g, ntime, signed = self.p, self.g, self.ntime, self.signed
        esz = p['esize']
        VS = 16 // esz
        final = p['dst_kind'] == 2
        nchan_band, nsrc, staged, region = (int(v) for v in p['hdr'][prog])
        # a pass whose source shares region 0 and has an odd number of levels lays
        # every other tile out mirrored (pk_tiles): region 0 <-> region 1
        assert region == 0 or (p['early'] and p['nlev'] % 2 == 1 and p['src_kind'] == 1)
        flip = region if (tile & 1) else 0
        bias = 128 * nchan_band if signed else 0
        t0 = g['tb'] + tile * p['T']
        nelem = p['data_bytes']          # one cell per BYTE offset keeps byte and word rows in one array
        data = np.full(nelem, POISON, np.int64) if esz == 2 else np.full(nelem, np.nan, np.float64)
        srcs = p['src'][prog]
        nbytes = 0
        for e in srcs[:nsrc]:
            row, y, z, w = (int(v) for v in e)
            assert w > 0 and w % VS == 0 and z % 16 == 0
            ts = t0 + y
            if p['src_kind'] == 0:
                t = np.arange(ts, ts + w)
                ok = (t >= 0) & (t < ntime)
                vals = np.full(w, 128 if signed else 0, np.int64)
                vals[ok] = self.xi[row, t[ok]]
                data[z:z + w] = vals                       # one byte per cell, misalignment 0 in this model
            else:
                c0 = ts - self.tb_prev
                assert c0 >= 0 and c0 % VS == 0 and c0 + w <= self.ws_prev_width
                cols = (c0 + np.arange(w)) % self.ring_prev
                vals = self.ws_prev[row, cols]
                assert region == 0 or z + w * esz <= region
                z += flip
                data[z:z + w * esz:esz] = vals
                nbytes += w * esz
        if p['src_kind'] != 0:
            assert nbytes == staged
        if nsrc < p['src_slots']:
            assert srcs[nsrc][3] == 0
        for lev in range(1, p['nlev'] + 1):
            aflip, oflip = (flip, -flip) if lev & 1 else (-flip, flip)
            for warp in range(p['nwarp']):
                oplist = p['ops'][prog, lev - 1, warp]
                skip = False
                for m, op in enumerate(oplist):
                    if skip:
                        skip = False
                        continue
                    dst, ay, bz, ctl = (int(v) for v in op)
                    if ctl == 0:
                        break
                    n = (ctl >> NVEC_SHIFT) * VS
                    assert n <= 32 * p['lv'] * VS
                    if ctl & GROUP4:
                        # steps 1+2 fused: two slots, four input channels
                        assert p['src_kind'] == 0 and lev == 1 and p['fused']
                        dst1, f2, f3, ctl2 = (int(v) for v in oplist[m + 1])
                        assert ctl2 & GROUP4
                        skip = True

                        def chan(field, back):
                            k, e = field & 0xFFF, (field & 0xFFFFFFFF) >> 12
                            assert k < nsrc and e - back >= 0 and e + n <= int(srcs[k][3])
                            base = int(srcs[k][2])
                            return data[base + e - back: base + e - back + n]
                        s0 = chan(ay, 0) + chan(bz, 0)
                        mask = (ctl >> WO_SHIFT) & 3
                        for d, where in ((0, dst), (1, dst1)):
                            if mask & (1 << d):
                                r = s0 + chan(f2, d) + chan(f3, d)
                                assert (r >= 0).all() and (r < 65536).all() and where % 16 == 0
                                data[where: where + n * esz: esz] = r
                        continue
                    if ctl & BYTES:
                        assert p['src_kind'] == 0 and lev == 1
                        def rd(field):
                            k, e = field & 0xFFF, (field & 0xFFFFFFFF) >> 12
                            assert k < nsrc
                            base = int(srcs[k][2])
                            assert e + n <= int(srcs[k][3])
                            return data[base + e: base + e + n]
                        a = np.zeros(n, np.int64) if ctl & NO_A else rd(ay)
                        b = np.zeros(n, np.int64) if ctl & NO_B else rd(bz)
                    else:
                        sub = ((ctl >> WO_SHIFT) & 3) * (4 // esz) + ((ctl >> H_SHIFT) & 1)
                        if esz == 4:
                            assert (ctl >> H_SHIFT) & 1 == 0
                        assert ay % 16 == 0 and bz % 16 == 0
                        ay += aflip
                        bz += aflip
                        assert (ctl & NO_A) or 0 <= ay < 2 * region or not region
                        a = np.zeros(n, data.dtype) if ctl & NO_A else data[ay: ay + n * esz: esz]
                        b = np.zeros(n, data.dtype) if ctl & NO_B else data[bz + sub * esz: bz + (sub + n) * esz: esz]
                    if esz == 4:
                        r = (a.astype(np.float32) + b.astype(np.float32)).astype(np.float64)
                    else:
                        r = a + b
                    if not (ctl & STORE_G):
                        assert dst % 16 == 0 and lev < p['nlev']
                        dst += oflip
                        assert dst >= 0
                        if esz == 2:
                            assert (r >= 0).all() and (r < 65536).all()
                        data[dst: dst + n * esz: esz] = r
                        continue
                    assert n == p['T'] and lev == p['nlev']
                    if esz == 2:
                        assert (r >= 0).all()
                        if p['dst_kind'] == 0:
                            assert (r < 65536).all()
                            val = r
                        else:
                            val = (r - bias).astype(np.float64)
                    else:
                        val = r
                    if final:
                        d = dst
                        t = np.arange(t0, t0 + n)
                        ok = (t >= d) & (t < ntime)
                        assert not np.isnan(val[ok]).any()
                        self.out[d, t[ok] - d] = val[ok].astype(np.float32)
                    else:
                        cols = (t0 - g['tb'] + np.arange(n)) % self.ring
                        self.ws[dst, cols] = val


def run_schedule(x, passes, out):
    """x [nchan, ntime] int8/uint8 -> writes out [max_delay, ntime] float32,
    pass by pass over linear workspaces (the separate-launch form)."""
    nchan, ntime = x.shape
    signed = x.dtype == np.int8
    geo = geometry(passes, ntime)
    xi = x.astype(np.int64) + (128 if signed else 0)
    ws_prev, tb_prev, width_prev = None, 0, 0
    for k, p in enumerate(passes):
        g = geo[k]
        width = g['te'] - g['tb']
        is_float = p['dst_kind'] == 1 or p['esize'] == 4
        ws = None
        if p['dst_kind'] != 2:
            ws = np.full((p['nrow_out'], width), np.nan if is_float else POISON, np.float64 if is_float else np.int64)
        m = Machine(p, g, x, xi, signed, ws_prev, tb_prev, max(width_prev, 1), ws, max(width, 1), out)
        m.ws_prev_width = width_prev
        for prog in range(p['nprog']):
            for tile in range(g['nt']):
                m.run(prog, tile)
        ws_prev, tb_prev, width_prev = ws, g['tb'], width
    return out


CASES = [
    # nchan, max_delay, f0, df, ntime, dtype
    (16, 12, 1000.0, 10.0, 300, np.int8),
    (17, 9, 60.0, -0.5, 257, np.uint8),          # odd channel count, reversed band
    (64, 50, 1200.0, 3.0, 1500, np.int8),
    (100, 37, 400.0, 0.25, 900, np.uint8),
    (256, 130, 1000.0, 1.5, 1100, np.int8),
    (1024, 300, 1000.0, 400. / 1024, 800, np.int8),   # crosses the 16-bit limit (steps 9, 10 in fp32)
]


@pytest.mark.parametrize("nchan,md,f0,df,ntime,dtype", CASES)
def test_packed_tables_reproduce_the_oracle(nchan, md, f0, df, ntime, dtype):
    passes = query(nchan, md, f0, df)
    if not passes:
        pytest.skip("integer schedule does not apply to this plan")
    rng = np.random.default_rng(nchan * 7 + md)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max + 1, size=(nchan, ntime)).astype(dtype)
    x[:, :5] = info.min          # extremes next to the t < 0 edge
    x[::3, 7:40] = info.max
    gold = np.full((md, ntime), -12345.0, np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    got = np.full((md, ntime), -12345.0, np.float32)
    run_schedule(x, passes, got)
    assert np.array_equal(got.view(np.uint32), gold.view(np.uint32))


def test_packed_schedule_of_the_baseline_plan():
    """Config 2's plan (4096 chan, max_delay 794): structure + one short gulp."""
    nchan, md, f0, df = 4096, 794, 1000.0, 400. / 4096
    passes = query(nchan, md, f0, df)
    assert [(p['s0'], p['s1'], p['esize']) for p in passes] == [(1, 5, 2), (6, 9, 2), (10, 12, 4)]
    assert passes[0]['src_kind'] == 0 and passes[1]['dst_kind'] == 1 and passes[2]['dst_kind'] == 2
    for p in passes:
        assert p['smem_bytes'] <= 111 * 1024 and p['T'] >= 256     # two or three CTAs per SM
    ntime = 1200
    rng = np.random.default_rng(5)
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    gold = np.zeros((md, ntime), np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    got = np.zeros((md, ntime), np.float32)
    run_schedule(x, passes, got)
    assert np.array_equal(got.view(np.uint32), gold.view(np.uint32))


@pytest.mark.parametrize("nchan,md,f0,df,ntime,nrank", [
    (256, 130, 1000.0, 1.5, 1100, 2),
    (256, 130, 1000.0, 1.5, 1100, 8),
    (1024, 300, 1000.0, 400. / 1024, 800, 2),      # split at the 16-bit limit (fp32 exchange rows)
    (64, 50, 1200.0, -3.0, 700, 2),                # reversed band
])
def test_sharded_schedule_tables_reproduce_the_oracle(nchan, md, f0, df, ntime, nrank, monkeypatch):
    """bfFdmtShardInit cuts the passes at the step that has `nrank` sub-bands and
    runs ONE pass above it; the tables of that schedule (whatever rank runs
    which program) must still be the transform."""
    sx = int(np.log2(nchan // nrank))
    monkeypatch.setenv('BFB_FDMT_PACKED_FORCE_END', str(sx))
    passes = query(nchan, md, f0, df)
    assert passes and passes[-2]['s1'] == sx and passes[-1]['s0'] == sx + 1
    rng = np.random.default_rng(nchan + nrank)
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    gold = np.full((md, ntime), -12345.0, np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    got = np.full((md, ntime), -12345.0, np.float32)
    run_schedule(x, passes, got)
    assert np.array_equal(got.view(np.uint32), gold.view(np.uint32))


def test_plans_beyond_exact_fp32_integers_keep_the_float_schedule():
    # 255 * nchan must stay below 2**24 for the integer argument to hold
    assert query(70000, 8, 1000.0, 0.001) == []
    assert query(65536, 8, 1000.0, 0.001) != []


# ---------------------------------------------------------------------------
# The persistent single-launch form: claim order, waits, ring workspaces
# ---------------------------------------------------------------------------
def mega_query(nchan, md, f0, df, ntime):
    h = np.zeros(32, np.int64)
    lp = h.ctypes.data_as(ctypes.POINTER(ctypes.c_long))
    assert _bf.bfFdmtPackedMegaQuery(nchan, md, f0, df, -2.0, ntime, lp, None) == 0
    if h[0] == 0:
        return None
    npass, lag, ipr, nchunk, C, t_ref = (int(v) for v in h[1:7])
    tmpl = np.zeros(ipr, np.int32)
    assert _bf.bfFdmtPackedMegaQuery(nchan, md, f0, df, -2.0, ntime, lp, _ip(tmpl)) == 0
    per = [dict(tb=int(h[7 + 3 * k]), nt=int(h[8 + 3 * k]), ring=int(h[9 + 3 * k])) for k in range(npass)]
    return dict(npass=npass, lag=lag, ipr=ipr, nchunk=nchunk, C=C, t_ref=t_ref, per=per, tmpl=tmpl)


def run_persistent(x, passes, mg, out):
    """Executes the items in claim order (one at a time), with the kernel's own
    wait rules and ring workspaces.  Asserts that (1) nothing an item waits for
    is still unfinished when the item is claimed -- so the claimed prefix can
    always make progress, whatever the interleaving on the device -- and (2) the
    waits cover the true dependencies: every producer tile whose columns the
    item stages, and every consumer tile that still has to read the ring columns
    the item overwrites."""
    nchan, ntime = x.shape
    signed = x.dtype == np.int8
    xi = x.astype(np.int64) + (128 if signed else 0)
    n = mg['npass']
    C, t_ref, lag, nchunk = mg['C'], mg['t_ref'], mg['lag'], mg['nchunk']
    geo = geometry(passes, ntime)
    for k in range(n):
        assert geo[k]['tb'] == mg['per'][k]['tb'] and geo[k]['nt'] == mg['per'][k]['nt']
    T = [p['T'] for p in passes]
    tb = [g['tb'] for g in geo]
    nt = [g['nt'] for g in geo]

    def ifirst(k, j):
        xx = t_ref + j * C - tb[k]
        i = 0 if xx <= 0 else -(-xx // T[k])
        return min(i, nt[k])

    def chunk_of(k, i):
        return (tb[k] + i * T[k] - t_ref) // C

    # workspaces (rings) and machines
    ws, machines = [], []
    for k, p in enumerate(passes):
        ring = mg['per'][k]['ring']
        if p['dst_kind'] == 2:
            ws.append(None)
        else:
            is_float = p['dst_kind'] == 1 or p['esize'] == 4
            ws.append(np.full((p['nrow_out'], ring), np.nan if is_float else POISON,
                              np.float64 if is_float else np.int64))
    for k, p in enumerate(passes):
        m = Machine(p, geo[k], x, xi, signed, ws[k - 1] if k else None, tb[k - 1] if k else 0,
                    mg['per'][k - 1]['ring'] if k else 1, ws[k], mg['per'][k]['ring'] if ws[k] is not None else 1, out)
        m.ws_prev_width = 1 << 60
        machines.append(m)
    # exact source reach of every program: (min first-sample offset, max end offset) relative to t0
    reach = []
    for p in passes:
        r = []
        for prog in range(p['nprog']):
            nsrc = int(p['hdr'][prog][1])
            e = p['src'][prog][:nsrc]
            r.append((int(e[:, 1].min()), int((e[:, 1] + e[:, 3]).max())))
        reach.append(r)
    done = np.zeros((n, nchunk), np.int64)
    nround = nchunk + lag * (n - 1)
    executed = 0
    amin = [min(r[0] for r in reach[k]) for k in range(n)]
    for rnd in range(nround):
        for e in mg['tmpl']:
            e = int(e)
            k, prog = e >> 24, e & 0xFFFFFF
            j = rnd - lag * k
            if j < 0 or j >= nchunk:
                continue
            i0, i1 = ifirst(k, j), ifirst(k, j + 1)
            waited = {}
            if i0 < i1:
                t_lo, t_hi = tb[k] + i0 * T[k], tb[k] + i1 * T[k]
                if k > 0:
                    lo, hi = t_lo - passes[k]['lookback'] - tb[k - 1], t_hi - 1 - tb[k - 1]
                    ilo, ihi = (0 if lo <= 0 else lo // T[k - 1]), min(hi // T[k - 1], nt[k - 1] - 1)
                    waited[k - 1] = (chunk_of(k - 1, ilo), chunk_of(k - 1, ihi))
                    # true producers of this program's staged columns
                    a, b = reach[k][prog]
                    for ip in range(max(0, (t_lo + a - tb[k - 1]) // T[k - 1]),
                                    (t_hi - T[k] + b - 1 - tb[k - 1]) // T[k - 1] + 1):
                        assert waited[k - 1][0] <= chunk_of(k - 1, ip) <= waited[k - 1][1]
                if k + 1 < n:
                    ring = mg['per'][k]['ring']
                    lo = t_lo - ring - tb[k + 1]
                    hi = t_hi - 1 - ring + passes[k + 1]['lookback'] - tb[k + 1]
                    if hi >= 0:
                        ilo, ihi = (0 if lo <= 0 else lo // T[k + 1]), min(hi // T[k + 1], nt[k + 1] - 1)
                        if ilo <= ihi:
                            waited[k + 1] = (chunk_of(k + 1, ilo), chunk_of(k + 1, ihi))
                    # true readers of the overwritten columns: consumer tiles whose staged
                    # times intersect [t_lo - ring, t_hi - ring)
                    old_lo, old_hi = t_lo - ring, t_hi - ring
                    for ic in range(nt[k + 1]):
                        tc = tb[k + 1] + ic * T[k + 1]
                        if tc + amin[k + 1] >= old_hi:
                            break
                        if tc + T[k + 1] <= old_lo:
                            continue
                        w = waited.get(k + 1)
                        assert w is not None and w[0] <= chunk_of(k + 1, ic) <= w[1]
            for kk, (jlo, jhi) in waited.items():
                for jj in range(jlo, jhi + 1):
                    if 0 <= jj < nchunk:
                        assert done[kk, jj] == passes[kk]['nprog'], "item claimed before what it waits for"
            for i in range(i0, i1):
                machines[k].run(prog, i)
                executed += 1
            done[k, j] += 1
    assert executed == sum(nt[k] * passes[k]['nprog'] for k in range(n))
    assert all((done[k] == passes[k]['nprog']).all() for k in range(n))


@pytest.mark.parametrize("nchan,md,f0,df,ntime,knobs", [
    (64, 50, 1200.0, 3.0, 5000, dict(BFB_FDMT_PACKED_CHUNK='800')),
    (256, 130, 1000.0, 1.5, 9000, dict(BFB_FDMT_PACKED_CHUNK='1000', BFB_FDMT_PACKED_TCAP='256,256,256')),
    (1024, 300, 1000.0, 400. / 1024, 7000, dict(BFB_FDMT_PACKED_CHUNK='740')),
    (100, 37, 400.0, 0.25, 12000, dict(BFB_FDMT_PACKED_CHUNK='300', BFB_FDMT_PACKED_TCAP='128,96,64')),
])
def test_persistent_schedule_is_deadlock_free_and_exact(nchan, md, f0, df, ntime, knobs, monkeypatch):
    monkeypatch.setenv('BFB_FDMT_PACKED_MEGA', '1')
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    passes = query(nchan, md, f0, df)
    mg = mega_query(nchan, md, f0, df, ntime)
    assert passes and mg is not None
    # the rings must really wrap for the test to mean something (except tiny plans)
    rng = np.random.default_rng(ntime)
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    gold = np.full((md, ntime), -12345.0, np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    got = np.full((md, ntime), -12345.0, np.float32)
    run_persistent(x, passes, mg, got)
    assert np.array_equal(got.view(np.uint32), gold.view(np.uint32))
    if len(passes) > 1:
        assert any(mg['per'][k]['ring'] < geometry(passes, ntime)[k]['te'] - geometry(passes, ntime)[k]['tb']
                   for k in range(len(passes) - 1))


@pytest.mark.parametrize("nrank", [2, 4, 8])
def test_packed_tables_of_the_config5_subband_plans(nrank):
    """The per-GPU plans `bench.py --gpus N` builds for BASELINE config 5
    (nchan/N channels, f0_g, max_delay_g): every one of them takes the integer
    schedule and its tables reproduce the oracle."""
    import bench
    for g in range(nrank):
        sb = bench.subband(g, nrank)
        nchan, md, f0, df = sb['nchan'], sb['max_delay'], sb['f0'], sb['df']
        passes = query(nchan, md, f0, df)
        assert passes, (nrank, g)
        ntime = md + 700
        rng = np.random.default_rng(100 * nrank + g)
        x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
        gold = np.zeros((md, ntime), np.float32)
        ofdmt.fdmt(x, md, f0, df, out=gold)
        got = np.zeros((md, ntime), np.float32)
        run_schedule(x, passes, got)
        assert np.array_equal(got.view(np.uint32), gold.view(np.uint32)), (nrank, g)


def test_packed_tables_of_random_plans():
    """Forty random geometries (odd channel counts, reversed and narrow bands,
    bank depths from 1 to a few hundred, low frequencies where the delays pile
    up in the first channels): whenever bfFdmtInit would take the integer
    schedule its tables must be the transform."""
    rng = np.random.default_rng(20240923)
    ran = 0
    for case in range(40):
        nchan = int(rng.integers(2, 400))
        md = int(rng.integers(1, 260))
        f0 = float(rng.choice([40.0, 150.0, 400.0, 1000.0, 1400.0, 4000.0]))
        frac = float(rng.uniform(0.02, 0.6))                 # bandwidth as a fraction of f0
        df = f0 * frac / nchan * (1 if rng.random() < 0.7 else -1)
        if df < 0:
            f0 = f0 * (1 + frac)                             # keep every channel above zero
        ntime = md + int(rng.integers(40, 500))
        dtype = np.int8 if rng.random() < 0.6 else np.uint8
        try:
            passes = query(nchan, md, f0, df)
        except Exception:
            passes = None                                    # a plan the reference would refuse as well
        if not passes:
            continue
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max + 1, size=(nchan, ntime)).astype(dtype)
        gold = np.full((md, ntime), -7.0, np.float32)
        ofdmt.fdmt(x, md, f0, df, out=gold)
        got = np.full((md, ntime), -7.0, np.float32)
        run_schedule(x, passes, got)
        assert np.array_equal(got.view(np.uint32), gold.view(np.uint32)), (case, nchan, md, f0, df, ntime, dtype)
        ran += 1
    assert ran >= 25
