"""CPU check of the integer "chain" FDMT schedule (csrc/fdmt_chain.cuh).

bfFdmtChainQuery returns the very tables bfFdmtExecute uploads for 1-byte
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

LEVEL_MASK, LOADA, NO_A, NO_B, STORE_S, STORE_G = 7, 8, 16, 32, 64, 128
POISON = -(1 << 40)


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def query(nchan, md, f0, df):
    hdr = np.zeros(16, np.int32)
    assert _bf.bfFdmtChainQuery(nchan, md, f0, df, -2.0, -1, _ip(hdr), None, None, None) == 0
    passes = []
    for k in range(int(hdr[0])):
        h = np.zeros(16, np.int32)
        assert _bf.bfFdmtChainQuery(nchan, md, f0, df, -2.0, k, _ip(h), None, None, None) == 0
        keys = ('s0 s1 nlev esize src_kind dst_kind T nprog nwarp slots src_slots smem_elems '
                'lookback nrow_out smem_bytes nops').split()
        p = dict(zip(keys, (int(v) for v in h)))
        ops = np.zeros((p['nprog'], p['nlev'], p['nwarp'], p['slots'], 4), np.int32)
        src = np.zeros((p['nprog'], p['src_slots'], 4), np.int32)
        ph = np.zeros((p['nprog'], 4), np.int32)
        assert _bf.bfFdmtChainQuery(nchan, md, f0, df, -2.0, k, _ip(h), _ip(ops), _ip(src), _ip(ph)) == 0
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


def run_schedule(x, passes, out):
    """x [nchan, ntime] int8/uint8 -> writes out [max_delay, ntime] float32."""
    nchan, ntime = x.shape
    signed = x.dtype == np.int8
    geo = geometry(passes, ntime)
    xi = x.astype(np.int64) + (128 if signed else 0)
    ws_prev, tb_prev = None, 0
    for k, p in enumerate(passes):
        g = geo[k]
        VS = 16 // p['esize']
        final = p['dst_kind'] == 2
        if not final:
            ws = np.full((p['nrow_out'], g['te'] - g['tb']), np.nan if (p['dst_kind'] == 1 or p['esize'] == 4) else POISON,
                         np.float64 if (p['dst_kind'] == 1 or p['esize'] == 4) else np.int64)
        for prog in range(p['nprog']):
            nchan_band, nsrc, staged, _ = (int(v) for v in p['hdr'][prog])
            bias = 128 * nchan_band if signed else 0
            for tile in range(g['nt']):
                t0 = g['tb'] + tile * p['T']
                if p['esize'] == 2:
                    data = np.full(p['smem_elems'], POISON, np.int64)
                else:
                    data = np.full(p['smem_elems'], np.nan, np.float64)
                nbytes = 0
                for e in p['src'][prog][:nsrc]:
                    row, y, z, w = (int(v) for v in e)
                    assert w > 0 and w % VS == 0 and z % VS == 0
                    ts = t0 + y
                    if p['src_kind'] == 0:
                        t = np.arange(ts, ts + w)
                        ok = (t >= 0) & (t < ntime)
                        vals = np.full(w, 128 if signed else 0, np.int64)
                        vals[ok] = xi[row, t[ok]]
                    else:
                        c0 = ts - tb_prev
                        assert c0 >= 0 and c0 % VS == 0 and c0 + w <= ws_prev.shape[1]
                        vals = ws_prev[row, c0:c0 + w]
                    data[z:z + w] = vals
                    nbytes += w * p['esize']
                assert nbytes == staged
                assert p['src'][prog][nsrc][3] == 0 if nsrc < p['src_slots'] else True
                regs = {}
                for lev in range(1, p['nlev'] + 1):
                    for warp in range(p['nwarp']):
                        for op in p['ops'][prog, lev - 1, warp]:
                            dst, a_off, b_off, ctl = (int(v) for v in op)
                            if ctl == 0:
                                break
                            l = ctl & LEVEL_MASK
                            n = (ctl >> 16) * VS
                            assert 1 <= l <= lev and n <= 32 * 3 * VS
                            if ctl & LOADA:
                                assert a_off % VS == 0
                                a = data[a_off:a_off + n].copy()
                            elif ctl & NO_A:
                                a = np.zeros(n, data.dtype)
                            else:
                                a = regs[(warp, l - 1)]
                                assert len(a) == n
                            b = np.zeros(n, data.dtype) if ctl & NO_B else data[b_off:b_off + n]
                            if p['esize'] == 4:
                                r = (a.astype(np.float32) + b.astype(np.float32)).astype(np.float64)
                            else:
                                r = a + b
                            regs[(warp, l)] = r
                            if ctl & STORE_S:
                                assert dst % VS == 0
                                if p['esize'] == 2:
                                    assert (r >= 0).all() and (r < 65536).all()
                                data[dst:dst + n] = r
                            elif ctl & STORE_G:
                                assert n == p['T'] and l == p['nlev']
                                if p['esize'] == 2:
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
                                    out[d, t[ok] - d] = val[ok].astype(np.float32)
                                else:
                                    ws[dst, t0 - g['tb']:t0 - g['tb'] + n] = val
        if not final:
            ws_prev, tb_prev = ws, g['tb']
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
def test_chain_tables_reproduce_the_oracle(nchan, md, f0, df, ntime, dtype):
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


def test_chain_schedule_of_the_baseline_plan():
    """Config 2's plan (4096 chan, max_delay 794): structure + one short gulp."""
    nchan, md, f0, df = 4096, 794, 1000.0, 400. / 4096
    passes = query(nchan, md, f0, df)
    assert [(p['s0'], p['s1'], p['esize']) for p in passes] == [(1, 5, 2), (6, 9, 2), (10, 12, 4)]
    assert passes[0]['src_kind'] == 0 and passes[1]['dst_kind'] == 1 and passes[2]['dst_kind'] == 2
    for p in passes:
        assert p['smem_bytes'] <= 113 * 1024 and p['T'] >= 256
    ntime = 1200
    rng = np.random.default_rng(5)
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    gold = np.zeros((md, ntime), np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    got = np.zeros((md, ntime), np.float32)
    run_schedule(x, passes, got)
    assert np.array_equal(got.view(np.uint32), gold.view(np.uint32))


def test_plans_beyond_exact_fp32_integers_keep_the_float_schedule():
    # 255 * nchan must stay below 2**24 for the integer argument to hold
    assert query(70000, 8, 1000.0, 0.001) == []
    assert query(65536, 8, 1000.0, 0.001) != []
