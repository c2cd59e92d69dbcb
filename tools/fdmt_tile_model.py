"""Planner prototype: smem footprint / redundancy of (band, delay-block, time-tile)
FDMT tiles spanning levels s0..s1 (input = level s0-1 state).  Analysis only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import fdmt as o

w = bench.workload(0)
p = o.FdmtPlan(w['nchan'], w['max_delay'], w['f0'], w['df'])

def band_of(step, row):
    ro = p.row_offsets[step]
    return int(np.searchsorted(ro, row, side='right') - 1)

def analyse(s0, s1, D, T, per_row=True, verbose=False):
    """returns (max smem floats for ping-pong, staged input floats per tile-column, compute rows)"""
    ro1 = p.row_offsets[s1]
    tot_in = 0; tot_comp = 0; max_smem = 0; ntile = 0
    for B in range(len(ro1) - 1):
        nd = ro1[B + 1] - ro1[B]
        for d_lo in range(0, nd, D):
            d_hi = min(nd, d_lo + D)
            # need[level] = dict row -> [smin, smax]
            cur = {ro1[B] + d: [0, 0] for d in range(d_lo, d_hi)}
            foot = []
            def footprint(need):
                if per_row:
                    return sum(T + (b - a) for a, b in need.values())
                lo = min(a for a, b in need.values()); hi = max(b for a, b in need.values())
                return len(need) * (T + hi - lo)
            foot.append(footprint(cur))
            comp = 0
            for s in range(s1, s0 - 1, -1):
                comp += footprint(cur)
                nxt = {}
                src = p.srcrows[s]; dly = p.delays[s]
                for r, (a, b) in cur.items():
                    for k, sh in ((0, 0), (1, int(dly[r]))):
                        q = int(src[r, k])
                        if q < 0: continue
                        e = nxt.get(q)
                        if e is None: nxt[q] = [a + sh, b + sh]
                        else:
                            e[0] = min(e[0], a + sh); e[1] = max(e[1], b + sh)
                cur = nxt
                foot.append(footprint(cur))
            smem = max(foot[i] + foot[i + 1] for i in range(len(foot) - 1))
            max_smem = max(max_smem, smem)
            tot_in += foot[-1]; tot_comp += comp; ntile += 1
    uniq_in = p.nrow[s0 - 1] * T
    uniq_comp = sum(p.nrow[s] for s in range(s0, s1 + 1)) * T
    return dict(s0=s0, s1=s1, D=D, T=T, ntile=ntile, smem_KB=max_smem * 4 / 1024,
                in_redund=tot_in / uniq_in, comp_redund=tot_comp / uniq_comp)

if __name__ == '__main__':
    for (s0, s1) in [(6, 8), (6, 9), (9, 12), (10, 12), (6, 12), (7, 9), (8, 10), (11, 12), (6,7),(8,9),(10,11)]:
        for D in (8, 16, 32, 64):
            for T in (128, 256, 512):
                r = analyse(s0, s1, D, T)
                if r['smem_KB'] <= 220:
                    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
