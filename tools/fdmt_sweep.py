"""Sweeps the FDMT tunables (environment variables read at bfFdmtInit) on the
bench workload in one process.  Prints ms per gulp for each setting."""
import itertools
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import bifrost_b200 as bf

w = bench.workload(0)
x = bench.make_input(w, 1)
d_in = bf.asarray(x, space='cuda')
d_out = bf.zeros((w['max_delay'], w['ntime']), 'f32', 'cuda')
stream = torch.cuda.current_stream()
bf.device.set_stream(stream.cuda_stream)


def run(env):
    for k in list(os.environ):
        if k.startswith('BFB_FDMT_'):
            del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    plan = bf.fdmt.Fdmt()
    plan.init(w['nchan'], w['max_delay'], w['f0'], w['df'])
    for _ in range(3):
        plan.execute(d_in, d_out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); plan.execute(d_in, d_out); e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


base = run({})
ref = d_out.copy(space='system') if hasattr(d_out, 'copy') else None
print(json.dumps(dict(env={}, ms=round(base, 4))), flush=True)
grid = []
for cap in (74, 110):
    for d in (12, 16, 20, 24, 32):
        for split in ('', '8'):
            e = dict(BFB_FDMT_TILE_SMEM_KB=cap, BFB_FDMT_TILE_D=d)
            if split:
                e['BFB_FDMT_SPLIT'] = split
            grid.append(e)
for extra in (dict(BFB_FDMT_K=4, BFB_FDMT_TILE_SMEM_KB=74), dict(BFB_FDMT_K=4, BFB_FDMT_SPLIT='8', BFB_FDMT_TILE_SMEM_KB=74),
              dict(BFB_FDMT_TILE_THREADS=192, BFB_FDMT_TILE_SMEM_KB=74), dict(BFB_FDMT_TILE_THREADS=192, BFB_FDMT_TILE_SMEM_KB=55, BFB_FDMT_TILE_D=12)):
    grid.append(dict(extra))
res = []
for e in grid:
    try:
        ms = run(e)
    except Exception as ex:
        ms = None
    res.append((ms, e))
    print(json.dumps(dict(env=e, ms=None if ms is None else round(ms, 4))), flush=True)
res = [r for r in res if r[0] is not None]
res.sort(key=lambda r: r[0])
print('BEST', json.dumps([dict(ms=round(m, 4), env=e) for m, e in res[:8]]))
