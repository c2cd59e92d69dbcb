"""Worker of tests/test_fdmt_sharding_cpu.py::test_two_rank_gloo_exchange: run under
torch.distributed.run (gloo).  Each rank owns the channels of its sub-band,
produces that sub-band's rows of the split step with the tile-pass tables
(numpy interpreter standing in for the kernels), all-gathers the split-step rows
(the traffic that becomes NVLink peer reads on the GPUs), runs ITS share of the
final pass's programs and sends its output rows to rank 0, which compares the
assembled dispersion bank with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_fdmt_tiles_cpu import query, oracle_states, run_pass, F32  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    nchan, md, ntime = 128, 90, 700
    f0, df = 1200., 300. / nchan
    x = np.random.default_rng(77).integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    plan, states = oracle_states(x, md, f0, df)          # rank 0 uses it as the arbiter only
    sx = int(np.log2(nchan // world))
    last = plan.nstep - 1
    ro = plan.row_offsets[sx]

    # ---- local part: steps 2..sx restricted to the programs of the own sub-band.
    # (step 1 comes from the raw pass on the GPU; here the oracle's step-1 rows of the
    # rank's own channels stand in for it -- rows of other ranks are NaN)
    c_lo, c_hi = rank * nchan // world, (rank + 1) * nchan // world
    own1 = np.full_like(states[1], np.nan)
    r1 = plan.row_offsets[1]
    own1[r1[c_lo // 2]:r1[c_hi // 2]] = states[1][r1[c_lo // 2]:r1[c_hi // 2]]
    tp = query(nchan, md, f0, df, 2, sx, 8, 4, raw=False)
    ntile = -(-ntime // tp['T'])
    mine = dict(tp)
    keep = [p for p in range(tp['nprog'])
            if ro[rank] <= int(tp['items'][p, tp['nphase'] - 1].reshape(-1, 4)[0][0]) < ro[rank + 1]]
    mine['items'], mine['nprog'] = tp['items'][keep], len(keep)
    local = run_pass(mine, own1, None, ntime, range(ntile))
    shard = np.zeros((ro[rank + 1] - ro[rank], ntime), F32)
    for row, cells in local.items():
        for t, v in cells.items():
            shard[row - ro[rank], t] = v
    assert len(local) == ro[rank + 1] - ro[rank]

    # ---- exchange: every rank needs (parts of) every sub-band for its delay blocks
    nmax = max(ro[g + 1] - ro[g] for g in range(world))
    buf = torch.zeros((nmax, ntime), dtype=torch.float32)
    buf[:shard.shape[0]] = torch.from_numpy(shard)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    full = np.concatenate([parts[g].numpy()[:ro[g + 1] - ro[g]] for g in range(world)], 0)

    # ---- final pass: programs dealt out round-robin
    tpf = query(nchan, md, f0, df, sx + 1, last, 8, 4, raw=False)
    minef = dict(tpf)
    keepf = [p for p in range(tpf['nprog']) if p % world == rank]
    minef['items'], minef['nprog'] = tpf['items'][keepf], len(keepf)
    ntilef = -(-ntime // tpf['T'])
    outrows = run_pass(minef, full, None, ntime, range(ntilef))
    out = np.full((plan.nrow[last], ntime), np.nan, F32)
    for row, cells in outrows.items():
        for t, v in cells.items():
            out[row, t] = v
    gathered = [torch.zeros((plan.nrow[last], ntime), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))
    if rank == 0:
        # a rank leaves NaN in the rows it does not own; genuine NaNs (t < delay cells of the
        # reference's step-0 rows) are NaN in every copy, so "first non-NaN" assembles the bank
        bank = np.full_like(out, np.nan)
        for g in range(world):
            bank = np.where(np.isnan(bank), gathered[g].numpy(), bank)
        ok = np.array_equal(bank.view(np.uint32), states[last].view(np.uint32))
        print('SHARDED_FDMT_OK' if ok else 'SHARDED_FDMT_MISMATCH', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
