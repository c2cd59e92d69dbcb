"""Worker of tests/test_fdmt_sharding_cpu.py::test_two_rank_gloo_packed_schedule:
the SHARDED packed-integer schedule (bfFdmtShardInit's tables, obtained on the
CPU through BFB_FDMT_PACKED_FORCE_END) executed by two gloo processes with the
numpy interpreter of tests/test_fdmt_packed_cpu.py standing in for the kernels.

Rank g runs only the programs of its own sub-tree in the passes up to the cut,
the cut-step workspace rows are exchanged (all_gather: the NCCL broadcasts of
bifrost_b200/fdmt_sharded.py), every rank runs its share (p % world) of the
last pass's programs, rank 0 assembles the bank and compares it with the oracle
bit for bit.  A row read before its owner produced it is POISON / NaN and
shows up as a mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_fdmt_packed_cpu import query, geometry, Machine, POISON, STORE_G  # noqa: E402
from oracle import fdmt as ofdmt  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    nchan, md, f0, df, ntime = 256, 130, 1000.0, 1.5, 1100
    sx = int(np.log2(nchan // world))
    os.environ['BFB_FDMT_PACKED_FORCE_END'] = str(sx)
    passes = query(nchan, md, f0, df)
    assert passes[-2]['s1'] == sx and passes[-1]['s0'] == sx + 1
    x = np.random.default_rng(9).integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    xi = x.astype(np.int64) + 128
    cpr = nchan // world
    geo = geometry(passes, ntime)
    npass = len(passes)

    # ---- ownership: pass 0 by the input channels a program stages, later local
    # passes by the owner of the workspace rows they stage
    row_owner = None
    prog_owner = []
    for k, p in enumerate(passes[:-1]):
        owners = []
        new_row_owner = {}
        for prog in range(p['nprog']):
            nsrc = int(p['hdr'][prog][1])
            rows = [int(r) for r in p['src'][prog][:nsrc, 0]]
            own = {r // cpr for r in rows} if k == 0 else {row_owner[r] for r in rows}
            assert len(own) == 1, "a program of a local pass must stay inside one rank's sub-tree"
            o = own.pop()
            owners.append(o)
            top = p['ops'][prog, p['nlev'] - 1].reshape(-1, 4)
            for op in top:
                if int(op[3]) & STORE_G:
                    new_row_owner[int(op[0])] = o
        prog_owner.append(owners)
        row_owner = new_row_owner
    # the rows of the cut step lie rank by rank, in order (the exchange is one block per rank)
    cut_rows = sorted(row_owner)
    assert [row_owner[r] for r in cut_rows] == sorted(row_owner[r] for r in cut_rows)

    # ---- local passes on the own programs only
    ws_prev, tb_prev, width_prev = None, 0, 0
    out = np.full((md, ntime), np.nan, np.float32)
    for k, p in enumerate(passes[:-1]):
        g = geo[k]
        width = g['te'] - g['tb']
        is_float = p['dst_kind'] == 1 or p['esize'] == 4
        ws = np.full((p['nrow_out'], width), np.nan if is_float else POISON, np.float64 if is_float else np.int64)
        m = Machine(p, g, x, xi, True, ws_prev, tb_prev, max(width_prev, 1), ws, max(width, 1), out)
        m.ws_prev_width = width_prev
        for prog in range(p['nprog']):
            if prog_owner[k][prog] != rank:
                continue
            for tile in range(g['nt']):
                m.run(prog, tile)
        ws_prev, tb_prev, width_prev = ws, g['tb'], width

    # ---- exchange of the cut-step rows: every rank contributes the rows it owns
    mine = np.array([row_owner[r] == rank for r in range(ws_prev.shape[0])])
    send = torch.from_numpy(np.where(mine[:, None], ws_prev, 0).astype(np.float64))
    parts = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    full = np.zeros_like(ws_prev)
    for gq in range(world):
        sel = np.array([row_owner[r] == gq for r in range(ws_prev.shape[0])])
        full[sel] = parts[gq].numpy()[sel].astype(ws_prev.dtype)

    # ---- last pass: programs dealt out round-robin
    p, g = passes[-1], geo[-1]
    m = Machine(p, g, x, xi, True, full, tb_prev, max(width_prev, 1), None, 1, out)
    m.ws_prev_width = width_prev
    for prog in range(p['nprog']):
        if prog % world != rank:
            continue
        for tile in range(g['nt']):
            m.run(prog, tile)
    gathered = [torch.zeros((md, ntime), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))
    if rank == 0:
        bank = np.full((md, ntime), np.nan, np.float32)
        nwriters = np.zeros((md, ntime), np.int32)
        for gq in range(world):
            part = gathered[gq].numpy()
            have = ~np.isnan(part)
            nwriters += have
            bank = np.where(have, part, bank)
        gold = np.full((md, ntime), np.nan, np.float32)
        ofdmt.fdmt(x, md, f0, df, out=gold)
        written = ~np.isnan(gold)
        ok = (nwriters[written] == 1).all() and np.array_equal(bank[written].view(np.uint32), gold[written].view(np.uint32))
        print('SHARDED_PACKED_OK' if ok else 'SHARDED_PACKED_MISMATCH', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
