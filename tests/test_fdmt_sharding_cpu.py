"""Groundwork for the cross-GPU full-band FDMT, on the
CPU: the tile-pass tables that bfFdmtExecute already uses decompose by program
ownership.  With R ranks owning the R sub-bands of the step where the merge tree
has R bands,
  * every program of the passes up to that step reads and writes rows of ONE
    rank's sub-tree only (no exchange), and
  * the final pass, with its delay-block programs dealt out to the ranks, reads
    the step rows of all ranks (the only cross-rank traffic) and reproduces the
    single-GPU full-band transform bit for bit.
The numpy interpreter of tests/test_fdmt_tiles_cpu.py stands in for the kernel."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_fdmt_tiles_cpu import query, oracle_states, F32  # noqa: E402


def band_of(plan, step, row):
    return int(np.searchsorted(plan.row_offsets[step], row, side='right') - 1)


def programs(tp):
    """(rows staged, rows written) per program, from the item tables."""
    out = []
    for prog in range(tp['nprog']):
        st = tp['items'][prog, 0].reshape(-1, 4)
        staged = sorted(int(r[0]) for r in st if r[3] > 0)
        last = tp['items'][prog, tp['nphase'] - 1].reshape(-1, 4)
        written = sorted(int(r[0]) for r in last if (r[3] & 0xFFFF) > 0)
        out.append((staged, written))
    return out


@pytest.mark.parametrize("nchan,md,nrank", [(256, 120, 4), (128, 90, 2), (256, 200, 8)])
def test_tile_passes_decompose_by_sub_band_ownership(nchan, md, nrank):
    f0, df, ntime = 1000., 400. / nchan, 900
    rng = np.random.default_rng(nchan + nrank)
    x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    plan, states = oracle_states(x, md, f0, df)
    sx = int(np.log2(nchan // nrank))                    # step with `nrank` sub-bands
    assert len(plan.row_offsets[sx]) - 1 == nrank
    chan_per_rank = nchan // nrank

    # ---- pass "local": steps 2..sx (step 1 comes from the raw pass; same ownership argument)
    tp = query(nchan, md, f0, df, 2, sx, 8, 4, raw=False)
    assert tp is not None
    for staged, written in programs(tp):
        owner = {band_of(plan, sx, r) for r in written}
        assert len(owner) == 1                            # a program belongs to one sub-band ...
        g = owner.pop()
        # ... and every row it stages (step 1) lies in that rank's channel range
        lo = plan.row_offsets[1][g * chan_per_rank // 2]
        hi = plan.row_offsets[1][(g + 1) * chan_per_rank // 2]
        assert all(lo <= r < hi for r in staged)

    # ---- final pass: steps sx+1 .. last, programs dealt out round-robin
    last = plan.nstep - 1
    tpf = query(nchan, md, f0, df, sx + 1, last, 8, 4, raw=False)
    assert tpf is not None
    progs = programs(tpf)
    remote_rows, local_rows = 0, 0
    covered = set()
    for p, (staged, written) in enumerate(progs):
        g = p % nrank                                     # the rank that runs this program
        for r in staged:
            if band_of(plan, sx, r) == g:
                local_rows += 1
            else:
                remote_rows += 1
        covered.update(written)
    assert covered == set(range(plan.nrow[last]))         # together the ranks produce every output row
    assert remote_rows > 0 and local_rows > 0
    # the exchange volume is bounded by the step-sx state (times the window redundancy)
    assert remote_rows + local_rows <= 3 * plan.nrow[sx] * max(1, tpf['nprog'] // 8 + 1)


def test_sharded_final_pass_equals_full_band_transform():
    """Runs the final pass program by program from per-rank copies of the step-sx
    rows (each rank holds ONLY its own sub-band; a read outside it is served from
    the owner) and compares with the oracle's last state."""
    from test_fdmt_tiles_cpu import run_pass
    nchan, md, nrank, ntime = 128, 90, 4, 700
    f0, df = 1200., 300. / nchan
    rng = np.random.default_rng(77)
    x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    plan, states = oracle_states(x, md, f0, df)
    sx = int(np.log2(nchan // nrank))
    last = plan.nstep - 1
    ro = plan.row_offsets[sx]
    # per-rank HBM: rows of the own sub-band, NaN elsewhere (a wrong-owner read would poison the result)
    shards = []
    for g in range(nrank):
        s = np.full_like(states[sx], np.nan)
        s[ro[g]:ro[g + 1]] = states[sx][ro[g]:ro[g + 1]]
        shards.append(s)
    gathered = np.full_like(states[sx], np.nan)          # what peer reads assemble, row by owner
    for g in range(nrank):
        gathered[ro[g]:ro[g + 1]] = shards[g][ro[g]:ro[g + 1]]
    tpf = query(nchan, md, f0, df, sx + 1, last, 8, 4, raw=False)
    ntile = -(-ntime // tpf['T'])
    written = run_pass(tpf, gathered, None, ntime, range(ntile))
    for row, cells in written.items():
        ts = np.array(sorted(cells))
        got = np.array([cells[t] for t in ts], F32)
        np.testing.assert_array_equal(got.view(np.uint32), states[last][row][ts].view(np.uint32))
    assert len(written) == plan.nrow[last]


def test_two_rank_gloo_exchange():
    """The same decomposition as two gloo processes: local passes on the own
    sub-band, all-gather of the split-step rows (the future NVLink peer reads),
    each rank's share of the final programs, assembled bank == oracle."""
    import socket
    import subprocess
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fdmt_shard_worker.py')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), worker],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'SHARDED_FDMT_OK' in out.stdout


def test_two_rank_gloo_packed_schedule():
    """The schedule bfFdmtShardInit builds (packed-integer passes cut at the
    step with `world` sub-bands), run by two gloo ranks with the numpy
    interpreter: local programs touch one rank's sub-tree only, the cut-step
    rows form one block per rank, and the ranks' delay blocks assemble the
    oracle's bank bit for bit."""
    import socket
    import subprocess
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fdmt_packed_shard_worker.py')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), worker],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'SHARDED_PACKED_OK' in out.stdout, out.stdout[-2000:]
