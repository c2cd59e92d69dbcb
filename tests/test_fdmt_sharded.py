"""Full-band FDMT across ranks (bfFdmtShard*, bifrost_b200/fdmt_sharded.py;
SURVEY 8f.1).

* one GPU: `nrank` sharded plans run on the same device, one after the other,
  into one workspace and one output -- no exchange is needed then, so this
  checks everything except NCCL: the forced pass boundary, the per-rank
  program lists, the channel offset of the sub-band input, the disjointness of
  the phase-1 rows.  The assembled bank must equal the oracle and
  bfFdmtExecute bit for bit.
* two GPUs (skipped when the box has one): the same through
  torch.distributed/NCCL (tests/fdmt_shard_gpu_worker.py under torchrun), each
  rank holding only its own channels.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import bifrost_b200 as bf
from bifrost_b200.fdmt import Fdmt
from oracle import fdmt as ofdmt

HERE = os.path.dirname(os.path.abspath(__file__))


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


CASES = [
    # nchan, max_delay, f0, df, ntime, nrank, dtype
    (256, 130, 1000.0, 1.5, 3001, 2, np.int8),
    (256, 130, 1000.0, 1.5, 3001, 8, np.uint8),
    (1024, 300, 1000.0, 400. / 1024, 2500, 2, np.int8),     # fp32 exchange rows
    (64, 50, 1200.0, -3.0, 1777, 4, np.int8),               # reversed band: rank 0 holds the LAST input channels
    (4096, 794, 1000.0, 400. / 4096, 2600, 8, np.int8),     # BASELINE config 2's plan, config 5's rank count
    (4096, 794, 1000.0, 400. / 4096, 2600, 2, np.int8),
    (4096, 794, 1000.0, 400. / 4096, 2600, 4, np.int8),     # cut at step 10: passes (1-5)(6-9)(10)(11-12)
]


@pytest.mark.gpu
@pytest.mark.parametrize("nchan,md,f0,df,ntime,nrank,dtype", CASES)
def test_sharded_plans_assemble_the_full_band_bank(nchan, md, f0, df, ntime, nrank, dtype):
    import torch
    rng = np.random.default_rng(nchan + nrank)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max + 1, size=(nchan, ntime)).astype(dtype)
    gold = np.full((md, ntime), -999.0, np.float32)
    ofdmt.fdmt(x, md, f0, df, out=gold)
    # single-plan result (what one GPU computes)
    one = Fdmt()
    one.init(nchan, md, f0, df)
    d_x = bf.asarray(x, space='cuda')
    d_one = bf.asarray(np.full((md, ntime), -999.0, np.float32), space='cuda')
    one.execute(d_x, d_one)
    assert same_bits(np.asarray(d_one.copy('system')), gold)
    # nrank sharded plans on this one device
    cpr = nchan // nrank
    plans = []
    for g in range(nrank):
        p = Fdmt()
        p.init(nchan, md, f0, df)
        p.shard_init(g, nrank)
        plans.append(p)
    t_out = torch.full((md, ntime), -999.0, dtype=torch.float32, device='cuda')
    a_out = bf.ndarray(base=t_out)
    subs = []
    for g in range(nrank):
        c0 = g * cpr if df > 0 else nchan - (g + 1) * cpr
        subs.append(bf.asarray(np.ascontiguousarray(x[c0:c0 + cpr]), space='cuda'))
    need = max(p.shard_workspace_size(subs[g], a_out) for g, p in enumerate(plans))
    ws = torch.empty((need,), dtype=torch.uint8, device='cuda')
    ws.fill_(0xFF)                                     # NaN / 65535: a row nobody produced poisons the result
    lay = plans[0].shard_layout(ntime)
    assert lay['nrank'] == nrank and lay['row_start'][0] == 0 and lay['row_start'][-1] == lay['nrow']
    assert all(p.shard_layout(ntime) == lay for p in plans)
    for g, p in enumerate(plans):
        p.shard_execute(0, subs[g], a_out, ws.data_ptr(), need)
    written = np.zeros(md, np.int32)
    for g, p in enumerate(plans):
        before = t_out.clone()
        p.shard_execute(1, subs[g], a_out, ws.data_ptr(), need)
        torch.cuda.synchronize()
        changed = (before != t_out).any(dim=1).cpu().numpy()
        mine = np.zeros(md, bool)
        for d0, nd, owner in lay['blocks']:
            if owner == g:
                mine[d0:d0 + nd] = True
        assert not (changed & ~mine).any(), "phase 1 wrote a delay block of another rank"
        written += mine
    assert (written == 1).all(), "the delay blocks of the ranks must tile the bank exactly once"
    assert same_bits(t_out.cpu().numpy(), gold)
    # ---- the same without an exchange: every "rank" keeps its own workspace and
    # phase 1 fetches each cut-step row from the workspace of its owner
    # (bfFdmtShardExecutePeers; on several GPUs those are peer mappings)
    wss = []
    for g, p in enumerate(plans):
        w = torch.empty((need,), dtype=torch.uint8, device='cuda')
        w.fill_(0xFF)
        p.shard_execute(0, subs[g], a_out, w.data_ptr(), need)
        wss.append(w)
    ptrs = [w.data_ptr() for w in wss]
    for tma in ('0', '1'):                 # remote rows by plain loads (default) / by TMA like the local ones
        os.environ['BFB_FDMT_PEER_TMA'] = tma
        try:
            t_out.fill_(-999.0)
            for g, p in enumerate(plans):
                p.shard_execute_peers(subs[g], a_out, wss[g].data_ptr(), need, ptrs)
            torch.cuda.synchronize()
        finally:
            del os.environ['BFB_FDMT_PEER_TMA']
        assert same_bits(t_out.cpu().numpy(), gold), tma


@pytest.mark.gpu
def test_shard_init_rejects_what_it_cannot_split():
    p = Fdmt()
    p.init(256, 100, 1000.0, 1.0)
    for bad in (3, 1, 16):
        with pytest.raises(Exception):
            p.shard_init(0, bad)
    with pytest.raises(Exception):
        p.shard_init(2, 2)
    q = Fdmt()
    q.init(100, 37, 400.0, 0.25)                     # 100 channels: no step with 4 equal sub-bands
    with pytest.raises(Exception):
        q.shard_init(0, 4)


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fdmt_over_nccl(world):
    """One process per GPU, NCCL exchange of the cut-step rows, gather on rank 0,
    bank == oracle bit for bit (worker prints the verdict)."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(HERE, 'fdmt_shard_gpu_worker.py')],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'SHARDED_FDMT_GPU_OK' in out.stdout, out.stdout[-2000:]
