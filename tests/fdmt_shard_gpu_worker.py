"""Worker of tests/test_fdmt_sharded.py::test_sharded_fdmt_over_nccl (run under
torch.distributed.run, one process per GPU, NCCL).  Every rank holds ONLY its
own channels of one synthetic gulp; ShardedFdmt runs phase 0, the NCCL exchange
of the cut-step rows, phase 1 and the gather; rank 0 compares the assembled
bank with the oracle (and with windows of the C oracle for the BASELINE-sized
case) bit for bit and prints the verdict, the exchange volume and timings."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bifrost_b200 as bf  # noqa: E402
from bifrost_b200.fdmt_sharded import ShardedFdmt  # noqa: E402
from oracle import fdmt as ofdmt  # noqa: E402


def main():
    with_peer = '--peer' in sys.argv
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    bf.device.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    report = []
    for nchan, md, f0, df, ntime in [(256, 130, 1000.0, 1.5, 3001),
                                     (512, 200, 1200.0, -0.5, 2222),
                                     (4096, 794, 1000.0, 400. / 4096, 20000)]:
        if nchan // world < 2:
            continue
        x = np.random.default_rng(nchan).integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
        sf = ShardedFdmt().init(nchan, md, f0, df)
        c0, c1 = sf.channels(df > 0)
        x_local = torch.from_numpy(np.ascontiguousarray(x[c0:c1])).cuda()      # the rank never sees the other channels
        out = torch.full((md, ntime), -999.0, dtype=torch.float32, device='cuda')
        sf.execute(x_local, out, gather_to=0)
        torch.cuda.synchronize()
        # timing of the three parts (device events, max over ranks)
        a_in, a_out = bf.ndarray(base=x_local), bf.ndarray(base=out)
        ws, need = sf._workspace(a_in, a_out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        sf.plan.shard_execute(0, a_in, a_out, ws.data_ptr(), need)
        ev[1].record()
        nbytes = sf.exchange(ws, ntime)
        ev[2].record()
        sf.plan.shard_execute(1, a_in, a_out, ws.data_ptr(), need)
        ev[3].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the same without an exchange: phase 1 reads the peers' rows in place (NVLink)
        out_p = torch.full((md, ntime), -999.0, dtype=torch.float32, device='cuda')
        tp = torch.zeros(1, device='cuda')
        if with_peer:
            sf.execute(x_local, out_p, gather_to=0, peer=True)
            torch.cuda.synchronize()
            evp = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            dist.barrier()
            torch.cuda.synchronize()
            evp[0].record()
            sf.execute(x_local, out_p, peer=True)
            evp[1].record()
            torch.cuda.synchronize()
            tp = torch.tensor([evp[0].elapsed_time(evp[1])], device='cuda')
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        if rank == 0:
            gold = np.full((md, ntime), -999.0, np.float32)
            ofdmt.fdmt(x, md, f0, df, out=gold)
            got = out.cpu().numpy()
            same = np.array_equal(got.view(np.uint32), gold.view(np.uint32))
            same_peer = (not with_peer) or np.array_equal(out_p.cpu().numpy().view(np.uint32), gold.view(np.uint32))
            ok = ok and same and same_peer
            report.append(dict(nchan=nchan, max_delay=md, ntime=ntime, world=world, same_bits=bool(same),
                               same_bits_peer_access=bool(same_peer) if with_peer else None,
                               peer_mode=('tma' if os.environ.get('BFB_FDMT_PEER_TMA') == '1' else 'ldg') if with_peer else None,
                               ms_whole_gulp_peer_access=float(tp[0]) if with_peer else None,
                               exchange_bytes_received_per_rank=int(nbytes),
                               ms_phase0=float(t[0]), ms_exchange=float(t[1]), ms_phase1=float(t[2])))
    if rank == 0:
        for r in report:
            print(json.dumps(r), flush=True)
        print('SHARDED_FDMT_GPU_OK' if ok and report else 'SHARDED_FDMT_GPU_MISMATCH', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
