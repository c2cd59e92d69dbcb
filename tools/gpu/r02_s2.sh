#!/bin/bash
# round-2 GPU check S2 (2 GPUs): peer-access phase of the sharded FDMT (own CUDA IPC mappings)
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^\[W\|^W0\|\*\*\*\|OMP_NUM\|frame #\|^\s*$'
echo "== peer access, remote rows by LDG"; timeout -s KILL 300 $W --master-port 29545 tests/fdmt_shard_gpu_worker.py --peer > gpurun_out/r02_peer_ldg.log 2>&1
grep -v "$F" gpurun_out/r02_peer_ldg.log | grep -i "error\|SHARDED\|same_bits\|Traceback\|File \"/\|raise\|illegal" | head -14 | cut -c1-600
echo "== peer access, remote rows by TMA"; BFB_FDMT_PEER_TMA=1 timeout -s KILL 300 $W --master-port 29546 tests/fdmt_shard_gpu_worker.py --peer > gpurun_out/r02_peer_tma.log 2>&1
grep -v "$F" gpurun_out/r02_peer_tma.log | grep -i "error\|SHARDED\|same_bits\|Traceback\|File \"/\|raise\|illegal" | head -14 | cut -c1-600
