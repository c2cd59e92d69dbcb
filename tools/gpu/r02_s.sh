#!/bin/bash
# round-2 GPU check S (2 GPUs): sharded full-band FDMT over NCCL and over peer memory, bench.py config 5
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 300 python -m pytest tests/test_fdmt_sharded.py -x -q -m gpu -k "assemble and 256" 2>&1 | tail -3
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^\[W\|^W0\|\*\*\*\|OMP_NUM\|frame #\|^\s*$'
echo "== NCCL exchange"; timeout -s KILL 300 $W --master-port 29544 tests/fdmt_shard_gpu_worker.py 2>&1 | grep -v "$F" | tail -5 | cut -c1-700 | tee gpurun_out/r02_sharded_2gpu.txt
echo "== peer access, remote rows by LDG"; timeout -s KILL 300 $W --master-port 29545 tests/fdmt_shard_gpu_worker.py --peer 2>&1 | grep -v "$F" | tail -5 | cut -c1-700 | tee -a gpurun_out/r02_sharded_2gpu.txt
echo "== peer access, remote rows by TMA"; BFB_FDMT_PEER_TMA=1 timeout -s KILL 300 $W --master-port 29546 tests/fdmt_shard_gpu_worker.py --peer 2>&1 | grep -v "$F" | tail -5 | cut -c1-700 | tee -a gpurun_out/r02_sharded_2gpu.txt
BENCH_VERBOSE=1 timeout -s KILL 420 $W --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench rc=$?"; tail -c 3500 gpurun_out/r02_bench_2gpu.json; grep -v "$F" gpurun_out/r02_bench_2gpu.err | tail -4 | cut -c1-300
