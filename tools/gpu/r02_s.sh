#!/bin/bash
# round-2 GPU check S (2 GPUs): sharded full-band FDMT over NCCL and over peer memory, bench.py config 5
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -6
timeout -s KILL 300 python -m pytest tests/test_fdmt_sharded.py -x -q -m gpu -k "assemble and 256" 2>&1 | tail -4
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/fdmt_shard_gpu_worker.py 2>&1 | grep -v "^\[W\|^W0\|\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/r02_sharded_2gpu.txt
BENCH_VERBOSE=1 timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench rc=$?"; tail -c 4500 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err
