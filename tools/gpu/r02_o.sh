#!/bin/bash
# round-2 GPU check O (2 GPUs): sharded full-band FDMT over NCCL, bench.py config 5
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 600 python -m pytest tests/test_fdmt_sharded.py -x -q -m gpu -k "nccl" 2>&1 | tail -8
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/fdmt_shard_gpu_worker.py 2>&1 | grep -v "^\[W\|^W0\|\*\*\*" | tail -6
BENCH_VERBOSE=1 timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err
