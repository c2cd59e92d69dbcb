#!/bin/bash
# round-2 GPU check K (2 GPUs): bench.py config 5 under torchrun, NCCL scatter / gather
nvidia-smi --query-gpu=index,name --format=csv,noheader
BENCH_VERBOSE=1 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 5000 gpurun_out/r02_bench_2gpu.json; tail -20 gpurun_out/r02_bench_2gpu.err
