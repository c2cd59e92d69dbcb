#!/bin/bash
# round-2 GPU check S3 (2 GPUs): bench.py config 5 incl. the full-band transform over peer memory
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^\[W\|^W0\|\*\*\*\|OMP_NUM\|frame #\|^\s*$'
BENCH_VERBOSE=1 timeout -s KILL 420 $W --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --fullband-peer > gpurun_out/r02_bench_2gpu_peer.json 2> gpurun_out/r02_bench_2gpu_peer.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu_peer.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['fullband'])[:1500])
PY
grep -v "$F" gpurun_out/r02_bench_2gpu_peer.err | tail -3 | cut -c1-300
BFB_FDMT_PEER_TMA=0 BENCH_VERBOSE=1 timeout -s KILL 420 $W --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --fullband-peer > gpurun_out/r02_bench_2gpu_peer_ldg.json 2> /dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu_peer_ldg.json').read().strip().splitlines()[-1])
print('LDG mode:', json.dumps(d['fullband']['peer_access'])[:600])
PY
