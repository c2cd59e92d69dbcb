#!/bin/bash
# round-2 GPU check W (2 GPUs): the driver's N = 2 command after the bench.py refactor (shared config object)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_w_2gpu.json 2> gpurun_out/bench_w_2gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_w_2gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'scaling', 'gpu_launches')}, d['comm'], d['parity'], d['e2e']['value'])
print(d['config'])
print(json.dumps(d['fullband'])[:700])
print(d['replicas'])
PY
grep -v '^\[W\|^W0\|\*\*\*\|OMP_NUM\|frame #\|^\s*$' gpurun_out/bench_w_2gpu.err | tail -4 | cut -c1-300
