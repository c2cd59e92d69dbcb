#!/bin/bash
# round-2 GPU check V (1 GPU, last call of the round): the tests of what changed since check U
# (native rings in device memory, C client, executor logs / frame-axis views, reference-size FFTs,
# sigproc sink), smoke, the bench line with the whole-gulp CPU sample, then as much of the rest of
# the GPU suite as the remaining time allows.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T0=$(date +%s)
grep MemAvailable /proc/meminfo; nproc
timeout -s KILL 200 python -m pytest tests/test_ring.py tests/test_cabi.py tests/test_blocks_gpu.py tests/test_io_formats.py tests/test_overlay.py -m gpu -x -q 2>&1 | tail -6
echo "== changed-path tests done at $(( $(date +%s) - T0 )) s"
timeout -s KILL 240 python -m pytest tests/test_fft.py -m gpu -x -q --durations=6 2>&1 | tail -14
echo "== fft tests done at $(( $(date +%s) - T0 )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_v_reference.json 2>> gpurun_out/bench_v.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['roofline'].get('traffic_over_algorithmic'),
      d['e2e']['value'], d['cpu_baseline'], d['chain']['value'], d['clocks'], d['parity'], d.get('host_numa'))
r = json.loads(open('gpurun_out/bench_v_reference.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['ms_per_step'], r['cpu_baseline']['sample'], 'same config:', r['config'] == d['config'],
      'e2e ratio', d['e2e']['value'] / r['value'])
PY
tail -3 gpurun_out/bench_v.err
echo "== bench done at $(( $(date +%s) - T0 )) s"
LEFT=$(( 470 - ( $(date +%s) - T0 ) ))
if [ $LEFT -gt 40 ]; then
  timeout -s KILL $LEFT python -m pytest tests -m gpu -x -q --deselect tests/test_fft.py --ignore=tests/test_ring.py --ignore=tests/test_cabi.py --ignore=tests/test_blocks_gpu.py --ignore=tests/test_io_formats.py --ignore=tests/test_overlay.py --ignore=tests/test_fft.py 2>&1 | tail -5
fi
echo "== all done at $(( $(date +%s) - T0 )) s"
