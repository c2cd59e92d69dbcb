#!/bin/bash
# end-of-round sanity: the things the driver runs (gpu tests, smoke, bench both arms) + durations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -22
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['roofline'].get('traffic_over_algorithmic'),
      d['e2e']['value'], d['cpu_baseline']['value'], d['chain']['value'], d['clocks'], d['parity'])
r = json.loads(open('gpurun_out/bench_final_reference.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['unit'], 'e2e ratio', d['e2e']['value'] / r['value'])
PY
