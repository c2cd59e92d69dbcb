#!/bin/bash
# end-of-round sanity: the three things the driver runs (gpu tests, smoke, bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['chain']['value'], d['clocks'])
PY
