#!/bin/bash
# round-2 GPU check I: whole GPU suite, smoke(), bench.py at N=1 (both arms), launch list under ncu
echo "== full GPU suite"
timeout -s KILL 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== smoke"
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1"
BENCH_VERBOSE=1 timeout -s KILL 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 6000 gpurun_out/r02_bench.json; tail -15 gpurun_out/r02_bench.err
echo "== bench reference arm"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>&1; tail -c 1500 gpurun_out/r02_bench_reference_arm.json
echo "== launch list of the bench command"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic --no-gpu-reference > gpurun_out/bench_under_ncu.log 2>&1
grep -c "fdmt" gpurun_out/r02_bench_launches.csv
