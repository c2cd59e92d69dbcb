#!/bin/bash
# round-2 GPU check P (1 GPU): the whole -m gpu suite, N=1 bench record, ncu capture of the FDMT passes
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed -s 9 -c 3 -f -o gpurun_out/r02_packed_prof4 python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_packed_prof4.log 2>&1
ls -la gpurun_out/r02_packed_prof4.ncu-rep
