#!/bin/bash
# round-2 GPU check A: FDMT parity (all schedules), old vs chain timing, launch list
timeout 900 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/fdmt_time.py --check "" "BFB_FDMT_CHAIN=0" > gpurun_out/r02_fdmt_time1.jsonl 2>gpurun_out/r02_fdmt_time1.err
cat gpurun_out/r02_fdmt_time1.jsonl; tail -5 gpurun_out/r02_fdmt_time1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_chain_launches.csv python tools/fdmt_time.py --nrep 2 "" > /dev/null 2>&1
grep -c fdmt gpurun_out/r02_chain_launches.csv
