#!/bin/bash
# round-2 GPU check J: micro-optimised op loop -- parity, timing, profile for profiles/
timeout -s KILL 1500 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 600 python tools/fdmt_time.py --check "" "BFB_FDMT_PACKED_WAVES=4" "BFB_FDMT_PACKED_WAVES=16" \
  "BFB_FDMT_PACKED_SMEM_KB=74,74,74" "BFB_FDMT_PACKED_MEGA=1" "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_SMEM_KB=74,74,74" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_SMEM_KB=74,74,74 BFB_FDMT_PACKED_LAG=3" \
  "BFB_FDMT_PACKED=0" > gpurun_out/r02_fdmt_time8.jsonl 2>gpurun_out/r02_fdmt_time8.err
cat gpurun_out/r02_fdmt_time8.jsonl; tail -3 gpurun_out/r02_fdmt_time8.err
timeout -s KILL 300 python tools/fdmt_time.py --md 204 "" "BFB_FDMT_PACKED=0" 2>&1 | tail -2
timeout -s KILL 300 python tools/fdmt_time.py --md 1621 --f0 1200 --bw 300 "" "BFB_FDMT_PACKED=0" 2>&1 | tail -2
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed -s 9 -c 3 -f -o gpurun_out/r02_packed_prof3 python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_packed_prof3.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:fdmt_packed_mega -s 3 -c 1 -f -o gpurun_out/r02_mega_prof3 python tools/fdmt_time.py --nrep 2 "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_SMEM_KB=74,74,74" > gpurun_out/r02_mega_prof3.log 2>&1
ls -la gpurun_out/*prof3.ncu-rep
