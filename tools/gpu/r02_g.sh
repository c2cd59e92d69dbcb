#!/bin/bash
# round-2 GPU check G: FDMT parity with the new knobs, then prefetch / vectors-per-lane sweep
echo "== correlator vs reference library"
timeout -s KILL 300 python -m pytest tests/test_linalg.py -x -q -m gpu -k reference_library 2>&1 | tail -30
echo "== spectrometer"
timeout -s KILL 600 python -m pytest tests/test_spectrometer.py tests/test_blocks_gpu.py -q -m gpu 2>&1 | tail -5
echo "== FDMT parity"
timeout -s KILL 1500 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -6
echo "== timing"
timeout -s KILL 900 python tools/fdmt_time.py --check "" "BFB_FDMT_PACKED_PREFETCH=0,0,0" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110" \
  "BFB_FDMT_PACKED_LV=3,5,3 BFB_FDMT_PACKED_SMEM_KB=74,110,74" \
  "BFB_FDMT_PACKED_LV=3,3,5 BFB_FDMT_PACKED_SMEM_KB=74,74,110" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110 BFB_FDMT_PACKED_WAVES=8" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110 BFB_FDMT_PACKED_PREFETCH=1,1,1" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,200,200" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,200,200 BFB_FDMT_PACKED_PREFETCH=1,1,1" \
  "BFB_FDMT_PACKED_PREFETCH=1,1,1 BFB_FDMT_PACKED_SMEM_KB=74,110,110" \
  "BFB_FDMT_PACKED_SMEM_KB=74,110,110" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110 BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=2560 BFB_FDMT_PACKED_LAG=3" \
  "BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110 BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=2560 BFB_FDMT_PACKED_LAG=4" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=2048 BFB_FDMT_PACKED_LAG=3" \
  > gpurun_out/r02_fdmt_time6.jsonl 2>gpurun_out/r02_fdmt_time6.err
cat gpurun_out/r02_fdmt_time6.jsonl; tail -5 gpurun_out/r02_fdmt_time6.err
echo "== launch lists (LV 3,5,5)"
BFB_FDMT_PACKED_LV=3,5,5 BFB_FDMT_PACKED_SMEM_KB=74,110,110 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 12 --csv --log-file gpurun_out/r02_packed_launches3.csv python tools/fdmt_time.py --nrep 1 "" > /dev/null 2>&1
grep fdmt gpurun_out/r02_packed_launches3.csv | tail -15 | cut -d, -f5,13-
