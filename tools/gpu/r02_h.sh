#!/bin/bash
# round-2 GPU check H: persistent kernel with producers first in each round: lag / ring / chunk sweep
timeout -s KILL 300 python -m pytest tests/test_linalg.py -q -m gpu -k reference_library 2>&1 | tail -3
K="BFB_FDMT_PACKED_MEGA=1"
timeout -s KILL 900 python tools/fdmt_time.py --check "BFB_FDMT_PACKED_SMEM_KB=74,110,110" \
  "$K" "$K BFB_FDMT_PACKED_RING_EXTRA=1" "$K BFB_FDMT_PACKED_RING_EXTRA=2" "$K BFB_FDMT_PACKED_RING_EXTRA=3" \
  "$K BFB_FDMT_PACKED_LAG=3" "$K BFB_FDMT_PACKED_LAG=3 BFB_FDMT_PACKED_RING_EXTRA=1" "$K BFB_FDMT_PACKED_LAG=3 BFB_FDMT_PACKED_RING_EXTRA=2" \
  "$K BFB_FDMT_PACKED_CHUNK=1536 BFB_FDMT_PACKED_RING_EXTRA=2" "$K BFB_FDMT_PACKED_CHUNK=1536 BFB_FDMT_PACKED_LAG=3 BFB_FDMT_PACKED_RING_EXTRA=2" \
  "$K BFB_FDMT_PACKED_CHUNK=3072 BFB_FDMT_PACKED_RING_EXTRA=1" "$K BFB_FDMT_PACKED_CHUNK=3072 BFB_FDMT_PACKED_RING_EXTRA=2" \
  "$K BFB_FDMT_PACKED_CHUNK=4096 BFB_FDMT_PACKED_RING_EXTRA=1" \
  > gpurun_out/r02_fdmt_time7.jsonl 2>gpurun_out/r02_fdmt_time7.err
cat gpurun_out/r02_fdmt_time7.jsonl; tail -5 gpurun_out/r02_fdmt_time7.err
echo "== traffic of the best two"
for knobs in "BFB_FDMT_PACKED_RING_EXTRA=2" "BFB_FDMT_PACKED_LAG=3 BFB_FDMT_PACKED_RING_EXTRA=2"; do
  env $K $knobs timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 4 --csv --log-file gpurun_out/r02_mega_launches_tmp.csv python tools/fdmt_time.py --nrep 1 "" > /dev/null 2>&1
  echo "$knobs"; grep fdmt gpurun_out/r02_mega_launches_tmp.csv | tail -4 | cut -d, -f13-
done
