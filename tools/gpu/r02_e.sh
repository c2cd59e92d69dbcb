#!/bin/bash
# round-2 GPU check E: whole GPU suite, then FDMT timing (fused steps 1+2, persistent kernel with chunk items)
echo "== full GPU suite"
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== timing"
timeout -s KILL 600 python tools/fdmt_time.py --check "" "BFB_FDMT_PACKED=0" "BFB_FDMT_PACKED_FUSE=0" \
  "BFB_FDMT_PACKED_MEGA=1" "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=4096" "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=1024" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=8192" \
  "BFB_FDMT_PACKED_SMEM_KB=56,74,74" "BFB_FDMT_PACKED_SMEM_KB=110,110,110" "BFB_FDMT_PACKED_SMEM_KB=56,110,110" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_SMEM_KB=110,110,110" \
  "BFB_FDMT_PACKED_WAVES=8" "BFB_FDMT_PACKED_WAVES=16" \
  "BFB_FDMT_PACKED_SPLIT=6,9" "BFB_FDMT_PACKED_SPLIT=5,8" "BFB_FDMT_PACKED_SPLIT=4,7,9" \
  "BFB_FDMT_PACKED_WARPS=4,8,8 BFB_FDMT_PACKED_SMEM_KB=36,74,74" \
  > gpurun_out/r02_fdmt_time4.jsonl 2>gpurun_out/r02_fdmt_time4.err
cat gpurun_out/r02_fdmt_time4.jsonl; tail -5 gpurun_out/r02_fdmt_time4.err
echo "== md sweep"
timeout -s KILL 300 python tools/fdmt_time.py --md 204 "" "BFB_FDMT_PACKED=0" "BFB_FDMT_PACKED_MEGA=1" 2>&1 | tail -3
timeout -s KILL 300 python tools/fdmt_time.py --md 1621 --f0 1200 --bw 300 "" "BFB_FDMT_PACKED=0" "BFB_FDMT_PACKED_MEGA=1" 2>&1 | tail -3
echo "== launch lists"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_packed_launches2.csv python tools/fdmt_time.py --nrep 1 "" > /dev/null 2>&1
grep fdmt gpurun_out/r02_packed_launches2.csv | tail -9 | cut -d, -f5,13-
echo "== ncu full"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed -s 9 -c 3 -f -o gpurun_out/r02_packed_prof2 python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_packed_prof2.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed_mega -s 3 -c 1 -f -o gpurun_out/r02_mega_prof2 python tools/fdmt_time.py --nrep 2 "BFB_FDMT_PACKED_MEGA=1" > gpurun_out/r02_mega_prof2.log 2>&1
ls -la gpurun_out/*prof2.ncu-rep
