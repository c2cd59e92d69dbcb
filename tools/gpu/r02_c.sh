#!/bin/bash
# round-2 GPU check C: chain schedule without register chains (KD=1) -- parity, sweep, launch list, ncu
timeout 900 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/fdmt_time.py --check "" "BFB_FDMT_CHAIN=0" \
  "BFB_FDMT_CHAIN_SMEM_KB=74,74,74" "BFB_FDMT_CHAIN_WAVES=8" "BFB_FDMT_CHAIN_WAVES=2" "BFB_FDMT_CHAIN_WAVES=1" \
  "BFB_FDMT_CHAIN_D=64,48,48 BFB_FDMT_CHAIN_SMEM_KB=110,200,200" \
  "BFB_FDMT_CHAIN_D=64,16,16 BFB_FDMT_CHAIN_SMEM_KB=74,74,74" \
  "BFB_FDMT_CHAIN_D=64,12,12 BFB_FDMT_CHAIN_SMEM_KB=56,56,56 BFB_FDMT_CHAIN_WARPS=4,4,4" \
  "BFB_FDMT_CHAIN_KD=2,2,2" "BFB_FDMT_CHAIN_KD=5,4,3 BFB_FDMT_CHAIN_JR=4,4,4" \
  "BFB_FDMT_CHAIN_SPLIT=6,9" "BFB_FDMT_CHAIN_SPLIT=4,9" "BFB_FDMT_CHAIN_SPLIT=5,8" "BFB_FDMT_CHAIN_SPLIT=3,6,9" \
  "BFB_FDMT_CHAIN_TCAP=512,512,256" \
  > gpurun_out/r02_fdmt_time2.jsonl 2>gpurun_out/r02_fdmt_time2.err
cat gpurun_out/r02_fdmt_time2.jsonl; tail -5 gpurun_out/r02_fdmt_time2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_chain_launches2.csv python tools/fdmt_time.py --nrep 1 "" > /dev/null 2>&1
grep fdmt gpurun_out/r02_chain_launches2.csv | tail -3 | cut -d, -f5,8,9,15-
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_chain -s 9 -c 3 -f -o gpurun_out/r02_chain_prof2 python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_chain_prof2.log 2>&1
ls -la gpurun_out/r02_chain_prof2.ncu-rep
