#!/bin/bash
# round-2 GPU check M: vector diagonal store, early prefetch (mirrored tiles), 12/16-warp CTAs
timeout -s KILL 900 python tools/fdmt_time.py --check "BFB_FDMT_PACKED=0" "" "BFB_FDMT_PACKED_EARLY=0" \
  "BFB_FDMT_PACKED_WARPS=8,12,12" "BFB_FDMT_PACKED_WARPS=8,16,16" "BFB_FDMT_PACKED_WARPS=8,12,8" "BFB_FDMT_PACKED_WARPS=8,8,12" \
  "BFB_FDMT_PACKED_WARPS=8,16,12" "BFB_FDMT_PACKED_WARPS=8,12,16" \
  "BFB_FDMT_PACKED_WARPS=8,12,12 BFB_FDMT_PACKED_EARLY=0" \
  "BFB_FDMT_PACKED_WARPS=8,12,12 BFB_FDMT_PACKED_D=64,32,32 BFB_FDMT_PACKED_SMEM_KB=74,112,112" \
  > gpurun_out/r02_fdmt_time9.jsonl 2>gpurun_out/r02_fdmt_time9.err
cat gpurun_out/r02_fdmt_time9.jsonl; tail -3 gpurun_out/r02_fdmt_time9.err
timeout -s KILL 1500 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -4
