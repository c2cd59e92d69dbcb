#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fdmt_tile -s 3 -c 3 \
  -o gpurun_out/r01_fdmt_tiles -f python tools/profile_fdmt.py 2 > gpurun_out/fdmt_tiles_prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/fdmt_tiles_prof.log
