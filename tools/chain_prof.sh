#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -s 5 -c 13 \
  -o gpurun_out/r01_chain_ops -f python tools/profile_chain.py > gpurun_out/chain_prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/chain_prof.log
