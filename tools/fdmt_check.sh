#!/bin/bash
# quick FDMT parity + timing on the GPU box
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_fdmt.py -m gpu -x -q 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 300 python tools/bench_ops.py --ops fdmt 2>&1 | tail -1 | cut -c1-140; }
run BFB_X=0
run BFB_FDMT_TILES_PER_CTA=4
run BFB_FDMT_TILES_PER_CTA=1
run BFB_FDMT_K=4 BFB_FDMT_TILES_PER_CTA=4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_tile -s 3 -c 3 \
  -o gpurun_out/r01_fdmt_tiles4 -f python tools/profile_fdmt.py 2 > gpurun_out/fdmt_tiles_prof.log 2>&1; echo "ncu rc=$?"
