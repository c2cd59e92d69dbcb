#!/bin/bash
# ncu full capture of the three chain kernels (one launch each, after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_chain -s 9 -c 3 -f -o gpurun_out/r02_chain_prof python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_chain_prof.log 2>&1
tail -3 gpurun_out/r02_chain_prof.log
ls -la gpurun_out/r02_chain_prof.ncu-rep
