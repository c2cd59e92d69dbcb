#!/bin/bash
# One GPU-box visit: tests, bench (both arms), ncu launch list of the bench
# command, and --set full captures of the dominant kernels.  Outputs in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/bench_ops.py --ops fdmt,fdmt_scaling,correlate,transpose,fft,detect,reduce,accumulate,fftsizes --nframe 16 > gpurun_out/bench_ops.jsonl 2>&1; tail -18 gpurun_out/bench_ops.jsonl | cut -c1-200
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
cat gpurun_out/bench_r01.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>&1
cat gpurun_out/bench_r01_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:fdmt_ -s 3 -c 3 --csv --log-file gpurun_out/r01_fdmt_dram.csv python tools/profile_fdmt.py 2 \
  > gpurun_out/fdmt_dram.log 2>&1; echo "ncu dram rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fdmt_tile -s 3 -c 3 \
  -o gpurun_out/r01_fdmt_tiles -f python tools/profile_fdmt.py 2 > gpurun_out/fdmt_tiles_full.log 2>&1; echo "ncu tiles rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spectrometer|corr_tc" -s 1 -c 2 \
  -o gpurun_out/r01_spec_corr -f python tools/profile_ops.py > gpurun_out/spec_corr_full.log 2>&1; echo "ncu spec/corr rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls gpurun_out | wc -l
