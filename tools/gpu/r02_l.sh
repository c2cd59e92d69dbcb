#!/bin/bash
# per-op timings (incl. cuFFT and the reference's kernels where oracle/_ref has them)
mkdir -p gpurun_out
timeout 900 python tools/bench_ops.py > gpurun_out/r02_bench_ops.jsonl 2> gpurun_out/r02_bench_ops.err
echo "rc=$?"; tail -40 gpurun_out/r02_bench_ops.jsonl; tail -5 gpurun_out/r02_bench_ops.err
