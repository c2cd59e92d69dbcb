#!/bin/bash
# round-2 GPU check Q (1 GPU): new paths (general bfMap, packed quantize, tensor-core a.b, long real FFTs,
# ring executor), FDMT byte pass at four CTAs per SM + knob sweep, per-op timings
timeout -s KILL 1200 python -m pytest tests/test_quantize.py tests/test_map.py tests/test_linalg.py tests/test_fft.py \
  tests/test_blocks_gpu.py tests/test_io_formats.py tests/test_spectrometer.py tests/test_reduce.py tests/test_transpose.py \
  tests/test_unpack.py tests/test_detect_accumulate.py tests/test_overlay.py -q -m gpu 2>&1 | tail -25
timeout -s KILL 900 python tools/fdmt_time.py --check "BFB_FDMT_PACKED=0" "" \
  "BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110" \
  "BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110 BFB_FDMT_PACKED_MINB4=0" \
  "BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110 BFB_FDMT_PACKED_WAVES=16" \
  "BFB_FDMT_PACKED_D=64,28,28 BFB_FDMT_PACKED_SMEM_KB=74,113,113" \
  "BFB_FDMT_PACKED_D=64,20,20" "BFB_FDMT_PACKED_D=64,24,32 BFB_FDMT_PACKED_SMEM_KB=74,110,113" \
  "BFB_FDMT_PACKED_SPLIT=6,9" "BFB_FDMT_PACKED_SPLIT=4,9" "BFB_FDMT_PACKED_SPLIT=5,9,11" "BFB_FDMT_PACKED_SPLIT=5,8,9" \
  "BFB_FDMT_PACKED_SPLIT=4,9 BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110" \
  > gpurun_out/r02_fdmt_time10.jsonl 2>gpurun_out/r02_fdmt_time10.err
cat gpurun_out/r02_fdmt_time10.jsonl; tail -3 gpurun_out/r02_fdmt_time10.err
timeout -s KILL 300 python tools/fdmt_time.py --md 204 "" "BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110" 2>&1 | tail -2
timeout -s KILL 300 python tools/fdmt_time.py --md 1621 --f0 1200 --bw 300 "" "BFB_FDMT_PACKED_PREFETCH=0,0,0 BFB_FDMT_PACKED_SMEM_KB=56,110,110" 2>&1 | tail -2
timeout 900 python tools/bench_ops.py > gpurun_out/r02_bench_ops.jsonl 2> gpurun_out/r02_bench_ops.err
echo "bench_ops rc=$?"; tail -45 gpurun_out/r02_bench_ops.jsonl | cut -c1-400; tail -3 gpurun_out/r02_bench_ops.err
