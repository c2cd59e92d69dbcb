#!/bin/bash
# round-2 GPU check R (1 GPU): fixes of check Q, slowest tests, 3-CTA variants of the workspace passes,
# bench record with the 4-CTA byte pass, launch list, ncu of the new kernels
timeout -s KILL 900 python -m pytest tests/test_fft.py tests/test_linalg.py tests/test_quantize.py -q -m gpu --durations=12 2>&1 | tail -25
timeout -s KILL 600 python tools/fdmt_time.py --check "" "BFB_FDMT_PACKED_SMEM_KB=56,74,74 BFB_FDMT_PACKED_WARPS=8,8,8" \
  "BFB_FDMT_PACKED_SMEM_KB=56,74,110 BFB_FDMT_PACKED_WARPS=8,8,12" "BFB_FDMT_PACKED_SMEM_KB=56,110,74 BFB_FDMT_PACKED_WARPS=8,12,8" \
  "BFB_FDMT_PACKED_TCAP=736,720,288" "BFB_FDMT_PACKED_WAVES=16" \
  "BFB_FDMT_PACKED_CHUNKED=32768" "BFB_FDMT_PACKED_CHUNKED=16384" "BFB_FDMT_PACKED_CHUNKED=8192" "BFB_FDMT_PACKED_CHUNKED=4096" \
  "BFB_FDMT_PACKED_CHUNKED=8192 BFB_FDMT_PACKED_WAVES=2" "BFB_FDMT_PACKED_CHUNKED=16384 BFB_FDMT_PACKED_WAVES=4" 2>&1 | tail -14
for C in 16384 8192; do
  BFB_FDMT_PACKED_CHUNKED=$C timeout -s KILL 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fdmt_packed -s 200 -c 100 --csv \
    --log-file gpurun_out/r02_chunked_dram_$C.csv python tools/fdmt_time.py --nrep 3 "" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_chunked_dram_$C.csv')) if len(r)>5 and r[0].isdigit()]
tot=sum(float(r[-1].replace(',','')) for r in rows)
print('chunked C=$C: %d launches captured, %.3f GB dram (read+write, unit as reported: %s)' % (len(rows)//2, tot/1e9, rows[0][-2] if rows else '?'))
PY
done
timeout -s KILL 300 python tools/profile_ops.py beamform,time 2>&1 | tail -3
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
echo "bench rc=$?"; head -c 900 gpurun_out/r02_bench_c.json; echo; tail -2 gpurun_out/r02_bench_c.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_c.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed -s 9 -c 3 -f -o gpurun_out/r02_packed_prof5 python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_packed_prof5.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:ab_tc -s 2 -c 1 -f -o gpurun_out/r02_abtc_prof python tools/profile_ops.py beamform > gpurun_out/r02_abtc_prof.log 2>&1
ls -la gpurun_out/r02_packed_prof5.ncu-rep gpurun_out/r02_abtc_prof.ncu-rep gpurun_out/r02_bench_launches_c.csv
