#!/bin/bash
# round-2 GPU check D: packed schedule (TMA byte staging), then its persistent single-launch form
echo "== smoke: persistent kernel on a small gulp (hard timeout)"
timeout -s KILL 120 python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ['BFB_FDMT_PACKED_MEGA'] = '1'
import test_fdmt as T
from oracle import fdmt as ofdmt
rng = np.random.default_rng(0)
for (ntime, nchan, md) in [(5000, 64, 40), (20000, 256, 100), (3000, 1024, 200)]:
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    got = T.run_gpu(x, md, 1000., 400. / nchan)
    want = np.full((md, ntime), T.SENTINEL, np.float32)
    ofdmt.fdmt(x, md, 1000., 400. / nchan, out=want)
    T.assert_same_bits(got, want)
    print('mega ok', ntime, nchan, md, flush=True)
PY
echo "smoke rc=$?"
echo "== parity, all schedules"
timeout -s KILL 1200 python -m pytest tests/test_fdmt.py -x -q -m gpu 2>&1 | tail -8
echo "== timing"
timeout -s KILL 600 python tools/fdmt_time.py --check "" "BFB_FDMT_PACKED=0" \
  "BFB_FDMT_PACKED_MEGA=1" "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=4096" "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_CHUNK=1024" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_RING_EXTRA=2" \
  "BFB_FDMT_PACKED_SMEM_KB=56,56,56" "BFB_FDMT_PACKED_SMEM_KB=56,74,74" "BFB_FDMT_PACKED_SMEM_KB=110,110,110" \
  "BFB_FDMT_PACKED_WAVES=8" "BFB_FDMT_PACKED_WAVES=2" \
  "BFB_FDMT_PACKED_SPLIT=6,9" "BFB_FDMT_PACKED_SPLIT=4,9" "BFB_FDMT_PACKED_SPLIT=5,8" "BFB_FDMT_PACKED_SPLIT=3,6,9" \
  "BFB_FDMT_PACKED_MEGA=1 BFB_FDMT_PACKED_SMEM_KB=56,56,56" \
  > gpurun_out/r02_fdmt_time3.jsonl 2>gpurun_out/r02_fdmt_time3.err
cat gpurun_out/r02_fdmt_time3.jsonl; tail -5 gpurun_out/r02_fdmt_time3.err
echo "== launch lists"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_packed_launches.csv python tools/fdmt_time.py --nrep 1 "" > /dev/null 2>&1
grep fdmt gpurun_out/r02_packed_launches.csv | tail -9 | cut -d, -f5,13-
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4 --csv --log-file gpurun_out/r02_mega_launches.csv python tools/fdmt_time.py --nrep 1 "BFB_FDMT_PACKED_MEGA=1" > /dev/null 2>&1
grep fdmt gpurun_out/r02_mega_launches.csv | tail -3 | cut -d, -f5,13-
echo "== ncu full"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed -s 9 -c 3 -f -o gpurun_out/r02_packed_prof python tools/fdmt_time.py --nrep 2 "" > gpurun_out/r02_packed_prof.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdmt_packed_mega -s 3 -c 1 -f -o gpurun_out/r02_mega_prof python tools/fdmt_time.py --nrep 2 "BFB_FDMT_PACKED_MEGA=1" > gpurun_out/r02_mega_prof.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
