#!/bin/bash
# round-2 GPU check U (1 GPU): sharded plan at N = 4 of config 2's band; where the reference-library FDMT test spends its time
timeout -s KILL 300 python -m pytest tests/test_fdmt_sharded.py -x -q -m gpu -k "4096" 2>&1 | tail -3
timeout -s KILL 500 python - <<'PY'
import sys, time, ctypes
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import bifrost_b200 as bf
import reflib
from bifrost_b200.libbifrost import _check
t0 = time.time(); ref = reflib.load(); print('load', round(time.time() - t0, 2))
for (ntime, nchan, md, dtype) in [(333, 17, 40, np.uint16), (1500, 100, 64, np.int8), (2000, 256, 301, np.float32)]:
    x = np.zeros((nchan, ntime), dtype)
    d_in = bf.asarray(x, space='cuda'); d_out = bf.asarray(np.zeros((md, ntime), np.float32), space='cuda')
    plan = ctypes.c_void_p()
    t = time.time(); _check(ref.bfFdmtCreate(ctypes.byref(plan))); a = time.time() - t
    t = time.time(); _check(ref.bfFdmtInit(plan, nchan, md, 1000., 400. / nchan, -2.0, 2, None, None)); b = time.time() - t
    t = time.time(); _check(ref.bfFdmtExecute(plan, d_in.as_BFarray(), d_out.as_BFarray(), 0, None, None)); _check(ref.bfStreamSynchronize()); c = time.time() - t
    print((ntime, nchan, md, dtype.__name__), 'create %.2f init %.2f execute %.2f' % (a, b, c), flush=True)
PY
