"""Short FDMT run for ncu captures (not a benchmark: numbers under ncu are never reported)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import bifrost_b200 as bf
from bifrost_b200.fdmt import Fdmt
w = bench.workload(0)
x = bench.make_input(w, 1)
d_in = bf.asarray(x, space='cuda')
d_out = bf.zeros((w['max_delay'], w['ntime']), dtype='f32', space='cuda')
plan = Fdmt(); plan.init(w['nchan'], w['max_delay'], w['f0'], w['df'])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    plan.execute(d_in, d_out)
bf.device.stream_synchronize()
print('done')
