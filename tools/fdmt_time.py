"""Times bfFdmtExecute on BASELINE config 2 (4096 chan x 131072+max_delay int8)
under a list of environment settings, one JSON line each:

    python tools/fdmt_time.py [--md 794] [--check] "BFB_FDMT_CHAIN_D=64,24,24" "BFB_FDMT_CHAIN=0" ...

An argument is a space-separated set of NAME=VALUE knobs ("" = defaults).  The
knobs are read by bfFdmtInit, so every setting gets a fresh plan.  --check
compares each setting's output with the first one bit for bit.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import bifrost_b200 as bf  # noqa: E402
from bifrost_b200.fdmt import Fdmt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--md', type=int, default=794)
    ap.add_argument('--nchan', type=int, default=4096)
    ap.add_argument('--ntime', type=int, default=131072)
    ap.add_argument('--f0', type=float, default=1000.)
    ap.add_argument('--bw', type=float, default=400.)
    ap.add_argument('--nrep', type=int, default=10)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('knobs', nargs='*', default=[''])
    args = ap.parse_args()
    nchan, ntime, md = args.nchan, args.ntime + args.md, args.md
    rng = np.random.default_rng(1234)
    x = np.clip(np.rint(rng.normal(0, 20, size=(nchan, ntime))), -127, 127).astype(np.int8)
    stream = torch.cuda.current_stream()
    bf.device.set_stream(stream.cuda_stream)
    d_in = bf.asarray(x, space='cuda')
    d_out = bf.empty((md, ntime), 'f32', 'cuda')
    alg = ntime * (nchan + 4 * md)
    first = None
    for knob in args.knobs:
        env = dict(kv.split('=', 1) for kv in knob.split() if '=' in kv)
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            plan = Fdmt()
            plan.init(nchan, md, args.f0, args.bw / nchan)
            for _ in range(3):
                plan.execute(d_in, d_out)
            torch.cuda.synchronize()
            times = []
            for _ in range(args.nrep):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                plan.execute(d_in, d_out)
                e1.record(stream)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            line = dict(knobs=knob, ms=round(ms, 4), min_ms=round(min(times), 4),
                        alg_GBps=round(alg / ms / 1e6, 1), Gsamples_s=round(nchan * ntime / ms / 1e6, 1))
            if args.check:
                got = np.asarray(d_out.copy('system'))
                if first is None:
                    first = got
                else:
                    line['same_bits_as_first'] = bool(np.array_equal(got.view(np.uint32), first.view(np.uint32)))
            print(json.dumps(line), flush=True)
            del plan
        except Exception as e:      # a knob set that cannot be planned
            print(json.dumps(dict(knobs=knob, error=str(e))), flush=True)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


if __name__ == '__main__':
    main()
