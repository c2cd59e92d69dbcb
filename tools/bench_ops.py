"""Per-op timings of the C-ABI entry points on BASELINE-sized inputs, next to
the reference's own CUDA kernels (oracle/_ref/libbifrost_ref.so, same device
buffers, same harness) where that library provides the op.

    python tools/bench_ops.py [--nframe 32] [--ops fdmt,transpose,fft,...]

Prints one JSON line per op: ms (median of runs, CUDA events on the launching
stream), algorithmic GB/s = in+out bytes at the C ABI / time, and
the reference's time when available.  Not the headline benchmark (bench.py is).
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402
import bifrost_b200 as bf  # noqa: E402
from bifrost_b200.libbifrost import _bf, _check  # noqa: E402
import reflib  # noqa: E402


def timeit(fn, nrep=10, nwarm=3, stream=None):
    for _ in range(nwarm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(nrep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times))


def cufft_ms(nbatch, nfft, stream, shift=False):
    """cuFFT (through torch.fft: bench tooling only, never on the product path)
    on the same number of transforms: cf32 -> cf32 c2c, the kernel the reference
    calls for this op (src/fft.cu:224-239).  The reference additionally converts
    ci8 -> cf32 in a load callback, so cuFFT's own traffic is 16 B/sample."""
    x = torch.randn(nbatch, nfft, dtype=torch.complex64, device='cuda')
    fn = (lambda: torch.fft.fftshift(torch.fft.fft(x, dim=1), dim=1)) if shift else (lambda: torch.fft.fft(x, dim=1))
    ms = timeit(fn, nrep=5, stream=stream)
    del x
    return ms


def report(name, ms, nbytes, ref_ms=None, **extra):
    line = dict(op=name, ms=round(ms, 4), alg_GBps=round(nbytes / ms / 1e6, 1))
    if ref_ms is not None:
        line['ref_ms'] = round(ref_ms, 4)
        line['speedup_vs_ref'] = round(ref_ms / ms, 2)
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nframe', type=int, default=32)
    ap.add_argument('--ops', default='config1,fdmt,transpose,fft,detect,reduce,accumulate,correlate')
    args = ap.parse_args()
    ops = args.ops.split(',')
    stream = torch.cuda.current_stream()
    bf.device.set_stream(stream.cuda_stream)
    ref = reflib.load()
    if ref is not None:
        h = ctypes.c_void_p(stream.cuda_stream)
        ref.bfStreamSet(ctypes.byref(h))
    rng = np.random.default_rng(1234)

    if 'config1' in ops:
        # BASELINE config 1: f32 [4096 frames, 256 chan]
        a = bf.asarray(rng.normal(size=(4096, 256)).astype(np.float32), space='cuda')
        t = bf.empty((256, 4096), 'f32', 'cuda')
        ms = timeit(lambda: bf.transpose(t, a, (1, 0)), stream=stream)
        rms = None
        if ref is not None:
            ax = (ctypes.c_int * 2)(1, 0)
            rms = timeit(lambda: ref.bfTranspose(a.as_BFarray(), t.as_BFarray(), ax), stream=stream)
        report('config1.transpose f32[4096,256]', ms, 2 * 4 * 4096 * 256, rms)
        r1 = bf.empty((4096, 64), 'f32', 'cuda')
        ms = timeit(lambda: bf.reduce(a, r1, 'sum'), stream=stream)
        rms = timeit(lambda: ref.bfReduce(a.as_BFarray(), r1.as_BFarray(), 0), stream=stream) if ref else None
        report('config1.reduce chan/4', ms, 4 * 4096 * 256 * 1.25, rms)
        r2 = bf.empty((512, 256), 'f32', 'cuda')
        ms = timeit(lambda: bf.reduce(a, r2, 'sum'), stream=stream)
        rms = timeit(lambda: ref.bfReduce(a.as_BFarray(), r2.as_BFarray(), 0), stream=stream) if ref else None
        report('config1.reduce time/8', ms, 4 * 4096 * 256 * 1.125, rms)

    if 'fdmt' in ops:
        import bench
        w = bench.workload(0)
        x = bench.make_input(w, 1)
        d_in = bf.asarray(x, space='cuda')
        d_out = bf.zeros((w['max_delay'], w['ntime']), 'f32', 'cuda')
        plan = bf.fdmt.Fdmt()
        plan.init(w['nchan'], w['max_delay'], w['f0'], w['df'])
        ms = timeit(lambda: plan.execute(d_in, d_out), stream=stream)
        rms = None
        if ref is not None:
            rp = ctypes.c_void_p()
            _check(ref.bfFdmtCreate(ctypes.byref(rp)))
            _check(ref.bfFdmtInit(rp, w['nchan'], w['max_delay'], w['f0'], w['df'], -2.0, 2, None, None))
            h = ctypes.c_void_p(stream.cuda_stream)
            ref.bfFdmtSetStream(rp, ctypes.byref(h))
            st = ref.bfFdmtExecute(rp, d_in.as_BFarray(), d_out.as_BFarray(), 0, None, None)
            if st == 0:
                rms = timeit(lambda: ref.bfFdmtExecute(rp, d_in.as_BFarray(), d_out.as_BFarray(), 0, None, None),
                             nrep=5, stream=stream)
            ref.bfFdmtDestroy(rp)
        report('fdmt 4096x%d i8 md=%d' % (w['ntime'], w['max_delay']), ms,
               w['ntime'] * (w['nchan'] + 4 * w['max_delay']), rms,
               Msamples_per_s=round(w['nchan'] * bench.NTIME_OUT / ms / 1e3, 1))
        del d_in, d_out, plan

    if 'fdmt_scaling' in ops:
        # the same 4096-channel gulp at max_delay 204 / 1621
        import bench
        for md, f0, bw in ((204, 1000., 400.), (1621, 1200., 300.)):
            ntime = bench.NTIME_OUT + md
            x = rng.integers(-64, 64, size=(4096, ntime), dtype=np.int8)
            d_in = bf.asarray(x, space='cuda')
            d_out = bf.zeros((md, ntime), 'f32', 'cuda')
            plan = bf.fdmt.Fdmt()
            plan.init(4096, md, f0, bw / 4096)
            ms = timeit(lambda: plan.execute(d_in, d_out), stream=stream)
            report('fdmt 4096x%d i8 md=%d' % (ntime, md), ms, ntime * (4096 + 4 * md),
                   Msamples_per_s=round(4096 * bench.NTIME_OUT / ms / 1e3, 1))
            del d_in, d_out, plan

    # ---- GUPPI chain (config 3): ci8 [nframe, 4096 chan, 4096 fine_time, 2 pol]
    nframe, nchan, nfft, npol = args.nframe, 4096, 4096, 2
    I = nframe * nchan * nfft * npol * 2         # bytes of ci8 in the gulp
    chain = [o for o in ops if o in ('transpose', 'fft', 'detect', 'reduce', 'accumulate')]
    if chain:
        raw = rng.integers(-127, 128, size=(nframe, nchan, nfft, npol, 2), dtype=np.int8)
        x = raw.view(bf.DataType('ci8').as_numpy_dtype()).reshape(nframe, nchan, nfft, npol)
        d_x = bf.asarray(x, space='cuda')
        d_t = bf.empty((nframe, npol, nchan, nfft), 'ci8', 'cuda')
        ms = timeit(lambda: bf.transpose(d_t, d_x, (0, 3, 1, 2)), nrep=5, stream=stream)
        report('chain.transpose ci8 [t,f,ft,p]->[t,p,f,ft]', ms, 2 * I, Msamples_per_s=round(nframe * nchan * nfft / ms / 1e3, 1))
        del d_x
        d_f = bf.empty((nframe, npol, nchan, nfft), 'cf32', 'cuda')
        plan = bf.fft.Fft()
        plan.init(d_t, d_f, axes=[3], apply_fftshift=True)
        ms = timeit(lambda: plan.execute(d_t, d_f), nrep=5, stream=stream)
        del d_t
        cu = cufft_ms(nframe * npol * nchan, nfft, stream)
        report('chain.fft ci8->cf32 n=4096 fftshift', ms, 5 * I, Msamples_per_s=round(nframe * nchan * nfft / ms / 1e3, 1),
               cufft_c2c_cf32_ms=round(cu, 4), cufft_GBps=round(16 * nframe * npol * nchan * nfft / cu / 1e6, 1),
               speedup_vs_cufft=round(cu / ms, 2))
        d_d = bf.empty((nframe, 4, nchan, nfft), 'f32', 'cuda')
        ms = timeit(lambda: bf.detect(d_f, d_d, 'stokes', 1), nrep=5, stream=stream)
        report('chain.detect stokes', ms, 8 * I, Msamples_per_s=round(nframe * nchan * nfft / ms / 1e3, 1))
        del d_f
        d_dm = d_d.reshape(nframe, 4, nchan * nfft)
        d_r = bf.empty((nframe, 4, nchan * nfft // 4), 'f32', 'cuda')
        ms = timeit(lambda: bf.reduce(d_dm, d_r, 'sum'), nrep=5, stream=stream)
        rms = timeit(lambda: ref.bfReduce(d_dm.as_BFarray(), d_r.as_BFarray(), 0), nrep=5, stream=stream) if ref else None
        report('chain.reduce freq/4', ms, 5 * I, rms, Msamples_per_s=round(nframe * nchan * nfft / ms / 1e3, 1))
        del d_d, d_dm
        d_acc = bf.zeros((1, 4, nchan * nfft // 4), 'f32', 'cuda')
        one = d_r[0:1]
        ms = timeit(lambda: bf.accumulate(one, d_acc, 1.0), stream=stream)
        report('chain.accumulate one frame', ms, 3 * 4 * 4 * nchan * nfft // 4)
        del d_r, d_acc

    if 'fftsizes' in ops:
        # c2c forward FFTs of other lengths (ci8 -> cf32), 2^28 complex samples each
        for nfft_ in (256, 1024, 2048, 8192, 16384, 131072):
            nbatch = (1 << 27) // nfft_
            raw = rng.integers(-127, 128, size=(nbatch, nfft_, 2), dtype=np.int8)
            xi = raw.view(bf.DataType('ci8').as_numpy_dtype()).reshape(nbatch, nfft_)
            d_i = bf.asarray(xi, space='cuda')
            d_o = bf.empty((nbatch, nfft_), 'cf32', 'cuda')
            plan = bf.fft.Fft()
            plan.init(d_i, d_o, axes=[1], apply_fftshift=False)
            ms = timeit(lambda: plan.execute(d_i, d_o), nrep=5, stream=stream)
            del d_i, d_o, plan
            cu = cufft_ms(nbatch, nfft_, stream)
            report('fft ci8->cf32 n=%d batch=%d' % (nfft_, nbatch), ms, 10 * nbatch * nfft_,
                   Msamples_per_s=round(nbatch * nfft_ / ms / 1e3, 1),
                   cufft_c2c_cf32_ms=round(cu, 4), cufft_GBps=round(16 * nbatch * nfft_ / cu / 1e6, 1),
                   speedup_vs_cufft=round(cu / ms, 2))

    if 'correlate' in ops:
        nchan_c, nstand, npol_c = 512, 256, 2
        n = nstand * npol_c
        la = bf.linalg.LinAlg()
        for ntime in (512, 2048, 8192):
            raw = rng.integers(-127, 128, size=(ntime, nchan_c, n, 2), dtype=np.int8)
            x = raw.view(bf.DataType('ci8').as_numpy_dtype()).reshape(ntime, nchan_c, n)
            d_x = bf.asarray(x, space='cuda')
            xv = d_x.transpose(1, 0, 2)               # [chan, time, stand*pol] view
            d_c = bf.zeros((nchan_c, n, n), 'cf32', 'cuda')
            try:
                la.matmul(1, None, xv, 0, d_c)
            except RuntimeError as e:
                print(json.dumps(dict(op='correlate', error=str(e))))
                break
            ms = timeit(lambda: la.matmul(1, None, xv, 0, d_c), nrep=5, stream=stream)
            rms = None
            if ref is not None:
                rl = ctypes.c_void_p()
                _check(ref.bfLinAlgCreate(ctypes.byref(rl)))
                st = ref.bfLinAlgMatMul(rl, 1.0, None, xv.as_BFarray(), 0.0, d_c.as_BFarray())
                if st == 0:
                    rms = timeit(lambda: ref.bfLinAlgMatMul(rl, 1.0, None, xv.as_BFarray(), 0.0, d_c.as_BFarray()),
                                 nrep=3, stream=stream)
                ref.bfLinAlgDestroy(rl)
            flops = nchan_c * ntime * n * (n + 1) / 2 * 8
            report('correlate ci8 ntime=%d n=%d nchan=%d' % (ntime, n, nchan_c), ms,
                   2 * ntime * nchan_c * n + 8 * nchan_c * n * (n + 1) / 2, rms,
                   TFLOPs=round(flops / ms / 1e9, 1))
            del d_x, d_c


if __name__ == '__main__':
    main()
