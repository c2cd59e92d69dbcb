#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Bifrost hot path.

Metric (BASELINE.json): Msamples/s through the FDMT on a synthetic 4096-chan
int8 filterbank (config 2: max_dm=100 -> max_delay=794, 131072 output samples
per gulp), plus % of the HBM roofline.  One "step" = one gulp through
bfFdmtExecute.  `value` is timed with the input resident in HBM; `e2e` is the
same call made from HOST buffers (pinned) with the H2D copy of the gulp and
the D2H read of the dispersion bank inside the timed region.  The second half
of the metric (FFT -> detect -> reduce -> accumulate on the ci8 GUPPI gulp,
config 3) is reported in the `chain` object (one fused kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun): rank g processes its own 4096-channel sub-band (weak
scaling, no data-path collective -- SURVEY 8e); time = max over ranks.
`--impl reference` times the CPU restatement of the reference algorithm
(oracle/) on the host cores of this box on a bounded sample of the workload.
`--dry-run` exercises the multi-rank host logic (sharding, barrier, max over
ranks) on the gloo backend with no GPU work (tests/test_bench_cpu.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCHAN = 4096
NTIME_OUT = 131072
# bounded CPU sample: 1/4 of the gulp's time span, all 4096 channels (env override: tests only)
CPU_SAMPLE_NTIME = int(os.environ.get('BENCH_CPU_SAMPLE_NTIME', 32768))
F0_MHZ = 1000.0
BW_MHZ = 400.0
DT_S = 256e-6
MAX_DM = 100.0
KDM = 4.148741601e3
METRIC = 'Msamples/s through FDMT+FFT->detect->reduce on 4096-chan ci8; % HBM roofline'


def max_delay_for(f0, df, nchan, dt, max_dm):
    """blocks/fdmt.py:79-81 of the reference."""
    rel = (f0 ** -2 - (f0 + nchan * df) ** -2)
    return int(np.ceil(abs(KDM / dt * max_dm * rel)))


def workload(rank=0):
    df = BW_MHZ / NCHAN
    f0 = F0_MHZ + rank * NCHAN * df
    md = max_delay_for(F0_MHZ, df, NCHAN, DT_S, MAX_DM)      # same bank depth on every rank
    return dict(nchan=NCHAN, ntime=NTIME_OUT + md, max_delay=md, f0=f0, df=df)


def make_input(w, seed, ntime=None):
    """round(N(0,20)) clipped to int8 plus three dispersed pulses (BASELINE.md cfg 2)."""
    ntime = ntime or w['ntime']
    rng = np.random.default_rng(seed)
    x = np.empty((w['nchan'], ntime), np.int8)
    step = 256
    blk = rng.normal(0, 20, size=(step, ntime)).astype(np.float32)
    np.rint(blk, out=blk)
    np.clip(blk, -127, 127, out=blk)
    blk = blk.astype(np.int8)
    for i, c0 in enumerate(range(0, w['nchan'], step)):
        # one noise block reused with a different cyclic time shift per block
        x[c0:c0 + step] = np.roll(blk, 7919 * i, axis=1)[:min(step, w['nchan'] - c0)]
    f = w['f0'] + w['df'] * np.arange(w['nchan'])
    fmax = f[-1]
    rel = (f ** -2 - fmax ** -2) / (f[0] ** -2 - fmax ** -2)
    for frac_dm, t0 in ((0.2, ntime // 5), (0.5, ntime // 2), (0.9, (3 * ntime) // 4)):
        tt = t0 + np.rint(rel * frac_dm * (w['max_delay'] - 1)).astype(np.int64)
        ok = tt < ntime
        x[np.arange(w['nchan'])[ok], tt[ok]] = 100
    return x


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smmax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smmax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(smmax)) if smmax else None,
                    samples=len(sm), reasons=sorted(reasons))


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return json.load(open(path)), 'measured'
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), 'fallback'


# --------------------------------------------------------------------------- CPU arm
def cpu_fdmt_sample(w, ntime_sample, threads=None):
    """Times the oracle (CPU restatement of the reference algorithm) on a
    bounded sample: the full 4096 channels, `ntime_sample` time samples.
    Returns (Msamples/s, kind, cores, seconds)."""
    try:
        from oracle import fdmt_c
        have_c = fdmt_c.available()
    except Exception:
        have_c = False
    if not have_c:
        ntime_sample = min(ntime_sample, 8192)        # the numpy port is ~10x slower
    x = make_input(w, 4321, ntime=ntime_sample + w['max_delay'])
    nsamp = w['nchan'] * ntime_sample
    if have_c:
        from oracle import fdmt_c
        cores = threads or os.cpu_count()
        plan = fdmt_c.Plan(w['nchan'], w['max_delay'], w['f0'], w['df'])
        out = np.zeros((w['max_delay'], x.shape[1]), np.float32)
        plan.execute(x, out, threads=cores)            # warm (page-in, thread pool)
        t0 = time.perf_counter()
        plan.execute(x, out, threads=cores)
        dt = time.perf_counter() - t0
        return nsamp / dt / 1e6, 'port', cores, dt
    from oracle import fdmt as ofdmt
    plan = ofdmt.FdmtPlan(w['nchan'], w['max_delay'], w['f0'], w['df'])
    t0 = time.perf_counter()
    ofdmt.fdmt(x, w['max_delay'], w['f0'], w['df'], plan=plan)
    dt = time.perf_counter() - t0
    return nsamp / dt / 1e6, 'port', 1, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w = workload(0)
    ntime_sample = CPU_SAMPLE_NTIME
    vals, secs = [], []
    cores = 1
    for i in range(args.warmup + args.steps):
        v, kind, cores, dt = cpu_fdmt_sample(w, ntime_sample)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = float(np.mean(vals))
    sample = (f"oracle CPU FDMT (oracle/fdmt_c.c, pthreads) on {w['nchan']} chan x {ntime_sample} samples "
              f"(1/{NTIME_OUT // ntime_sample} of the gulp) per step, {cores} threads")
    line = dict(impl='reference', metric=METRIC, value=value, unit='Msamples/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=float(np.mean(secs) * 1e3),
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic',
                config=dict(workload='BASELINE config 2: FDMT max_dm=100 on 4096-chan x 128k-sample '
                                     'int8 filterbank (bounded sample, see cpu_baseline.sample)'),
                cpu_baseline=dict(value=value, unit='Msamples/s', cores=cores, kind='port',
                                  sample=sample),
                e2e=dict(value=value, unit='Msamples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------- dry run
def run_dry(args, rank, world):
    """Multi-rank host logic without a GPU: sharding, barrier, max over ranks."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('gloo')
    w = workload(rank)
    ms = 1.0 + 0.5 * rank                      # pretend rank r needs 1 + r/2 ms per step
    f0s = [w['f0']]
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([w['f0']], dtype=torch.float64))
        f0s = [float(p.item()) for p in parts]
    if rank == 0:
        print(json.dumps(dict(dry_run=True, n_gpus=world, ms_per_step=ms,
                              value=w['nchan'] * NTIME_OUT * world / (ms * 1e-3) / 1e6,
                              subband_f0_mhz=f0s, max_delay=w['max_delay'])))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- GPU arm
def log(*a):
    if os.environ.get('BENCH_VERBOSE'):
        print('[bench]', *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-chain', action='store_true')
    ap.add_argument('--dry-run', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if args.dry_run:
        run_dry(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import bifrost_b200 as bf
    from bifrost_b200.fdmt import Fdmt

    torch.cuda.set_device(local_rank)
    bf.device.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    stream = torch.cuda.current_stream()
    bf.device.set_stream(stream.cuda_stream)

    log('torch/bf imported, device set')
    w = workload(rank)
    nchan, ntime, md = w['nchan'], w['ntime'], w['max_delay']
    x_host = make_input(w, 1234 + rank)
    pinned_in = bf.empty((nchan, ntime), dtype='i8', space='cuda_host')
    np.copyto(np.asarray(pinned_in), x_host)
    pinned_out = bf.empty((md, ntime), dtype='f32', space='cuda_host')
    d_in = bf.empty((nchan, ntime), dtype='i8', space='cuda')
    d_out = bf.empty((md, ntime), dtype='f32', space='cuda')
    bf.copy_array(d_in, pinned_in)
    bf.memset_array(d_out, 0)
    log('buffers ready')
    plan = Fdmt()
    plan.init(nchan, md, w['f0'], w['df'])
    ws_size = plan.get_workspace_size(d_in, d_out)
    ws = bf.empty((ws_size,), dtype='u8', space='cuda')

    def step_resident():
        plan.execute_workspace(d_in, d_out, ws.ctypes.data, ws_size)

    def step_e2e():
        bf.copy_array(d_in, pinned_in)            # H2D of this gulp (pinned, async on the stream)
        plan.execute_workspace(d_in, d_out, ws.ctypes.data, ws_size)
        bf.copy_array(pinned_out, d_out)          # D2H of the dispersion bank (+ stream sync)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nwarm, nstep, sampler=None):
        for _ in range(nwarm):
            fn()
        barrier()
        if sampler:
            sampler.start()
            time.sleep(0.1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = bf.launch_count()
        ev0.record(stream)
        for _ in range(nstep):
            fn()
        ev1.record(stream)
        barrier()
        launches = bf.launch_count() - launches0
        clocks = None
        if sampler:
            # the timed region can be shorter than nvidia-smi's sampling period:
            # keep the same load running (untimed) until a few samples exist
            t_end = time.time() + 1.5
            while len(sampler.lines) < 8 and time.time() < t_end:
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
            clocks = sampler.stop()
            clocks['window'] = 'timed steps + untimed post-roll of the same steps (nvidia-smi -lms 50)'
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    log('workspace', ws_size)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, clocks = timed(step_resident, args.warmup, args.steps, sampler)
    ms_step = ms_total / args.steps
    log('resident ms/step', ms_step)
    e2e_steps = max(4, min(args.steps, 10))
    ms_e2e_total, _, _ = timed(step_e2e, 2, e2e_steps)
    ms_e2e_serial = ms_e2e_total / e2e_steps

    # The same three calls as a host pipeline, the way bifrost runs blocks: one
    # thread (and thread-local stream) per stage -- copy(H2D) -> fdmt -> copy(D2H)
    # -- double-buffered, each stage synchronising its own stream per gulp
    # (pipeline.py:628 of the reference).  Every gulp's H2D and D2H is inside
    # the timed region; consecutive gulps overlap on the copy engines.
    import queue
    d_in2 = [d_in, bf.empty((nchan, ntime), dtype='i8', space='cuda')]
    d_out2 = [d_out, bf.empty((md, ntime), dtype='f32', space='cuda')]
    ws2 = [ws, bf.empty((ws_size,), dtype='u8', space='cuda')]

    def run_pipeline(nstep):
        q_in_free, q_in_full = queue.Queue(), queue.Queue()
        q_out_free, q_out_full = queue.Queue(), queue.Queue()
        for i in range(2):
            q_in_free.put(i)
            q_out_free.put(i)
        errors = []

        def stage(fn):
            def body():
                try:
                    bf.device.set_device(local_rank)
                    fn()
                except Exception as e:      # surfaces in the main thread
                    errors.append(e)
            return threading.Thread(target=body)

        def h2d():
            for _ in range(nstep):
                i = q_in_free.get()
                bf.copy_array(d_in2[i], pinned_in)
                bf.device.stream_synchronize()
                q_in_full.put(i)

        def compute():
            for _ in range(nstep):
                i = q_in_full.get()
                o = q_out_free.get()
                plan.execute_workspace(d_in2[i], d_out2[o], ws2[o].ctypes.data, ws_size)
                bf.device.stream_synchronize()
                q_in_free.put(i)
                q_out_full.put(o)

        def d2h():
            for _ in range(nstep):
                o = q_out_full.get()
                bf.copy_array(pinned_out, d_out2[o])
                bf.device.stream_synchronize()
                q_out_free.put(o)

        threads = [stage(h2d), stage(compute), stage(d2h)]
        barrier()
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        return dt * 1e3

    try:
        run_pipeline(2)                                # warm-up
        ms_pipe = run_pipeline(e2e_steps) / e2e_steps
    except Exception as e:                             # keep the serial figure; every rank still reduces
        log('pipelined e2e failed:', e)
        ms_pipe = float('inf')
    if world > 1:
        t = torch.tensor([ms_pipe], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_pipe = float(t.item())
    ms_e2e = min(ms_pipe, ms_e2e_serial)
    del d_in2, d_out2, ws2

    log('e2e ms/step', ms_e2e)
    samples_per_step = nchan * NTIME_OUT * world            # pol not counted
    value = samples_per_step / (ms_step * 1e-3) / 1e6
    e2e_value = samples_per_step / (ms_e2e * 1e-3) / 1e6

    # Roofline of the op (all kernels of one bfFdmtExecute): compulsory bytes
    # ntime*(nchan*1 + max_delay*4) per gulp, per GPU.
    peaks, peak_kind = measured_peaks()
    alg_bytes = ntime * (nchan * 1 + md * 4)
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    roofline = dict(bound='hbm', kernel='bfFdmtExecute = 3 x fdmt_tile_kernel (raw head 1..5, pass 6..9, final pass 10..12); '
                           'bytes and time are those of the whole call',
                    achieved=achieved, peak=peaks['hbm_gbs'], peak_source=peak_kind,
                    unit='GB/s', frac=achieved / peaks['hbm_gbs'],
                    algorithmic_bytes=alg_bytes, traffic=None)
    prof = os.path.join(ROOT, 'profiles', 'fdmt_traffic.json')
    if os.path.exists(prof):
        try:
            roofline['traffic'] = json.load(open(prof)).get('dram_bytes_per_call')
        except Exception:
            pass

    # ---- second half of the metric: the fused GUPPI chain (config 3), per rank
    chain = None
    if not args.no_chain:
        del ws, d_out
        nframe, cchan, nfft, f_avg = 32, 4096, 4096, 4
        nbyte = nframe * cchan * nfft * 4
        raw = torch.randint(-127, 128, (nbyte,), dtype=torch.int8, device='cuda')
        xg = bf.empty((nframe, cchan, nfft, 2), 'ci8', 'cuda')
        from bifrost_b200.libbifrost import _bf, _check
        _check(_bf.bfMemcpy(xg.ctypes.data, xg.as_BFarray().space, raw.data_ptr(),
                            xg.as_BFarray().space, nbyte))
        torch.cuda.synchronize()
        del raw
        og = bf.zeros((4, cchan * nfft // f_avg), 'f32', 'cuda')
        ms_c, _, _ = timed(lambda: bf.spectrometer(xg, og, nfft, f_avg, 0.0), 3, 10)
        ms_c /= 10
        gbs = nbyte / (ms_c * 1e-3) / 1e9
        chain = dict(workload='BASELINE config 3: GUPPI ci8 [32 frames, 4096 chan, 4096 fine_time, 2 pol] -> '
                              'fft(fine_time, fftshift) -> stokes -> reduce(f_avg=4) -> accumulate(32 frames); '
                              'bfSpectrometerFused, one kernel per gulp per GPU',
                     ms_per_gulp=ms_c, value=nframe * cchan * nfft * world / (ms_c * 1e-3) / 1e6,
                     unit='Msamples/s', hbm_GBps=gbs, hbm_frac=gbs / peaks['hbm_gbs'],
                     note='fp32-issue bound (2 x 4096-pt FFT per 16 KB read); see DESIGN.md')
        log('chain ms/gulp', ms_c)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, kind, cores, dt = cpu_fdmt_sample(w, CPU_SAMPLE_NTIME)
        cpu = dict(value=v, unit='Msamples/s', cores=cores, kind=kind,
                   sample=f"oracle CPU FDMT: {nchan} chan x {CPU_SAMPLE_NTIME} samples "
                          f"(1/{NTIME_OUT // CPU_SAMPLE_NTIME} of the gulp), {dt:.2f} s, "
                          f"{cores} thread(s) of {os.cpu_count()} host cores")

    line = dict(metric=METRIC, value=value, unit='Msamples/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='BASELINE config 2: bfFdmtExecute, max_dm=100 '
                                     f'(max_delay={md}) on {nchan}-chan x {NTIME_OUT}(+{md})-sample '
                                     'int8 filterbank per GPU; f0=1000 MHz, bw=400 MHz, dt=256 us',
                            sharding='one 4096-chan sub-band per GPU, no collective' if world > 1
                                     else 'single GPU',
                            l2='input 540 MB + output 419 MB per step exceed the 126 MB L2'),
                roofline=roofline, cpu_baseline=cpu,
                e2e=dict(value=e2e_value, unit='Msamples/s', ms_per_step=ms_e2e,
                         ms_per_step_serial=ms_e2e_serial, ms_per_step_pipelined=ms_pipe,
                         how='host buffers -> bf.copy_array(H2D) -> Fdmt.execute -> bf.copy_array(D2H) per gulp; '
                             'pipelined = one host thread + stream per stage, double-buffered (wall clock, '
                             'device synchronised both sides); serial = one stream',
                         h2d_bytes_per_step=int(nchan * ntime), d2h_bytes_per_step=int(md * ntime * 4)),
                gpu_launches=int(launches), clocks=clocks, chain=chain)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
