"""CPU-side checks of bench.py's host logic and of the C oracle used as the
CPU baseline: the C restatement (oracle/fdmt_c.c) must equal the numpy oracle
bit for bit, the workload sharding must give disjoint sub-bands, and the
multi-rank path (barrier + max over ranks) must run at world_size 2 on gloo."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import fdmt as ofdmt  # noqa: E402
from oracle import fdmt_c  # noqa: E402


@pytest.mark.parametrize('nchan,max_delay,f0,df,ntime', [
    (64, 50, 1200.0, 1.5, 300),
    (256, 100, 1000.0, 400. / 256, 513),
    (100, 37, 60.0, -0.02, 200),
])
def test_c_oracle_equals_numpy_oracle(nchan, max_delay, f0, df, ntime):
    if not fdmt_c.available():
        pytest.skip('oracle/libfdmt_oracle.so not built')
    rng = np.random.default_rng(nchan)
    x = rng.integers(-128, 128, size=(nchan, ntime), dtype=np.int8)
    want = ofdmt.fdmt(x, max_delay, f0, df)
    plan = fdmt_c.Plan(nchan, max_delay, f0, df)
    for threads in (1, 3):
        got = np.zeros((max_delay, ntime), np.float32)
        plan.execute(x, got, threads=threads)
        np.testing.assert_array_equal(got, want)


def test_workload_shards_are_disjoint_subbands():
    ws = [bench.workload(r) for r in range(8)]
    assert ws[0]['max_delay'] == 794            # BASELINE config 2: max_dm=100 -> 794 delays
    for a, b in zip(ws[:-1], ws[1:]):
        assert abs(a['f0'] + a['nchan'] * a['df'] - b['f0']) < 1e-9
        assert a['max_delay'] == b['max_delay'] and a['ntime'] == b['ntime']


def test_input_is_deterministic_and_has_pulses():
    w = dict(bench.workload(0))
    w.update(nchan=512)
    a = bench.make_input(w, 7, ntime=4096)
    b = bench.make_input(w, 7, ntime=4096)
    assert a.dtype == np.int8 and a.shape == (512, 4096)
    np.testing.assert_array_equal(a, b)
    assert (a == 100).sum() >= 512


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_config5_subbands_partition_the_gulp():
    """Config 5: N sub-bands of 4096/N channels with their own headers; the
    per-sub-band bank depths follow blocks/fdmt.py:79-81 and add up to (about)
    the full band's."""
    full = bench.workload(0)
    for n in (2, 4, 8):
        subs = [bench.subband(g, n) for g in range(n)]
        assert [s['chan0'] for s in subs] == [g * 4096 // n for g in range(n)]
        assert all(s['nchan'] == 4096 // n and s['ntime'] == full['ntime'] for s in subs)
        for a, b in zip(subs[:-1], subs[1:]):
            assert abs(a['f0'] + a['nchan'] * a['df'] - b['f0']) < 1e-9
            assert a['max_delay'] > b['max_delay']          # nu^-2: the low sub-bands are the deep ones
        assert full['max_delay'] <= sum(s['max_delay'] for s in subs) <= full['max_delay'] + n


def test_dry_run_single():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--dry-run'],
                         capture_output=True, text=True, timeout=300, check=True)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['n_gpus'] == 1 and line['ms_per_step'] == 1.0
    assert line['config'] == bench.bench_config(1) and 'config 2' in line['config']['workload']


def test_dry_run_world_size_2_gloo():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--dry-run']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1                       # rank 0 alone prints
    line = json.loads(lines[0])
    assert line['n_gpus'] == 2 and line['scaling'] == 'strong'
    assert line['config'] == bench.bench_config(2) and 'config 5' in line['config']['workload']
    assert line['config']['subband_max_delay'] == line['subband_max_delay']
    assert line['ms_per_step'] == 1.5            # max over ranks, not mean / rank 0
    assert line['subband_f0_mhz'] == [1000.0, 1200.0] and line['subband_nchan'] == [2048, 2048]
    # one gulp whatever N is: the value is that gulp's samples over the slowest rank's time
    assert abs(line['value'] - 4096 * 131072 / 1.5e-3 / 1e6) < 1e-3
    # the banks of both ranks landed at their offsets in rank 0's bank, nothing left unwritten
    assert line['gathered'] == [0.0, 1.0, 0.0]
    assert line['bank_offsets'][1] == line['subband_max_delay'][0]


def test_reference_arm_line(monkeypatch):
    """--impl reference prints the contract's JSON line (bounded sample shrunk here)."""
    if not fdmt_c.available():
        pytest.skip('oracle/libfdmt_oracle.so not built')
    env = dict(os.environ, BENCH_CPU_SAMPLE_NTIME='2048')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, env=env, check=True)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'Msamples/s'
    assert line['config'] == bench.bench_config(1)          # the GPU arm's config, verbatim
    assert line['metric'] == bench.METRIC and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['value'] == line['value']
    assert line['value'] > 0
