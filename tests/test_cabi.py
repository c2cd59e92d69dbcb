"""The boundary from C: include/bifrost/*.h compile as C99, and a plain C
program linked against libbifrost_b200.so (tests/cabi/ring_fdmt_client.c) moves
a filterbank through rings whose ringlets are the channels -- spans are then
[nchan][ntime] views with a ring stride between rows, the layout arrays from a
ring have -- and, on the GPU, through bfFdmtExecute with the
overlap of the reference's FDMT block (blocks/fdmt.py:112-124).  The output is
compared with the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from oracle import fdmt as ofdmt

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, 'bifrost_b200', 'lib')
HEADERS = ['ring', 'proclog', 'affinity', 'fdmt', 'fft', 'linalg', 'map', 'memory', 'array', 'common', 'cuda',
           'quantize', 'reduce', 'transpose', 'unpack']


def test_headers_are_valid_c99_and_cxx11(tmp_path):
    src = tmp_path / 'all_headers.c'
    src.write_text(''.join(f'#include <bifrost/{h}.h>\n' for h in HEADERS) +
                   'int main(void) { BFspan_info s; BFsequence_info q; BFarray a; (void)s; (void)q; (void)a;\n'
                   '  return sizeof(BFarray) == 168 ? 0 : 1; }\n')
    inc = '-I' + os.path.join(ROOT, 'include')
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', inc, '-fsyntax-only', str(src)])
    subprocess.check_call(['g++', '-std=c++11', '-x', 'c++', '-Wall', '-Werror', inc, '-fsyntax-only', str(src)])


@pytest.fixture(scope='module')
def client(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp('cabi') / 'ring_fdmt_client')
    subprocess.check_call(['gcc', '-std=c99', '-O1', '-Wall', '-Werror', '-I' + os.path.join(ROOT, 'include'),
                           os.path.join(HERE, 'cabi', 'ring_fdmt_client.c'), '-o', exe,
                           '-L' + LIBDIR, '-lbifrost_b200', '-Wl,-rpath,' + LIBDIR])
    return exe


def run_client(client, tmp_path, space, op, x, gulp, overlap, max_delay=0, f0=0., df=0.):
    fin, fout = str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')
    x.tofile(fin)
    nchan, ntime = x.shape
    res = subprocess.run([client, space, op, str(nchan), str(ntime), str(gulp), str(overlap), str(max_delay),
                          repr(float(f0)), repr(float(df)), fin, fout], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.startswith('OK')
    raw = open(fout, 'rb').read()
    nrow = max_delay if op == 'fdmt' else nchan
    dtype = np.float32 if op == 'fdmt' else np.int8
    chunks, pos = [], 0
    while pos < len(raw):
        nout = int(np.frombuffer(raw, np.int64, 1, pos)[0])
        pos += 8
        n = nrow * nout * np.dtype(dtype).itemsize
        chunks.append(np.frombuffer(raw, dtype, nrow * nout, pos).reshape(nrow, nout))
        pos += n
    return chunks


def test_c_client_moves_spans_through_system_rings(client, tmp_path):
    """No GPU needed: the same program with a plain copy in place of the transform.
    61 channels as ringlets, gulps that do not divide the ring, a ragged end."""
    rng = np.random.default_rng(1)
    nchan, ntime, gulp, overlap = 61, 5000, 700, 130
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    chunks = run_client(client, tmp_path, 'system', 'copy', x, gulp, overlap)
    assert [c.shape[1] for c in chunks] == [700] * 6 + [ntime - 6 * 700 - overlap]
    np.testing.assert_array_equal(np.concatenate(chunks, axis=1), x[:, :ntime - overlap])


@pytest.mark.gpu
def test_c_client_runs_the_fdmt_block_through_device_rings(client, tmp_path):
    rng = np.random.default_rng(2)
    nchan, ntime, gulp, md = 256, 9000, 2048, 130
    f0, df = 1000.0, 400. / 256
    x = np.clip(np.rint(rng.normal(0, 20, size=(nchan, ntime))), -127, 127).astype(np.int8)
    chunks = run_client(client, tmp_path, 'cuda', 'fdmt', x, gulp, md, md, f0, df)
    got = np.concatenate(chunks, axis=1)
    want = np.zeros((md, ntime), np.float32)
    ofdmt.fdmt(x, md, f0, df, out=want)
    assert got.shape == (md, ntime - md)
    assert np.array_equal(got.view(np.uint32), want[:, :ntime - md].view(np.uint32))
