"""The INTEGRATION.md overlay, run for real: a ctypes namespace bound to a
stand-in for the STOCK libbifrost (tests/overlay/stock_standin.c: ring /
proclog / affinity symbols of SURVEY 8(b) plus hot-path symbols that only count
their calls and return UNSUPPORTED) gets its hot-path symbols rebound to
libbifrost_b200.so by bifrost_b200.overlay.apply.  Afterwards the ring symbols
are still the stand-in's, the hot-path calls are served by this library, and
(GPU) a gulp goes through the overlaid namespace bit-exactly."""
import ctypes
import os
import subprocess
import types

import numpy as np
import pytest

import bifrost_b200 as bf
from bifrost_b200 import overlay
from bifrost_b200.libbifrost import _PROTOTYPES, BFarray

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
B200 = os.path.join(ROOT, 'bifrost_b200', 'lib', 'libbifrost_b200.so')


@pytest.fixture(scope='module')
def stock(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('overlay') / 'libbifrost_stock_standin.so')
    subprocess.check_call(['gcc', '-shared', '-fPIC', '-O1', '-o', so, os.path.join(HERE, 'overlay', 'stock_standin.c')])
    lib = ctypes.CDLL(so, mode=os.RTLD_LOCAL)
    # what the reference's generated binding looks like: a namespace of typed functions
    ns = types.SimpleNamespace()
    for name in overlay.HOT_PATH + overlay.MIRRORED + (
            'bfRingCreate', 'bfRingDestroy', 'bfRingResize', 'bfRingGetName', 'bfRingGetSpace',
            'bfProcLogCreate', 'bfProcLogDestroy', 'bfProcLogUpdate', 'bfAffinityGetCore', 'bfAffinitySetCore'):
        fn = getattr(lib, name, None)
        if fn is None:
            continue
        fn.restype = ctypes.c_int
        if name in _PROTOTYPES:                       # the stock prototypes of the hot path
            fn.restype, fn.argtypes = _PROTOTYPES[name]
        setattr(ns, name, fn)
    ns._lib = lib
    return ns


def test_overlay_rebinds_the_hot_path_and_nothing_else(stock):
    ring_create = stock.bfRingCreate
    before = stock._lib.standin_hot_calls()
    plan = ctypes.c_void_p()
    assert stock.bfFdmtCreate(ctypes.byref(plan)) == 7 and stock._lib.standin_hot_calls() == before + 1
    done = overlay.apply(stock, B200)
    assert set(done) == set(overlay.HOT_PATH)
    # hot path: served by libbifrost_b200.so now (stand-in counter stays put)
    calls = stock._lib.standin_hot_calls()
    assert stock.bfFdmtCreate(ctypes.byref(plan)) == 0 and plan.value
    assert stock.bfFdmtDestroy(plan) == 0
    h = ctypes.c_void_p()
    assert stock.bfLinAlgCreate(ctypes.byref(h)) == 0 and stock.bfLinAlgDestroy(h) == 0
    assert stock._lib.standin_hot_calls() == calls
    assert stock.bfFdmtExecute.argtypes == _PROTOTYPES['bfFdmtExecute'][1]
    # everything else: still the stock library's
    assert stock.bfRingCreate is ring_create
    ring = ctypes.c_void_p()
    assert stock.bfRingCreate(ctypes.byref(ring), b'overlay_ring', 2) == 0
    name, space = ctypes.c_char_p(), ctypes.c_int()
    assert stock.bfRingGetName(ring, ctypes.byref(name)) == 0 and name.value == b'overlay_ring'
    assert stock.bfRingGetSpace(ring, ctypes.byref(space)) == 0 and space.value == 2
    assert stock.bfRingDestroy(ring) == 0
    assert stock.bfAffinitySetCore(3) == 0
    core = ctypes.c_int()
    assert stock.bfAffinityGetCore(ctypes.byref(core)) == 0 and core.value == 3
    # the per-thread stream / device setters reach both libraries
    sets = stock._lib.standin_stream_sets()
    null_stream = ctypes.c_void_p(0)
    assert stock.bfStreamSet(ctypes.byref(null_stream)) == 0
    assert stock._lib.standin_stream_sets() == sets + 1


@pytest.mark.gpu
def test_a_gulp_through_the_overlaid_namespace(stock):
    """FDMT + transpose + reduce called through the (overlaid) stock namespace
    with plain BFarray structs: results equal the oracle."""
    from oracle import fdmt as ofdmt
    overlay.apply(stock, B200)
    rng = np.random.default_rng(4)
    nchan, ntime, md = 64, 3000, 40
    x = rng.integers(-128, 128, size=(nchan, ntime)).astype(np.int8)
    d_in = bf.asarray(x, space='cuda')
    d_out = bf.asarray(np.full((md, ntime), -7.0, np.float32), space='cuda')
    plan = ctypes.c_void_p()
    assert stock.bfFdmtCreate(ctypes.byref(plan)) == 0
    assert stock.bfFdmtInit(plan, nchan, md, 1000., 400. / nchan, -2.0, 2, None, None) == 0
    assert stock.bfFdmtExecute(plan, d_in.as_BFarray(), d_out.as_BFarray(), 0, None, None) == 0
    bf.device.stream_synchronize()
    assert stock.bfFdmtDestroy(plan) == 0
    want = np.full((md, ntime), -7.0, np.float32)
    ofdmt.fdmt(x, md, 1000., 400. / nchan, out=want)
    assert np.array_equal(np.asarray(d_out.copy('system')).view(np.uint32), want.view(np.uint32))
    a = rng.normal(size=(128, 64)).astype(np.float32)
    d_a = bf.asarray(a, space='cuda')
    d_t = bf.empty((64, 128), 'f32', 'cuda')
    axes = (ctypes.c_int * 2)(1, 0)
    assert stock.bfTranspose(d_a.as_BFarray(), d_t.as_BFarray(), axes) == 0
    d_r = bf.empty((64, 32), 'f32', 'cuda')
    assert stock.bfReduce(d_t.as_BFarray(), d_r.as_BFarray(), 0) == 0
    bf.device.stream_synchronize()
    np.testing.assert_array_equal(np.asarray(d_r.copy('system')), a.T.reshape(64, 32, 4).sum(-1, dtype=np.float32))
