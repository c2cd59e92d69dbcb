"""The product's host FDMT plan (bifrost_b200/csrc/fdmt_plan.hpp, exposed by
bfFdmtPlanQuery) must produce exactly the tables of the oracle's restatement of
src/fdmt.cu:338-530 -- they define which samples are summed."""
import ctypes

import numpy as np
import pytest

from bifrost_b200.libbifrost import _bf, _check
from oracle.fdmt import FdmtPlan

CASES = [
    # nchan, max_delay, f0, df        (test/test_fdmt.py:66-103 shapes + BASELINE configs)
    (128, 200, 1000., 400. / 128),
    (2, 20, 1000., 200.),
    (32, 2, 1000., 400. / 32),
    (32, 1, 1000., 400. / 32),
    (33, 65, 1000., 400. / 33),
    (4096, 794, 1000., 400. / 4096),     # BASELINE config 2
    (4096, 204, 1000., 400. / 4096),
    (512, 130, 1350., 400. / 4096),      # one sub-band of config 5
    (200, 37, 1400., -1.5),              # negative df: reversed band
    (1000, 300, 60., 0.024),             # LWA-like low band
    (7, 9, 1000., 10.), (5, 3, 300., 7.), (3, 50, 1200., 100.),
]


def query(nchan, md, f0, df, ex=-2.0):
    n = ctypes.c_int()
    _check(_bf.bfFdmtPlanQuery(nchan, md, f0, df, ex, -1, ctypes.byref(n), None))
    out = []
    for s in range(n.value):
        nr = ctypes.c_int()
        _check(_bf.bfFdmtPlanQuery(nchan, md, f0, df, ex, s, ctypes.byref(nr), None))
        cnt = nchan if s == 0 else nr.value
        rows = (ctypes.c_int * (3 * cnt))()
        _check(_bf.bfFdmtPlanQuery(nchan, md, f0, df, ex, s, ctypes.byref(nr), rows))
        out.append((nr.value, np.array(rows).reshape(cnt, 3)))
    return out


@pytest.mark.parametrize("nchan,md,f0,df", CASES)
def test_plan_tables_match_oracle(nchan, md, f0, df):
    p = FdmtPlan(nchan, md, f0, df)
    q = query(nchan, md, f0, df)
    assert len(q) == p.nstep
    assert p.nrow[-1] == md
    for s in range(p.nstep):
        assert q[s][0] == p.nrow[s]
        if s == 0:
            np.testing.assert_array_equal(q[0][1][:, 0], p.row_offsets[0][:-1])
            np.testing.assert_array_equal(q[0][1][:, 1], np.diff(p.row_offsets[0]))
        else:
            np.testing.assert_array_equal(q[s][1][:, :2], p.srcrows[s])
            np.testing.assert_array_equal(q[s][1][:, 2], p.delays[s])


def test_baseline_config2_row_counts():
    """the 13-step plan of the headline config."""
    p = FdmtPlan(4096, 794, 1000., 400. / 4096)
    assert p.nrow == [8192, 4096, 2048, 1456, 1125, 958, 877, 834, 814, 804, 798, 796, 794]


def test_source_rows_stay_inside_parent_bands():
    for (nchan, md, f0, df) in CASES:
        p = FdmtPlan(nchan, md, f0, df)
        for s in range(1, p.nstep):
            src = p.srcrows[s]
            assert src.max() < p.nrow[s - 1]
            assert (p.delays[s] >= 0).all() and p.delays[s].max() < max(md, 1)
